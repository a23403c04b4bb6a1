//! `Postnet` replaces `postnet: ort::Session` of xd-tts's `Tacotron2` (/root/reference src/tacotron2/mod.rs:146; loaded
//! at :256-259, run at :344-357, output name "mel_outputs_postnet"); `Decoder` replaces the `decoder` session and the
//! per-frame loop around it (:145, :251-254, :272-342).  Both are thin wrappers over include/xdtts_b200.h.
use anyhow::{bail, Result};
use ndarray::{Array2, ArrayView2};
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_float, c_int, c_ulonglong};
use std::path::Path;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct XdttsPostnetOpts {
    pub precision: c_int, // 0 bf16x3 on the tensor cores (fp32-class, default), 1 single bf16 pass, 2 fp32 CUDA cores
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct XdttsDecoderOpts {
    pub gate_threshold: c_float, // 0: 0.6 (src/tacotron2/mod.rs:279)
    pub max_steps: c_int,        // 0: 1000 (src/tacotron2/mod.rs:280)
    pub prenet_dropout: c_int,   // 0: on (the exported graph draws a mask per step), 1: off
    pub seed: c_ulonglong,
}
#[repr(C)]
pub struct XdttsPostnet {
    _private: [u8; 0],
}
#[repr(C)]
pub struct XdttsDecoder {
    _private: [u8; 0],
}

extern "C" {
    fn xdtts_postnet_create_from_onnx(path: *const c_char, opts: *const XdttsPostnetOpts, device: c_int, out: *mut *mut XdttsPostnet) -> c_int;
    fn xdtts_postnet_infer(h: *mut XdttsPostnet, mel: *const c_float, t: c_int, out: *mut c_float) -> c_int;
    fn xdtts_postnet_destroy(h: *mut XdttsPostnet);
    fn xdtts_decoder_create_from_onnx(path: *const c_char, opts: *const XdttsDecoderOpts, device: c_int, out: *mut *mut XdttsDecoder) -> c_int;
    fn xdtts_decoder_max_steps(h: *const XdttsDecoder) -> c_int;
    fn xdtts_decoder_infer_batch(h: *mut XdttsDecoder, memory: *const *const c_float, processed: *const *const c_float, t_enc: c_int,
                                 unpadded_len: *const c_int, b: c_int, out_mels: *const *mut c_float, n_frames: *mut c_int,
                                 out_gates: *const *mut c_float, out_align: *const *mut c_float) -> c_int;
    fn xdtts_decoder_destroy(h: *mut XdttsDecoder);
    fn xdtts_last_error() -> *const c_char;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(xdtts_last_error()).to_string_lossy().into_owned() }
}
fn device() -> c_int {
    std::env::var("XDTTS_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0)
}

pub struct Postnet {
    h: *mut XdttsPostnet,
}
unsafe impl Send for Postnet {}
unsafe impl Sync for Postnet {}

impl Postnet {
    /// src/tacotron2/mod.rs:256-259: `commit_from_file(path.join("postnet.onnx"))` -- the library reads the Conv /
    /// BatchNormalization initializers itself and checks that the graph is the Tacotron2 postnet.
    pub fn from_onnx(path: impl AsRef<Path>) -> Result<Self> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes())?;
        let mut h = std::ptr::null_mut();
        let rc = unsafe { xdtts_postnet_create_from_onnx(c.as_ptr(), std::ptr::null(), device(), &mut h) };
        if rc != 0 {
            bail!("xdtts_postnet_create_from_onnx failed ({rc}): {}", last_error());
        }
        Ok(Self { h })
    }

    /// src/tacotron2/mod.rs:347-355: `postnet.run(inputs![mel])["mel_outputs_postnet"]`, `[80, T]` in and out
    pub fn run(&self, mel: &Array2<f32>) -> Result<Array2<f32>> {
        let mel = mel.as_standard_layout();
        let (c, t) = mel.dim();
        let mut out = Array2::<f32>::zeros((c, t));
        let rc = unsafe { xdtts_postnet_infer(self.h, mel.as_ptr(), t as c_int, out.as_mut_ptr()) };
        if rc != 0 {
            bail!("xdtts_postnet_infer failed ({rc}): {}", last_error());
        }
        Ok(out)
    }
}
impl Drop for Postnet {
    fn drop(&mut self) {
        unsafe { xdtts_postnet_destroy(self.h) }
    }
}

pub struct Decoder {
    h: *mut XdttsDecoder,
}
unsafe impl Send for Decoder {}
unsafe impl Sync for Decoder {}

impl Decoder {
    /// src/tacotron2/mod.rs:251-254: `commit_from_file(path.join("decoder_iter.onnx"))`
    pub fn from_onnx(path: impl AsRef<Path>) -> Result<Self> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes())?;
        let mut h = std::ptr::null_mut();
        let rc = unsafe { xdtts_decoder_create_from_onnx(c.as_ptr(), std::ptr::null(), device(), &mut h) };
        if rc != 0 {
            bail!("xdtts_decoder_create_from_onnx failed ({rc}): {}", last_error());
        }
        Ok(Self { h })
    }

    /// The whole loop of `run_decoder` (src/tacotron2/mod.rs:272-342) in one call: encoder outputs in, `[80, T]` out
    /// (the layout of `mel_spec.t()` at :345).  `unpadded_len` is the mask boundary of `DecoderState::new` (:228-229).
    pub fn run(&self, memory: ArrayView2<f32>, processed: ArrayView2<f32>, unpadded_len: usize) -> Result<Array2<f32>> {
        let (memory, processed) = (memory.as_standard_layout(), processed.as_standard_layout());
        let cap = unsafe { xdtts_decoder_max_steps(self.h) } as usize;
        let mut buf = vec![0f32; 80 * cap];
        let (mut n, len) = (0 as c_int, unpadded_len as c_int);
        let rc = unsafe {
            xdtts_decoder_infer_batch(self.h, &memory.as_ptr(), &processed.as_ptr(), memory.nrows() as c_int, &len, 1, &buf.as_mut_ptr(),
                                      &mut n, std::ptr::null(), std::ptr::null())
        };
        if rc != 0 {
            bail!("xdtts_decoder_infer_batch failed ({rc}): {}", last_error());
        }
        buf.truncate(80 * n as usize);
        Ok(Array2::from_shape_vec((80, n as usize), buf)?)
    }
}
impl Drop for Decoder {
    fn drop(&mut self) {
        unsafe { xdtts_decoder_destroy(self.h) }
    }
}
