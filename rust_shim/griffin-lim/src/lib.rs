//! `griffin-lim` as xd-tts uses it (call sites: /root/reference src/tacotron2/mod.rs:67-68,453,456 and
//! src/lib.rs:5,35,51,141-155), implemented on the B200 by `libxdtts_b200.so` (include/xdtts_b200.h).
//!
//! Only the items xd-tts touches exist: `mel::create_mel_filter_bank`, `GriffinLim::new`, `GriffinLim::infer`.
//! Differences from the crate this replaces, all reachable through `GriffinLim::with_options`:
//!  * the mel -> linear lift is `max(0, pinv(basis) . exp(mel)) ^ power` by default; `Lift::Nnls` (or the cargo feature
//!    `nnls-lift`, or XDTTS_B200_LIFT=nnls) selects the non-negative least-squares lift librosa's `mel_to_stft` performs;
//!  * n_fft = 2 (K - 1) must be a power of two in [64, 4096] (anything else fails in `new` with the library's message); any
//!    `noverlap` in [0, n_fft) is taken.  hop == n_fft / 4 at n_fft 512 / 1024 / 2048 -- what `create_griffin_lim` configures,
//!    1024 / 768 -- runs the fused one-launch-per-iteration kernel, every other geometry the un-fused kernels (same results
//!    contract, ~6x the memory traffic).
use anyhow::{bail, Result};
use ndarray::{Array1, Array2};
use std::ffi::CStr;
use std::os::raw::{c_char, c_float, c_int, c_ulonglong};

/// include/xdtts_b200.h `xdtts_gl_opts` (all zero = the library's defaults)
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct XdttsGlOpts {
    pub delog: c_int,        // 0 exp (Tacotron2 ln-mel), 1 10^x, 2 none
    pub pad_mode: c_int,     // 0 reflect (librosa 0.9.2), 1 constant
    pub normalise: c_int,    // 0 peak-normalise to [-1, 1] (src/lib.rs:155 scales by i16::MAX), 1 none
    pub run_frames: c_int,   // 0 auto
    pub seed: c_ulonglong,
    pub persistent: c_int,
    pub lift: c_int,         // 0 pseudo-inverse, 1 NNLS
    pub nnls_iters: c_int,
    pub fixed_seed: c_int,   // 0: a new phase field per call (as the crate: rand per call), 1: reproducible
    pub exponent: c_int,     // 0: S = x^power, 1: S = x^(1/power)
}

#[repr(C)]
pub struct XdttsGl {
    _private: [u8; 0],
}

extern "C" {
    fn xdtts_mel_filter_bank(sr: c_float, n_fft: c_int, n_mels: c_int, fmin: c_float, fmax: c_float, out: *mut c_float) -> c_int;
    fn xdtts_gl_create(basis: *const c_float, n_mels: c_int, k: c_int, noverlap: c_int, power: c_float, n_iter: c_int,
                       momentum: c_float, opts: *const XdttsGlOpts, device: c_int, out: *mut *mut XdttsGl) -> c_int;
    fn xdtts_gl_destroy(h: *mut XdttsGl);
    fn xdtts_gl_out_len(h: *const XdttsGl, t: c_int) -> c_int;
    fn xdtts_gl_infer(h: *mut XdttsGl, mel: *const c_float, t: c_int, init_phase: *const c_float, out: *mut c_float, out_len: c_int) -> c_int;
    fn xdtts_last_error() -> *const c_char;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(xdtts_last_error()).to_string_lossy().into_owned() }
}

pub mod mel {
    use super::*;

    /// Slaney-scale, Slaney-normalised triangular filters, `[n_mels, n_fft / 2 + 1]` -- the signature of the call at
    /// xd-tts src/tacotron2/mod.rs:453: `create_mel_filter_bank(22050.0, 1024, 80, 0.0, Some(8000.0))`.
    pub fn create_mel_filter_bank(sample_rate: f32, n_fft: usize, n_mels: usize, fmin: f32, fmax: Option<f32>) -> Array2<f32> {
        let mut out = Array2::<f32>::zeros((n_mels, n_fft / 2 + 1));
        let rc = unsafe { xdtts_mel_filter_bank(sample_rate, n_fft as c_int, n_mels as c_int, fmin, fmax.unwrap_or(-1.0), out.as_mut_ptr()) };
        assert_eq!(rc, 0, "xdtts_mel_filter_bank: {}", last_error());
        out
    }
}

#[derive(Clone, Copy, PartialEq, Eq)]
pub enum Lift {
    PseudoInverse,
    Nnls,
}

pub struct GriffinLim {
    h: *mut XdttsGl,
    n_mels: usize,
}

// the library serialises calls on one handle (a mutex inside xdtts_gl), so sharing &GriffinLim between threads is sound
unsafe impl Send for GriffinLim {}
unsafe impl Sync for GriffinLim {}

impl GriffinLim {
    /// xd-tts src/tacotron2/mod.rs:456: `GriffinLim::new(mel_basis, 1024 - 256, 1.7, 30, 0.99)?`
    pub fn new(mel_basis: Array2<f32>, noverlap: usize, power: f32, iter: usize, momentum: f32) -> Result<Self> {
        let mut opts = XdttsGlOpts::default();
        let env_nnls = std::env::var("XDTTS_B200_LIFT").map(|v| v.eq_ignore_ascii_case("nnls")).unwrap_or(false);
        if cfg!(feature = "nnls-lift") || env_nnls {
            opts.lift = 1;
        }
        if let Some(seed) = std::env::var("XDTTS_B200_SEED").ok().and_then(|s| s.parse().ok()) {
            opts.seed = seed;
        }
        Self::with_options(mel_basis, noverlap, power, iter, momentum, opts)
    }

    pub fn with_options(mel_basis: Array2<f32>, noverlap: usize, power: f32, iter: usize, momentum: f32, opts: XdttsGlOpts) -> Result<Self> {
        let basis = mel_basis.as_standard_layout();
        let (n_mels, k) = basis.dim();
        let device = std::env::var("XDTTS_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            xdtts_gl_create(basis.as_ptr(), n_mels as c_int, k as c_int, noverlap as c_int, power, iter as c_int, momentum, &opts, device, &mut h)
        };
        if rc != 0 {
            bail!("xdtts_gl_create failed ({rc}): {}", last_error());
        }
        Ok(Self { h, n_mels })
    }

    /// xd-tts src/lib.rs:141: `self.vocoder.infer(&spectrogram)?` -- `[n_mels, T]` ln-mel in, `hop * (T - 1)` samples
    /// out, peak-normalised to [-1, 1] (the caller multiplies by i16::MAX, src/lib.rs:155)
    pub fn infer(&self, mel: &Array2<f32>) -> Result<Array1<f32>> {
        let mel = mel.as_standard_layout();
        let (rows, t) = mel.dim();
        if rows != self.n_mels {
            bail!("mel has {rows} rows, the vocoder was built for {}", self.n_mels);
        }
        let n = unsafe { xdtts_gl_out_len(self.h, t as c_int) };
        if n < 0 {
            bail!("xdtts_gl_out_len failed ({n}): {}", last_error());
        }
        let mut out = Array1::<f32>::zeros(n as usize);
        let rc = unsafe { xdtts_gl_infer(self.h, mel.as_ptr(), t as c_int, std::ptr::null(), out.as_mut_ptr(), n) };
        if rc != 0 {
            bail!("xdtts_gl_infer failed ({rc}): {}", last_error());
        }
        Ok(out)
    }
}

impl Drop for GriffinLim {
    fn drop(&mut self) {
        unsafe { xdtts_gl_destroy(self.h) }
    }
}
