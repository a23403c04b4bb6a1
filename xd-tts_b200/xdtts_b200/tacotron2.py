"""Mirror of the postnet tail of the reference's `tacotron2` module
(/root/reference src/tacotron2/mod.rs): `Tacotron2::load` opens `postnet.onnx` as an ort::Session
(:256-259) and `run_decoder` ends by running it on the decoder mel `[1, 80, T]` and taking
"mel_outputs_postnet" (:344-357).  Here the session is `Postnet`: the same five
Conv1d(k=5) + BatchNorm layers (tanh after the first four) plus the residual add, executed by the
CUDA library; weights are passed as arrays (the ONNX initializers) because the LFS object is not
available.

    post = tacotron2.Postnet.from_layers(layers)          # layers: list of dicts w,b,gamma,beta,mean,var
    mel_outputs_postnet = post.run(mel)                   # [80, T] -> [80, T]
    audio = tacotron2.infer_tail(post, vocoder, mel)      # postnet -> lift -> Griffin-Lim on the device

There is no CPU path: every call goes through include/xdtts_b200.h.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _ffi
from ._ffi import PostnetOpts, XdttsError, check, fptr, fptr_array, load_library

PRECISION_BF16X3, PRECISION_BF16, PRECISION_FP32 = 0, 1, 2
BN_EPS = 1e-5   # torch.nn.BatchNorm1d default, as exported to ONNX


def _as_f32_2d(a, rows, what):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[0] != rows:
        raise XdttsError(_ffi.ERR_SHAPE, "%s must be [%d, T], got %s" % (what, rows, a.shape))
    return a


class Postnet:
    """Device-side postnet (the `postnet` session of Tacotron2, src/tacotron2/mod.rs:146,256-259)."""

    def __init__(self, handle, channels, precision):
        self._h = handle
        self.channels = list(channels)
        self.n_mels = self.channels[0]
        self.precision = precision

    @classmethod
    def from_layers(cls, layers, *, eps=BN_EPS, precision=PRECISION_BF16X3, device=0):
        """layers[i]: dict with 'w' [Cout, Cin, 5], optional 'b' [Cout] and the BatchNorm arrays
        'gamma', 'beta', 'mean', 'var' [Cout] (all four or none)."""
        lib = load_library()
        n = len(layers)
        ws = [np.ascontiguousarray(l["w"], dtype=np.float32) for l in layers]
        for i, w in enumerate(ws):
            if w.ndim != 3:
                raise XdttsError(_ffi.ERR_SHAPE, "layers[%d]['w'] must be [Cout, Cin, k]" % i)
            if i and w.shape[1] != ws[i - 1].shape[0]:
                raise XdttsError(_ffi.ERR_SHAPE, "layers[%d] takes %d channels, layers[%d] makes %d" % (i, w.shape[1], i - 1, ws[i - 1].shape[0]))
        if not ws:
            raise XdttsError(_ffi.ERR_BAD_ARG, "no layers")
        channels = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        ksize = ws[0].shape[2]
        if any(w.shape[2] != ksize for w in ws):
            raise XdttsError(_ffi.ERR_SHAPE, "all layers must share one kernel size")

        keep = []   # keep the float32 copies alive across the call

        def column(name):
            arr = (ctypes.POINTER(ctypes.c_float) * n)()
            for i, l in enumerate(layers):
                v = l.get(name)
                if v is None:
                    arr[i] = None
                    continue
                v = np.ascontiguousarray(v, dtype=np.float32)
                if v.shape != (channels[i + 1],):
                    raise XdttsError(_ffi.ERR_SHAPE, "layers[%d]['%s'] must be [%d]" % (i, name, channels[i + 1]))
                keep.append(v)
                arr[i] = fptr(v)
            return arr

        ch = (ctypes.c_int * (n + 1))(*channels)
        opts = PostnetOpts(int(precision))
        h = ctypes.c_void_p()
        check(lib.xdtts_postnet_create(n, ch, int(ksize), fptr_array(ws), column("b"), column("gamma"), column("beta"),
                                       column("mean"), column("var"), float(eps), ctypes.byref(opts), int(device),
                                       ctypes.byref(h)))
        return cls(h, channels, int(precision))

    @classmethod
    def load(cls, path, *, precision=PRECISION_BF16X3, device=0):
        """The postnet part of Tacotron2::load (src/tacotron2/mod.rs:256-259): `path` is postnet.onnx."""
        lib = load_library()
        opts = PostnetOpts(int(precision))
        h = ctypes.c_void_p()
        check(lib.xdtts_postnet_create_from_onnx(str(path).encode(), ctypes.byref(opts), int(device), ctypes.byref(h)))
        layers = read_onnx_postnet(path)
        return cls(h, [layers[0]["w"].shape[1]] + [l["w"].shape[0] for l in layers], int(precision))

    def close(self):
        if self._h:
            load_library().xdtts_postnet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, mel):
        """postnet.run(inputs![mel])["mel_outputs_postnet"] (src/tacotron2/mod.rs:347-355): [C, T] -> [C, T]."""
        m = _as_f32_2d(mel, self.n_mels, "mel")
        out = np.empty_like(m)
        check(load_library().xdtts_postnet_infer(self._h, fptr(m), m.shape[1], fptr(out)))
        return out

    def run_batch(self, mels):
        ms = [_as_f32_2d(a, self.n_mels, "mel") for a in mels]
        outs = [np.empty_like(m) for m in ms]
        ts = (ctypes.c_int * len(ms))(*[m.shape[1] for m in ms])
        check(load_library().xdtts_postnet_infer_batch(self._h, fptr_array(ms), ts, len(ms), fptr_array(outs)))
        return outs

    def plan(self, frame_counts):
        return PostnetPlan(self, frame_counts)


def write_npy(path, spectrogram):
    """ndarray_npy::write_npy(path, &spectrogram) (src/lib.rs:132): the --output-spectrogram dump."""
    a = np.ascontiguousarray(spectrogram, dtype=np.float32)
    if a.ndim != 2:
        raise XdttsError(_ffi.ERR_SHAPE, "spectrogram must be 2-D")
    check(load_library().xdtts_npy_write_f32(str(path).encode(), fptr(a), a.shape[0], a.shape[1]))


def read_npy(path):
    """A mel dumped by the reference's `xd_tts --output-spectrogram` (or by numpy) -> float32 [rows, cols]."""
    lib = load_library()
    r, c = ctypes.c_int(), ctypes.c_int()
    check(lib.xdtts_npy_read_f32(str(path).encode(), None, 0, ctypes.byref(r), ctypes.byref(c)))
    out = np.empty((r.value, c.value), dtype=np.float32)
    check(lib.xdtts_npy_read_f32(str(path).encode(), fptr(out), out.size, ctypes.byref(r), ctypes.byref(c)))
    return out


def read_onnx_postnet(path):
    """Layers of a postnet.onnx as `from_layers` takes them (host only; the library's own ONNX reader)."""
    lib = load_library()
    m = ctypes.c_void_p()
    check(lib.xdtts_onnx_postnet_open(str(path).encode(), ctypes.byref(m)))
    try:
        n = lib.xdtts_onnx_postnet_n_layers(m)
        if n < 0:
            check(n)
        layers = []
        for i in range(n):
            co, ci, k, hb, hbn = (ctypes.c_int() for _ in range(5))
            eps = ctypes.c_float()
            check(lib.xdtts_onnx_postnet_layer_info(m, i, ctypes.byref(co), ctypes.byref(ci), ctypes.byref(k), ctypes.byref(hb),
                                                    ctypes.byref(hbn), ctypes.byref(eps)))
            layer = {"eps": eps.value}
            names = [("w", (co.value, ci.value, k.value), True), ("b", (co.value,), bool(hb.value))]
            names += [(nm, (co.value,), bool(hbn.value)) for nm in ("gamma", "beta", "mean", "var")]
            for which, (nm, shape, present) in enumerate(names):
                if present:
                    a = np.empty(shape, dtype=np.float32)
                    check(lib.xdtts_onnx_postnet_layer_copy(m, i, which, fptr(a)))
                    layer[nm] = a
            layers.append(layer)
        return layers
    finally:
        lib.xdtts_onnx_postnet_close(m)


class PostnetPlan:
    """Device-resident postnet batch; `run(feed=gl_plan)` hands the result to a vocoder plan in HBM."""

    def __init__(self, post, frame_counts):
        self.post = post
        self.ts = [int(t) for t in frame_counts]
        p = ctypes.c_void_p()
        t_arr = (ctypes.c_int * len(self.ts))(*self.ts)
        check(load_library().xdtts_postnet_plan_create(post._h, t_arr, len(self.ts), ctypes.byref(p)))
        self._p = p

    def close(self):
        if self._p:
            load_library().xdtts_postnet_plan_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, mels):
        ms = [_as_f32_2d(a, self.post.n_mels, "mel") for a in mels]
        if [m.shape[1] for m in ms] != self.ts:
            raise XdttsError(_ffi.ERR_SHAPE, "inputs do not match the plan's frame counts")
        check(load_library().xdtts_postnet_plan_upload(self._p, fptr_array(ms)))

    def upload_ptrs(self, ptr_array):
        check(load_library().xdtts_postnet_plan_upload(self._p, ptr_array))

    def run(self, feed=None):
        """-> device milliseconds.  feed: a griffin_lim.GlPlan of the same frame counts."""
        ms = ctypes.c_float()
        check(load_library().xdtts_postnet_plan_run(self._p, None if feed is None else feed._p, ctypes.byref(ms)))
        return ms.value

    def download(self):
        outs = [np.empty((self.post.n_mels, t), dtype=np.float32) for t in self.ts]
        check(load_library().xdtts_postnet_plan_download(self._p, fptr_array(outs)))
        return outs


def infer_tail_batch(post, vocoder, mels, init_phases=None, return_mels=False):
    """The tail of XdTts::infer for B utterances (src/lib.rs:123 postnet + :141 vocoder.infer):
    decoder mels -> postnet -> mel-to-linear lift -> Griffin-Lim, intermediates stay in HBM."""
    lib = load_library()
    ms = [_as_f32_2d(a, post.n_mels, "mel") for a in mels]
    ts = [m.shape[1] for m in ms]
    waves = [np.empty(max(vocoder.hop * (t - 1), 0), dtype=np.float32) for t in ts]
    out_mels = [np.empty_like(m) for m in ms] if return_mels else None
    phs = None
    if init_phases is not None:
        phs = [_as_f32_2d(a, vocoder.k_bins, "init_phase") for a in init_phases]
        if [a.shape[1] for a in phs] != ts:
            raise XdttsError(_ffi.ERR_SHAPE, "init_phases do not match the inputs' frame counts")
    t_arr = (ctypes.c_int * len(ts))(*ts)
    check(lib.xdtts_tail_infer_batch(post._h, vocoder._h, fptr_array(ms), t_arr, len(ms),
                                     None if phs is None else fptr_array(phs),
                                     None if out_mels is None else fptr_array(out_mels), fptr_array(waves)))
    return (waves, out_mels) if return_mels else waves


def infer_tail(post, vocoder, mel, init_phase=None):
    return infer_tail_batch(post, vocoder, [mel], None if init_phase is None else [init_phase])[0]
