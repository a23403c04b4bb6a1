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
from ._ffi import DecoderOpts, DecoderWeights, PostnetOpts, XdttsError, check, fptr, fptr_array, load_library

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


GATE_THRESHOLD, MAX_DECODER_STEPS = 0.6, 1000   # src/tacotron2/mod.rs:279-280
_DECODER_SHAPES = dict(
    prenet1=(256, 80), prenet2=(256, 256), att_w_ih=(4096, 768), att_w_hh=(4096, 1024), att_b_ih=(4096,), att_b_hh=(4096,),
    query=(128, 1024), v=(128,), loc_conv=(32, 2, 31), loc_dense=(128, 32), dec_w_ih=(4096, 1536), dec_w_hh=(4096, 1024),
    dec_b_ih=(4096,), dec_b_hh=(4096,), proj_w=(80, 1536), proj_b=(80,), gate_w=(1536,), gate_b=(1,))


class Decoder:
    """Device-side decoder loop: the `decoder` session of Tacotron2 plus the loop of `run_decoder`
    (src/tacotron2/mod.rs:145,251-254,272-342).  One persistent kernel runs all steps; `run` / `run_batch`
    take the encoder outputs and return the spectrogram `[80, T]` the postnet consumes (:345)."""

    def __init__(self, handle, max_steps):
        self._h = handle
        self.max_steps = max_steps

    @classmethod
    def from_weights(cls, weights, *, gate_threshold=GATE_THRESHOLD, max_steps=MAX_DECODER_STEPS, prenet_dropout=True, seed=0,
                     device=0):
        """weights: dict of float32 arrays in the PyTorch layouts named in include/xdtts_b200.h
        (xdtts_decoder_weights); gate_w may be [1, 1536]."""
        lib = load_library()
        keep, w = [], DecoderWeights()
        for name, shape in _DECODER_SHAPES.items():
            if name not in weights:
                raise XdttsError(_ffi.ERR_BAD_ARG, "decoder weights: '%s' is missing" % name)
            a = np.ascontiguousarray(weights[name], dtype=np.float32)
            if a.size != int(np.prod(shape)):
                raise XdttsError(_ffi.ERR_SHAPE, "decoder weights: '%s' must be %s, got %s" % (name, shape, a.shape))
            keep.append(a)
            setattr(w, name, fptr(a))
        opts = DecoderOpts(float(gate_threshold), int(max_steps), 0 if prenet_dropout else 1, int(seed))
        h = ctypes.c_void_p()
        check(lib.xdtts_decoder_create(ctypes.byref(w), ctypes.byref(opts), int(device), ctypes.byref(h)))
        return cls(h, lib.xdtts_decoder_max_steps(h))

    @classmethod
    def from_onnx(cls, path, *, gate_threshold=GATE_THRESHOLD, max_steps=MAX_DECODER_STEPS, prenet_dropout=True, seed=0, device=0):
        """The decoder part of Tacotron2::load (src/tacotron2/mod.rs:251-254): `path` is decoder_iter.onnx."""
        lib = load_library()
        opts = DecoderOpts(float(gate_threshold), int(max_steps), 0 if prenet_dropout else 1, int(seed))
        h = ctypes.c_void_p()
        check(lib.xdtts_decoder_create_from_onnx(str(path).encode(), ctypes.byref(opts), int(device), ctypes.byref(h)))
        return cls(h, lib.xdtts_decoder_max_steps(h))

    def close(self):
        if self._h:
            load_library().xdtts_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_batch(self, memories, processed_memories, unpadded_lens, return_aux=False):
        """run_decoder for B utterances: memories[b] [t_enc, 512], processed_memories[b] [t_enc, 128] (one common
        t_enc), unpadded_lens[b] -> list of [80, T_b] float32 (+ gate logits [T_b], alignments [T_b, t_enc])."""
        lib = load_library()
        mem = [np.ascontiguousarray(a, dtype=np.float32) for a in memories]
        pm = [np.ascontiguousarray(a, dtype=np.float32) for a in processed_memories]
        n = len(mem)
        if n == 0 or len(pm) != n or len(unpadded_lens) != n:
            raise XdttsError(_ffi.ERR_BAD_ARG, "memories, processed_memories and unpadded_lens must have one entry per utterance")
        t_enc = mem[0].shape[0]
        for a, b in zip(mem, pm):
            if a.ndim != 2 or a.shape != (t_enc, 512) or b.shape != (t_enc, 128):
                raise XdttsError(_ffi.ERR_SHAPE, "memory must be [t_enc, 512] and processed_memory [t_enc, 128] with one common t_enc")
        outs = [np.empty(80 * self.max_steps, dtype=np.float32) for _ in range(n)]
        gates = [np.empty(self.max_steps, dtype=np.float32) for _ in range(n)] if return_aux else None
        aligns = [np.empty(self.max_steps * t_enc, dtype=np.float32) for _ in range(n)] if return_aux else None
        lens = (ctypes.c_int * n)(*[int(x) for x in unpadded_lens])
        nf = (ctypes.c_int * n)()
        check(lib.xdtts_decoder_infer_batch(self._h, fptr_array(mem), fptr_array(pm), t_enc, lens, n, fptr_array(outs), nf,
                                            None if gates is None else fptr_array(gates),
                                            None if aligns is None else fptr_array(aligns)))
        mels = [o[:80 * nf[b]].reshape(80, nf[b]).copy() for b, o in enumerate(outs)]
        if return_aux:
            return (mels, [g[:nf[b]].copy() for b, g in enumerate(gates)],
                    [a[:nf[b] * t_enc].reshape(nf[b], t_enc).copy() for b, a in enumerate(aligns)])
        return mels

    def run(self, memory, processed_memory, unpadded_len):
        """Tacotron2::run_decoder up to the postnet (src/tacotron2/mod.rs:272-345): -> [80, T]."""
        return self.run_batch([memory], [processed_memory], [unpadded_len])[0]

    def info(self, nb, t_enc):
        """-> dict(weight_bytes_per_step, smem_resident_bytes, handovers_per_step, polled_cells) of a launch for nb utterances"""
        a = (ctypes.c_longlong * 4)()
        check(load_library().xdtts_decoder_info(self._h, int(nb), int(t_enc), a))
        return dict(weight_bytes_per_step=a[0], smem_resident_bytes=a[1], handovers_per_step=a[2], polled_cells=bool(a[3]))

    def last_timing(self):
        """-> (device milliseconds of the decoder kernel, steps executed) of the last call"""
        ms, steps = ctypes.c_float(), ctypes.c_int()
        check(load_library().xdtts_decoder_last_timing(self._h, ctypes.byref(ms), ctypes.byref(steps)))
        return ms.value, steps.value


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


def synthesize_batch(decoder, post, vocoder, memories, processed_memories, unpadded_lens, init_phases=None, return_mels=False):
    """Everything `XdTts::infer` does after the encoder session, for B utterances (src/lib.rs:123 `model.infer`
    = decoder loop + postnet, src/tacotron2/mod.rs:272-357, then :141 `vocoder.infer`): encoder outputs ->
    decoder loop -> postnet -> mel-to-linear lift -> Griffin-Lim -> float32 samples.  The frame counts come out of
    the decoder's stop rule, so the tail runs as one batch of ragged lengths."""
    mels = decoder.run_batch(memories, processed_memories, unpadded_lens)
    short = [i for i, m in enumerate(mels) if m.shape[1] < 2]
    if short:
        raise XdttsError(_ffi.ERR_SHAPE, "utterances %s stopped after one frame: no samples to vocode" % short)
    return infer_tail_batch(post, vocoder, mels, init_phases, return_mels=return_mels)


DECODER_TENSORS = ("prenet1", "prenet2", "att_w_ih", "att_w_hh", "att_b_ih", "att_b_hh", "query", "v", "loc_conv", "loc_dense",
                   "dec_w_ih", "dec_w_hh", "dec_b_ih", "dec_b_hh", "proj_w", "proj_b", "gate_w", "gate_b")


def read_onnx_decoder(path):
    """(dims, weights) of a decoder_iter.onnx: the library's own ONNX reader, host only.  weights: flat float32 arrays in
    the order and layouts of include/xdtts_b200.h `xdtts_decoder_weights` (LSTM gate order i, f, g, o)."""
    lib = load_library()
    m = ctypes.c_void_p()
    check(lib.xdtts_onnx_decoder_open(str(path).encode(), ctypes.byref(m)))
    try:
        d = (ctypes.c_int * 10)()
        check(lib.xdtts_onnx_decoder_dims(m, d))
        dims = dict(zip(("n_mel", "prenet", "enc", "att_rnn", "dec_rnn", "att_dim", "loc_f", "loc_k", "lstm_form", "has_dropout"), list(d)))
        weights = {}
        for i, name in enumerate(DECODER_TENSORS):
            n = lib.xdtts_onnx_decoder_tensor(m, i, None, 0)
            if n < 0:
                check(int(n))
            a = np.empty(int(n), np.float32)
            got = lib.xdtts_onnx_decoder_tensor(m, i, fptr(a), int(n))
            if got < 0:
                check(int(got))
            weights[name] = a
        return dims, weights
    finally:
        lib.xdtts_onnx_decoder_close(m)
