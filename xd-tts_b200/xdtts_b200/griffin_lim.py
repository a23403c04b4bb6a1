"""Mirror of the `griffin_lim` crate surface that xd-tts uses (external crate griffin-lim 0.2.0,
call sites /root/reference src/tacotron2/mod.rs:67-68,453,456 and src/lib.rs:5,35,51,141):

    griffin_lim::mel::create_mel_filter_bank(f32, usize, usize, f32, Option<f32>) -> Array2<f32>
    griffin_lim::GriffinLim::new(Array2<f32>, usize, f32, usize, f32) -> Result<GriffinLim>
    GriffinLim::infer(&self, &Array2<f32>) -> Result<samples>

plus the batch / device-plan entry points the B200 back end adds.  Shapes and argument order are
the reference's; errors surface as XdttsError (anyhow::Error in the Rust shim).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _ffi
from ._ffi import GlOpts, XdttsError, check, fptr, fptr_array, load_library, sptr_array

DELOG_EXP, DELOG_POW10, DELOG_NONE = 0, 1, 2
PAD_REFLECT, PAD_CONSTANT = 0, 1
NORM_PEAK, NORM_NONE = 0, 1
LIFT_PINV, LIFT_NNLS = 0, 1


class mel:  # noqa: N801  (module-like namespace, as in griffin_lim::mel)
    @staticmethod
    def create_mel_filter_bank(sample_rate, n_fft, n_mels, fmin, fmax=None):
        """[n_mels, n_fft//2+1] float32 Slaney filterbank (call site src/tacotron2/mod.rs:453)."""
        lib = load_library()
        out = np.empty((int(n_mels), int(n_fft) // 2 + 1), dtype=np.float32)
        check(lib.xdtts_mel_filter_bank(float(sample_rate), int(n_fft), int(n_mels), float(fmin),
                                        -1.0 if fmax is None else float(fmax), fptr(out)))
        return out


def pinv(a):
    """Pseudo-inverse [cols, rows] of a row-major float32 matrix, as the lift uses it."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty((a.shape[1], a.shape[0]), dtype=np.float32)
    check(load_library().xdtts_pinv(fptr(a), a.shape[0], a.shape[1], fptr(out)))
    return out


def _as_f32_2d(a, rows, what):
    a = np.ascontiguousarray(a, dtype=np.float32)   # the shim's as_standard_layout()
    if a.ndim != 2 or a.shape[0] != rows:
        raise XdttsError(_ffi.ERR_SHAPE, "%s must be [%d, T], got %s" % (what, rows, a.shape))
    return a


class GriffinLim:
    """Owns the device-side vocoder state (pseudo-inverse, twiddles, cached plans)."""

    def __init__(self, handle, n_mels, k_bins, hop, n_iter):
        self._h = handle
        self.n_mels, self.k_bins, self.hop, self.n_iter = n_mels, k_bins, hop, n_iter

    @classmethod
    def new(cls, mel_basis, noverlap, power, iter, momentum, *, delog=DELOG_EXP, pad_mode=PAD_REFLECT,  # noqa: A002
            normalise=NORM_PEAK, seed=0, run_frames=0, persistent=False, lift=LIFT_PINV, nnls_iters=0, fixed_seed=False,
            exponent=0, device=0):
        """GriffinLim::new(mel_basis, noverlap, power, iter, momentum) (src/tacotron2/mod.rs:456)."""
        lib = load_library()
        basis = np.ascontiguousarray(mel_basis, dtype=np.float32)
        if basis.ndim != 2:
            raise XdttsError(_ffi.ERR_SHAPE, "mel_basis must be 2-D [n_mels, K]")
        opts = GlOpts(int(delog), int(pad_mode), int(normalise), int(run_frames), int(seed), int(bool(persistent)), int(lift),
                      int(nnls_iters), int(bool(fixed_seed)), int(exponent))
        h = ctypes.c_void_p()
        check(lib.xdtts_gl_create(fptr(basis), basis.shape[0], basis.shape[1], int(noverlap), float(power), int(iter),
                                  float(momentum), ctypes.byref(opts), int(device), ctypes.byref(h)))
        n_fft = 2 * (basis.shape[1] - 1)
        return cls(h, basis.shape[0], basis.shape[1], n_fft - int(noverlap), int(iter))

    def close(self):
        if self._h:
            load_library().xdtts_gl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def out_len(self, n_frames):
        rc = load_library().xdtts_gl_out_len(self._h, int(n_frames))
        if rc < 0:
            check(rc)
        return rc

    def pinv(self):
        out = np.empty((self.k_bins, self.n_mels), dtype=np.float32)
        check(load_library().xdtts_gl_get_pinv(self._h, fptr(out)))
        return out

    def infer(self, mel_spectrogram, init_phase=None):
        """GriffinLim::infer(&mel) (src/lib.rs:141): [n_mels, T] -> float32 samples [hop*(T-1)]."""
        lib = load_library()
        m = _as_f32_2d(mel_spectrogram, self.n_mels, "mel")
        t = m.shape[1]
        ph = None if init_phase is None else _as_f32_2d(init_phase, self.k_bins, "init_phase")
        if ph is not None and ph.shape[1] != t:
            raise XdttsError(_ffi.ERR_SHAPE, "init_phase has %d frames, mel has %d" % (ph.shape[1], t))
        n = self.hop * (t - 1) if t >= 1 else 0
        out = np.empty(max(n, 0), dtype=np.float32)
        check(lib.xdtts_gl_infer(self._h, fptr(m), t, None if ph is None else fptr(ph), fptr(out), n))
        return out

    def _batch(self, fn, ins, rows, what, init_phases):
        lib = load_library()
        ins = [_as_f32_2d(a, rows, what) for a in ins]
        ts = [a.shape[1] for a in ins]
        outs = [np.empty(max(self.hop * (t - 1), 0), dtype=np.float32) for t in ts]
        phs = None
        if init_phases is not None:
            phs = [_as_f32_2d(a, self.k_bins, "init_phase") for a in init_phases]
            if [a.shape[1] for a in phs] != ts:
                raise XdttsError(_ffi.ERR_SHAPE, "init_phases do not match the inputs' frame counts")
        t_arr = (ctypes.c_int * len(ts))(*ts)
        check(getattr(lib, fn)(self._h, fptr_array(ins), t_arr, len(ins), None if phs is None else fptr_array(phs),
                               fptr_array(outs)))
        return outs

    def infer_batch(self, mels, init_phases=None):
        """B utterances in one device pass (the reference loops over them, src/lib.rs:83-104)."""
        return self._batch("xdtts_gl_infer_batch", mels, self.n_mels, "mel", init_phases)

    def infer_batch_pcm16(self, mels, init_phases=None):
        """infer_batch + the caller's `(sample * i16::MAX as f32) as i16` loop (src/lib.rs:153-157) on the device."""
        lib = load_library()
        ins = [_as_f32_2d(a, self.n_mels, "mel") for a in mels]
        ts = [a.shape[1] for a in ins]
        outs = [np.empty(max(self.hop * (t - 1), 0), dtype=np.int16) for t in ts]
        phs = None
        if init_phases is not None:
            phs = [_as_f32_2d(a, self.k_bins, "init_phase") for a in init_phases]
            if [a.shape[1] for a in phs] != ts:
                raise XdttsError(_ffi.ERR_SHAPE, "init_phases do not match the inputs' frame counts")
        t_arr = (ctypes.c_int * len(ts))(*ts)
        check(lib.xdtts_gl_infer_batch_pcm16(self._h, fptr_array(ins), t_arr, len(ins),
                                             None if phs is None else fptr_array(phs), sptr_array(outs)))
        return outs

    def from_magnitude_batch(self, mags, init_phases=None):
        """Griffin-Lim proper from linear magnitudes [K, T] (no mel -> linear lift)."""
        return self._batch("xdtts_gl_from_mag_batch", mags, self.k_bins, "magnitude", init_phases)

    def plan(self, frame_counts):
        return GlPlan(self, frame_counts)

    def pipe(self, frame_counts, depth=2, postnet=None, want_mels=False):
        """Streaming vocoder for a sequence of batches of one shape (xdtts_pipe)."""
        return Pipe(self, frame_counts, depth, postnet, want_mels)


class GlPlan:
    """Device-resident batch (buffers + CUDA graph) for pipelines and benchmarking."""

    def __init__(self, voc, frame_counts):
        lib = load_library()
        self.voc = voc
        self.ts = [int(t) for t in frame_counts]
        p = ctypes.c_void_p()
        t_arr = (ctypes.c_int * len(self.ts))(*self.ts)
        check(lib.xdtts_gl_plan_create(voc._h, t_arr, len(self.ts), ctypes.byref(p)))
        self._p = p

    def close(self):
        if self._p:
            load_library().xdtts_gl_plan_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a = (ctypes.c_int * 4)()
        check(load_library().xdtts_gl_plan_info(self._p, a))
        return dict(n_runs=a[0], run_frames=a[1], ctas=a[2], total_frames=a[3])

    def lift_ms(self):
        """Device time of the mel -> linear step of the last run(RUN_NO_GRAPH) of this plan."""
        ms = ctypes.c_float()
        check(load_library().xdtts_gl_plan_lift_ms(self._p, ctypes.byref(ms)))
        return ms.value

    def time_lift(self, reps=20):
        """Device time per launch (ms) of the mel -> linear step, launched reps times back to back after a warm-up."""
        ms = ctypes.c_float()
        check(load_library().xdtts_gl_plan_time_lift(self._p, int(reps), ctypes.byref(ms)))
        return ms.value

    def is_persistent(self):
        """True when run() vocodes the plan with the single cooperative launch (all runs resident at once)."""
        rc = load_library().xdtts_gl_plan_is_persistent(self._p)
        if rc < 0:
            check(rc)
        return bool(rc)

    def upload(self, kind, arrays):
        rows = self.voc.n_mels if kind == 0 else self.voc.k_bins
        arrays = [_as_f32_2d(a, rows, "input") for a in arrays]
        if [a.shape[1] for a in arrays] != self.ts:
            raise XdttsError(_ffi.ERR_SHAPE, "inputs do not match the plan's frame counts")
        check(load_library().xdtts_gl_plan_upload(self._p, int(kind), fptr_array(arrays)))

    def upload_ptrs(self, kind, ptr_array):
        check(load_library().xdtts_gl_plan_upload(self._p, int(kind), ptr_array))

    def run(self, flags=0):
        """-> (ms_total, ms_iter, n_iter_launches), device times from CUDA events."""
        a, b, n = ctypes.c_float(), ctypes.c_float(), ctypes.c_int()
        check(load_library().xdtts_gl_plan_run(self._p, int(flags), ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
        return a.value, b.value, n.value

    def download(self):
        outs = [np.empty(self.voc.hop * (t - 1), dtype=np.float32) for t in self.ts]
        check(load_library().xdtts_gl_plan_download(self._p, fptr_array(outs)))
        return outs

    def download_pcm16(self):
        outs = [np.empty(self.voc.hop * (t - 1), dtype=np.int16) for t in self.ts]
        check(load_library().xdtts_gl_plan_download_pcm16(self._p, sptr_array(outs)))
        return outs

    def download_ptrs(self, ptr_array):
        check(load_library().xdtts_gl_plan_download(self._p, ptr_array))

    def peek(self, what):
        m, tt = self.voc.k_bins - 1, sum(self.ts)
        shape = {0: (tt, m), 1: (tt,), 2: (tt, m, 2)}[what]
        out = np.empty(shape, dtype=np.float32)
        check(load_library().xdtts_gl_plan_peek(self._p, what, fptr(out), out.size))
        return out


def _pinned(shape, dtype=np.float32):
    """numpy view of pinned host memory from xdtts_host_alloc -> (array, pointer to free)"""
    lib = load_library()
    n = int(np.prod(shape))
    ptr = lib.xdtts_host_alloc(max(n, 1) * 4)
    if not ptr:
        raise MemoryError("xdtts_host_alloc(%d bytes)" % (n * 4))
    buf = (ctypes.c_float * n).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), ptr


class Pipe:
    """xdtts_pipe: `depth` batches of one shape in flight, PCIe copies overlapped with the kernels of the
    neighbouring batches.  push() stages a batch in pinned host memory and enqueues it; results come back in
    push order from pop() / flush() (and from push() itself when it has to free a slot first).

    With `postnet` (a tacotron2.Postnet) the pipe runs the whole tail of XdTts::infer: decoder mels ->
    postnet -> lift -> Griffin-Lim; want_mels also returns "mel_outputs_postnet".
    """

    def __init__(self, voc, frame_counts, depth=2, postnet=None, want_mels=False):
        lib = load_library()
        self.voc, self.postnet = voc, postnet
        self.ts = [int(t) for t in frame_counts]
        self.depth = int(depth)
        self.want_mels = bool(want_mels)
        if self.want_mels and postnet is None:
            raise XdttsError(_ffi.ERR_BAD_ARG, "want_mels needs a postnet")
        q = ctypes.c_void_p()
        t_arr = (ctypes.c_int * len(self.ts))(*self.ts)
        check(lib.xdtts_pipe_create(voc._h, None if postnet is None else postnet._h, t_arr, len(self.ts), self.depth,
                                    ctypes.byref(q)))
        self._q = q
        self._ptrs = []
        self._slots = []
        for _ in range(self.depth):
            sl = dict(mel=[self._pin((voc.n_mels, t)) for t in self.ts], phase=None,
                      wave=[self._pin((voc.hop * (t - 1),)) for t in self.ts],
                      omel=[self._pin((voc.n_mels, t)) for t in self.ts] if self.want_mels else None)
            self._slots.append(sl)
        self._head = 0
        self._inflight = []   # slot indices, oldest first

    def _pin(self, shape):
        a, ptr = _pinned(shape)
        self._ptrs.append(ptr)
        return a

    def close(self):
        if getattr(self, "_q", None):
            lib = load_library()
            lib.xdtts_pipe_destroy(self._q)
            self._q = None
            for ptr in self._ptrs:
                lib.xdtts_host_free(ptr)
            self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _result(self, i):
        sl = self._slots[i]
        waves = [a.copy() for a in sl["wave"]]
        return (waves, [a.copy() for a in sl["omel"]]) if self.want_mels else waves

    def pop(self):
        """Wait for the oldest batch in flight -> its waveforms (and mels), or None when the pipe is empty."""
        rc = load_library().xdtts_pipe_pop(self._q)
        if rc < 0:
            check(rc)
        if rc == 0:
            return None
        return self._result(self._inflight.pop(0))

    def push(self, mels, init_phases=None):
        """Enqueue one batch; returns the result of the oldest batch if a slot had to be freed, else None."""
        lib = load_library()
        done = self.pop() if len(self._inflight) == self.depth else None
        sl = self._slots[self._head]
        ins = [_as_f32_2d(a, self.voc.n_mels, "mel") for a in mels]
        if [a.shape[1] for a in ins] != self.ts:
            raise XdttsError(_ffi.ERR_SHAPE, "mels do not match the pipe's frame counts")
        for dst, src in zip(sl["mel"], ins):
            dst[...] = src
        phs = None
        if init_phases is not None:
            phs = [_as_f32_2d(a, self.voc.k_bins, "init_phase") for a in init_phases]
            if [a.shape[1] for a in phs] != self.ts:
                raise XdttsError(_ffi.ERR_SHAPE, "init_phases do not match the pipe's frame counts")
            if sl["phase"] is None:
                sl["phase"] = [self._pin((self.voc.k_bins, t)) for t in self.ts]
            for dst, src in zip(sl["phase"], phs):
                dst[...] = src
        check(lib.xdtts_pipe_push(self._q, fptr_array(sl["mel"]), None if phs is None else fptr_array(sl["phase"]),
                                  fptr_array(sl["omel"]) if self.want_mels else None, fptr_array(sl["wave"])))
        self._inflight.append(self._head)
        self._head = (self._head + 1) % self.depth
        return done

    def flush(self):
        """Wait for everything in flight -> list of results, oldest first."""
        out = []
        while self._inflight:
            out.append(self.pop())
        return out

    def pending(self):
        rc = load_library().xdtts_pipe_pending(self._q)
        if rc < 0:
            check(rc)
        return rc


class GriffinLimPool:
    """One vocoder per GPU behind one handle (xdtts_pool_*): `infer_batch` shards the utterances over the devices
    (longest first to the least-loaded device, no data-path exchange -- SURVEY.md section 8e) and returns the
    waveforms in the caller's order.  Same constructor arguments as GriffinLim.new, plus the device list."""

    def __init__(self, handle, n_mels, k_bins, hop):
        self._h, self.n_mels, self.k_bins, self.hop = handle, n_mels, k_bins, hop

    @classmethod
    def new(cls, mel_basis, noverlap, power, iter, momentum, *, devices=None, delog=DELOG_EXP, pad_mode=PAD_REFLECT,  # noqa: A002
            normalise=NORM_PEAK, seed=0, run_frames=0, lift=LIFT_PINV, nnls_iters=0, fixed_seed=False, exponent=0):
        lib = load_library()
        basis = np.ascontiguousarray(mel_basis, dtype=np.float32)
        if basis.ndim != 2:
            raise XdttsError(_ffi.ERR_SHAPE, "mel_basis must be 2-D [n_mels, K]")
        opts = GlOpts(int(delog), int(pad_mode), int(normalise), int(run_frames), int(seed), 0, int(lift), int(nnls_iters),
                      int(bool(fixed_seed)), int(exponent))
        devs = list(devices) if devices is not None else []
        d_arr = (ctypes.c_int * max(len(devs), 1))(*devs) if devs else None
        h = ctypes.c_void_p()
        check(lib.xdtts_pool_create(fptr(basis), basis.shape[0], basis.shape[1], int(noverlap), float(power), int(iter),
                                    float(momentum), ctypes.byref(opts), d_arr, len(devs), ctypes.byref(h)))
        return cls(h, basis.shape[0], basis.shape[1], 2 * (basis.shape[1] - 1) - int(noverlap))

    def close(self):
        if self._h:
            load_library().xdtts_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def n_devices(self):
        return load_library().xdtts_pool_n_devices(self._h)

    def assignment(self, frame_counts):
        ts = [int(t) for t in frame_counts]
        t_arr = (ctypes.c_int * len(ts))(*ts)
        out = (ctypes.c_int * len(ts))()
        check(load_library().xdtts_pool_assignment(self._h, t_arr, len(ts), out))
        return list(out)

    def _run(self, fn, ins, rows, what, init_phases):
        lib = load_library()
        ins = [_as_f32_2d(a, rows, what) for a in ins]
        ts = [a.shape[1] for a in ins]
        outs = [np.empty(max(self.hop * (t - 1), 0), dtype=np.float32) for t in ts]
        phs = None
        if init_phases is not None:
            phs = [_as_f32_2d(a, self.k_bins, "init_phase") for a in init_phases]
            if [a.shape[1] for a in phs] != ts:
                raise XdttsError(_ffi.ERR_SHAPE, "init_phases do not match the inputs' frame counts")
        t_arr = (ctypes.c_int * len(ts))(*ts)
        check(getattr(lib, fn)(self._h, fptr_array(ins), t_arr, len(ins), None if phs is None else fptr_array(phs), fptr_array(outs)))
        return outs

    def infer_batch(self, mels, init_phases=None):
        return self._run("xdtts_pool_infer_batch", mels, self.n_mels, "mel", init_phases)

    def from_magnitude_batch(self, mags, init_phases=None):
        return self._run("xdtts_pool_from_mag_batch", mags, self.k_bins, "magnitude", init_phases)
