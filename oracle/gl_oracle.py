"""numpy restatement of the Griffin-Lim vocoder path of xd-tts (TEST INFRASTRUCTURE).

PARITY UNPINNED -- see ``oracle/__init__.py``.  The reference calls an
un-vendored crate (``griffin-lim 0.2.0`` @ e6415314, ``Cargo.lock:666-680``);
this file restates the librosa-0.9.2 algorithm that crate is a port of
(``slides/vocoding.typ:50``, ``scripts/requirements.txt:12``), with the
parameters of the reference's call sites:

* ``create_mel_filter_bank(22050.0, 1024, 80, 0.0, Some(8000.0))``
  -- ``src/tacotron2/mod.rs:453``
* ``GriffinLim::new(mel_basis, 1024 - 256, 1.7, 30, 0.99)``
  -- ``src/tacotron2/mod.rs:456``
* ``self.vocoder.infer(&spectrogram)`` -> samples scaled by ``i16::MAX``
  -- ``src/lib.rs:141-155``

Every function takes ``dtype`` (np.float32 = what the reference computes in,
np.float64 = the high-precision yardstick of SURVEY.md section 0.8).
Random numbers are never drawn here: the initial phase is an argument.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as _fft

TINY32 = float(np.finfo(np.float32).tiny)

# ----------------------------------------------------------------------------
# mel filterbank -- griffin_lim::mel::create_mel_filter_bank
# (call site src/tacotron2/mod.rs:453; librosa.filters.mel, htk=False, norm='slaney')
# ----------------------------------------------------------------------------
_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / _F_SP
    log = _MIN_LOG_MEL + np.log(np.maximum(f, 1e-300) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(f >= _MIN_LOG_HZ, log, lin)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = _F_SP * m
    log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, log, lin)


def create_mel_filter_bank(sr, n_fft, n_mels, fmin=0.0, fmax=None, dtype=np.float32):
    """[n_mels, n_fft//2+1] Slaney-scale, Slaney-normalised triangular filters."""
    if fmax is None:
        fmax = sr / 2.0
    k = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, k)
    mel_pts = np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2)
    mel_f = mel_to_hz(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(dtype)


# ----------------------------------------------------------------------------
# mel -> linear lift -- GriffinLim::infer step 1
# ----------------------------------------------------------------------------
def pinv_basis(mel_basis):
    """Moore-Penrose pseudo-inverse [K, n_mels] of the filterbank, fp64."""
    return np.linalg.pinv(np.asarray(mel_basis, dtype=np.float64))


DELOG_EXP, DELOG_POW10, DELOG_NONE = 0, 1, 2


def delog(mel, mode=DELOG_EXP, dtype=np.float32):
    mel = np.asarray(mel, dtype=dtype)
    if mode == DELOG_EXP:
        return np.exp(mel)
    if mode == DELOG_POW10:
        return np.power(dtype(10.0), mel)
    return mel.copy()


def lift_pinv_clamp(mel, mel_basis, power=1.7, delog_mode=DELOG_EXP, dtype=np.float32):
    """S[K,T] = max(0, pinv(basis) . delog(mel)) ** power  (build contract, north star)."""
    p = pinv_basis(mel_basis).astype(dtype)
    m = delog(mel, delog_mode, dtype)
    s = p @ m
    np.maximum(s, 0, out=s)
    return np.power(s, dtype(power)).astype(dtype)


def lift_nnls(mel, mel_basis, power=1.7, delog_mode=DELOG_EXP, maxiter=15000):
    """librosa.util.nnls-style lift (lstsq init -> clip -> L-BFGS-B), fp64.

    Only used to *report* the gap to the pseudo-inverse contract;
    never on a timed or parity path.
    """
    import scipy.optimize

    a = np.asarray(mel_basis, dtype=np.float64)
    b = delog(mel, delog_mode, np.float64)
    x0 = np.linalg.lstsq(a, b, rcond=None)[0]
    np.clip(x0, 0, None, out=x0)
    shape = x0.shape

    def f(x):
        x = x.reshape(shape)
        d = a @ x - b
        return 0.5 * np.sum(d * d), (a.T @ d).ravel()

    x, _, _ = scipy.optimize.fmin_l_bfgs_b(
        f, x0.ravel(), bounds=[(0, None)] * x0.size, m=10, factr=1e7, pgtol=1e-5, maxiter=maxiter
    )
    return np.power(x.reshape(shape), power)


def lift_nnls_fista(mel, mel_basis, power=1.7, delog_mode=DELOG_EXP, max_iter=300, pgtol=3e-6, dtype=np.float64):
    """The device's NNLS lift (xdtts_gl_opts.lift = 1), restated: librosa's `nnls` problem -- minimise
    0.5 |A x - b|^2 over x >= 0 from x0 = clip(pinv(A) b) -- solved per frame by accelerated projected
    gradient (FISTA, step 1/L, L = 1.001 sigma_max(A)^2, gradient restart), at most max_iter iterations, a
    frame stops once its projected-gradient step is below pgtol / L.  Returns (x ** power, iterations used)."""
    a = np.asarray(mel_basis, dtype=dtype)
    b = delog(mel, delog_mode, np.float32).astype(dtype)
    big_l = dtype(np.linalg.norm(np.asarray(mel_basis, np.float64), 2) ** 2 * 1.001)
    x = np.maximum(pinv_basis(mel_basis).astype(dtype) @ b, 0)
    y = x.copy()
    n = x.shape[1]
    tk = np.ones(n, dtype)
    active = np.ones(n, bool)
    used = np.zeros(n, int)
    for _ in range(max_iter):
        if not active.any():
            break
        idx = np.where(active)[0]
        ya, xa = y[:, idx], x[:, idx]
        g = a.T @ (a @ ya - b[:, idx])
        xn = np.maximum(ya - g / big_l, 0)
        dot = (g * (xn - xa)).sum(0)
        step = np.abs(xn - ya).max(0)
        restart = dot > 0
        tn = (1 + np.sqrt(1 + 4 * tk[idx] ** 2)) / 2
        beta = np.where(restart, 0, (tk[idx] - 1) / tn)
        tk[idx] = np.where(restart, 1, tn)
        y[:, idx] = xn + beta * (xn - xa)
        x[:, idx] = xn
        used[idx] += 1
        active[idx[step < pgtol / big_l]] = False
    return np.power(x, dtype(power)).astype(dtype), used


# ----------------------------------------------------------------------------
# STFT / ISTFT -- librosa 0.9.2 semantics
# ----------------------------------------------------------------------------
def hann_periodic(n, dtype=np.float32):
    i = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * i / n)).astype(dtype)


PAD_REFLECT, PAD_CONSTANT = 0, 1


def stft(y, n_fft, hop, pad_mode=PAD_REFLECT, dtype=np.float32, workers=1):
    """center=True STFT; returns [K, T] complex, T = 1 + len(y)//hop."""
    y = np.asarray(y, dtype=dtype)
    mode = "reflect" if pad_mode == PAD_REFLECT else "constant"
    yp = np.pad(y, n_fft // 2, mode=mode)
    t = 1 + (len(yp) - n_fft) // hop
    frames = np.lib.stride_tricks.as_strided(
        yp, shape=(t, n_fft), strides=(hop * yp.strides[0], yp.strides[0]), writeable=False
    )
    w = hann_periodic(n_fft, dtype)
    spec = _fft.rfft(frames * w[None, :], axis=1, workers=workers)
    return np.ascontiguousarray(spec.T)


def window_sumsquare(n_fft, hop, n_frames, dtype=np.float32):
    w2 = hann_periodic(n_fft, dtype) ** 2
    x = np.zeros(n_fft + hop * (n_frames - 1), dtype=dtype)
    for i in range(n_frames):
        x[i * hop : i * hop + n_fft] += w2
    return x


def istft(spec, n_fft, hop, dtype=np.float32, workers=1):
    """Inverse of :func:`stft`: irfft -> window -> overlap-add (ascending frame
    order) -> divide by the window sum-square where > tiny -> trim n_fft//2."""
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    spec = np.asarray(spec, dtype=cdtype)
    k, t = spec.shape
    assert k == n_fft // 2 + 1
    w = hann_periodic(n_fft, dtype)
    frames = _fft.irfft(spec.T, n=n_fft, axis=1, workers=workers).astype(dtype) * w[None, :]
    y = np.zeros(n_fft + hop * (t - 1), dtype=dtype)
    if n_fft % hop == 0:
        r = n_fft // hop
        yb = y.reshape(t - 1 + r, hop)
        fb = frames.reshape(t, r, hop)
        for d in range(r - 1, -1, -1):  # oldest frame first == ascending frame order
            yb[d : d + t] += fb[:, d, :]
    else:
        for i in range(t):
            y[i * hop : i * hop + n_fft] += frames[i]
    wss = window_sumsquare(n_fft, hop, t, dtype)
    nz = wss > np.finfo(dtype).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2 : len(y) - n_fft // 2]


# ----------------------------------------------------------------------------
# initial phase: counter-based uniform "turns" u in [0,1), shared bit-for-bit
# with the CUDA library (xd-tts_b200/csrc/gl_phase.cuh)
# ----------------------------------------------------------------------------
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def phase_turns(seed, utt, k_bins, n_frames):
    """u[k, t] = top 24 bits of splitmix64(seed, utt, t*K + k) * 2**-24, float32."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + _GOLD * np.uint64(utt + 1)
        t = np.arange(n_frames, dtype=np.uint64)[None, :]
        k = np.arange(k_bins, dtype=np.uint64)[:, None]
        idx = t * np.uint64(k_bins) + k
        z = base + _GOLD * (idx + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) * (2.0**-24)).astype(np.float32)


def turns_to_angles(turns, dtype=np.float32):
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    th = 2.0 * np.pi * np.asarray(turns, dtype=np.float64)
    return (np.cos(th) + 1j * np.sin(th)).astype(cdtype)


# ----------------------------------------------------------------------------
# Griffin-Lim -- GriffinLim::infer steps 2-4 (SURVEY.md section 8 rows a6-a8;
# librosa.griffinlim: momentum update alpha = m/(1+m), eps = tiny, final ISTFT)
# ----------------------------------------------------------------------------
def griffin_lim(
    s_mag,
    turns0,
    n_iter,
    momentum,
    n_fft,
    hop,
    pad_mode=PAD_REFLECT,
    dtype=np.float32,
    checkpoints=None,
    workers=1,
):
    """Returns the waveform after ``n_iter`` iterations; ``checkpoints`` (a dict
    keyed by iteration index) is filled with (waveform_k, rebuilt_k) where
    waveform_0 is the ISTFT of the initial guess."""
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    s_mag = np.asarray(s_mag, dtype=dtype)
    angles = turns_to_angles(turns0, dtype)
    alpha = dtype(momentum / (1.0 + momentum))
    tiny = dtype(TINY32)
    rebuilt = np.zeros(s_mag.shape, dtype=cdtype)
    for it in range(n_iter):
        tprev = rebuilt
        inverse = istft(s_mag * angles, n_fft, hop, dtype, workers)
        if checkpoints is not None and it in checkpoints:
            checkpoints[it] = (inverse.copy(), tprev.copy())
        rebuilt = stft(inverse, n_fft, hop, pad_mode, dtype, workers)
        angles = rebuilt - alpha * tprev
        angles = (angles / (np.abs(angles) + tiny)).astype(cdtype)
    y = istft(s_mag * angles, n_fft, hop, dtype, workers)
    if checkpoints is not None and n_iter in checkpoints:
        checkpoints[n_iter] = (y.copy(), rebuilt.copy())
    return y


def gl_one_iteration(s_mag, y_prev, rebuilt_prev, momentum, n_fft, hop, pad_mode=PAD_REFLECT, dtype=np.float32):
    """One teacher-forced iteration: (y_k, R_k) -> (y_{k+1}, R_{k+1}).  ``momentum``
    may be 0 to model the first iteration (tprev == 0)."""
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    alpha = dtype(momentum / (1.0 + momentum))
    rebuilt = stft(y_prev, n_fft, hop, pad_mode, dtype)
    angles = rebuilt - alpha * np.asarray(rebuilt_prev, dtype=cdtype)
    angles = (angles / (np.abs(angles) + dtype(TINY32))).astype(cdtype)
    return istft(np.asarray(s_mag, dtype=dtype) * angles, n_fft, hop, dtype), rebuilt


NORM_PEAK, NORM_NONE = 0, 1


def peak_normalise(y):
    m = np.max(np.abs(y)) if len(y) else 0
    return y / m if m > 0 else y


def infer(
    mel,
    mel_basis,
    noverlap,
    power,
    n_iter,
    momentum,
    turns0,
    delog_mode=DELOG_EXP,
    pad_mode=PAD_REFLECT,
    normalise=NORM_PEAK,
    dtype=np.float32,
    workers=1,
):
    """``GriffinLim::infer`` (``src/lib.rs:141``): mel [n_mels,T] -> samples [hop*(T-1)]."""
    k = mel_basis.shape[1]
    n_fft = 2 * (k - 1)
    hop = n_fft - noverlap
    s = lift_pinv_clamp(mel, mel_basis, power, delog_mode, dtype)
    y = griffin_lim(s, turns0, n_iter, momentum, n_fft, hop, pad_mode, dtype, workers=workers)
    if normalise == NORM_PEAK:
        y = peak_normalise(y)
    return y.astype(dtype)


# ----------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench
# ----------------------------------------------------------------------------
def synth_mel(seed, n_mels, n_frames):
    """ln-mel in U(-8, 0), the range Tacotron2 emits."""
    return np.random.default_rng(seed).uniform(-8.0, 0.0, (n_mels, n_frames)).astype(np.float32)


def synth_speech_like_mag(seed, n_fft, hop, n_frames, sr=22050.0):
    """|STFT| of a harmonic stack (f0 ~ 120 Hz, slow vibrato) + noise: a
    speech-like linear magnitude for parity tests (uniform spectra under-state
    fp32 error growth)."""
    rng = np.random.default_rng(seed)
    n = hop * (n_frames - 1)
    t = np.arange(n) / sr
    f0 = 120.0 * (1.0 + 0.05 * np.sin(2 * np.pi * 3.0 * t))
    ph = 2 * np.pi * np.cumsum(f0) / sr
    y = sum((1.0 / h) * np.sin(h * ph) for h in range(1, 30))
    y = y * (0.6 + 0.4 * np.sin(2 * np.pi * 1.7 * t)) + 0.02 * rng.standard_normal(n)
    return np.abs(stft(y, n_fft, hop, dtype=np.float64)).astype(np.float32)


# ----------------------------------------------------------------------------
# 16-bit PCM -- the caller's loop, src/lib.rs:153-157:
#     for sample in &audio { wav.write_sample((sample * i16::MAX as f32) as i16) }
# Rust float -> int `as` casts truncate toward zero, saturate at the bounds and map NaN to 0.
# ----------------------------------------------------------------------------
def pcm16(y):
    x = np.asarray(y, dtype=np.float32) * np.float32(32767.0)
    x = np.where(np.isnan(x), np.float32(0.0), x)
    return np.trunc(np.clip(x, -32768.0, 32767.0)).astype(np.int16)
