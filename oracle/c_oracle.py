"""ctypes wrapper over oracle/c/gl_oracle.c (TEST INFRASTRUCTURE, see that file).

Used by tests/ as a second checker and by bench.py as the timed CPU arm
(``cpu_baseline.kind == "port"``).  PARITY UNPINNED, see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgl_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "c", "gl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def use_native_build():
    """Timed CPU arm only (bench.py): rebuild the port with -march=native ON THE HOST THAT RUNS IT (the portable build
    that travels with the repo targets x86-64-v3) and load that one.  Falls back to the portable build."""
    global _lib, _SO
    src = os.path.join(_HERE, "c", "gl_oracle.c")
    out = os.path.join(_HERE, "_build", "libgl_oracle_native.so")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-std=gnu11", "-shared", "-o", out, src, "-lm"],
                              stderr=subprocess.DEVNULL)
        _SO, _lib = out, None
        lib()
        return True
    except Exception:  # noqa: BLE001
        _SO, _lib = os.path.join(_HERE, "_build", "libgl_oracle.so"), None
        return False


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.oracle_gl_from_mag.argtypes = [fp, fp] + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_int, fp]
        _lib.oracle_lift_pinv.argtypes = [fp, fp] + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, fp]
        _lib.oracle_infer.argtypes = (
            [fp, fp, fp] + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_int, ctypes.c_float] + [ctypes.c_int] * 3 + [fp]
        )
        _lib.oracle_infer_batch.argtypes = (
            [fp, fp, fp] + [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_int, ctypes.c_float] + [ctypes.c_int] * 3 + [fp]
        )
        _lib.oracle_stft.argtypes = [fp] + [ctypes.c_int] * 4 + [fp]
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def gl_from_mag(s_mag, turns, hop, n_iter, momentum, pad_constant=0):
    s_mag = np.ascontiguousarray(s_mag, dtype=np.float32)
    turns = np.ascontiguousarray(turns, dtype=np.float32)
    k, t = s_mag.shape
    out = np.empty(hop * (t - 1), dtype=np.float32)
    lib().oracle_gl_from_mag(_p(s_mag), _p(turns), k, t, hop, n_iter, momentum, pad_constant, _p(out))
    return out


def lift_pinv(pinv, mel, power, delog=0):
    pinv = np.ascontiguousarray(pinv, dtype=np.float32)
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    k, m = pinv.shape
    t = mel.shape[1]
    out = np.empty((k, t), dtype=np.float32)
    lib().oracle_lift_pinv(_p(pinv), _p(mel), k, m, t, power, delog, _p(out))
    return out


def infer(pinv, mel, turns, hop, power, n_iter, momentum, delog=0, pad_constant=0, normalise=0):
    pinv = np.ascontiguousarray(pinv, dtype=np.float32)
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    turns = np.ascontiguousarray(turns, dtype=np.float32)
    k, m = pinv.shape
    t = mel.shape[1]
    out = np.empty(hop * (t - 1), dtype=np.float32)
    lib().oracle_infer(_p(pinv), _p(mel), _p(turns), k, m, t, hop, power, n_iter, momentum, delog, pad_constant, normalise, _p(out))
    return out


def infer_batch(pinv, mels, turns, hop, power, n_iter, momentum, delog=0, pad_constant=0, normalise=0):
    """mels [B, n_mels, T], turns [B, K, T] -> [B, hop*(T-1)], one thread per utterance."""
    pinv = np.ascontiguousarray(pinv, dtype=np.float32)
    mels = np.ascontiguousarray(mels, dtype=np.float32)
    turns = np.ascontiguousarray(turns, dtype=np.float32)
    k, m = pinv.shape
    b, _, t = mels.shape
    out = np.empty((b, hop * (t - 1)), dtype=np.float32)
    lib().oracle_infer_batch(_p(pinv), _p(mels), _p(turns), b, k, m, t, hop, power, n_iter, momentum, delog, pad_constant,
                             normalise, _p(out))
    return out


def stft(y, n_fft, hop, pad_constant=0):
    y = np.ascontiguousarray(y, dtype=np.float32)
    t = 1 + len(y) // hop
    k = n_fft // 2 + 1
    out = np.empty((t, k, 2), dtype=np.float32)
    lib().oracle_stft(_p(y), len(y), n_fft, hop, pad_constant, _p(out))
    return (out[..., 0] + 1j * out[..., 1]).T.astype(np.complex64)


def num_threads():
    return lib().oracle_num_threads()
