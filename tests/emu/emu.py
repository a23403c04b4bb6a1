"""ctypes wrapper of the CPU lane-program emulator (tests/emu/gl_emu.cpp) -- TEST INFRASTRUCTURE."""
import ctypes
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libgl_emu.so")
_lib = None
_fp = ctypes.POINTER(ctypes.c_float)


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


def gl_from_mag(s_mag, turns, n_iter, momentum, run_frames, pad_mode=0, seed=0):
    """-> (waveform, R [T, M] complex packed as the kernel stores it, max|y|, n_runs)"""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
        _lib.emu_gl_from_mag.argtypes = [_fp, _fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_ulonglong, _fp, _fp, _fp]
    s_mag = np.ascontiguousarray(s_mag, np.float32)
    turns = None if turns is None else np.ascontiguousarray(turns, np.float32)
    k, t = s_mag.shape
    hop = (k - 1) // 2
    out = np.zeros(hop * (t - 1), np.float32)
    r = np.zeros((t, k - 1, 2), np.float32)
    pk = np.zeros(1, np.float32)
    n = _lib.emu_gl_from_mag(_p(s_mag), _p(turns), k, t, n_iter, momentum, run_frames, pad_mode, seed, _p(out), _p(r), _p(pk))
    if n < 0:
        raise ValueError("emulator rejected the input (%d)" % n)
    return out, r[..., 0] + 1j * r[..., 1], float(pk[0]), n
