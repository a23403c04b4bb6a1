"""Where the end-to-end time of xdtts_gl_infer_batch goes (pinned host buffers): upload / run / download."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import _ffi, griffin_lim  # noqa: E402

lib = _ffi.load_library()
b, t, n_fft, it = bench.CONFIGS["cfg2"]
hop = n_fft // 4
basis = griffin_lim.mel.create_mel_filter_bank(bench.SR, n_fft, bench.N_MELS, 0.0, bench.FMAX)
voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, bench.POWER, it, bench.MOMENTUM)
mels = bench.synth_batch(b, t, 1234)
pin_in = [bench.pinned_array(lib, (bench.N_MELS, t)) for _ in range(b)]
for (a, _), m in zip(pin_in, mels):
    a[...] = m
pin_out = [bench.pinned_array(lib, (hop * (t - 1),)) for _ in range(b)]
pin_pcm = [bench.pinned_array(lib, (hop * (t - 1) // 2 + 1,)) for _ in range(b)]
in_ptrs = _ffi.fptr_array([a for a, _ in pin_in])
out_ptrs = _ffi.fptr_array([a for a, _ in pin_out])
pcm_ptrs = _ffi.sptr_array([a.view(np.int16) for a, _ in pin_pcm])
t_arr = (ctypes.c_int * b)(*([t] * b))
plan = voc.plan([t] * b)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


print("upload   %.3f ms" % timed(lambda: (plan.upload_ptrs(0, in_ptrs), plan.run(0))) + "  (upload + run)")
print("run      %.3f ms" % timed(lambda: plan.run(0)))
print("download %.3f ms" % timed(lambda: plan.download_ptrs(out_ptrs)))
print("e2e f32  %.3f ms" % timed(lambda: _ffi.check(lib.xdtts_gl_infer_batch(voc._h, in_ptrs, t_arr, b, None, out_ptrs))))
print("e2e pcm  %.3f ms" % timed(lambda: _ffi.check(lib.xdtts_gl_infer_batch_pcm16(voc._h, in_ptrs, t_arr, b, None, pcm_ptrs))))
