"""Where the time of a blocking batch call goes with pinned vs pageable host buffers: upload, run, download, each timed
on the host around the plan-level C calls (cfg2 shape)."""
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
plan = voc.plan([t] * b)
mels = bench.synth_batch(b, t, 1234)
out_len = hop * (t - 1)
pin_in = [bench.pinned_array(lib, (80, t)) for _ in range(b)]
for (a, _), m in zip(pin_in, mels):
    a[...] = m
pin_out = [bench.pinned_array(lib, (out_len,)) for _ in range(b)]
pg_in = [np.array(m) for m in mels]
pg_out = [np.ones(out_len, np.float32) for _ in range(b)]
sets = {"pinned": (_ffi.fptr_array([a for a, _ in pin_in]), _ffi.fptr_array([a for a, _ in pin_out])),
        "pageable": (_ffi.fptr_array(pg_in), _ffi.fptr_array(pg_out))}
for name, (ip, op) in sets.items():
    acc = [0.0, 0.0, 0.0, 0.0]
    n = 12
    for i in range(n + 3):
        t0 = time.perf_counter()
        _ffi.check(lib.xdtts_gl_plan_upload(plan._p, 0, ip))
        t1 = time.perf_counter()
        ms = ctypes.c_float()
        _ffi.check(lib.xdtts_gl_plan_run(plan._p, 0, ctypes.byref(ms), None, None))
        t2 = time.perf_counter()
        _ffi.check(lib.xdtts_gl_plan_download(plan._p, op))
        t3 = time.perf_counter()
        if i >= 3:
            acc[0] += t1 - t0
            acc[1] += t2 - t1
            acc[2] += t3 - t2
            acc[3] += ms.value * 1e-3
    print("%s: upload %.3f ms, run %.3f ms (device %.3f ms), download %.3f ms, total %.3f ms" % (
        name, acc[0] / n * 1e3, acc[1] / n * 1e3, acc[3] / n * 1e3, acc[2] / n * 1e3, sum(acc[:3]) / n * 1e3))
# raw host copy rates for reference
src = np.ones(out_len * b, np.float32)
dst = np.empty_like(src)
dst[...] = 0
t0 = time.perf_counter()
for _ in range(5):
    np.copyto(dst, src)
print("single-thread numpy copy of %.1f MB: %.2f GB/s" % (src.nbytes / 1e6, 5 * src.nbytes / (time.perf_counter() - t0) / 1e9))
