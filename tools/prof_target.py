"""Short command for ncu: one kernel-by-kernel pass of a bench configuration (no graph, no CPU leg).

    ncu ... python tools/prof_target.py [cfg2|cfg5] [passes]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import _ffi, griffin_lim  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
b, t, n_fft, it = bench.CONFIGS[cfg]
hop = n_fft // 4
basis = griffin_lim.mel.create_mel_filter_bank(bench.SR, n_fft, bench.N_MELS, 0.0, bench.FMAX)
voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, bench.POWER, it, bench.MOMENTUM)
plan = voc.plan([t] * b)
plan.upload(0, bench.synth_batch(b, t, 1234))
for _ in range(passes):
    ms, mi, n = plan.run(_ffi.RUN_NO_GRAPH)
    print("pass: %.3f ms total, %.4f ms per steady-state launch (%d launches)" % (ms, mi / max(n, 1), n))
