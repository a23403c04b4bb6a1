"""Per-step device times of the resident cfg2 pass, with and without the bench's NVML clock sampler running:
    python tests/gpu_tools/step_times.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import griffin_lim  # noqa: E402

b, t, n_fft, it = bench.CONFIGS["cfg2"]
hop = n_fft // 4
basis = griffin_lim.mel.create_mel_filter_bank(bench.SR, n_fft, bench.N_MELS, 0.0, bench.FMAX)
voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, bench.POWER, it, bench.MOMENTUM)
plan = voc.plan([t] * b)
plan.upload(0, bench.synth_batch(b, t, 1234))
for _ in range(5):
    plan.run(0)
for label in ("no sampler", "sampler", "no sampler"):
    s = bench.ClockSampler(0) if label == "sampler" else None
    if s:
        s.start()
    ms = [plan.run(0)[0] for _ in range(30)]
    if s:
        print("clocks", s.stop())
    print("%-10s mean %.3f min %.3f max %.3f  " % (label, np.mean(ms), min(ms), max(ms)) + " ".join("%.2f" % x for x in ms))
