"""Latency of small vocodes (the reference's own use: one utterance per call, cfg1 = 80x200, 30 iterations):
persistent single-launch kernel vs one launch per iteration (CUDA graph), device time and end-to-end call time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import _ffi, griffin_lim  # noqa: E402

basis = griffin_lim.mel.create_mel_filter_bank(bench.SR, 1024, 80, 0.0, bench.FMAX)
for b, t, it in ((1, 200, 30), (1, 1000, 30), (4, 500, 60), (32, 1000, 60)):
    voc = griffin_lim.GriffinLim.new(basis, 768, bench.POWER, it, bench.MOMENTUM, persistent=True)
    mels = bench.synth_batch(b, t, 1)
    plan = voc.plan([t] * b)
    plan.upload(0, mels)
    res = {}
    for name, fl in (("persistent", 0), ("per-launch graph", _ffi.RUN_PER_LAUNCH)):
        for _ in range(3):
            plan.run(fl)
        res[name] = min(plan.run(fl)[0] for _ in range(10))
    t0 = time.perf_counter()
    for _ in range(20):
        voc.infer_batch(mels)
    e2e = (time.perf_counter() - t0) / 20 * 1e3
    print("B=%d T=%d it=%d: persistent %.3f ms, per-launch graph %.3f ms (device); infer_batch call %.3f ms; runs %d, persistent=%s"
          % (b, t, it, res["persistent"], res["per-launch graph"], e2e, plan.info()["n_runs"], plan.is_persistent()))
