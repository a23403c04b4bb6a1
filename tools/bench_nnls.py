"""Cost of the NNLS lift (xdtts_gl_opts.lift = 1) next to the pseudo-inverse lift on a bench configuration.

    python tools/bench_nnls.py [cfg2|cfg5]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import griffin_lim  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
b, t, n_fft, it = bench.CONFIGS[cfg]
hop = n_fft // 4
basis = griffin_lim.mel.create_mel_filter_bank(bench.SR, n_fft, bench.N_MELS, 0.0, bench.FMAX)
rng = np.random.default_rng(0)
for name, mels in (("uniform ln-mel U(-8,0) (bench input, far outside the filterbank's range)", bench.synth_batch(b, t, 1234)),
                   ("smooth spectra pushed through the filterbank (speech-like)",
                    [np.log(np.maximum(basis @ np.abs(np.cumsum(rng.standard_normal((n_fft // 2 + 1, t)), 0) * 0.05 + 1.0), 1e-5)).astype(np.float32)
                     for _ in range(b)])):
    res = {}
    for lift in (0, 1):
        voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, bench.POWER, 0, bench.MOMENTUM, lift=lift)   # 0 iterations: lift + one ISTFT
        plan = voc.plan([t] * b)
        plan.upload(0, mels)
        for _ in range(2):
            plan.run(0)
        res[lift] = min(plan.run(0)[0] for _ in range(5))
    print("%s %s: pinv lift pass %.3f ms, NNLS lift pass %.3f ms -> NNLS adds %.3f ms for %d frames (%.1f M frames/s)" % (
        cfg, name, res[0], res[1], res[1] - res[0], b * t, b * t / (res[1] - res[0]) / 1e3))
