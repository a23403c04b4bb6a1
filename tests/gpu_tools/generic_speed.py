"""Step time of the un-fused Griffin-Lim path on the cfg2 batch (32 x 1000 frames, 60 iterations): the shipped geometry forced
through it (XDTTS_GL_GENERIC=1) next to the fused kernel, and a hop the fused kernel does not cover.
    python tests/gpu_tools/generic_speed.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import griffin_lim  # noqa: E402

b, t, n_fft, it = bench.CONFIGS["cfg2"]
basis = griffin_lim.mel.create_mel_filter_bank(bench.SR, n_fft, bench.N_MELS, 0.0, bench.FMAX)
mels = bench.synth_batch(b, t, 1234)
for label, hop, env in (("fused, hop 256", 256, None), ("un-fused, hop 256", 256, "1"), ("un-fused, hop 200", 200, None), ("un-fused, hop 512", 512, None)):
    if env:
        os.environ["XDTTS_GL_GENERIC"] = env
    voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, bench.POWER, it, bench.MOMENTUM)
    os.environ.pop("XDTTS_GL_GENERIC", None)
    plan = voc.plan([t] * b)
    plan.upload(0, mels)
    for _ in range(2):
        plan.run(0)
    ms = min(plan.run(0)[0] for _ in range(3))
    print("%-18s %8.3f ms per step  %.2f M frames/s" % (label, ms, b * t / ms / 1e3))
