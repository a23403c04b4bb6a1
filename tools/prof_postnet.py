"""Short command for ncu / timing: the postnet on the cfg3 batch (32 x 80x1000), each precision mode.

    python tools/prof_postnet.py [passes] [precision ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import tacotron2  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
modes = [int(x) for x in sys.argv[2:]] or [0, 1]
b, t = 32, 1000
layers = bench.synth_postnet_layers()
mels = bench.synth_batch(b, t, 1234)
flop = 8683520.0 * b * t
for prec in modes:
    post = tacotron2.Postnet.from_layers(layers, precision=prec)
    plan = post.plan([t] * b)
    plan.upload(mels)
    best = 1e9
    for _ in range(passes):
        ms = plan.run()
        best = min(best, ms)
        print("precision %d: %.3f ms  -> %.1f algorithmic TFLOP/s" % (prec, ms, flop / ms / 1e9))
    print("precision %d best %.3f ms, %.2f M frames/s" % (prec, best, b * t / best / 1e3))
