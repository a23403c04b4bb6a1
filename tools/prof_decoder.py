"""Short command for ncu: one decoder-loop launch (batch 1 or N, 300 steps).

    ncu --set full --import-source on -k regex:dec_persist -c 1 ... python tools/prof_decoder.py [batch] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from xdtts_b200 import tacotron2  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dec = tacotron2.Decoder.from_weights(bench.synth_decoder_weights(), gate_threshold=0.999999, max_steps=steps, seed=1)
enc = [bench.synth_encoder_outputs(100 + i, bench.DEC_T_ENC) for i in range(nb)]
for _ in range(2):
    dec.run_batch([m for m, _ in enc], [p for _, p in enc], [bench.DEC_UNPADDED] * nb)
    ms, n = dec.last_timing()
    print("decoder batch %d: %.3f ms for %d steps = %.2f us/step" % (nb, ms, n, ms * 1e3 / n))
