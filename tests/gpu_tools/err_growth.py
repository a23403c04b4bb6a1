"""Per-iteration error of the GPU path vs the fp64 oracle, next to the CPU fp32 oracle's (same input, same phase)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))
from oracle import gl_oracle as o  # noqa: E402
from xdtts_b200 import griffin_lim  # noqa: E402


def rel(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.abs(b).max())


for n_fft in (512, 1024, 2048):
    hop, k, t = n_fft // 4, n_fft // 2 + 1, 120
    basis = o.create_mel_filter_bank(22050.0, n_fft, 80, 0.0, 8000.0)
    for kind in ("speech", "uniform"):
        if kind == "speech":
            s = o.synth_speech_like_mag(7, n_fft, hop, t)
        else:
            s = o.lift_pinv_clamp(o.synth_mel(903, 80, t), basis, 1.7, dtype=np.float32)
        tu = o.phase_turns(11, 0, k, t)
        row = []
        for it in (0, 1, 2, 4, 8, 12):
            ref = o.griffin_lim(s, tu, it, 0.99, n_fft, hop, dtype=np.float64)
            c32 = o.griffin_lim(s, tu, it, 0.99, n_fft, hop, dtype=np.float32)
            voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, 1.7, it, 0.99, normalise=griffin_lim.NORM_NONE)
            (y,) = voc.from_magnitude_batch([s], [tu])
            row.append("it%-2d gpu %.1e cpu32 %.1e" % (it, rel(y, ref), rel(c32, ref)))
        print(n_fft, kind, " | ".join(row))
