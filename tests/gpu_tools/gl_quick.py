"""Quick GPU check of a library build: parity on a small case + steady-state kernel time of cfg2 / cfg5.

    [XDTTS_B200_LIB=path/to/variant.so] python tests/gpu_tools/gl_quick.py [cfg2 cfg5 ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

import bench  # noqa: E402
from oracle import gl_oracle as o  # noqa: E402
from xdtts_b200 import _ffi, griffin_lim  # noqa: E402

tag = os.path.basename(os.environ.get("XDTTS_B200_LIB", "default"))
for n_fft, t in ((1024, 77), (2048, 40)):
    hop, k = n_fft // 4, n_fft // 2 + 1
    s = o.synth_speech_like_mag(7, n_fft, hop, t)
    tu = o.phase_turns(11, 0, k, t)
    ref = o.griffin_lim(s, tu, 4, 0.99, n_fft, hop, dtype=np.float64)
    basis = o.create_mel_filter_bank(22050.0, n_fft, 80, 0.0, 8000.0)
    voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, 1.7, 4, 0.99, normalise=griffin_lim.NORM_NONE, run_frames=9)
    (y,) = voc.from_magnitude_batch([s], [tu])
    err = float(np.sqrt(np.mean((y - ref) ** 2)) / np.abs(ref).max())
    print("%s parity n_fft=%d rel rms %.2e %s" % (tag, n_fft, err, "OK" if err < 2e-6 else "FAIL"))
# lift accuracy against the fp64 oracle (the gate of tests/test_gpu_gl.py::test_lift_matches_oracle is 1e-5 of full scale)
for seed in (5, 6):
    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    mel = o.synth_mel(seed, 80, 150)
    voc = griffin_lim.GriffinLim.new(basis, 768, 1.7, 0, 0.99)
    plan = voc.plan([150])
    plan.upload(0, [mel])
    plan.run(0)
    got = np.concatenate([plan.peek(0).T, plan.peek(1)[None, :]], 0)
    ref = o.lift_pinv_clamp(mel, basis, 1.7, dtype=np.float64)
    print("%s lift max err / full scale %.2e" % (tag, np.abs(got - ref).max() / ref.max()))
peak = bench.measured_peak()[0]
for cfg in (sys.argv[1:] or ["cfg2", "cfg5"]):
    b, t, n_fft, it = bench.CONFIGS[cfg]
    hop, k = n_fft // 4, n_fft // 2 + 1
    basis = griffin_lim.mel.create_mel_filter_bank(bench.SR, n_fft, bench.N_MELS, 0.0, bench.FMAX)
    voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, bench.POWER, it, bench.MOMENTUM)
    plan = voc.plan([t] * b)
    plan.upload(0, bench.synth_batch(b, t, 1234))
    for _ in range(3):
        plan.run(0)
    tot = min(plan.run(0)[0] for _ in range(5))
    best = 1e9
    for _ in range(4):
        _, mi, n = plan.run(_ffi.RUN_NO_GRAPH)
        best = min(best, mi / n)
    lift = min(plan.lift_ms() for _ in range(1)) if hasattr(plan, "lift_ms") else float("nan")
    lifts = []
    for _ in range(5):
        plan.run(_ffi.RUN_NO_GRAPH)
        lifts.append(plan.lift_ms())
    lift1 = min(lifts)                                   # one launch between two events: + ~7 us of launch / event latency
    lift = min(plan.time_lift(20) for _ in range(3)) if hasattr(plan, "time_lift") else lift1
    print("%s %s: lift %.1f us = %.0f GB/s (%.3f of peak); single bracketed launch %.1f us" % (tag, cfg, lift * 1e3, b * t * 4 * (80 + k) / lift / 1e6, b * t * 4 * (80 + k) / lift / 1e6 / peak, lift1 * 1e3))
    alg = b * t * (20 * k + 8 * hop)
    info = plan.info()
    print("%s %s: step %.3f ms (%.2f M frames/s), iteration kernel %.2f us, %.0f GB/s = %.3f of %.0f, runs %d x %d frames, %d CTAs"
          % (tag, cfg, tot, b * t / tot / 1e3, best * 1e3, alg / best / 1e6, alg / best / 1e6 / peak, peak, info["n_runs"],
             info["run_frames"], info["ctas"]))
