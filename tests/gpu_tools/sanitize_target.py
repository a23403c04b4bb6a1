"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tests/gpu_tools/sanitize_target.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

from oracle import gl_oracle as o  # noqa: E402
from oracle import postnet_oracle as po  # noqa: E402
from xdtts_b200 import griffin_lim, tacotron2  # noqa: E402

layers = po.synth_weights(seed=7)
for n_fft, ts in ((1024, [37, 4, 150]), (2048, [21, 9]), (512, [30])):
    hop, k = n_fft // 4, n_fft // 2 + 1
    basis = o.create_mel_filter_bank(22050.0, n_fft, 80, 0.0, 8000.0)
    voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, 1.7, 3, 0.99, run_frames=5, fixed_seed=True)
    mels = [o.synth_mel(i, 80, t) for i, t in enumerate(ts)]
    ys = voc.infer_batch(mels)
    ys2 = voc.infer_batch(mels, [o.phase_turns(0, i, k, t) for i, t in enumerate(ts)])
    assert all(np.array_equal(a, b) for a, b in zip(ys, ys2))
    if n_fft == 1024:
        for prec in (0, 1, 2):
            post = tacotron2.Postnet.from_layers(layers, precision=prec)
            outs = post.run_batch(mels)
            assert all(np.isfinite(x).all() for x in outs)
        post = tacotron2.Postnet.from_layers(layers)
        w = tacotron2.infer_tail_batch(post, voc, mels)
        assert all(np.isfinite(x).all() for x in w)
        pipe = voc.pipe(ts, depth=2, postnet=post)
        got = [r for r in (pipe.push(mels) for _ in range(3)) if r is not None] + pipe.flush()
        assert len(got) == 3 and all(np.array_equal(a, b) for a, b in zip(got[0], w))
        pipe.close()
# the un-fused path (gl_generic.cu): a hop that is not n_fft / 4, and an n_fft outside the fused kernel's
for n_fft, hop, ts in ((1024, 200, [9, 31]), (256, 64, [12])):
    basis = o.create_mel_filter_bank(22050.0, n_fft, 40, 0.0, 8000.0)
    voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, 1.7, 2, 0.99, fixed_seed=True)
    ys = voc.infer_batch([o.synth_mel(i, 40, t) for i, t in enumerate(ts)])
    assert all(y.shape == (hop * (t - 1),) and np.isfinite(y).all() for y, t in zip(ys, ts))
# decoder loop: batch of 3 (template NB = 4), 6 steps, spin barriers under the sanitizer are slow but finite
from oracle import decoder_oracle as d  # noqa: E402

dec = tacotron2.Decoder.from_weights(d.synth_weights(11), gate_threshold=0.999999, max_steps=6, seed=1)
enc = [d.synth_encoder_outputs(40 + i, 20) for i in range(3)]
out = dec.run_batch([m for m, _ in enc], [p for _, p in enc], [20, 17, 11])
assert all(x.shape == (80, 6) and np.isfinite(x).all() for x in out)
one = dec.run(enc[0][0], enc[0][1], 20)                    # batch 1: the whole-CTA energies / softmax forms
assert np.array_equal(one, out[0])
enc8 = [d.synth_encoder_outputs(60 + i, 20) for i in range(8)]
out8 = dec.run_batch([m for m, _ in enc8], [p for _, p in enc8], [20 - i for i in range(8)])   # template NB = 8
assert all(x.shape == (80, 6) and np.isfinite(x).all() for x in out8)
print("sanitize target ok")
