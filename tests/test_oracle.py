"""CPU tests: pin the oracle (numpy + C port) against the committed golden
vectors, which were produced by independent implementations (torch.stft/istft,
torchaudio filterbank, torch conv1d) -- see oracle/make_golden.py.  The
reference itself has no vectors for this path (PARITY UNPINNED)."""
import os

import numpy as np
import pytest

from oracle import c_oracle, gl_oracle as o, postnet_oracle as p


def rel_rms(a, b):
    s = max(float(np.abs(b).max()), 1e-30)
    return float(np.sqrt(np.mean(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) / s) ** 2)))


def test_melbank_matches_torchaudio(golden_dir):
    g = np.load(os.path.join(golden_dir, "melbank.npz"))
    for n_fft in (1024, 2048):
        fb = o.create_mel_filter_bank(22050.0, n_fft, 80, 0.0, 8000.0)
        assert fb.shape == (80, n_fft // 2 + 1) and fb.dtype == np.float32
        assert np.abs(fb - g[f"torchaudio_{n_fft}"]).max() < 2e-7
        assert np.array_equal(fb, g[f"oracle_{n_fft}"])
    fb = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    # facts recorded in SURVEY.md A.2
    assert abs(fb.max() - 0.026493) < 1e-5
    assert int((fb.sum(0) == 0).sum()) == 142
    assert (fb.sum(1) > 0).all()


def test_stft_istft_roundtrip_and_shapes():
    rng = np.random.default_rng(0)
    for n_fft, hop, t in ((1024, 256, 50), (2048, 512, 20), (512, 128, 9)):
        y = rng.standard_normal(hop * (t - 1)).astype(np.float32)
        x = o.stft(y, n_fft, hop)
        assert x.shape == (n_fft // 2 + 1, t)
        yi = o.istft(x, n_fft, hop)
        assert yi.shape == y.shape
        assert np.abs(yi - y).max() < 5e-6
        xc = c_oracle.stft(y, n_fft, hop)
        x64 = o.stft(y, n_fft, hop, dtype=np.float64)
        assert np.abs(xc - x64).max() / np.abs(x64).max() < 5e-7


def test_window_sumsquare_facts():
    # 1.5 in the interior, 1.25 at the first kept sample
    for n_fft in (1024, 2048):
        hop = n_fft // 4
        w = o.window_sumsquare(n_fft, hop, 12, np.float64)
        assert abs(w[n_fft // 2] - 1.25) < 1e-12
        assert np.abs(w[n_fft : -n_fft] - 1.5).max() < 1e-12


def test_cfg1_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cfg1_gl.npz"))
    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    assert np.array_equal(o.synth_mel(1234, 80, 200), g["mel"])
    assert np.array_equal(o.phase_turns(4321, 0, 513, 200), g["turns"])
    s = o.lift_pinv_clamp(g["mel"], basis, 1.7)
    big = g["s_mag64"] > 1e-3 * g["s_mag64"].max()
    assert np.abs(s[big] / g["s_mag64"][big] - 1).max() < 2e-3  # fp32 GEMM with cancellation
    y64 = o.griffin_lim(g["s_mag"], g["turns"], 30, 0.99, 1024, 256, dtype=np.float64)
    assert rel_rms(y64, g["y_torch64"]) < 1e-6   # independent torch implementation
    assert rel_rms(y64, g["y_fp64"]) < 1e-6
    y32 = o.griffin_lim(g["s_mag"], g["turns"], 30, 0.99, 1024, 256)
    # fp32 Griffin-Lim is self-amplifying: envelope, not equality
    assert rel_rms(y32, g["y_fp64"]) < 1e-3
    yc = c_oracle.gl_from_mag(g["s_mag"], g["turns"], 256, 30, 0.99)
    assert rel_rms(yc, g["y_fp64"]) < 1e-3


def test_teacher_forced_iterations(golden_dir):
    g = np.load(os.path.join(golden_dir, "speech48_ckpt.npz"))
    assert rel_rms(g["y10"], g["y10_torch64"]) < 1e-6
    for k0, k1 in ((0, 1), (1, 2)):
        m = 0.0 if k0 == 0 else 0.99
        y, r = o.gl_one_iteration(g["s_mag"], g[f"y{k0}"], g[f"r{k0}"], m, 1024, 256)
        assert rel_rms(y, g[f"y{k1}"]) < 2e-6
        assert np.abs(r - g[f"r{k1}"]).max() / np.abs(g[f"r{k1}"]).max() < 2e-6
    # C port: few iterations stay inside the fp32 envelope
    for it in (0, 1, 2, 5):
        yc = c_oracle.gl_from_mag(g["s_mag"], g["turns"], 256, it, 0.99)
        assert rel_rms(yc, g[f"y{it}"]) < 2e-5, it


def test_n2048_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "n2048_gl.npz"))
    y = o.griffin_lim(g["s_mag"], g["turns"], 8, 0.99, 2048, 512, dtype=np.float64)
    assert rel_rms(y, g["y_torch64"]) < 1e-6
    yc = c_oracle.gl_from_mag(g["s_mag"], g["turns"], 512, 8, 0.99)
    assert rel_rms(yc, g["y_fp64"]) < 1e-4


def test_infer_port_matches_numpy():
    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    pinv = o.pinv_basis(basis).astype(np.float32)
    mel = o.synth_mel(3, 80, 40)
    u = o.phase_turns(1, 0, 513, 40)
    a = o.infer(mel, basis, 768, 1.7, 4, 0.99, u)
    b = c_oracle.infer(pinv, mel, u, 256, 1.7, 4, 0.99)
    assert abs(np.abs(a).max() - 1.0) < 1e-6 and a.shape == (256 * 39,)
    assert rel_rms(b, a) < 2e-5


def test_phase_turns_range_and_determinism():
    u = o.phase_turns(42, 5, 513, 33)
    assert u.dtype == np.float32 and u.shape == (513, 33)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert np.array_equal(u, o.phase_turns(42, 5, 513, 33))
    assert not np.array_equal(u, o.phase_turns(42, 6, 513, 33))
    assert abs(float(u.mean()) - 0.5) < 0.01


def test_postnet_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "postnet.npz"))
    layers = p.synth_weights(7)
    out = p.postnet(g["mel"], layers)
    assert out.shape == (80, 96)
    assert np.abs(out - g["out_torch32"]).max() < 2e-5
    assert np.abs(out - g["out_oracle64"]).max() < 2e-5
    # folded-BN form is the same function
    x = g["mel"].astype(np.float64)
    for i, l in enumerate(layers):
        w, b = p.fold_bn(l)
        x = p.conv1d_same(x, w, b)
        if i < 4:
            x = np.tanh(x)
    assert np.abs(g["mel"] + x - g["out_oracle64"]).max() < 2e-6


def test_pcm16_cast_semantics():
    """Rust `(sample * i16::MAX as f32) as i16` (src/lib.rs:155): truncation, saturation, NaN -> 0."""
    y = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 1.5, -1.5, 3.0517578e-05, -3.0517578e-05, np.nan, 0.99999, 2.0e-5], np.float32)
    want = np.array([0, 32767, -32767, 16383, -16383, 32767, -32768, 0, 0, 0, 32766, 0], np.int16)
    assert np.array_equal(o.pcm16(y), want)


def test_decoder_oracle_matches_torch_fixture(golden_dir):
    """oracle/decoder_oracle.py vs the decoder composed from torch modules (nn.LSTMCell, F.conv1d, masked
    softmax) in fp64, 16 free-running steps with the seeded prenet dropout (oracle/make_golden.py)."""
    import os

    from oracle import decoder_oracle as d

    g = np.load(os.path.join(golden_dir, "decoder.npz"))
    wt = d.synth_weights(int(g["weights_seed"]))
    mel, gate, align = d.run_decoder(wt, g["memory"], g["processed"], int(g["unpadded_len"]), seed=int(g["seed"]),
                                     gate_threshold=2.0, max_steps=16, return_aux=True)
    assert np.abs(mel - g["mel_torch64"]).max() < 1e-12
    assert np.abs(gate - g["gate_torch64"]).max() < 1e-12
    assert np.abs(align - g["align_torch64"]).max() < 1e-7          # stored as float32
    # fp32 evaluation of the same graph stays within 1e-4 of fp64 over 16 steps (the device tolerance's scale)
    mel32 = d.run_decoder(wt, g["memory"], g["processed"], int(g["unpadded_len"]), seed=int(g["seed"]), gate_threshold=2.0,
                          max_steps=16, dtype=np.float32)
    assert np.abs(mel32 - g["mel_torch64"]).max() < 1e-4
    # stop rule: the frame that fires the gate is kept (src/tacotron2/mod.rs:312-324)
    thr = 0.25
    full, gates, _ = d.run_decoder(wt, g["memory"], g["processed"], 19, seed=3, gate_threshold=2.0, max_steps=40, return_aux=True)
    fired = np.where(1 / (1 + np.exp(-gates)) > thr)[0]
    cut = d.run_decoder(wt, g["memory"], g["processed"], 19, seed=3, gate_threshold=thr, max_steps=40)
    assert cut.shape[0] == (fired[0] + 1 if len(fired) else 40)
    assert np.array_equal(cut, full[:cut.shape[0]])
    # dropout mask: Bernoulli(1/2), distinct per step / utterance, reproducible
    k = np.stack([d.dropout_keep(1, 0, i) for i in range(50)])
    assert 0.45 < k.mean() < 0.55 and not np.array_equal(k[0], k[1])
    assert np.array_equal(d.dropout_keep(1, 0, 7), k[7]) and not np.array_equal(d.dropout_keep(1, 1, 7), k[7])


def test_nnls_fista_lift_solves_librosas_problem():
    """The restated device algorithm (FISTA with restart) reaches the objective of librosa's own solver
    (lstsq start -> clip -> L-BFGS-B, oracle lift_nnls) on the same non-negative least-squares problem; the
    minimiser itself is not unique (80 equations, 513 unknowns), so the comparison is on the mel residual."""
    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    a = basis.astype(np.float64)
    for mel in (o.synth_mel(1234, 80, 6), np.log(np.maximum(a @ o.synth_speech_like_mag(5, 1024, 256, 6).astype(np.float64), 1e-5)).astype(np.float32)):
        b = np.exp(mel.astype(np.float32)).astype(np.float64)
        x_f, used = o.lift_nnls_fista(mel, basis, power=1.0, max_iter=400)
        x_l = o.lift_nnls(mel, basis, power=1.0)
        x_p = o.lift_pinv_clamp(mel, basis, power=1.0, dtype=np.float64)
        f = lambda x: 0.5 * ((a @ x - b) ** 2).sum(0)   # noqa: E731
        assert (x_f >= 0).all() and used.max() <= 400
        assert np.all(f(x_f) <= f(x_l) * 1.005 + 1e-8 * (b ** 2).sum(0))   # as good as L-BFGS-B at librosa's tolerances
        assert np.all(f(x_f) <= f(x_p) + 1e-15)                 # and never worse than the clipped pseudo-inverse start
