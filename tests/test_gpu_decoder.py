"""GPU parity tests of the Tacotron2 decoder loop (xdtts_decoder_*), through the C ABI, against the numpy
oracle (oracle/decoder_oracle.py) and the committed torch fixture (tests/golden/decoder.npz).
Floating point, fp32 on the device vs fp64 oracle: tolerances are written at each assertion."""
import os

import numpy as np
import pytest

from oracle import decoder_oracle as d

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def taco(built):
    from xdtts_b200 import tacotron2

    return tacotron2


@pytest.fixture(scope="module")
def weights():
    return d.synth_weights(11)


def test_golden_fixture_16_steps(taco, weights, golden_dir):
    """Free-running 16 steps vs the torch (nn.LSTMCell / F.conv1d) fp64 fixture: mel, gate, alignment."""
    g = np.load(os.path.join(golden_dir, "decoder.npz"))
    dec = taco.Decoder.from_weights(weights, gate_threshold=0.999999, max_steps=16, seed=int(g["seed"]))
    mels, gates, aligns = dec.run_batch([g["memory"]], [g["processed"]], [int(g["unpadded_len"])], return_aux=True)
    assert mels[0].shape == (80, 16)
    assert np.abs(mels[0].T - g["mel_torch64"]).max() < 2e-4
    assert np.abs(gates[0] - g["gate_torch64"]).max() < 2e-4
    assert np.abs(aligns[0] - g["align_torch64"]).max() < 1e-5
    assert np.all(aligns[0][:, int(g["unpadded_len"]):] == 0)          # masked positions
    assert np.abs(aligns[0].sum(1) - 1).max() < 1e-5


@pytest.mark.parametrize("dropout", [True, False])
def test_stop_rule_and_batch_in_lockstep(taco, weights, dropout):
    """Utterances of one batch stop at their own frame (sigmoid(gate) > threshold, firing frame kept,
    src/tacotron2/mod.rs:312-324); a batch returns what single calls return; all match the oracle."""
    thr = 0.25
    cases = [(31 + i, 40, 40 - 3 * i) for i in range(5)]          # (seed of encoder outputs, t_enc, unpadded)
    enc = [d.synth_encoder_outputs(s, t) for s, t, _ in cases]
    lens = [u for _, _, u in cases]
    dec = taco.Decoder.from_weights(weights, gate_threshold=thr, max_steps=60, prenet_dropout=dropout, seed=9)
    mels, gates, _ = dec.run_batch([m for m, _ in enc], [p for _, p in enc], lens, return_aux=True)
    stops = set()
    for b, ((mem, pm), u) in enumerate(zip(enc, lens)):
        ref, ref_gates, _ = d.run_decoder(weights, mem, pm, u, seed=9, utt=b, dropout=dropout, gate_threshold=thr, max_steps=60,
                                          return_aux=True)
        margin = np.abs(ref_gates - np.log(thr / (1 - thr))).min()
        assert margin > 1e-3, "test case sits on the stop threshold; pick another seed"
        assert mels[b].shape == (80, ref.shape[0]), (b, mels[b].shape, ref.shape)
        assert np.abs(mels[b].T - ref).max() < 5e-4
        assert np.abs(gates[b] - ref_gates).max() < 5e-4
        stops.add(ref.shape[0])
    assert len(stops) > 1, "the cases should stop at different frames"
    # utterance index selects the dropout stream: a single call on utterance 0 equals the batch's row 0
    single = dec.run(enc[0][0], enc[0][1], lens[0])
    assert np.array_equal(single, mels[0])


def test_more_than_one_group_and_max_steps(taco, weights):
    """11 utterances = one group of 8 + one of 3; nothing fires -> exactly max_steps frames."""
    enc = [d.synth_encoder_outputs(70 + i, 24) for i in range(11)]
    dec = taco.Decoder.from_weights(weights, gate_threshold=0.999999, max_steps=12, seed=2)
    mels = dec.run_batch([m for m, _ in enc], [p for _, p in enc], [24 - (i % 5) for i in range(11)])
    for b in (0, 7, 8, 10):
        ref = d.run_decoder(weights, enc[b][0], enc[b][1], 24 - (b % 5), seed=2, utt=b, gate_threshold=2.0, max_steps=12)
        assert mels[b].shape == (80, 12)
        assert np.abs(mels[b].T - ref).max() < 2e-4
    ms, steps = dec.last_timing()
    assert ms > 0 and steps == 12


def test_reference_shapes_t_enc_100(taco, weights):
    """The reference's fixed chunk: 100 padded positions (src/tacotron2/mod.rs:361-368), 77 real ones."""
    mem, pm = d.synth_encoder_outputs(5, 100)
    dec = taco.Decoder.from_weights(weights, gate_threshold=0.999999, max_steps=30, seed=4)
    mel = dec.run(mem, pm, 77)
    ref = d.run_decoder(weights, mem, pm, 77, seed=4, gate_threshold=2.0, max_steps=30)
    assert np.abs(mel.T - ref).max() < 3e-4


def test_synthesize_encoder_outputs_to_waveform(taco, weights):
    """decoder loop -> postnet -> lift -> Griffin-Lim on the device == the same chain of oracles, per utterance
    (different stop frames -> ragged tail batch)."""
    from oracle import gl_oracle as o
    from oracle import postnet_oracle as po
    from xdtts_b200 import griffin_lim

    layers = po.synth_weights(seed=7)
    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    voc = griffin_lim.GriffinLim.new(basis, 768, 1.7, 6, 0.99, seed=5)
    post = taco.Postnet.from_layers(layers)
    dec = taco.Decoder.from_weights(weights, gate_threshold=0.25, max_steps=40, prenet_dropout=False)
    enc = [d.synth_encoder_outputs(31 + i, 40) for i in (1, 2, 3)]
    lens = [37, 34, 31]
    waves, mels = taco.synthesize_batch(dec, post, voc, [m for m, _ in enc], [p for _, p in enc], lens, return_mels=True)
    frames = set()
    for b, ((mem, pm), u) in enumerate(zip(enc, lens)):
        dmel = d.run_decoder(weights, mem, pm, u, dropout=False, gate_threshold=0.25, max_steps=40).T      # [80, T]
        pmel = po.postnet(dmel.astype(np.float32), layers, dtype=np.float64).astype(np.float32)
        assert mels[b].shape == pmel.shape
        assert np.abs(mels[b] - pmel).max() < 2e-3
        t = pmel.shape[1]
        frames.add(t)
        assert waves[b].shape == (256 * (t - 1),) and np.isfinite(waves[b]).all()
        ref = o.infer(mels[b], basis, 768, 1.7, 6, 0.99, o.phase_turns(5, b, 513, t), dtype=np.float64)
        assert float(np.sqrt(np.mean((waves[b] - ref) ** 2))) < 1e-4
    assert len(frames) > 1


def test_longest_encoder_and_full_group(taco, weights):
    """The largest shapes one launch takes: 8 utterances in lockstep, 512 encoder positions (shared memory at its
    maximum), ragged unpadded lengths including 1."""
    enc = [d.synth_encoder_outputs(200 + i, 512) for i in range(8)]
    lens = [512, 400, 257, 256, 33, 32, 2, 1]
    dec = taco.Decoder.from_weights(weights, gate_threshold=0.999999, max_steps=3, seed=8)
    mels, gates, aligns = dec.run_batch([m for m, _ in enc], [p for _, p in enc], lens, return_aux=True)
    for b in (0, 2, 4, 6, 7):
        ref, _, ref_al = d.run_decoder(weights, enc[b][0], enc[b][1], lens[b], seed=8, utt=b, gate_threshold=2.0, max_steps=3,
                                       return_aux=True)
        assert np.abs(mels[b].T - ref).max() < 2e-4, b
        assert np.abs(aligns[b] - ref_al).max() < 1e-5, b
        assert np.all(aligns[b][:, lens[b]:] == 0)
    with pytest.raises(Exception):
        dec.run(np.zeros((513, 512), np.float32), np.zeros((513, 128), np.float32), 10)      # t_enc > 512


def test_errors(taco, weights):
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_SHAPE, XdttsError

    dec = taco.Decoder.from_weights(weights, max_steps=8)
    mem, pm = d.synth_encoder_outputs(1, 16)
    with pytest.raises(XdttsError) as e:
        dec.run(mem, pm, 17)                      # unpadded_len > t_enc
    assert e.value.code == ERR_SHAPE
    with pytest.raises(XdttsError) as e:
        dec.run(mem, pm, 0)
    assert e.value.code == ERR_SHAPE
    with pytest.raises(XdttsError) as e:
        dec.run(mem[:, :100], pm, 16)
    assert e.value.code == ERR_SHAPE
    bad = dict(weights)
    bad["query"] = bad["query"][:64]
    with pytest.raises(XdttsError) as e:
        taco.Decoder.from_weights(bad)
    assert e.value.code == ERR_SHAPE
    bad = dict(weights)
    bad["v"] = bad["v"].copy()
    bad["v"][3] = np.nan
    with pytest.raises(XdttsError) as e:
        taco.Decoder.from_weights(bad)
    assert e.value.code == ERR_BAD_ARG
    with pytest.raises(XdttsError) as e:
        taco.Decoder.from_weights(weights, gate_threshold=1.5)
    assert e.value.code == ERR_BAD_ARG


@pytest.mark.parametrize("use_lstm_op", [False, True])
def test_decoder_from_onnx_equals_from_weights(taco, weights, tmp_path, use_lstm_op):
    """xdtts_decoder_create_from_onnx (Tacotron2::load for decoder_iter.onnx, src/tacotron2/mod.rs:251-254) on a FULL-SIZE
    file written here by PyTorch's ONNX serializer from the same weights -- with the LSTM cells as ONNX LSTM operators
    (gate order i, o, f, c: NVIDIA's export script) and as their Gemm / Split decomposition -- gives, bit for bit, the
    decoder built from the arrays."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import make_foreign_decoder_onnx as gen

    path = tmp_path / "decoder_iter.onnx"
    gen.build(gen.FULL, str(path), weights=weights, use_lstm_op=use_lstm_op)
    assert path.stat().st_size > 70e6                                   # 18.2 M fp32 weights
    dims, got = taco.read_onnx_decoder(path)
    assert dims["lstm_form"] == int(use_lstm_op) and dims["has_dropout"] == 1
    for name in taco.DECODER_TENSORS:
        assert np.array_equal(got[name], np.asarray(weights[name], np.float32).ravel()), name
    mem, pm = d.synth_encoder_outputs(5, 30)
    a = taco.Decoder.from_onnx(path, gate_threshold=0.999999, max_steps=12, seed=4)
    b = taco.Decoder.from_weights(weights, gate_threshold=0.999999, max_steps=12, seed=4)
    ya, yb = a.run(mem, pm, 27), b.run(mem, pm, 27)
    assert ya.shape == (80, 12) and np.array_equal(ya, yb)
    a.close()
    b.close()
