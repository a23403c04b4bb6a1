"""GPU parity tests of the postnet (Tacotron2 postnet session, /root/reference
src/tacotron2/mod.rs:344-357) through the C ABI, against the numpy oracle (unfused BatchNorm, fp64)
and the committed golden fixture (torch conv1d + batch_norm, tests/golden/postnet.npz).

Tolerances (floating point path, stated per precision mode; values are ln-mels of magnitude ~1-8):
  fp32 CUDA-core kernel   max abs error <= 1e-4   (SURVEY.md 8d gate 5 "fp32 path")
  bf16x3 tensor-core      max abs error <= 2e-4   (three-pass split, fp32 accumulate in TMEM)
  bf16 tensor-core        max abs error <= 6e-2   (single pass, 8-bit mantissa inputs)"""
import os

import numpy as np
import pytest

from oracle import gl_oracle as o
from oracle import postnet_oracle as po

pytestmark = pytest.mark.gpu

TOL = {0: 2e-4, 1: 6e-2, 2: 1e-4}


@pytest.fixture(scope="module")
def taco(built):
    from xdtts_b200 import tacotron2

    return tacotron2


@pytest.fixture(scope="module")
def layers():
    return po.synth_weights(seed=7)


@pytest.mark.parametrize("precision", [2, 0, 1])
def test_golden_fixture(taco, layers, golden_dir, precision):
    g = np.load(os.path.join(golden_dir, "postnet.npz"))
    post = taco.Postnet.from_layers(layers, precision=precision)
    out = post.run(g["mel"])
    assert out.shape == g["mel"].shape and out.dtype == np.float32 and np.isfinite(out).all()
    assert np.abs(out - g["out_oracle64"]).max() < TOL[precision]
    assert np.abs(out - g["out_torch32"]).max() < TOL[precision] + 2e-5
    # the residual is really added: out - mel is the conv stack, not zero
    assert np.abs(out - g["mel"]).max() > 1e-2


@pytest.mark.parametrize("precision", [0, 2])
@pytest.mark.parametrize("t", [1, 2, 5, 127, 128, 129, 300])
def test_lengths_and_tile_edges(taco, layers, precision, t):
    """T around the 128-row tile size, and shorter than the receptive field (21 frames)."""
    mel = o.synth_mel(40 + t, 80, t)
    post = taco.Postnet.from_layers(layers, precision=precision)
    ref = po.postnet(mel, layers, dtype=np.float64)
    out = post.run(mel)
    assert np.abs(out - ref).max() < TOL[precision], t


@pytest.mark.parametrize("precision", [0, 1])
def test_ragged_batch_equals_single_calls(taco, layers, precision):
    """Zero rows between stacked utterances are the conv padding: no leakage across utterances."""
    ts = [3, 130, 17, 256, 64, 1, 99]
    mels = [o.synth_mel(200 + i, 80, t) for i, t in enumerate(ts)]
    post = taco.Postnet.from_layers(layers, precision=precision)
    batch = post.run_batch(mels)
    for m, b in zip(mels, batch):
        single = post.run(m)
        assert np.array_equal(single, b)
        assert np.abs(b - po.postnet(m, layers, dtype=np.float64)).max() < TOL[precision]
    again = post.run_batch(mels)       # buffers are reused: the gap rows must still be zero
    for a, b in zip(batch, again):
        assert np.array_equal(a, b)
    longer = post.run_batch([o.synth_mel(1, 80, 400)])   # another shape in between, then back
    assert longer[0].shape == (80, 400)
    third = post.run_batch(mels)
    for a, b in zip(batch, third):
        assert np.array_equal(a, b)


def test_tensor_path_agrees_with_cuda_core_path(taco, layers):
    """On-device cross-check: tcgen05 bf16x3 vs the plain fp32 kernel on a cfg3-sized utterance."""
    mel = o.synth_mel(77, 80, 1000)
    a = taco.Postnet.from_layers(layers, precision=0).run(mel)
    b = taco.Postnet.from_layers(layers, precision=2).run(mel)
    assert np.abs(a - b).max() < 2e-4
    c = taco.Postnet.from_layers(layers, precision=1).run(mel)
    assert np.abs(c - b).max() < 6e-2
    assert np.abs(c - b).max() > np.abs(a - b).max()       # the split really buys precision


def test_other_architectures(taco):
    """No BatchNorm / no bias, other channel counts (N tile of 128, 64-channel padding, one layer)."""
    rng = np.random.default_rng(3)
    for channels in ([32, 384, 32], [80, 80], [48, 64, 256, 48]):
        layers = po.synth_weights(seed=11, channels=tuple(channels))
        layers[0].pop("b")
        for k in ("gamma", "beta", "mean", "var"):
            layers[-1].pop(k)
        mel = rng.uniform(-4, 1, (channels[0], 150)).astype(np.float32)
        full = []
        for l in layers:          # the oracle wants every key: identity BatchNorm / zero bias
            c = l["w"].shape[0]
            d = dict(w=l["w"], b=l.get("b", np.zeros(c, np.float32)), gamma=l.get("gamma", np.ones(c, np.float32)),
                     beta=l.get("beta", np.zeros(c, np.float32)), mean=l.get("mean", np.zeros(c, np.float32)),
                     var=l.get("var", np.ones(c, np.float32) - np.float32(po.BN_EPS)))
            full.append(d)
        ref = po.postnet(mel, full, dtype=np.float64)
        for precision in (0, 2):
            out = taco.Postnet.from_layers(layers, precision=precision).run(mel)
            assert np.abs(out - ref).max() < TOL[precision], (channels, precision)


def test_plan_feeds_vocoder_and_tail_call(taco, layers, built):
    """postnet -> lift -> Griffin-Lim without leaving HBM == the two public calls chained on the host."""
    from xdtts_b200 import griffin_lim

    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    voc = griffin_lim.GriffinLim.new(basis, 768, 1.7, 8, 0.99, run_frames=8)
    post = taco.Postnet.from_layers(layers, precision=0)
    ts = [40, 150]
    # decoder-like mels scaled so that postnet output stays in the ln-mel range
    mels = [o.synth_mel(300 + i, 80, t) for i, t in enumerate(ts)]
    phs = [o.phase_turns(5, i, 513, t) for i, t in enumerate(ts)]
    waves, pmels = taco.infer_tail_batch(post, voc, mels, phs, return_mels=True)
    chained_mels = post.run_batch(mels)
    chained = voc.infer_batch(chained_mels, phs)
    for a, b in zip(pmels, chained_mels):
        assert np.array_equal(a, b)
    for a, b, t in zip(waves, chained, ts):
        assert a.shape == (256 * (t - 1),)
        assert np.array_equal(a, b)
    # and against the oracle end to end (fp64 postnet -> fp64 vocoder), same initial phase
    for m, ph, w in zip(mels, phs, waves):
        pm = po.postnet(m, layers, dtype=np.float64).astype(np.float32)
        ref = o.infer(pm, basis, 768, 1.7, 8, 0.99, ph, dtype=np.float64)
        assert float(np.sqrt(np.mean((w - ref) ** 2))) < 1e-3
    # plan API: run(feed=...) then the vocoder plan
    pp, gp = post.plan(ts), voc.plan(ts)
    pp.upload(mels)
    ms = pp.run(feed=gp)
    assert ms > 0
    gp.upload(2, phs)
    gp.run(2)
    for a, b in zip(gp.download(), waves):
        assert np.array_equal(a, b)
    single = taco.infer_tail(post, voc, mels[0], phs[0])
    assert np.array_equal(single, waves[0])


def test_pipe_equals_blocking_calls_bitwise(taco, layers, built):
    """xdtts_pipe (copies overlapped with the neighbouring batches' kernels) returns, batch for batch and in
    push order, the bits of the blocking calls -- vocoder alone and postnet + vocoder, explicit and seeded phase."""
    from xdtts_b200 import griffin_lim
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_SHAPE, XdttsError

    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    voc = griffin_lim.GriffinLim.new(basis, 768, 1.7, 6, 0.99, seed=3, fixed_seed=True)
    post = taco.Postnet.from_layers(layers, precision=0)
    ts = [33, 90, 5]
    batches = [[o.synth_mel(1000 + 10 * j + i, 80, t) for i, t in enumerate(ts)] for j in range(5)]
    phases = [[o.phase_turns(j, i, 513, t) for i, t in enumerate(ts)] for j in range(5)]
    # vocoder alone, depth 2, explicit phase for even batches and the seeded generator for odd ones
    want = [voc.infer_batch(m, ph if j % 2 == 0 else None) for j, (m, ph) in enumerate(zip(batches, phases))]
    pipe = voc.pipe(ts, depth=2)
    got = []
    for j, (m, ph) in enumerate(zip(batches, phases)):
        r = pipe.push(m, ph if j % 2 == 0 else None)
        assert (r is None) == (j < 2)
        if r is not None:
            got.append(r)
    assert pipe.pending() == 2
    got += pipe.flush()
    assert pipe.pending() == 0 and pipe.pop() is None and len(got) == 5
    for a, b in zip(got, want):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    pipe.close()
    # the whole tail, depth 3, with the postnet mels returned
    want = [taco.infer_tail_batch(post, voc, m, ph, return_mels=True) for m, ph in zip(batches, phases)]
    pipe = voc.pipe(ts, depth=3, postnet=post, want_mels=True)
    got = [r for r in (pipe.push(m, ph) for m, ph in zip(batches, phases)) if r is not None] + pipe.flush()
    assert len(got) == 5
    for (gw, gm), (ww, wm) in zip(got, want):
        for x, y in zip(gw, ww):
            assert np.array_equal(x, y)
        for x, y in zip(gm, wm):
            assert np.array_equal(x, y)
    pipe.close()
    # pageable host buffers straight through the C ABI (staged copies), depth 1
    import ctypes

    from xdtts_b200._ffi import check, fptr_array, load_library

    lib = load_library()
    q = ctypes.c_void_p()
    t_arr = (ctypes.c_int * len(ts))(*ts)
    check(lib.xdtts_pipe_create(voc._h, None, t_arr, len(ts), 1, ctypes.byref(q)))
    outs = [np.zeros(256 * (t - 1), np.float32) for t in ts]
    check(lib.xdtts_pipe_push(q, fptr_array(batches[0]), fptr_array(phases[0]), None, fptr_array(outs)))
    check(lib.xdtts_pipe_flush(q))
    for x, y in zip(outs, voc.infer_batch(batches[0], phases[0])):
        assert np.array_equal(x, y)
    assert lib.xdtts_pipe_push(q, fptr_array(batches[0]), None, fptr_array(outs), fptr_array(outs)) == ERR_BAD_ARG  # out_mels without postnet
    lib.xdtts_pipe_destroy(q)
    with pytest.raises(XdttsError) as e:
        voc.pipe([33, 1], depth=2)          # one frame = no samples
    assert e.value.code == ERR_SHAPE
    with pytest.raises(XdttsError) as e:
        voc.pipe(ts, depth=0)
    assert e.value.code == ERR_BAD_ARG


def test_errors(taco, layers):
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_SHAPE, XdttsError

    post = taco.Postnet.from_layers(layers)
    with pytest.raises(XdttsError) as e:
        post.run(np.zeros((79, 10), np.float32))
    assert e.value.code == ERR_SHAPE
    with pytest.raises(XdttsError) as e:
        post.run(np.zeros((80, 0), np.float32))
    assert e.value.code == ERR_SHAPE
    out = post.run(np.zeros((80, 7), np.float32))           # zeros in -> the bias response, finite
    assert np.isfinite(out).all()
    p = post.plan([10])
    with pytest.raises(XdttsError) as e:
        p.upload([np.zeros((80, 11), np.float32)])
    assert e.value.code == ERR_SHAPE
    assert ERR_BAD_ARG < 0


def test_load_from_onnx_file(taco, layers, tmp_path):
    """Tacotron2::load's postnet session (src/tacotron2/mod.rs:256-259): same bits as passing the arrays."""
    from onnx_writer import postnet_model

    path = tmp_path / "postnet.onnx"
    path.write_bytes(postnet_model(layers, eps=po.BN_EPS))
    mel = o.synth_mel(9, 80, 200)
    a = taco.Postnet.load(path).run(mel)
    b = taco.Postnet.from_layers(layers).run(mel)
    assert np.array_equal(a, b)
    assert np.abs(a - po.postnet(mel, layers, dtype=np.float64)).max() < TOL[0]


def test_cfg3_full_size_batch(taco, layers):
    """BASELINE.json configs[2] shape (32 x 80x1000): tensor path vs the CUDA-core fp32 path on the whole batch,
    and no cross-utterance leakage at full tile occupancy."""
    b, t = 32, 1000
    mels = [o.synth_mel(1234 + i, 80, t) for i in range(b)]
    tc = taco.Postnet.from_layers(layers, precision=0)
    outs = tc.run_batch(mels)
    ref = taco.Postnet.from_layers(layers, precision=2).run_batch(mels)
    worst = max(float(np.abs(a - r).max()) for a, r in zip(outs, ref))
    assert worst < 2e-4, worst
    assert np.array_equal(tc.run(mels[17]), outs[17])
    assert np.abs(outs[5] - po.postnet(mels[5], layers, dtype=np.float64)).max() < TOL[0]


def test_load_files_written_by_pytorch(taco, golden_dir):
    """Postnet.load (xdtts_postnet_create_from_onnx) on ONNX files serialised by PyTorch's exporter -- BatchNorm kept and
    BatchNorm folded -- reproduces PyTorch's own output of that module (tests/make_foreign_onnx.py)."""
    import os

    g = np.load(os.path.join(golden_dir, "postnet_torch_export.npz"))
    for tag in ("bn", "fused"):
        post = taco.Postnet.load(os.path.join(golden_dir, "postnet_torch_export_%s.onnx" % tag))
        y = post.run(g["x"])
        assert y.shape == g["y"].shape
        assert np.abs(y - g["y"]).max() < 2e-4, tag
        post.close()


@pytest.mark.gpu
def test_tail_pipe_and_pool_on_the_unfused_path(taco, layers, built):
    """The callers of the vocoder do not care which kernels a plan runs: the fused tail call, the streaming pipe and the
    multi-GPU pool give the bits of the plain calls on a geometry the fused kernel does not cover (hop 200) and on a batch
    with a three-frame utterance (which sends the shipped geometry through the un-fused kernels too)."""
    from xdtts_b200 import griffin_lim

    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    post = taco.Postnet.from_layers(layers, precision=0)
    for hop, ts in ((200, [12, 30]), (256, [3, 25])):
        voc = griffin_lim.GriffinLim.new(basis, 1024 - hop, 1.7, 3, 0.99)
        mels = [o.synth_mel(500 + i, 80, t) for i, t in enumerate(ts)]
        phs = [o.phase_turns(6, i, 513, t) for i, t in enumerate(ts)]
        waves, pmels = taco.infer_tail_batch(post, voc, mels, phs, return_mels=True)
        chained = voc.infer_batch(post.run_batch(mels), phs)
        for a, b, t in zip(waves, chained, ts):
            assert a.shape == (hop * (t - 1),) and np.array_equal(a, b)
        pm = po.postnet(mels[1], layers, dtype=np.float64).astype(np.float32)
        ref = o.infer(pm, basis, 1024 - hop, 1.7, 3, 0.99, phs[1], dtype=np.float64)
        assert float(np.sqrt(np.mean((waves[1] - ref) ** 2))) < 1e-3
        pipe = voc.pipe(ts, depth=2)
        fixed = griffin_lim.GriffinLim.new(basis, 1024 - hop, 1.7, 3, 0.99, seed=9, fixed_seed=True)
        want = fixed.infer_batch(mels)
        pipe2 = fixed.pipe(ts, depth=2)
        got = [r for r in (pipe2.push(mels) for _ in range(2)) if r is not None] + pipe2.flush()
        assert len(got) == 2 and all(np.array_equal(a, b) for a, b in zip(got[0], want))
        pipe.close()
        pipe2.close()
        pool = griffin_lim.GriffinLimPool.new(basis, 1024 - hop, 1.7, 3, 0.99, seed=9, fixed_seed=True)
        for a, b in zip(pool.infer_batch(mels), want):
            assert np.array_equal(a, b)
        pool.close()
