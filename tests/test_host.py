"""CPU tests of the product's host side: the C ABI library loads and exports every symbol of
include/xdtts_b200.h, host math (filterbank, pseudo-inverse), argument validation, and the kernel's
lane program run by the CPU emulator against the oracle.  No CUDA compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import rel_rms
from oracle import gl_oracle as o


@pytest.fixture(scope="module")
def lib(built):
    from xdtts_b200 import _ffi

    return _ffi.load_library()


def test_library_exports_every_declared_symbol(lib, built):
    from xdtts_b200 import _ffi

    hdr = open(os.path.join(os.path.dirname(built.__file__), "include", "xdtts_b200.h")).read()
    declared = set(re.findall(r"\b(xdtts_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "header parse failed"
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    assert b"sm_100a" in lib.xdtts_version()


def test_header_is_plain_c(built, tmp_path):
    """The boundary is a C ABI: include/xdtts_b200.h must compile as C99 (no C++, no CUDA or torch types)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    inc = os.path.join(os.path.dirname(built.__file__), "include")
    src = tmp_path / "hdr.c"
    src.write_text('#include "xdtts_b200.h"\nint main(void) { xdtts_gl_opts o = {0}; xdtts_decoder_opts d = {0}; (void)o; (void)d; return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)


def test_c_client_links_and_calls(built, lib, tmp_path):
    """A plain C program links against the shared library and calls it (the host-only entry points: no GPU needed)."""
    import shutil
    import subprocess

    from xdtts_b200 import _ffi, griffin_lim

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    inc = os.path.join(os.path.dirname(built.__file__), "include")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    src = tmp_path / "client.c"
    src.write_text(r"""
#include <stdio.h>
#include <stdlib.h>
#include "xdtts_b200.h"
int main(void) {
    float* fb = (float*)malloc(sizeof(float) * 80 * 513);
    double sum = 0.0;
    int i, rc = xdtts_mel_filter_bank(22050.0f, 1024, 80, 0.0f, 8000.0f, fb);
    if (rc != XDTTS_OK) { printf("error %d: %s\n", rc, xdtts_last_error()); return 1; }
    for (i = 0; i < 80 * 513; i++) sum += fb[i];
    printf("%s %.9f\n", xdtts_version(), sum);
    rc = xdtts_mel_filter_bank(22050.0f, 1024, 80, 9000.0f, 8000.0f, fb);     /* fmax <= fmin */
    printf("%d %s\n", rc, xdtts_last_error());
    free(fb);
    return 0;
}
""")
    exe = tmp_path / "client"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lxdtts_b200",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    ref = griffin_lim.mel.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0).astype(np.float64).sum()
    assert out[0].startswith("xdtts_b200") and abs(float(out[0].split()[-1]) - ref) < 1e-6
    assert out[1].startswith("-1 ") and "fmax" in out[1]


def test_mel_filter_bank_matches_golden(lib, golden_dir):
    from xdtts_b200 import griffin_lim

    g = np.load(os.path.join(golden_dir, "melbank.npz"))
    for n_fft in (1024, 2048):
        fb = griffin_lim.mel.create_mel_filter_bank(22050.0, n_fft, 80, 0.0, 8000.0)
        assert fb.shape == (80, n_fft // 2 + 1) and fb.dtype == np.float32
        assert np.abs(fb - g["torchaudio_%d" % n_fft]).max() < 2e-7     # independent implementation
        assert np.abs(fb - g["oracle_%d" % n_fft]).max() < 1e-8
    # fmax = None -> sr/2
    fb = griffin_lim.mel.create_mel_filter_bank(16000.0, 512, 40, 20.0, None)
    assert np.abs(fb - o.create_mel_filter_bank(16000.0, 512, 40, 20.0, None)).max() < 1e-8


def test_pinv_matches_numpy(lib):
    from xdtts_b200 import griffin_lim

    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    p = griffin_lim.pinv(basis)
    ref = np.linalg.pinv(basis.astype(np.float64))
    assert p.shape == (513, 80)
    assert np.abs(p - ref).max() / np.abs(ref).max() < 1e-6
    assert (p[0] == 0).all() and (p[372:] == 0).all()       # empty columns
    # rank-deficient input (duplicated + zero row) still gives the Moore-Penrose inverse
    rng = np.random.default_rng(0)
    a = rng.standard_normal((6, 11)).astype(np.float32)
    a[3] = a[1]
    a[5] = 0
    p = griffin_lim.pinv(a)
    ref = np.linalg.pinv(a.astype(np.float64), rcond=1e-6)
    assert np.abs(p - ref).max() < 1e-5


def test_argument_validation_without_gpu(lib):
    from xdtts_b200 import griffin_lim
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_CUDA, ERR_SHAPE, ERR_UNSUPPORTED, XdttsError

    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    try:                                                                # hop != n_fft/4: the un-fused path takes it -- the only
        griffin_lim.GriffinLim.new(basis, 1024 - 200, 1.7, 30, 0.99)    # way this call fails is the missing GPU
    except XdttsError as err:
        assert err.code == ERR_CUDA
    with pytest.raises(XdttsError) as e:
        griffin_lim.GriffinLim.new(basis[:, :400], 768, 1.7, 30, 0.99)  # n_fft = 798
    assert e.value.code == ERR_UNSUPPORTED
    with pytest.raises(XdttsError) as e:
        griffin_lim.GriffinLim.new(basis, 768, -1.0, 30, 0.99)
    assert e.value.code == ERR_BAD_ARG
    with pytest.raises(XdttsError) as e:
        griffin_lim.GriffinLim.new(basis, 768, 1.7, 30, float("nan"))
    assert e.value.code == ERR_BAD_ARG
    with pytest.raises(XdttsError) as e:
        griffin_lim.GriffinLim.new(basis[0], 768, 1.7, 30, 0.99)
    assert e.value.code == ERR_SHAPE
    bad = basis.copy()
    bad[3, 7] = np.inf
    with pytest.raises(XdttsError) as e:
        griffin_lim.GriffinLim.new(bad, 768, 1.7, 30, 0.99)
    assert e.value.code == ERR_BAD_ARG
    with pytest.raises(XdttsError):
        griffin_lim.mel.create_mel_filter_bank(22050.0, 1024, 80, 9000.0, 8000.0)
    assert b"fmax" in lib.xdtts_last_error()


def test_decoder_argument_validation_without_gpu(lib):
    import ctypes

    from oracle import decoder_oracle as d
    from xdtts_b200 import tacotron2
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_CUDA, ERR_SHAPE, XdttsError

    h = ctypes.c_void_p()
    assert lib.xdtts_decoder_create(None, None, 0, ctypes.byref(h)) == ERR_BAD_ARG
    assert lib.xdtts_decoder_max_steps(None) == ERR_BAD_ARG
    assert lib.xdtts_decoder_infer_batch(None, None, None, 10, None, 1, None, None, None, None) == ERR_BAD_ARG
    lib.xdtts_decoder_destroy(None)
    wt = d.synth_weights(11)
    bad = dict(wt)
    del bad["gate_b"]
    with pytest.raises(XdttsError) as e:
        tacotron2.Decoder.from_weights(bad)
    assert e.value.code == ERR_BAD_ARG
    bad = dict(wt)
    bad["dec_w_ih"] = bad["dec_w_ih"][:, :1024]
    with pytest.raises(XdttsError) as e:
        tacotron2.Decoder.from_weights(bad)
    assert e.value.code == ERR_SHAPE
    with pytest.raises(XdttsError) as e:
        tacotron2.Decoder.from_weights(wt, max_steps=-1)
    assert e.value.code == ERR_BAD_ARG
    import torch

    if not torch.cuda.is_available():          # no CPU fallback: a valid request fails for want of a device
        with pytest.raises(XdttsError) as e:
            tacotron2.Decoder.from_weights(wt)
        assert e.value.code == ERR_CUDA


def test_pipe_argument_validation_without_gpu(lib):
    import ctypes

    from xdtts_b200._ffi import ERR_BAD_ARG

    q = ctypes.c_void_p()
    ts = (ctypes.c_int * 2)(10, 12)
    assert lib.xdtts_pipe_create(None, None, ts, 2, 2, ctypes.byref(q)) == ERR_BAD_ARG
    assert not q.value
    assert lib.xdtts_pipe_push(None, None, None, None, None) == ERR_BAD_ARG
    assert lib.xdtts_pipe_pop(None) == ERR_BAD_ARG
    assert lib.xdtts_pipe_flush(None) == ERR_BAD_ARG
    assert lib.xdtts_pipe_pending(None) == ERR_BAD_ARG
    lib.xdtts_pipe_destroy(None)    # no-op


def test_no_cpu_fallback(lib):
    """Without a usable sm_100 device the product refuses to work instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from xdtts_b200 import griffin_lim
    from xdtts_b200._ffi import ERR_CUDA, XdttsError

    basis = o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    with pytest.raises(XdttsError) as e:
        griffin_lim.GriffinLim.new(basis, 768, 1.7, 30, 0.99)
    assert e.value.code == ERR_CUDA


# ---------------------------------------------------------------- lane program (CPU emulation)
@pytest.mark.parametrize("n_fft,hop,t", [(1024, 256, 37), (2048, 512, 21), (512, 128, 30), (512, 128, 10)])
def test_lane_program_matches_oracle(built, n_fft, hop, t):
    from emu import emu

    k = n_fft // 2 + 1
    s = o.synth_speech_like_mag(5, n_fft, hop, t)
    tu = o.phase_turns(9, 0, k, t)
    for it in (0, 1, 3):
        y64 = o.griffin_lim(s, tu, it, 0.99, n_fft, hop, dtype=np.float64)
        for run_frames in (1000, 8, 4):     # one run; several runs; shortest legal runs (5 frames: 4 is raised to 5)
            y, r, peak, n_runs = emu.gl_from_mag(s, tu, it, 0.99, run_frames)
            assert rel_rms(y, y64) < 5e-7, (it, run_frames)
            assert abs(peak - np.abs(y).max()) < 1e-6


def test_lane_program_rebuilt_spectrum_and_options(built):
    from emu import emu

    n_fft, hop, t = 1024, 256, 16
    s = o.synth_speech_like_mag(3, n_fft, hop, t)
    tu = o.phase_turns(1, 0, 513, t)
    # after 2 iterations R holds stft(y_0) (the last launch does not store): packed slot 0 = (R[0], R[M])
    ck = {0: None}
    o.griffin_lim(s, tu, 2, 0.99, n_fft, hop, dtype=np.float64, checkpoints=ck)
    r_ref = o.stft(ck[0][0], n_fft, hop, dtype=np.float64).T           # [T, K]
    _, r, _, _ = emu.gl_from_mag(s, tu, 2, 0.99, 5)
    scale = np.abs(r_ref).max()
    assert np.abs(r[:, 1:] - r_ref[:, 1:512]).max() / scale < 1e-6
    assert np.abs(r[:, 0].real - r_ref[:, 0].real).max() / scale < 1e-6
    assert np.abs(r[:, 0].imag - r_ref[:, 512].real).max() / scale < 1e-6
    # constant (zero) padding variant
    y64 = o.griffin_lim(s, tu, 2, 0.99, n_fft, hop, pad_mode=o.PAD_CONSTANT, dtype=np.float64)
    y, _, _, _ = emu.gl_from_mag(s, tu, 2, 0.99, 6, pad_mode=1)
    assert rel_rms(y, y64) < 5e-7
    # momentum 0 (classic Griffin-Lim)
    y64 = o.griffin_lim(s, tu, 3, 0.0, n_fft, hop, dtype=np.float64)
    y, _, _, _ = emu.gl_from_mag(s, tu, 3, 0.0, 6)
    assert rel_rms(y, y64) < 5e-7
    # seeded phase generator == oracle's phase_turns
    y_seed, _, _, _ = emu.gl_from_mag(s, None, 1, 0.99, 6, seed=1)
    y_expl, _, _, _ = emu.gl_from_mag(s, tu, 1, 0.99, 6)
    assert rel_rms(y_seed, y_expl) < 1e-6


def test_postnet_argument_validation_without_gpu(lib):
    from oracle import postnet_oracle as po
    from xdtts_b200 import tacotron2
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_CUDA, ERR_SHAPE, ERR_UNSUPPORTED, XdttsError

    layers = po.synth_weights(seed=7)
    with pytest.raises(XdttsError) as e:        # residual needs matching channel counts
        tacotron2.Postnet.from_layers(layers[:2])
    assert e.value.code == ERR_SHAPE
    bad = po.synth_weights(seed=7, channels=(80, 100, 80))      # 100 is not a multiple of 16
    with pytest.raises(XdttsError) as e:
        tacotron2.Postnet.from_layers(bad)
    assert e.value.code == ERR_UNSUPPORTED
    k3 = [dict(l, w=l["w"][:, :, :3].copy()) for l in layers]
    with pytest.raises(XdttsError) as e:
        tacotron2.Postnet.from_layers(k3)
    assert e.value.code == ERR_UNSUPPORTED
    partial = [dict(l) for l in layers]
    partial[1].pop("var")
    with pytest.raises(XdttsError) as e:
        tacotron2.Postnet.from_layers(partial)
    assert e.value.code == ERR_BAD_ARG
    with pytest.raises(XdttsError) as e:
        tacotron2.Postnet.from_layers(layers, precision=7)
    assert e.value.code == ERR_BAD_ARG
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(XdttsError) as e:    # valid arguments, no device: refuses, never computes on the CPU
            tacotron2.Postnet.from_layers(layers)
        assert e.value.code == ERR_CUDA


def test_onnx_postnet_reader(lib, tmp_path):
    """The library's own protobuf reader recovers every initializer of a hand-encoded postnet.onnx."""
    from onnx_writer import postnet_model
    from oracle import postnet_oracle as po
    from xdtts_b200 import tacotron2
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_SHAPE, ERR_UNSUPPORTED, XdttsError

    layers = po.synth_weights(seed=7)
    for raw in (True, False):                       # raw_data and packed float_data encodings
        path = tmp_path / ("postnet_%d.onnx" % raw)
        path.write_bytes(postnet_model(layers, eps=1e-5, raw=raw))
        got = tacotron2.read_onnx_postnet(path)
        assert len(got) == 5
        for a, b in zip(got, layers):
            assert abs(a["eps"] - 1e-5) < 1e-12
            for k in ("w", "b", "gamma", "beta", "mean", "var"):
                assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    # no BatchNormalization nodes (a fused export), no bias on one layer
    nb = [dict(w=l["w"], b=l["b"]) for l in layers]
    nb[2].pop("b")
    path = tmp_path / "fused.onnx"
    path.write_bytes(postnet_model(nb))
    got = tacotron2.read_onnx_postnet(path)
    assert all("gamma" not in l for l in got) and "b" not in got[2] and np.array_equal(got[4]["w"], layers[4]["w"])
    # error paths: the LFS pointer the reference repo actually ships, garbage, a missing file, a broken graph
    ptr = tmp_path / "pointer.onnx"
    ptr.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:7a65\nsize 17414016\n")
    with pytest.raises(XdttsError) as e:
        tacotron2.read_onnx_postnet(ptr)
    assert e.value.code == ERR_BAD_ARG and "LFS" in e.value.message
    junk = tmp_path / "junk.onnx"
    junk.write_bytes(bytes(range(256)) * 4)
    with pytest.raises(XdttsError):
        tacotron2.read_onnx_postnet(junk)
    with pytest.raises(XdttsError) as e:
        tacotron2.read_onnx_postnet(tmp_path / "missing.onnx")
    assert e.value.code == ERR_BAD_ARG
    bad = [dict(l) for l in layers]
    bad[1]["w"] = bad[1]["w"][:, :100].copy()      # layer 1 takes 100 channels, layer 0 makes 512
    path = tmp_path / "bad.onnx"
    path.write_bytes(postnet_model(bad))
    with pytest.raises(XdttsError) as e:
        tacotron2.read_onnx_postnet(path)
    assert e.value.code == ERR_SHAPE
    trunc = tmp_path / "trunc.onnx"
    trunc.write_bytes(postnet_model(layers)[:200000])
    with pytest.raises(XdttsError):
        tacotron2.read_onnx_postnet(trunc)
    # the graph must BE the postnet the device computes: another activation, a missing residual or a foreign operator is
    # rejected instead of loading into a network that computes something else (inference-mode Dropout is accepted)
    small = [dict(l) for l in po.synth_weights(seed=3, channels=(8, 16, 16, 8))]
    ok = tmp_path / "dropout.onnx"
    ok.write_bytes(postnet_model(small, dropout=True))
    assert len(tacotron2.read_onnx_postnet(ok)) == 3
    for kw, word in ((dict(activation="Relu"), "Relu"), (dict(activation=None), "Tanh"), (dict(residual=False), "residual"),
                     (dict(extra_op="Sigmoid"), "Sigmoid"), (dict(extra_op="Tanh"), "Tanh")):
        path = tmp_path / "wrong.onnx"
        path.write_bytes(postnet_model(small, **kw))
        with pytest.raises(XdttsError) as e:
            tacotron2.read_onnx_postnet(path)
        assert e.value.code == ERR_UNSUPPORTED and word in e.value.message, (kw, e.value.message)


def test_onnx_reader_on_files_written_by_pytorch(lib, golden_dir):
    """A FOREIGN encoder: tests/golden/postnet_torch_export_*.onnx were serialised by PyTorch's ONNX exporter
    (tests/make_foreign_onnx.py), not by tests/onnx_writer.py -- initializer names, packing, attribute order and the
    BatchNorm folding are the exporter's, not ours."""
    from xdtts_b200 import tacotron2

    g = np.load(os.path.join(golden_dir, "postnet_torch_export.npz"))
    bn = tacotron2.read_onnx_postnet(os.path.join(golden_dir, "postnet_torch_export_bn.onnx"))
    assert len(bn) == 5
    for i, l in enumerate(bn):
        for k, name in (("w", "w"), ("b", "b"), ("gamma", "gamma"), ("beta", "beta"), ("mean", "mean"), ("var", "var")):
            assert np.array_equal(l[k], g["%s%d" % (name, i)]), (i, k)
        assert abs(l["eps"] - 1e-5) < 1e-9
    fused = tacotron2.read_onnx_postnet(os.path.join(golden_dir, "postnet_torch_export_fused.onnx"))
    assert len(fused) == 5 and all("gamma" not in l for l in fused)
    for i, l in enumerate(fused):   # the exporter folded BatchNorm into the convolution: W' = W g / sqrt(var + eps), ...
        sc = g["gamma%d" % i].astype(np.float64) / np.sqrt(g["var%d" % i].astype(np.float64) + 1e-5)
        assert np.allclose(l["w"], g["w%d" % i] * sc[:, None, None], rtol=1e-5, atol=1e-6)
        assert np.allclose(l["b"], (g["b%d" % i] - g["mean%d" % i]) * sc + g["beta%d" % i], rtol=1e-5, atol=1e-6)
    # and the oracle run on the layers we read reproduces PyTorch's own output of that graph
    from oracle import postnet_oracle as po

    for layers in (bn, fused):
        y = po.postnet(g["x"], layers, dtype=np.float64)
        assert np.abs(y - g["y"]).max() < 1e-5


def test_decoder_onnx_reader_on_files_written_by_pytorch(lib, golden_dir, tmp_path):
    """decoder_iter.onnx (src/tacotron2/mod.rs:251-254): the reader finds every weight by its role in the dataflow, in both
    encodings of the LSTM cells -- ONNX LSTM operators (NVIDIA's export script; gate order i, o, f, c) and the Gemm / Split
    decomposition of today's PyTorch exporter -- on files serialised by PyTorch itself (tests/make_foreign_decoder_onnx.py)."""
    from xdtts_b200 import tacotron2
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_UNSUPPORTED, XdttsError

    want = np.load(os.path.join(golden_dir, "decoder_iter_small.npz"))
    for tag, form in (("cells", 0), ("lstm", 1)):
        dims, got = tacotron2.read_onnx_decoder(os.path.join(golden_dir, "decoder_iter_small_%s.onnx" % tag))
        assert dims == dict(n_mel=8, prenet=16, enc=24, att_rnn=32, dec_rnn=40, att_dim=12, loc_f=6, loc_k=7, lstm_form=form, has_dropout=1)
        for name in tacotron2.DECODER_TENSORS:
            assert np.array_equal(got[name], want[name].ravel()), (tag, name)
    # a postnet is not a decoder; the LFS pointer the reference ships; a decoder is not a postnet
    with pytest.raises(XdttsError) as e:
        tacotron2.read_onnx_decoder(os.path.join(golden_dir, "postnet_torch_export_bn.onnx"))
    assert e.value.code == ERR_UNSUPPORTED and "decoder_input" in e.value.message
    ptr = tmp_path / "decoder_iter.onnx"
    ptr.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:0000\nsize 72800000\n")
    with pytest.raises(XdttsError) as e:
        tacotron2.read_onnx_decoder(ptr)
    assert e.value.code == ERR_BAD_ARG and "LFS" in e.value.message
    with pytest.raises(XdttsError) as e:
        tacotron2.read_onnx_postnet(os.path.join(golden_dir, "decoder_iter_small_lstm.onnx"))
    assert e.value.code == ERR_UNSUPPORTED
    # truncated / mutated files never crash the reader
    raw = open(os.path.join(golden_dir, "decoder_iter_small_lstm.onnx"), "rb").read()
    rng = np.random.default_rng(1)
    for i in range(60):
        b = bytearray(raw[: int(rng.integers(50, len(raw)))] if i % 2 else raw)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        path = tmp_path / "mut.onnx"
        path.write_bytes(bytes(b))
        m = ctypes.c_void_p()
        rc = lib.xdtts_onnx_decoder_open(str(path).encode(), ctypes.byref(m))
        if rc == 0:
            lib.xdtts_onnx_decoder_close(m)
        else:
            assert rc in (ERR_BAD_ARG, ERR_UNSUPPORTED, -2)


def test_npy_io_matches_numpy(lib, tmp_path):
    """The reference dumps mels with ndarray_npy::write_npy (src/lib.rs:132): our writer/reader interoperate with numpy."""
    from xdtts_b200 import tacotron2
    from xdtts_b200._ffi import ERR_BAD_ARG, ERR_UNSUPPORTED, XdttsError

    rng = np.random.default_rng(1)
    for shape in ((80, 200), (80, 1), (1, 7), (513, 33)):
        a = rng.standard_normal(shape).astype(np.float32)
        p = tmp_path / ("a_%d_%d.npy" % shape)
        tacotron2.write_npy(p, a)
        b = np.load(p)                                   # numpy reads what we wrote
        assert b.dtype == np.float32 and b.shape == shape and np.array_equal(a, b)
        assert (p.stat().st_size - a.nbytes) % 64 == 0   # header padded as numpy / ndarray-npy do
        q = tmp_path / "np.npy"
        np.save(q, a)                                    # we read what numpy wrote
        assert np.array_equal(tacotron2.read_npy(q), a)
        np.save(q, np.asfortranarray(a))
        assert np.array_equal(tacotron2.read_npy(q), a)
    np.save(tmp_path / "v.npy", np.arange(5, dtype=np.float32))
    assert tacotron2.read_npy(tmp_path / "v.npy").shape == (1, 5)
    np.save(tmp_path / "d.npy", np.zeros((2, 2)))
    with pytest.raises(XdttsError) as e:
        tacotron2.read_npy(tmp_path / "d.npy")           # float64 is not what the reference writes
    assert e.value.code == ERR_UNSUPPORTED
    (tmp_path / "junk.npy").write_bytes(b"not numpy at all")
    with pytest.raises(XdttsError) as e:
        tacotron2.read_npy(tmp_path / "junk.npy")
    assert e.value.code == ERR_BAD_ARG


def test_file_readers_survive_malformed_input(lib, tmp_path):
    """postnet.onnx and .npy come from disk: mutated / truncated files must produce an error code (or parse), never a
    crash or a read out of bounds.  400 seeded mutants of each format."""
    import ctypes
    import random

    from onnx_writer import postnet_model
    from oracle import postnet_oracle as po
    from xdtts_b200._ffi import fptr

    rng = random.Random(7)
    good = bytes(postnet_model(po.synth_weights(seed=7, channels=(8, 16, 8))))
    path = str(tmp_path / "m.onnx").encode()

    def mutate(data):
        d = bytearray(data)
        mode = rng.randrange(4)
        if mode == 0:
            for _ in range(rng.randrange(1, 8)):
                d[rng.randrange(len(d))] = rng.randrange(256)
        elif mode == 1:
            d = d[:rng.randrange(len(d))]
        elif mode == 2:
            i = rng.randrange(len(d))
            d[i:i] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 16)))
        else:
            i = rng.randrange(min(len(d), 400))
            d[i:i + 1] = b"\xff\xff\xff\xff\xff\x7f"      # a huge varint where a length or key was
        return bytes(d)

    parsed = 0
    for _ in range(400):
        open(path, "wb").write(mutate(good))
        m = ctypes.c_void_p()
        rc = lib.xdtts_onnx_postnet_open(path, ctypes.byref(m))
        assert rc <= 0
        if rc == 0:
            parsed += 1
            n = lib.xdtts_onnx_postnet_n_layers(m)
            for i in range(max(n, 0)):
                co, ci, k, hb, hbn = (ctypes.c_int() for _ in range(5))
                eps = ctypes.c_float()
                lib.xdtts_onnx_postnet_layer_info(m, i, ctypes.byref(co), ctypes.byref(ci), ctypes.byref(k), ctypes.byref(hb),
                                                  ctypes.byref(hbn), ctypes.byref(eps))
                if 0 < co.value * ci.value * k.value < 10 ** 6:
                    buf = np.empty(co.value * ci.value * k.value, np.float32)
                    lib.xdtts_onnx_postnet_layer_copy(m, i, 0, fptr(buf))
            lib.xdtts_onnx_postnet_close(m)
    assert 0 < parsed < 400          # some mutants are harmless (payload bytes), most are rejected
    npy = str(tmp_path / "a.npy")
    np.save(npy, np.arange(12, dtype=np.float32).reshape(3, 4))
    good = open(npy, "rb").read()
    for _ in range(400):
        open(npy, "wb").write(mutate(good))
        r, c = ctypes.c_int(), ctypes.c_int()
        rc = lib.xdtts_npy_read_f32(npy.encode(), None, 0, ctypes.byref(r), ctypes.byref(c))
        assert rc <= 0
        if rc == 0 and 0 < r.value * c.value < 10 ** 6:
            buf = np.empty(r.value * c.value, np.float32)
            lib.xdtts_npy_read_f32(npy.encode(), fptr(buf), buf.size, ctypes.byref(r), ctypes.byref(c))
