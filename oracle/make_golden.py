"""Generates tests/golden/*.npz (TEST INFRASTRUCTURE; run in the build container).

The reference ships no golden vectors for this path (PARITY UNPINNED, see
oracle/__init__.py), so the fixtures are authored here from

  * the numpy oracle (oracle/gl_oracle.py, oracle/postnet_oracle.py), and
  * INDEPENDENT implementations present in this image -- torch.stft/torch.istft
    (fp64 Griffin-Lim loop), torchaudio.functional.melscale_fbanks,
    torch.nn.functional.conv1d/batch_norm --

so that tests/test_oracle.py can pin the oracle against something it did not
produce itself, on a box where torch/torchaudio may differ or be missing.

    python oracle/make_golden.py        # writes tests/golden/
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gl_oracle as o  # noqa: E402
from oracle import postnet_oracle as p  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def torch_gl64(s_mag, turns, n_iter, momentum, n_fft, hop):
    import torch

    s = torch.from_numpy(s_mag.astype(np.float64))
    ang = torch.from_numpy(o.turns_to_angles(turns, np.float64))
    w = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
    reb = torch.zeros_like(ang)
    a = momentum / (1 + momentum)
    for _ in range(n_iter):
        tp = reb
        inv = torch.istft(s * ang, n_fft, hop, window=w, center=True)
        reb = torch.stft(inv, n_fft, hop, window=w, center=True, pad_mode="reflect", return_complex=True)
        ang = reb - a * tp
        ang = ang / (ang.abs() + o.TINY32)
    return torch.istft(s * ang, n_fft, hop, window=w, center=True).numpy()


def torch_decoder64(wt, memory, processed, unpadded_len, keeps, n_steps):
    """The decoder step composed from torch modules (nn.LSTMCell, F.conv1d, F.linear, masked softmax) in
    fp64: an implementation independent of oracle/decoder_oracle.py, teacher-free (feeds its own output)."""
    import torch
    import torch.nn.functional as F

    W = {k: torch.from_numpy(np.asarray(v, np.float64)) for k, v in wt.items()}
    mem = torch.from_numpy(memory.astype(np.float64))[None]
    pm = torch.from_numpy(processed.astype(np.float64))[None]
    t_enc = memory.shape[0]
    att = torch.nn.LSTMCell(768, 1024).double()
    dec = torch.nn.LSTMCell(1536, 1024).double()
    with torch.no_grad():
        for cell, pre in ((att, "att"), (dec, "dec")):
            cell.weight_ih.copy_(W[pre + "_w_ih"]); cell.weight_hh.copy_(W[pre + "_w_hh"])
            cell.bias_ih.copy_(W[pre + "_b_ih"]); cell.bias_hh.copy_(W[pre + "_b_hh"])
        x = torch.zeros(1, 80, dtype=torch.float64)
        ha, ca, hd, cd = (torch.zeros(1, 1024, dtype=torch.float64) for _ in range(4))
        w = torch.zeros(1, t_enc, dtype=torch.float64)
        wc = torch.zeros(1, t_enc, dtype=torch.float64)
        ctx = torch.zeros(1, 512, dtype=torch.float64)
        mask = torch.arange(t_enc)[None] >= unpadded_len
        mels, gates, aligns = [], [], []
        for i in range(n_steps):
            h = x
            for layer, name in enumerate(("prenet1", "prenet2")):
                h = F.relu(F.linear(h, W[name])) * torch.from_numpy(keeps[i][layer].astype(np.float64))[None] * 2.0
            ha, ca = att(torch.cat([h, ctx], 1), (ha, ca))
            cat = torch.stack([w, wc], 1)                                       # [1, 2, t_enc]
            loc = F.conv1d(cat, W["loc_conv"], padding=15).transpose(1, 2)       # [1, t_enc, 32]
            pa = F.linear(loc, W["loc_dense"])
            e = F.linear(torch.tanh(F.linear(ha, W["query"])[:, None] + pa + pm), W["v"][None]).squeeze(-1)
            e = e.masked_fill(mask, -float("inf"))
            w = F.softmax(e, dim=1)
            ctx = torch.bmm(w[:, None], mem).squeeze(1)
            wc = wc + w
            hd, cd = dec(torch.cat([ha, ctx], 1), (hd, cd))
            hc = torch.cat([hd, ctx], 1)
            x = F.linear(hc, W["proj_w"], W["proj_b"])
            g = F.linear(hc, W["gate_w"], W["gate_b"])
            mels.append(x[0].numpy().copy()); gates.append(float(g[0, 0])); aligns.append(w[0].numpy().copy())
    return np.stack(mels), np.array(gates), np.stack(aligns)


def decoder_golden():
    """tests/golden/decoder.npz: 16 free-running decoder steps, t_enc 24 (19 unpadded), seeded dropout."""
    from oracle import decoder_oracle as d

    wt = d.synth_weights(11)
    memory, processed = d.synth_encoder_outputs(21, 24)
    n = 16
    keeps = [d.dropout_keep(5, 0, i) for i in range(n)]
    mel_t, gate_t, align_t = torch_decoder64(wt, memory, processed, 19, keeps, n)
    mel_o, gate_o, align_o = d.run_decoder(wt, memory, processed, 19, seed=5, utt=0, gate_threshold=2.0, max_steps=n,
                                           return_aux=True)
    np.savez_compressed(os.path.join(OUT, "decoder.npz"), memory=memory, processed=processed, unpadded_len=19, seed=5,
                        weights_seed=11, mel_torch64=mel_t, gate_torch64=gate_t, align_torch64=align_t.astype(np.float32),
                        mel_oracle64=mel_o, gate_oracle64=gate_o)
    print("decoder.npz: oracle vs torch max abs diff mel %.2e gate %.2e align %.2e" % (
        np.abs(mel_t - mel_o).max(), np.abs(gate_t - gate_o).max(), np.abs(align_t - align_o).max()))


def main():
    import torch
    import torch.nn.functional as F
    import torchaudio

    if "--decoder-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        return decoder_golden()

    os.makedirs(OUT, exist_ok=True)

    # --- filterbank: reference call site src/tacotron2/mod.rs:453 (+ the cfg5 variant)
    fb = {}
    for n_fft in (1024, 2048):
        k = n_fft // 2 + 1
        fb[f"torchaudio_{n_fft}"] = (
            torchaudio.functional.melscale_fbanks(k, 0.0, 8000.0, 80, 22050, norm="slaney", mel_scale="slaney").T.numpy()
        )
        fb[f"oracle_{n_fft}"] = o.create_mel_filter_bank(22050.0, n_fft, 80, 0.0, 8000.0)
    np.savez_compressed(os.path.join(OUT, "melbank.npz"), **fb)

    # --- cfg1: single 80x200 mel, 30 iterations (BASELINE.json configs[0])
    n_fft, hop, t = 1024, 256, 200
    basis = fb["oracle_1024"]
    mel = o.synth_mel(1234, 80, t)
    turns = o.phase_turns(4321, 0, 513, t)
    s = o.lift_pinv_clamp(mel, basis, 1.7)
    s64 = o.lift_pinv_clamp(mel, basis, 1.7, dtype=np.float64)
    y32 = o.griffin_lim(s, turns, 30, 0.99, n_fft, hop)
    y64 = o.griffin_lim(s, turns, 30, 0.99, n_fft, hop, dtype=np.float64)
    yt64 = torch_gl64(s, turns, 30, 0.99, n_fft, hop)
    np.savez_compressed(
        os.path.join(OUT, "cfg1_gl.npz"),
        mel=mel, turns=turns, s_mag=s, s_mag64=s64.astype(np.float32),
        y_fp32=y32, y_fp64=y64.astype(np.float32), y_torch64=yt64.astype(np.float32),
    )

    # --- short speech-like case with per-iteration checkpoints (teacher forcing)
    t = 48
    sm = o.synth_speech_like_mag(5, n_fft, hop, t)
    tu = o.phase_turns(99, 3, 513, t)
    ck = {k: None for k in (0, 1, 2, 5, 10)}
    o.griffin_lim(sm, tu, 10, 0.99, n_fft, hop, dtype=np.float64, checkpoints=ck)
    d = dict(s_mag=sm, turns=tu)
    for k, (y, r) in ck.items():
        d[f"y{k}"] = y.astype(np.float32)
        d[f"r{k}"] = r.astype(np.complex64)
    d["y10_torch64"] = torch_gl64(sm, tu, 10, 0.99, n_fft, hop).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "speech48_ckpt.npz"), **d)

    # --- n_fft 2048 / hop 512 short case (cfg5 geometry)
    n_fft2, hop2, t2 = 2048, 512, 24
    sm2 = o.synth_speech_like_mag(6, n_fft2, hop2, t2)
    tu2 = o.phase_turns(7, 0, 1025, t2)
    np.savez_compressed(
        os.path.join(OUT, "n2048_gl.npz"),
        s_mag=sm2, turns=tu2,
        y_fp64=o.griffin_lim(sm2, tu2, 8, 0.99, n_fft2, hop2, dtype=np.float64).astype(np.float32),
        y_torch64=torch_gl64(sm2, tu2, 8, 0.99, n_fft2, hop2).astype(np.float32),
    )

    # --- postnet: torch conv1d + batch_norm (independent) on seeded weights
    layers = p.synth_weights(7)
    melp = o.synth_mel(11, 80, 96)
    x = torch.from_numpy(melp)[None]
    for i, l in enumerate(layers):
        x = F.conv1d(x, torch.from_numpy(l["w"]), torch.from_numpy(l["b"]), padding=2)
        x = F.batch_norm(
            x, torch.from_numpy(l["mean"]), torch.from_numpy(l["var"]), torch.from_numpy(l["gamma"]),
            torch.from_numpy(l["beta"]), False, 0.0, p.BN_EPS,
        )
        if i < len(layers) - 1:
            x = torch.tanh(x)
    np.savez_compressed(
        os.path.join(OUT, "postnet.npz"),
        mel=melp, out_torch32=(torch.from_numpy(melp) + x[0]).numpy(),
        out_oracle64=p.postnet(melp, layers, dtype=np.float64).astype(np.float32),
    )
    decoder_golden()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
