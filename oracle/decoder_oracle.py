"""numpy restatement of the Tacotron2 decoder loop (TEST INFRASTRUCTURE).

PARITY UNPINNED -- see ``oracle/__init__.py``.  The reference drives
``models/tacotron2/decoder_iter.onnx`` through ONNX Runtime once per output frame
(``Tacotron2::run_decoder``, ``src/tacotron2/mod.rs:272-342``; state tensors and their sizes
``DecoderState::new`` ``:204-238``; stop rule ``:279-280,319-324``: sigmoid(gate) > 0.6 or 1000
steps, the frame that fires the gate is kept).  The ONNX file is a git-LFS pointer and ORT is absent, so
this restates the published graph: NVIDIA Tacotron2 ``Decoder.decode`` as wrapped by ``DecoderIter``
in ``export_tacotron2_onnx.py`` (named at ``src/tacotron2/mod.rs:137-138``):

    x      = prenet(decoder_input)            2 x [Linear(no bias) -> relu -> dropout(0.5, ALWAYS on)]
    h_a,c_a= LSTMCell_att([x, ctx], (h_a, c_a))                     768 -> 1024
    e[t]   = v . tanh(Wq h_a + Wld conv1d([w; w_cum], k=31)[:, t] + processed_memory[t])
    w      = softmax(e masked with -inf past the unpadded length);  ctx = w @ memory;  w_cum += w
    h_d,c_d= LSTMCell_dec([h_a, ctx], (h_d, c_d))                   1536 -> 1024
    mel    = Wp [h_d, ctx] + bp  (80);   gate = Wg [h_d, ctx] + bg  (1)

The exported prenet draws its dropout mask inside the graph (``torch.le(torch.rand(256), 0.5)``, scaled by
2), so the reference's decoder is stochastic by construction; here the mask comes from the counter-based
generator the device uses (splitmix64 of seed, utterance, step, layer, unit) or is switched off.
LSTM gate order is PyTorch's (i, f, g, o).  Weights are synthetic and seeded.
"""
from __future__ import annotations

import numpy as np

from . import gl_oracle as _gl

N_MEL, PRENET, ENC, ATT_RNN, DEC_RNN, ATT_DIM, LOC_F, LOC_K = 80, 256, 512, 1024, 1024, 128, 32, 31
GATE_THRESHOLD, MAX_STEPS = 0.6, 1000   # src/tacotron2/mod.rs:279-280


def synth_weights(seed=11, gate_bias=-2.0, gate_gain=6.0, gain=2.0):
    """Seeded random decoder in PyTorch layouts (float32).  `gain` (on the N(0, 1/fan_in) matrices) makes
    the recurrent state lively but bounded; gate_bias / gate_gain shape when the stop gate fires."""
    rng = np.random.default_rng(seed)

    def mat(rows, cols, g=gain):
        return (g * rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32)

    def vec(n, s=0.1):
        return (s * rng.standard_normal(n)).astype(np.float32)

    return dict(
        prenet1=mat(PRENET, N_MEL), prenet2=mat(PRENET, PRENET),
        att_w_ih=mat(4 * ATT_RNN, PRENET + ENC), att_w_hh=mat(4 * ATT_RNN, ATT_RNN),
        att_b_ih=vec(4 * ATT_RNN), att_b_hh=vec(4 * ATT_RNN),
        query=mat(ATT_DIM, ATT_RNN), v=mat(1, ATT_DIM, 2.0)[0],
        loc_conv=(rng.standard_normal((LOC_F, 2, LOC_K)) / np.sqrt(2 * LOC_K)).astype(np.float32),
        loc_dense=mat(ATT_DIM, LOC_F, 1.0),
        dec_w_ih=mat(4 * DEC_RNN, ATT_RNN + ENC), dec_w_hh=mat(4 * DEC_RNN, DEC_RNN),
        dec_b_ih=vec(4 * DEC_RNN), dec_b_hh=vec(4 * DEC_RNN),
        proj_w=mat(N_MEL, DEC_RNN + ENC), proj_b=vec(N_MEL),
        gate_w=mat(1, DEC_RNN + ENC, gate_gain), gate_b=np.array([gate_bias], np.float32),
    )


def synth_encoder_outputs(seed, t_enc):
    """memory [t_enc, 512] and processed_memory [t_enc, 128] as the encoder graph would emit them."""
    rng = np.random.default_rng(seed)
    memory = (0.5 * rng.standard_normal((t_enc, ENC))).astype(np.float32)
    processed = (0.5 * rng.standard_normal((t_enc, ATT_DIM))).astype(np.float32)
    return memory, processed


def dropout_keep(seed, utt, step):
    """[2, 256] keep mask (1.0 / 0.0) of the two prenet layers at `step`: u <= 0.5 with u from the same
    splitmix64 stream as the vocoder's phase generator, index = step * 512 + layer * 256 + unit."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + _gl._GOLD * np.uint64(utt + 1)
        idx = np.uint64(step) * np.uint64(2 * PRENET) + np.arange(2 * PRENET, dtype=np.uint64)
        z = base + _gl._GOLD * (idx + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * _gl._M1
        z = (z ^ (z >> np.uint64(27))) * _gl._M2
        z = z ^ (z >> np.uint64(31))
    u = ((z >> np.uint64(40)).astype(np.float64) * (2.0**-24)).astype(np.float32)
    return (u <= 0.5).astype(np.float32).reshape(2, PRENET)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _lstm_cell(w_ih, w_hh, b_ih, b_hh, x, h, c):
    g = w_ih @ x + b_ih + w_hh @ h + b_hh
    n = h.shape[0]
    i, f, gg, o = g[:n], g[n:2 * n], g[2 * n:3 * n], g[3 * n:]
    c = _sigmoid(f) * c + _sigmoid(i) * np.tanh(gg)
    return _sigmoid(o) * np.tanh(c), c


def new_state(t_enc, dtype=np.float64):
    """DecoderState::new (src/tacotron2/mod.rs:204-238): everything zero."""
    z = lambda n: np.zeros(n, dtype)  # noqa: E731
    return dict(x=z(N_MEL), h_a=z(ATT_RNN), c_a=z(ATT_RNN), h_d=z(DEC_RNN), c_d=z(DEC_RNN), w=z(t_enc), w_cum=z(t_enc),
                ctx=z(ENC))


def step(wt, st, memory, processed, unpadded_len, keep=None, dtype=np.float64):
    """One decoder_iter.onnx evaluation; updates `st` in place, returns (mel [80], gate logit)."""
    W = {k: np.asarray(v, dtype) for k, v in wt.items()}
    x = st["x"]
    for layer, name in enumerate(("prenet1", "prenet2")):
        x = np.maximum(W[name] @ x, 0)
        if keep is not None:
            x = x * keep[layer].astype(dtype) * dtype(2.0)
    st["h_a"], st["c_a"] = _lstm_cell(W["att_w_ih"], W["att_w_hh"], W["att_b_ih"], W["att_b_hh"],
                                      np.concatenate([x, st["ctx"]]), st["h_a"], st["c_a"])
    t_enc = memory.shape[0]
    cat = np.stack([st["w"], st["w_cum"]])                       # [2, t_enc]
    pad = np.pad(cat, ((0, 0), (LOC_K // 2, LOC_K // 2)))
    loc = np.empty((LOC_F, t_enc), dtype)
    for t in range(t_enc):                                       # cross-correlation, as torch conv1d
        loc[:, t] = np.einsum("fck,ck->f", W["loc_conv"], pad[:, t:t + LOC_K])
    pa = (W["loc_dense"] @ loc).T                                # [t_enc, 128]
    pq = W["query"] @ st["h_a"]
    e = np.tanh(pq[None, :] + pa + np.asarray(processed, dtype)) @ W["v"]
    e[unpadded_len:] = -np.inf
    p = np.exp(e - e[:unpadded_len].max())
    w = p / p.sum()
    st["w"] = w.astype(dtype)
    st["ctx"] = w @ np.asarray(memory, dtype)
    st["w_cum"] = st["w_cum"] + w
    st["h_d"], st["c_d"] = _lstm_cell(W["dec_w_ih"], W["dec_w_hh"], W["dec_b_ih"], W["dec_b_hh"],
                                      np.concatenate([st["h_a"], st["ctx"]]), st["h_d"], st["c_d"])
    hc = np.concatenate([st["h_d"], st["ctx"]])
    mel = W["proj_w"] @ hc + W["proj_b"]
    gate = float((W["gate_w"] @ hc + W["gate_b"])[0])
    st["x"] = mel
    return mel, gate


def run_decoder(wt, memory, processed, unpadded_len, seed=0, utt=0, dropout=True, gate_threshold=GATE_THRESHOLD,
                max_steps=MAX_STEPS, dtype=np.float64, return_aux=False):
    """Tacotron2::run_decoder (src/tacotron2/mod.rs:272-342) -> mel [T, 80] (before the transpose at :345).
    The frame whose gate fires is kept; at most max_steps frames."""
    st = new_state(memory.shape[0], dtype)
    wt = {k: np.asarray(v, dtype) for k, v in wt.items()}
    mels, gates, aligns = [], [], []
    for i in range(max_steps):
        keep = dropout_keep(seed, utt, i) if dropout else None
        mel, gate = step(wt, st, memory, processed, unpadded_len, keep, dtype)
        mels.append(mel.copy())
        gates.append(gate)
        aligns.append(st["w"].copy())
        if _sigmoid(gate) > gate_threshold or i + 1 == max_steps:
            break
    out = np.stack(mels)
    if return_aux:
        return out, np.array(gates), np.stack(aligns)
    return out
