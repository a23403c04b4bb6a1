"""decoder_iter.onnx written by PyTorch's ONNX serializer -- a FOREIGN encoder for the library's decoder reader
(xdtts_decoder_create_from_onnx / xdtts_onnx_decoder_*), like tests/make_foreign_onnx.py is for the postnet.

The module restates NVIDIA Tacotron2's `Decoder.decode` as `export_tacotron2_onnx.py` wraps it (`DecoderIter`; the
exporter the reference names at /root/reference src/tacotron2/mod.rs:137-138): prenet with the dropout mask drawn inside
the graph, attention LSTM cell, location-sensitive attention, decoder LSTM cell, projection + gate, and the eleven named
inputs / nine named outputs `Tacotron2::run_decoder` feeds and reads (src/tacotron2/mod.rs:285-341).

    python tests/make_foreign_decoder_onnx.py          # writes the SMALL fixture tests/golden/decoder_iter_small.onnx (+ .npz)
    build(dims, path)                                   # any size; the GPU test builds the full-size file in a temp dir
"""
import os
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
FULL = dict(n_mel=80, prenet=256, enc=512, att_rnn=1024, dec_rnn=1024, att_dim=128, loc_f=32, loc_k=31)
SMALL = dict(n_mel=8, prenet=16, enc=24, att_rnn=32, dec_rnn=40, att_dim=12, loc_f=6, loc_k=7)


class LinearNorm(nn.Module):
    def __init__(self, i, o, bias=True):
        super().__init__()
        self.linear_layer = nn.Linear(i, o, bias=bias)

    def forward(self, x):
        return self.linear_layer(x)


class ConvNorm(nn.Module):
    def __init__(self, i, o, k):
        super().__init__()
        self.conv = nn.Conv1d(i, o, k, padding=(k - 1) // 2, bias=False)

    def forward(self, x):
        return self.conv(x)


class LocationLayer(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.location_conv = ConvNorm(2, d["loc_f"], d["loc_k"])
        self.location_dense = LinearNorm(d["loc_f"], d["att_dim"], bias=False)

    def forward(self, cat):
        return self.location_dense(self.location_conv(cat).transpose(1, 2))


class Attention(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.query_layer = LinearNorm(d["att_rnn"], d["att_dim"], bias=False)
        self.v = LinearNorm(d["att_dim"], 1, bias=False)
        self.location_layer = LocationLayer(d)

    def forward(self, hidden, memory, processed_memory, cat, mask):
        pq = self.query_layer(hidden.unsqueeze(1))
        energies = self.v(torch.tanh(pq + self.location_layer(cat) + processed_memory)).squeeze(2)
        energies = energies.masked_fill(mask, -float("inf"))
        w = F.softmax(energies, dim=1)
        ctx = torch.bmm(w.unsqueeze(1), memory).squeeze(1)
        return ctx, w


class Prenet(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.layers = nn.ModuleList([LinearNorm(d["n_mel"], d["prenet"], bias=False), LinearNorm(d["prenet"], d["prenet"], bias=False)])

    def forward(self, x):
        for lin in self.layers:   # dropout stays on at inference; the export draws the mask inside the graph
            x = F.relu(lin(x))
            keep = torch.le(torch.rand_like(x[0]), 0.5).to(x.dtype)
            x = x * keep * 2.0
        return x


class Decoder(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.prenet = Prenet(d)
        self.attention_rnn = nn.LSTMCell(d["prenet"] + d["enc"], d["att_rnn"])
        self.attention_layer = Attention(d)
        self.decoder_rnn = nn.LSTMCell(d["att_rnn"] + d["enc"], d["dec_rnn"])
        self.linear_projection = LinearNorm(d["dec_rnn"] + d["enc"], d["n_mel"])
        self.gate_layer = LinearNorm(d["dec_rnn"] + d["enc"], 1, bias=True)


def lstmcell2lstm(cell):
    """what export_tacotron2_onnx.py does before exporting: the cell as a one-step nn.LSTM, so that the graph holds an
    ONNX `LSTM` operator (gate order i, o, f, c) instead of the cell's decomposition"""
    lstm = nn.LSTM(cell.input_size, cell.hidden_size, 1)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(cell.weight_ih)
        lstm.weight_hh_l0.copy_(cell.weight_hh)
        lstm.bias_ih_l0.copy_(cell.bias_ih)
        lstm.bias_hh_l0.copy_(cell.bias_hh)
    return lstm


class DecoderIter(nn.Module):
    def __init__(self, d, use_lstm_op=False):
        super().__init__()
        self.decoder = Decoder(d)
        self.use_lstm_op = use_lstm_op

    def convert_cells(self):
        self.att_lstm = lstmcell2lstm(self.decoder.attention_rnn)
        self.dec_lstm = lstmcell2lstm(self.decoder.decoder_rnn)

    def cell(self, which, x, h, c):
        if not self.use_lstm_op:
            return (self.decoder.attention_rnn if which == 0 else self.decoder.decoder_rnn)(x, (h, c))
        lstm = self.att_lstm if which == 0 else self.dec_lstm
        _, (h1, c1) = lstm(x.unsqueeze(0), (h.unsqueeze(0), c.unsqueeze(0)))
        return h1.squeeze(0), c1.squeeze(0)

    def forward(self, decoder_input, attention_hidden, attention_cell, decoder_hidden, decoder_cell, attention_weights,
                attention_weights_cum, attention_context, memory, processed_memory, mask):
        dec = self.decoder
        x = dec.prenet(decoder_input)
        attention_hidden, attention_cell = self.cell(0, torch.cat((x, attention_context), -1), attention_hidden, attention_cell)
        cat = torch.cat((attention_weights.unsqueeze(1), attention_weights_cum.unsqueeze(1)), dim=1)
        attention_context, attention_weights = dec.attention_layer(attention_hidden, memory, processed_memory, cat, mask)
        attention_weights_cum = attention_weights_cum + attention_weights
        decoder_hidden, decoder_cell = self.cell(1, torch.cat((attention_hidden, attention_context), -1), decoder_hidden, decoder_cell)
        hc = torch.cat((decoder_hidden, attention_context), dim=1)
        return (dec.linear_projection(hc), dec.gate_layer(hc), attention_hidden, attention_cell, decoder_hidden, decoder_cell,
                attention_weights, attention_weights_cum, attention_context)


IN_NAMES = ["decoder_input", "attention_hidden", "attention_cell", "decoder_hidden", "decoder_cell", "attention_weights",
            "attention_weights_cum", "attention_context", "memory", "processed_memory", "mask"]
OUT_NAMES = ["decoder_output", "gate_prediction", "out_attention_hidden", "out_attention_cell", "out_decoder_hidden", "out_decoder_cell",
             "out_attention_weights", "out_attention_weights_cum", "out_attention_context"]


def weights_of(m):
    """the arrays in the layout include/xdtts_b200.h `xdtts_decoder_weights` takes (PyTorch layouts, gate order i, f, g, o)"""
    d = m.decoder
    g = lambda t: t.detach().numpy().copy()  # noqa: E731
    return dict(prenet1=g(d.prenet.layers[0].linear_layer.weight), prenet2=g(d.prenet.layers[1].linear_layer.weight),
                att_w_ih=g(d.attention_rnn.weight_ih), att_w_hh=g(d.attention_rnn.weight_hh), att_b_ih=g(d.attention_rnn.bias_ih),
                att_b_hh=g(d.attention_rnn.bias_hh), query=g(d.attention_layer.query_layer.linear_layer.weight),
                v=g(d.attention_layer.v.linear_layer.weight)[0], loc_conv=g(d.attention_layer.location_layer.location_conv.conv.weight),
                loc_dense=g(d.attention_layer.location_layer.location_dense.linear_layer.weight),
                dec_w_ih=g(d.decoder_rnn.weight_ih), dec_w_hh=g(d.decoder_rnn.weight_hh), dec_b_ih=g(d.decoder_rnn.bias_ih),
                dec_b_hh=g(d.decoder_rnn.bias_hh), proj_w=g(d.linear_projection.linear_layer.weight),
                proj_b=g(d.linear_projection.linear_layer.bias), gate_w=g(d.gate_layer.linear_layer.weight)[0],
                gate_b=g(d.gate_layer.linear_layer.bias))


def build(dims, path, seed=3, t_enc=9, weights=None, use_lstm_op=False):
    """export a DecoderIter of these dimensions to `path`; returns the weight dict (xdtts_decoder_weights layouts)"""
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto   # needs the absent `onnx` package; no-op here
    torch.manual_seed(seed)
    m = DecoderIter(dims, use_lstm_op).eval()
    if weights is not None:   # load given arrays (the GPU test exports the bench's synthetic decoder)
        d = m.decoder
        with torch.no_grad():
            for t, k in ((d.prenet.layers[0].linear_layer.weight, "prenet1"), (d.prenet.layers[1].linear_layer.weight, "prenet2"),
                         (d.attention_rnn.weight_ih, "att_w_ih"), (d.attention_rnn.weight_hh, "att_w_hh"), (d.attention_rnn.bias_ih, "att_b_ih"),
                         (d.attention_rnn.bias_hh, "att_b_hh"), (d.attention_layer.query_layer.linear_layer.weight, "query"),
                         (d.attention_layer.location_layer.location_conv.conv.weight, "loc_conv"),
                         (d.attention_layer.location_layer.location_dense.linear_layer.weight, "loc_dense"),
                         (d.decoder_rnn.weight_ih, "dec_w_ih"), (d.decoder_rnn.weight_hh, "dec_w_hh"), (d.decoder_rnn.bias_ih, "dec_b_ih"),
                         (d.decoder_rnn.bias_hh, "dec_b_hh"), (d.linear_projection.linear_layer.weight, "proj_w"),
                         (d.linear_projection.linear_layer.bias, "proj_b"), (d.gate_layer.linear_layer.bias, "gate_b")):
                t.copy_(torch.from_numpy(np.asarray(weights[k], np.float32)))
            d.attention_layer.v.linear_layer.weight.copy_(torch.from_numpy(np.asarray(weights["v"], np.float32)).reshape(1, -1))
            d.gate_layer.linear_layer.weight.copy_(torch.from_numpy(np.asarray(weights["gate_w"], np.float32)).reshape(1, -1))
    if use_lstm_op:
        m.convert_cells()
    z = lambda *s: torch.zeros(*s)  # noqa: E731
    args = (z(1, dims["n_mel"]), z(1, dims["att_rnn"]), z(1, dims["att_rnn"]), z(1, dims["dec_rnn"]), z(1, dims["dec_rnn"]), z(1, t_enc),
            z(1, t_enc), z(1, dims["enc"]), torch.randn(1, t_enc, dims["enc"]), torch.randn(1, t_enc, dims["att_dim"]),
            torch.zeros(1, t_enc, dtype=torch.bool))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(m, args, path, opset_version=12, input_names=IN_NAMES, output_names=OUT_NAMES, dynamo=False,
                          dynamic_axes={"memory": {1: "t"}, "processed_memory": {1: "t"}, "mask": {1: "t"}, "attention_weights": {1: "t"},
                                        "attention_weights_cum": {1: "t"}})
    return weights_of(m)


if __name__ == "__main__":
    for tag, op in (("cells", False), ("lstm", True)):   # LSTM cells decomposed (today's exporter) / as ONNX LSTM operators (NVIDIA's script)
        out = os.path.join(HERE, "golden", "decoder_iter_small_%s.onnx" % tag)
        w = build(SMALL, out, use_lstm_op=op)
        print(out, os.path.getsize(out), "bytes")
    np.savez(os.path.join(HERE, "golden", "decoder_iter_small.npz"), **w)
