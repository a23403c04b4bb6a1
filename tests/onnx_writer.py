"""Minimal ONNX (protobuf wire format) writer for the tests -- TEST INFRASTRUCTURE.

The real `models/tacotron2/postnet.onnx` of the reference is a git-LFS pointer in this image and the
`onnx` package is absent, so the tests encode a structurally equivalent file by hand: the graph
torch.onnx emits for NVIDIA Tacotron2's Postnet (Conv -> BatchNormalization -> Tanh, x5 without the last
Tanh, then Add with the input; output "mel_outputs_postnet", /root/reference src/tacotron2/mod.rs:349)."""
import struct

import numpy as np


def _varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field, wire):
    return _varint((field << 3) | wire)


def _ld(field, payload):   # length-delimited
    return _key(field, 2) + _varint(len(payload)) + payload


def _str(field, s):
    return _ld(field, s.encode())


def _int(field, v):
    return _key(field, 0) + _varint(v)


def tensor(name, arr, raw=True, packed_dims=False):
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    if packed_dims:
        body = _ld(1, b"".join(_varint(d) for d in arr.shape))
    else:
        body = b"".join(_int(1, d) for d in arr.shape)
    body += _int(2, 1)   # FLOAT
    if raw:
        body += _str(8, name) + _ld(9, arr.tobytes())
    else:
        body += _ld(4, arr.tobytes()) + _str(8, name)   # packed float_data
    return body


def attr_float(name, v):
    return _str(1, name) + _key(2, 5) + struct.pack("<f", v) + _int(20, 1)


def attr_int(name, v):
    return _str(1, name) + _int(3, v) + _int(20, 2)


def attr_ints(name, vs, packed=False):
    body = _str(1, name)
    body += _ld(8, b"".join(_varint(v) for v in vs)) if packed else b"".join(_int(8, v) for v in vs)
    return body + _int(20, 7)


def node(op, inputs, outputs, attrs=(), name=""):
    body = b"".join(_str(1, i) for i in inputs) + b"".join(_str(2, o) for o in outputs)
    if name:
        body += _str(3, name)
    body += _str(4, op) + b"".join(_ld(5, a) for a in attrs)
    return body


def postnet_model(layers, eps=1e-5, prefix="postnet.convolutions", raw=True, with_bn=True, activation="Tanh", residual=True,
                  extra_op=None, dropout=False):
    """bytes of a ModelProto for the postnet with the given layers (dicts w,b,gamma,beta,mean,var).
    activation / residual / extra_op build graphs that are NOT the Tacotron2 postnet (the reader must reject them);
    dropout inserts the inference-mode Dropout nodes some exporters keep."""
    nodes, inits = [], []
    x = "mel"
    n = len(layers)
    for i, l in enumerate(layers):
        k = l["w"].shape[2]
        wn, bn_ = "%s.%d.0.conv.weight" % (prefix, i), "%s.%d.0.conv.bias" % (prefix, i)
        inits.append(tensor(wn, l["w"], raw=raw, packed_dims=bool(i & 1)))
        conv_in = [x, wn]
        if "b" in l:
            inits.append(tensor(bn_, l["b"], raw=raw))
            conv_in.append(bn_)
        y = "conv_%d" % i
        nodes.append(node("Conv", conv_in, [y], [attr_ints("dilations", [1]), attr_int("group", 1), attr_ints("kernel_shape", [k]),
                                                 attr_ints("pads", [k // 2, k // 2], packed=bool(i & 1)), attr_ints("strides", [1])],
                          name="Conv_%d" % i))
        if with_bn and "gamma" in l:
            names = ["%s.%d.1.%s" % (prefix, i, s) for s in ("weight", "bias", "running_mean", "running_var")]
            for nm, key in zip(names, ("gamma", "beta", "mean", "var")):
                inits.append(tensor(nm, l[key], raw=raw))
            z = "bn_%d" % i
            nodes.append(node("BatchNormalization", [y] + names, [z], [attr_float("epsilon", eps), attr_float("momentum", 0.9)]))
            y = z
        if i < n - 1 and activation:
            z = "tanh_%d" % i
            nodes.append(node(activation, [y], [z]))
            y = z
        if dropout:
            z = "drop_%d" % i
            nodes.append(node("Dropout", [y], [z]))
            y = z
        if extra_op and i == 1:
            z = "extra_%d" % i
            nodes.append(node(extra_op, [y], [z]))
            y = z
        x = y
    if residual:
        nodes.append(node("Add", ["mel", x], ["mel_outputs_postnet"]))
    else:
        nodes.append(node("Identity", [x], ["mel_outputs_postnet"]))
    graph = b"".join(_ld(1, nd) for nd in nodes) + _str(2, "torch_jit") + b"".join(_ld(5, t) for t in inits)
    graph += _ld(11, _str(1, "mel")) + _ld(12, _str(1, "mel_outputs_postnet"))   # ValueInfoProto names only
    model = _int(1, 7) + _str(2, "pytorch") + _str(3, "1.13") + _ld(7, graph) + _ld(8, _str(1, "") + _int(2, 13))
    return model
