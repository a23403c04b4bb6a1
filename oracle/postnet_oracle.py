"""numpy restatement of the Tacotron2 postnet (TEST INFRASTRUCTURE).

PARITY UNPINNED -- see ``oracle/__init__.py``.  The reference runs
``models/tacotron2/postnet.onnx`` through ONNX Runtime
(``src/tacotron2/mod.rs:344-357``; session load ``:256-259``); the weights are a
git-LFS pointer and ORT is absent, so this restates the published graph
(NVIDIA Tacotron2 ``Postnet`` as exported by ``export_tacotron2_onnx.py``, named
at ``src/tacotron2/mod.rs:137-138``):

    x0 = mel                     [80, T]
    x  = tanh(BN_i(Conv1d_i(x))) i = 0..3     Conv1d k=5, pad=2, bias
    x  = BN_4(Conv1d_4(x))                    channels 80->512->512->512->512->80
    out = mel + x                "mel_outputs_postnet"

BatchNorm in eval mode (running stats, eps 1e-5); dropout is the identity.
Weights are synthetic and seeded.
"""
from __future__ import annotations

import numpy as np

CHANNELS = (80, 512, 512, 512, 512, 80)
KSIZE = 5
BN_EPS = 1e-5


def synth_weights(seed=7, channels=CHANNELS):
    """conv W ~ N(0, 1/sqrt(Cin*5)), b ~ N(0, 0.1); BN gamma~U(.5,1.5) beta~N(0,.1)
    mean~N(0,.1) var~U(.5,1.5).  Returns a list of dicts of float32 arrays."""
    rng = np.random.default_rng(seed)
    layers = []
    for cin, cout in zip(channels[:-1], channels[1:]):
        layers.append(
            dict(
                w=(rng.standard_normal((cout, cin, KSIZE)) / np.sqrt(cin * KSIZE)).astype(np.float32),
                b=(0.1 * rng.standard_normal(cout)).astype(np.float32),
                gamma=rng.uniform(0.5, 1.5, cout).astype(np.float32),
                beta=(0.1 * rng.standard_normal(cout)).astype(np.float32),
                mean=(0.1 * rng.standard_normal(cout)).astype(np.float32),
                var=rng.uniform(0.5, 1.5, cout).astype(np.float32),
            )
        )
    return layers


def fold_bn(layer, eps=BN_EPS, dtype=np.float64):
    """W' = W*g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta."""
    scale = layer["gamma"].astype(dtype) / np.sqrt(layer["var"].astype(dtype) + eps)
    w = layer["w"].astype(dtype) * scale[:, None, None]
    b = (layer["b"].astype(dtype) - layer["mean"].astype(dtype)) * scale + layer["beta"].astype(dtype)
    return w, b


def conv1d_same(x, w, b):
    """x [Cin,T], w [Cout,Cin,5], b [Cout] -> [Cout,T]; zero padding 2."""
    cin, t = x.shape
    k = w.shape[2]
    p = k // 2
    xp = np.pad(x, ((0, 0), (p, p)))
    out = np.empty((w.shape[0], t), dtype=x.dtype)
    out[:] = b[:, None]
    for j in range(k):
        out += w[:, :, j] @ xp[:, j : j + t]
    return out


def postnet(mel, layers, eps=BN_EPS, dtype=np.float32):
    """mel [80,T] -> mel + Postnet(mel), unfused BN (the graph as exported)."""
    x = np.asarray(mel, dtype=dtype)
    n = len(layers)
    for i, l in enumerate(layers):
        bias = l["b"].astype(dtype) if l.get("b") is not None else np.zeros(l["w"].shape[0], dtype)
        y = conv1d_same(x, l["w"].astype(dtype), bias)
        if l.get("gamma") is not None:   # a layer without BatchNormalization: an export with the statistics folded in
            inv = 1.0 / np.sqrt(l["var"].astype(dtype) + dtype(eps))
            y = (y - l["mean"].astype(dtype)[:, None]) * (l["gamma"].astype(dtype) * inv)[:, None] + l[
                "beta"
            ].astype(dtype)[:, None]
        x = np.tanh(y) if i < n - 1 else y
    return (np.asarray(mel, dtype=dtype) + x).astype(dtype)
