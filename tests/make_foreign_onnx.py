"""Writes tests/golden/postnet_torch_export_{bn,fused}.onnx with PyTorch's own ONNX serializer -- a FOREIGN encoder for
xd-tts_b200/csrc/onnx_postnet.cu (tests/onnx_writer.py shares an author with the reader and therefore any misreading
of the wire format).  The TorchScript exporter serialises the graph in C++ (torch._C Graph._export_onnx); only its last,
optional step (inlining onnxscript functions) imports the `onnx` package, which this image lacks -- that step is
replaced by the identity.  Small channel counts keep the fixtures a few kB; the graph is NVIDIA Tacotron2's Postnet
(5 x ConvNorm + BatchNorm1d, tanh x 4, dropout off in eval) + the residual add of export_tacotron2_onnx.py.

    python tests/make_foreign_onnx.py        # run in the build container; commits the .onnx files and their tensors (.npz)
"""
import os
import warnings

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
CH = (16, 32, 32, 32, 32, 16)   # the device postnet takes channel counts that are multiples of 16


class ConvNorm(nn.Module):
    def __init__(self, cin, cout, k=5):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, padding=k // 2)

    def forward(self, x):
        return self.conv(x)


class Postnet(nn.Module):
    def __init__(self):
        super().__init__()
        self.convolutions = nn.ModuleList(
            [nn.Sequential(ConvNorm(a, b), nn.BatchNorm1d(b)) for a, b in zip(CH[:-1], CH[1:])])

    def forward(self, x):
        for i in range(len(self.convolutions) - 1):
            x = torch.nn.functional.dropout(torch.tanh(self.convolutions[i](x)), 0.5, self.training)
        return torch.nn.functional.dropout(self.convolutions[-1](x), 0.5, self.training)


class Wrap(nn.Module):
    def __init__(self):
        super().__init__()
        self.postnet = Postnet()

    def forward(self, mel):
        return mel + self.postnet(mel)


def main():
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto   # needs the absent `onnx` package; no-op here
    torch.manual_seed(7)
    m = Wrap().eval()
    with torch.no_grad():
        for seq in m.postnet.convolutions:   # non-trivial running statistics
            bn = seq[1]
            bn.running_mean.normal_(0, 0.1)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.1)
    x = torch.randn(1, CH[0], 37)
    with torch.no_grad():
        y = m(x)
    tensors = {"x": x[0].numpy(), "y": y[0].numpy()}
    for i, seq in enumerate(m.postnet.convolutions):
        tensors.update({"w%d" % i: seq[0].conv.weight.detach().numpy(), "b%d" % i: seq[0].conv.bias.detach().numpy(),
                        "gamma%d" % i: seq[1].weight.detach().numpy(), "beta%d" % i: seq[1].bias.detach().numpy(),
                        "mean%d" % i: seq[1].running_mean.numpy(), "var%d" % i: seq[1].running_var.numpy()})
    out_dir = os.path.join(HERE, "golden")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for tag, fold in (("fused", True), ("bn", False)):
            # do_constant_folding folds eval-mode BatchNorm into the Conv initializers (what exporters do by default)
            kw = dict(opset_version=12, input_names=["mel_outputs"], output_names=["mel_outputs_postnet"],
                      dynamic_axes={"mel_outputs": {2: "T"}, "mel_outputs_postnet": {2: "T"}}, dynamo=False,
                      do_constant_folding=fold)
            if not fold:
                kw["training"] = torch.onnx.TrainingMode.PRESERVE
            path = os.path.join(out_dir, "postnet_torch_export_%s.onnx" % tag)
            torch.onnx.export(m, (x,), path, **kw)
            print(path, os.path.getsize(path), "bytes")
    np.savez(os.path.join(out_dir, "postnet_torch_export.npz"), **tensors)


if __name__ == "__main__":
    main()
