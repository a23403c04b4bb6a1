"""xdtts_b200 -- host-side mirror of the reference's interfaces for the vocoding + postnet
hot path of xd-tts, over the C ABI of libxdtts_b200.so (include/xdtts_b200.h).

Mirrors, name for name, what the reference's Rust call sites use
(/root/reference src/tacotron2/mod.rs:441-458, src/lib.rs:141):

    from xdtts_b200 import griffin_lim
    basis   = griffin_lim.mel.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)
    vocoder = griffin_lim.GriffinLim.new(basis, 1024 - 256, 1.7, 30, 0.99)
    audio   = vocoder.infer(mel)            # mel [80, T] float32 -> [256 * (T - 1)] float32

The arithmetic runs only in the CUDA library; there is no CPU path.  Importing this package
does not need a GPU, creating a GriffinLim does.
"""
from ._ffi import XdttsError, load_library, version  # noqa: F401
from . import griffin_lim, tacotron2  # noqa: F401
