"""CPU oracle for the xd-tts vocoding + postnet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU arm.  The product path (``xd-tts_b200/``) never imports this package
and fails loudly when its CUDA library is missing.

PARITY UNPINNED: the reference (``/root/reference``) ships no golden vector,
known-answer test or fixture for this path (only a shape assertion,
``src/tacotron2/mod.rs:510-522``) and its arithmetic lives in two absent
third-party dependencies:

* ``griffin-lim 0.2.0`` @ e6415314cf3309787e54d9ff2768454373e82a5c
  (``Cargo.toml:20``, ``Cargo.lock:666-680``) -- "a port from librosa"
  (``slides/vocoding.typ:50``; librosa pinned to 0.9.2 in
  ``scripts/requirements.txt:12``);
* ONNX Runtime 1.17.0 executing ``models/tacotron2/postnet.onnx`` (git-LFS
  pointer only, ``models/tacotron2/postnet.onnx:1-3``).

The oracle therefore restates the *published* algorithms (librosa 0.9.2
``filters.mel`` / ``stft`` / ``istft`` / ``griffinlim``; NVIDIA Tacotron2
``Postnet``) anchored on the reference's own call sites
(``src/tacotron2/mod.rs:441-458``, ``src/lib.rs:141-155``,
``src/tacotron2/mod.rs:344-357``) and is cross-checked against the independent
implementations available in this image (``torch.stft`` / ``torch.istft`` /
``torchaudio.functional`` / ``torch.nn.functional.conv1d``), see
``tests/test_oracle.py`` and ``oracle/make_golden.py``.
"""
