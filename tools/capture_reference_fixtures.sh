#!/bin/bash
# Pins parity to the REFERENCE'S OWN output.  Run on a machine that has what the build image lacks: cargo, git-lfs,
# network access and an ONNX Runtime 1.17 shared library (README of xd-tts: ORT_DYLIB_PATH).  It builds the UNMODIFIED
# reference, synthesises a fixed text with --output-spectrogram (the reference's own mel dump, src/lib.rs:125-139,
# src/bin/app.rs:12-14) and stores
#     tests/golden/reference/mel.npy          [80, T] float32  "mel_outputs_postnet" as the reference computed it
#     tests/golden/reference/audio.wav        16-bit PCM, 22050 Hz, the reference's Griffin-Lim output
#     tests/golden/reference/meta.json        commit, crate revision, text, parameters
# tests/test_gpu_gl.py::test_reference_fixtures (skipped while the directory is empty) then checks, on the GPU:
#   * the vocoder run on mel.npy gives a waveform whose STFT magnitude matches the reference audio's (the phases are random
#     on both sides, so the comparison is spectral: log-magnitude distance and spectral convergence), for lift = pinv and
#     lift = nnls and both exponent conventions -- the combination that matches names what the crate really does
#     (SURVEY.md section 7 "unknowns in the external crate"), and
#   * length, peak and sample-rate conventions bit for bit.
# To pin the arithmetic exactly (not only spectrally), additionally dump the crate's initial phase: patch griffin-lim's
# random init to write `angles.npy` and pass it to xdtts_gl_infer as init_phase.
#
# usage: tools/capture_reference_fixtures.sh /path/to/xd-tts-checkout "Hello world from Rust"
set -euo pipefail
REF=${1:?path to a checkout of xd009642/xd-tts (with git-lfs objects)}
TEXT=${2:-"Hello world from Rust"}
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=$HERE/tests/golden/reference
command -v cargo >/dev/null || { echo "cargo not found: this script must run outside the build image" >&2; exit 2; }
: "${ORT_DYLIB_PATH:?set ORT_DYLIB_PATH to libonnxruntime.so 1.17 (xd-tts README)}"
mkdir -p "$OUT"
( cd "$REF" && git lfs pull && cargo build --release --bin xd_tts )
# app.rs: --input <text> --output <wav> --output-spectrogram <npy>
( cd "$REF" && RUST_LOG=xd_tts=info ./target/release/xd_tts --input "$TEXT" --output "$OUT/audio.wav" --output-spectrogram "$OUT/mel.npy" \
      2> "$OUT/run.log" )
python3 - "$REF" "$TEXT" "$OUT" <<'PY'
import json, subprocess, sys
ref, text, out = sys.argv[1:4]
commit = subprocess.check_output(["git", "-C", ref, "rev-parse", "HEAD"], text=True).strip()
lock = open(ref + "/Cargo.lock").read()
i = lock.find('name = "griffin-lim"')
meta = {"reference_commit": commit, "text": text, "griffin_lim_lock_entry": lock[i:i + 400].split("\n\n")[0],
        "params": {"sample_rate": 22050, "n_fft": 1024, "noverlap": 768, "n_mels": 80, "fmin": 0.0, "fmax": 8000.0, "power": 1.7,
                   "iter": 30, "momentum": 0.99}}
json.dump(meta, open(out + "/meta.json", "w"), indent=1)
print("wrote", out)
PY
