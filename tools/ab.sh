#!/bin/bash
# A/B of library builds on one box: every tuning build under _lib/variants plus the product library,
# each timed on cfg2 and cfg5 by tests/gpu_tools/gl_quick.py (parity on a small case first).
# Usage: gpurun --timeout 900 -- 'bash tools/ab.sh <tag> [pytest-expression]'
out=gpurun_out/${1:-ab}.log
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
for f in xd-tts_b200/xdtts_b200/_lib/variants/libxdtts_*.so ""; do
  cfgs="cfg2 cfg5"
  case "$f" in *libxdtts_8_*) cfgs=cfg2;; *libxdtts_16_*) cfgs=cfg5;; esac
  if [ -n "$f" ]; then export XDTTS_B200_LIB=$f; else unset XDTTS_B200_LIB; fi
  timeout 200 python tests/gpu_tools/gl_quick.py $cfgs 2>&1 | grep -E "parity|cfg|rror|Trace" >> $out
done
if [ -n "$2" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$2" 2>&1 | tail -5 >> $out; fi
cat $out
