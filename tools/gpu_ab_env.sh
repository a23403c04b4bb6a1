#!/bin/bash
# A/B of two library builds on the quick cfg2 / cfg5 timing: bash tools/gpu_ab_env.sh <libA> <libB> [rounds]
for i in $(seq 1 ${3:-2}); do
  for lib in "$1" "$2"; do
    XDTTS_B200_LIB=$lib timeout 200 python tests/gpu_tools/gl_quick.py cfg2 cfg5 2>&1 | grep -E "step"
  done
done
