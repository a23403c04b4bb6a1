#!/bin/bash
# lift kernel check: lift parity tests + quick kernel timing (lift and iteration kernel, CUDA events)
tag=${1:-lift}
out=gpurun_out/$tag
mkdir -p $out
timeout 150 python -m pytest tests -m gpu -x -q -k "${2:-lift or golden or tail or pipe or nnls}" 2>&1 | tail -15 > $out/pytest.log; cat $out/pytest.log
timeout 90 python tests/gpu_tools/gl_quick.py cfg2 cfg5 2>&1 | grep -E "parity|lift|cfg|rror|Trace" | tee $out/quick.log
XDTTS_LIFT_F32=1 timeout 90 python tests/gpu_tools/gl_quick.py cfg2 2>&1 | grep -E "lift" | sed 's/^/f32 fallback: /' | tee -a $out/quick.log
