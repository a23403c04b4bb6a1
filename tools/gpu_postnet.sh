#!/bin/bash
# postnet bring-up: CUDA-core path first, then the tcgen05 path, each under its own timeout
tag=${1:-pn}
out=gpurun_out/$tag
mkdir -p $out
export XDTTS_DEBUG=1
timeout 180 python -m pytest tests/test_gpu_postnet.py -x -q -k "golden_fixture and 2" > $out/pn_fp32.log 2>&1; echo "rc=$?" >> $out/pn_fp32.log; tail -15 $out/pn_fp32.log
timeout 180 python -m pytest tests/test_gpu_postnet.py -x -q -k "golden_fixture and 1" > $out/pn_bf16.log 2>&1; echo "rc=$?" >> $out/pn_bf16.log; tail -15 $out/pn_bf16.log
timeout 180 python -m pytest tests/test_gpu_postnet.py -x -q -k "golden_fixture and 0" > $out/pn_bf16x3.log 2>&1; echo "rc=$?" >> $out/pn_bf16x3.log; tail -15 $out/pn_bf16x3.log
timeout 600 python -m pytest tests/test_gpu_postnet.py -q > $out/pn_all.log 2>&1; echo "rc=$?" >> $out/pn_all.log; tail -40 $out/pn_all.log
