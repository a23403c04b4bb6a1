#!/bin/bash
tag=${1:-mix}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python tests/gpu_tools/latency_small.py 2>&1 | tee $out/latency.log
timeout 600 python -m pytest tests -m gpu -x -q -k "pytorch or pool" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-decoder --no-extras > $out/bench.json 2> $out/bench.err
python - <<PY
import json
d = json.load(open("$out/bench.json")); e = d["e2e"]
print("cfg2 value %.3fM e2e %.3fM blocking %.3fM pageable %.3fM (ratio %.3f)" % (d["value"]/1e6, e["value"]/1e6, e["blocking_call"]["value"]/1e6, e["blocking_call_pageable"]["value"]/1e6, e["blocking_call_pageable"]["ratio_to_pinned"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_iter_kernel -s 20 -c 1 -f -o $out/prof_gl_iter_cfg2 \
    python tools/prof_target.py cfg2 1 > $out/prof_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_iter_kernel -s 20 -c 1 -f -o $out/prof_gl_iter_cfg5 \
    python tools/prof_target.py cfg5 1 > $out/prof_cfg5.log 2>&1
ls $out
