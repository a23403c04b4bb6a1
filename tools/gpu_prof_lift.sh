#!/bin/bash
tag=${1:-plift}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_lift_tc -c 1 -f -o $out/prof_lift_cfg2 \
    python tools/prof_target.py cfg2 1 > $out/prof_lift.log 2>&1
ls -la $out
