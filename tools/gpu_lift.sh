#!/bin/bash
# lift kernel check: GL parity tests, quick kernel timing, and the ncu launch list of the first launches of a cfg2 pass
tag=${1:-lift}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "gl or GL or griffin or lift or tail or pipe" 2>&1 | tail -15 > $out/pytest.log; cat $out/pytest.log
timeout 200 python tests/gpu_tools/gl_quick.py cfg2 cfg5 2>&1 | grep -E "parity|cfg|rror|Trace" | tee $out/quick.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file $out/launches_head_cfg2.csv \
    python tools/prof_target.py cfg2 1 > $out/launches.log 2>&1
grep -E "gl_lift" $out/launches_head_cfg2.csv | cut -d, -f5,12- | head -4
XDTTS_LIFT_F32=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4 --csv --log-file $out/launches_head_cfg2_f32.csv \
    python tools/prof_target.py cfg2 1 > $out/launches_f32.log 2>&1
grep -E "gl_lift" $out/launches_head_cfg2_f32.csv | cut -d, -f5,12- | head -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4 --csv --log-file $out/launches_head_cfg5.csv \
    python tools/prof_target.py cfg5 1 > $out/launches5.log 2>&1
grep -E "gl_lift" $out/launches_head_cfg5.csv | cut -d, -f5,12- | head -4
