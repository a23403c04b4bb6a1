#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list + full captures of the top kernels, sanitizer.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_cfg2.json 2> $out/bench_cfg2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_cfg2.csv \
    python tools/prof_target.py cfg2 2 > $out/launches_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gl_iter_kernel -s 20 -c 2 -f -o $out/prof_gl_iter_cfg2 \
    python tools/prof_target.py cfg2 1 > $out/prof_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gl_iter_kernel -s 20 -c 1 -f -o $out/prof_gl_iter_cfg5 \
    python tools/prof_target.py cfg5 1 > $out/prof_cfg5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_lift_tc -c 1 -f -o $out/prof_lift_cfg2 \
    python tools/prof_target.py cfg2 1 > $out/prof_lift.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pn_conv_tc -s 1 -c 2 -f -o $out/prof_postnet_cfg3 \
    python tools/prof_postnet.py 1 0 > $out/prof_postnet.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_persist -c 1 -f -o $out/prof_decoder_b1 \
    python tools/prof_decoder.py 1 300 > $out/prof_decoder.log 2>&1
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/gpu_tools/sanitize_target.py > $out/sanitizer_$tool.log 2>&1; tail -2 $out/sanitizer_$tool.log
done
ls -la $out
