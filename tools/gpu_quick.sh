#!/bin/bash
# Short GPU check between full rounds: parity tests + cfg2/cfg3/cfg5 bench lines (no CPU leg, no ncu).
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
tag=${1:-quick}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
for cfg in cfg2 cfg3 cfg5; do
  timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu > $out/bench_$cfg.json 2> $out/bench_$cfg.err
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_$cfg.json"))
    if "decoder" in d:
        q = d["decoder"]
        print("decoder b1 %.1f us/step (%.0f GB/s, frac %.2f)  b8 %.1f us/step  %.0f / %.0f frames/s" % (q["b1"]["us_per_step"], q["achieved"], q["frac"], q["b8"]["us_per_step"], q["b1"]["frames_per_s"], q["b8"]["frames_per_s"]))
    print("$cfg value %.3fM  e2e %.3fM (blocking %.3fM)  kernel %.2f us frac %.3f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6,
          d["e2e"].get("blocking_call", {}).get("value", 0) / 1e6, d["roofline"]["kernel_ms"] * 1e3, d["roofline"]["frac"]))
except Exception as e:
    print("$cfg failed", e); print(open("$out/bench_$cfg.err").read()[-2000:])
PY
done
