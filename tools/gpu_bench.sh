#!/bin/bash
# GPU parity tests + the default bench line (all sub-objects) + the reference arm
tag=${1:-bench}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
timeout 900 python bench.py --steps ${2:-20} --warmup 5 > $out/bench_cfg2.json 2> $out/bench_cfg2.err; echo "bench rc=$?"
tail -5 $out/bench_cfg2.err
python - <<PY
import json
d = json.load(open("$out/bench_cfg2.json"))
def show(name, q):
    e = q["e2e"]
    print("%s value %.3fM  e2e %.3fM (blocking %.3fM, pageable %.3fM)  kernel %.2f us frac %.3f  lift %.1f us frac %.3f" % (
        name, q["value"] / 1e6, e["value"] / 1e6, e.get("blocking_call", {}).get("value", 0) / 1e6,
        e.get("blocking_call_pageable", {}).get("value", 0) / 1e6, q["roofline"]["kernel_ms"] * 1e3, q["roofline"]["frac"],
        q["lift"]["ms"] * 1e3, q["lift"]["frac"]))
show("cfg2", d)
for c in ("cfg3", "cfg5"):
    if c in d: show(c, d[c])
if "strong_scaling" in d: print("strong", json.dumps(d["strong_scaling"])[:400])
if "dropin" in d: print("dropin", json.dumps(d["dropin"]["cases"])[:900])
if "decoder" in d: print("decoder", {k: d["decoder"][k] for k in ("b1", "b8", "floors", "frac", "frac_of_handover_floor")})
if "postnet" in d.get("cfg3", {}): print("postnet", {k: d["cfg3"]["postnet"][k] for k in ("ms", "frac", "frac_executed")})
print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("candidates"))
print("clocks", d["clocks"], "launches", d["gpu_launches"])
PY
