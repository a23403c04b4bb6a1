#!/bin/bash
# multi-GPU check on N GPUs of one box: pool parity test, then the bench under torchrun (weak; the line carries the strong sub-object)
n=${1:-2}
tag=${2:-multi}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -x -q 2>&1 | tail -5 | tee $out/pytest_pool.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 \
    > $out/scale_n$n.json 2> $out/scale_n$n.err; echo "bench rc=$?"
tail -3 $out/scale_n$n.err
python - <<PY
import json
d = json.load(open("$out/scale_n$n.json"))
print("N=%d weak value %.3fM e2e %.3fM kernel %.2f us frac %.3f" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"]))
print("strong", json.dumps(d.get("strong_scaling"))[:500])
PY
