#!/bin/bash
# decoder loop check: parity tests + us/step at batch 1 / 2 / 4 / 8, with and without the shared-memory weight cache
tag=${1:-dec}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "decoder" 2>&1 | tail -8 | tee $out/pytest.log
for nb in 1 2 4 8; do
  timeout 200 python tools/prof_decoder.py $nb 300 2>&1 | tail -1 | sed 's/^/cache: /' | tee -a $out/times.log
  XDTTS_DEC_NO_CACHE=1 timeout 200 python tools/prof_decoder.py $nb 300 2>&1 | tail -1 | sed 's/^/no cache: /' | tee -a $out/times.log
done
