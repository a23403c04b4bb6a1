#!/bin/bash
# time every tuning build found under _lib/variants (see tools/build_variants.py)
out=gpurun_out/${1:-variants}.log
: > $out
for f in xd-tts_b200/xdtts_b200/_lib/variants/libxdtts_*.so; do
  r3=$(basename $f | cut -d_ -f2)
  cfg=cfg2; [ "$r3" = "16" ] && cfg=cfg5
  XDTTS_B200_LIB=$f timeout 120 python tests/gpu_tools/gl_quick.py $cfg 2>&1 | grep -E "parity|cfg|rror" >> $out
done
timeout 120 python tests/gpu_tools/gl_quick.py 2>&1 | grep -E "parity|cfg|rror" >> $out
cat $out
