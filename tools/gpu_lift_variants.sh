#!/bin/bash
# lift time of the timing-experiment builds (tools/build_lift_variants.py): which role costs what
for f in xd-tts_b200/xdtts_b200/_lib/variants/libxdtts_liftskip_*.so; do
  XDTTS_B200_LIB=$f timeout 120 python tests/gpu_tools/gl_quick.py ${@:-cfg2} 2>&1 | grep -E "cfg[0-9]: lift"
done
