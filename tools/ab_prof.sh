#!/bin/bash
# tools/ab.sh + one ncu --set full capture of the steady-state Griffin-Lim launch of cfg2 and cfg5 (product library)
tag=${1:-abp}
bash tools/ab.sh $tag "$2"
unset XDTTS_B200_LIB
mkdir -p gpurun_out/$tag
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_iter_kernel -s 20 -c 1 -f -o gpurun_out/$tag/prof_gl_iter_cfg2 \
    python tools/prof_target.py cfg2 1 > gpurun_out/$tag/prof_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_iter_kernel -s 20 -c 1 -f -o gpurun_out/$tag/prof_gl_iter_cfg5 \
    python tools/prof_target.py cfg5 1 > gpurun_out/$tag/prof_cfg5.log 2>&1
ls -la gpurun_out/$tag
