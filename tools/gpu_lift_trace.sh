XDTTS_LIFT_TRACE_PRINT=1 XDTTS_B200_LIB=xd-tts_b200/xdtts_b200/_lib/variants/libxdtts_liftskip_trace.so timeout 120 python tools/prof_target.py cfg2 2 2>&1 | grep -A12 "lift trace" | tail -30
XDTTS_LIFT_TRACE_PRINT=1 XDTTS_B200_LIB=xd-tts_b200/xdtts_b200/_lib/variants/libxdtts_liftskip_trace.so timeout 120 python tools/prof_target.py cfg5 2 2>&1 | grep -A18 "lift trace" | tail -20
