"""Timing-experiment builds of the lift kernel (gl_lift.cu with -DXDTTS_LIFT_SKIP=n), next to the product library:
    python tools/build_lift_variants.py 1 2 4 7 trace      ("trace": per-role clock stamps, printed with XDTTS_LIFT_TRACE_PRINT=1)
    XDTTS_B200_LIB=.../variants/libxdtts_liftskip_1.so python tests/gpu_tools/gl_quick.py cfg2"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

out_dir = os.path.join(g.LIB_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
g.build_cuda()
for n in sys.argv[1:]:
    obj = os.path.join(out_dir, "gl_lift_%s.o" % n)
    subprocess.check_call([g.NVCC] + g.NVCC_FLAGS + (["-DXDTTS_LIFT_TRACE"] if n == "trace" else ["-DXDTTS_LIFT_SKIP=%s" % n]) + ["-c", os.path.join(g.CSRC, "gl_lift.cu"), "-o", obj])
    others = [os.path.join(g.CSRC, "_obj", s.replace(".cu", ".o")) for s in g.CUDA_SOURCES if s != "gl_lift.cu"]
    lib = os.path.join(out_dir, "libxdtts_liftskip_%s.so" % n)
    subprocess.check_call([g.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, obj] + others +
                          ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    os.remove(obj)
    print("built", lib)
