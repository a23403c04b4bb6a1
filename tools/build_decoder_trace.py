"""Timing-experiment build of the decoder loop (decoder.cu with -DXDTTS_DEC_TRACE: clock stamps at every hand-over):
    python tools/build_decoder_trace.py
    XDTTS_B200_LIB=.../variants/libxdtts_dectrace.so python tools/prof_decoder.py 1 300"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build_cuda()
out_dir = os.path.join(g.LIB_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "decoder_trace.o")
subprocess.check_call([g.NVCC] + g.NVCC_FLAGS + ["-DXDTTS_DEC_TRACE", "-c", os.path.join(g.CSRC, "decoder.cu"), "-o", obj])
others = [os.path.join(g.CSRC, "_obj", s.replace(".cu", ".o")) for s in g.CUDA_SOURCES if s != "decoder.cu"]
lib = os.path.join(out_dir, "libxdtts_dectrace.so")
subprocess.check_call([g.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, obj] + others +
                      ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
os.remove(obj)
print("built", lib)
