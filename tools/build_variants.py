"""Tuning builds of the library with other Griffin-Lim launch shapes (warps per CTA, CTAs per SM,
exchange-buffer aliasing), written next to the product library so that they travel to the GPU box:

    python tools/build_variants.py 8:6,2,0 8:4,4,1 16:9,1,1 ...      # R3:WARPS,CTAS,ALIAS
    XDTTS_B200_LIB=xd-tts_b200/xdtts_b200/_lib/variants/libxdtts_8_4_4_1.so python tests/gpu_tools/gl_quick.py cfg2
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

out_dir = os.path.join(g.LIB_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
g.build_cuda()
procs = []
for spec in sys.argv[1:]:
    r3, rest = spec.split(":")
    w, c, a = rest.split(",")
    tag = "%s_%s_%s_%s%s" % (r3, w, c, a, os.environ.get("XDTTS_VARIANT_TAG", ""))
    obj = os.path.join(out_dir, "gl_iter_%s.o" % tag)
    defs = ["-DXDTTS_GL%s_WARPS=%s" % (r3, w), "-DXDTTS_GL%s_CTAS=%s" % (r3, c), "-DXDTTS_GL%s_ALIAS=%s" % (r3, a)]
    defs += [d for d in os.environ.get("XDTTS_VARIANT_DEFS", "").split() if d]
    cmd = [g.NVCC] + g.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-c", os.path.join(g.CSRC, "gl_iter.cu"), "-o", obj]
    procs.append((tag, obj, r3, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for tag, obj, r3, p in procs:
    out, _ = p.communicate()
    if p.returncode:
        print(out)
        raise SystemExit("variant %s failed" % tag)
    lines = out.splitlines()
    for i, ln in enumerate(lines):
        if "ILi%sELi2ELb1ELb0" % r3 in ln and "Compiling" in ln:   # the steady-state kernel of this geometry
            print(tag, "|", lines[i + 2].strip() if i + 2 < len(lines) else "", "|", lines[i + 3].strip() if i + 3 < len(lines) else "")
    others = [os.path.join(g.CSRC, "_obj", s.replace(".cu", ".o")) for s in g.CUDA_SOURCES if s != "gl_iter.cu"]
    lib = os.path.join(out_dir, "libxdtts_%s.so" % tag)
    subprocess.check_call([g.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, obj] + others +
                          ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    os.remove(obj)
    print("built", lib)
