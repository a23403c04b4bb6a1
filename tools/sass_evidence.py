#!/usr/bin/env python
"""SASS evidence that the kernels are Blackwell-native (B200_PROFILING.md "What proves a Blackwell-native kernel"):
per kernel of the shipped library, the count of tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM), TMA / bulk copies
(UTMALDG / UBLKCP), tcgen05 barriers (UTCBAR), legacy tensor ops (HMMA: must be 0) and registers.

    python tools/sass_evidence.py > profiles/r02_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

PAT = collections.OrderedDict([("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"),
                               ("UTMALDG (TMA tensor load)", r"\bUTMALDG"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"),
                               ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (legacy mma.sync)", r"\bHMMA"), ("LDGSTS (cp.async)", r"\bLDGSTS")])
sass = subprocess.run(["cuobjdump", "-sass", g.LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", g.LIB], capture_output=True, text=True).stdout
regs = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", ln)
    if m and cur:
        regs[cur] = int(m.group(1))
rows = collections.OrderedDict()
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", ln):
        rows[cur]["instr"] += 1
        for name, pat in PAT.items():
            if re.search(pat, ln):
                rows[cur][name] += 1
dem = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence, %s (`cuobjdump -sass`, sm_100a)\n" % os.path.relpath(g.LIB, ROOT))
print("| kernel | SASS instr | regs | " + " | ".join(PAT) + " |")
print("|---|---|---|" + "---|" * len(PAT))
for (k, c), d in zip(rows.items(), dem):
    d = d.replace("(anonymous namespace)::", "").replace("xdtts::", "")   # before cutting the parameter list at its "("
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", d))
    print("| `%s` | %d | %s | " % (name[:70], c["instr"], regs.get(k, "?")) + " | ".join(str(c[n]) for n in PAT) + " |")
