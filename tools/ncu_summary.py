#!/usr/bin/env python
"""Summarise what a tools/gpu_round.sh call brought back (gpurun_out/<tag>/) into profiles/<tag>_*.

    python tools/ncu_summary.py <tag>

Writes profiles/<tag>_launches.md (per-kernel share of a pass, from the gpu__time_duration launch
list), profiles/<tag>_<report>.md (key `ncu --set full` metrics + top stall sites per kernel) and
copies the bench JSON lines.  Runs here, without a GPU (ncu -i only reads the report).
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarise_report(rep, dst):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}
    lines = ["# %s" % os.path.basename(rep), "",
             "`ncu --set full --clock-control none --import-source on` (cold-cache, serialised replays; "
             "durations are NOT bench numbers).", ""]
    for r in data:
        lines += ["## %s  (launch id %s)" % (r[ix["Kernel Name"]], r[ix["ID"]]), "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in ix and r[ix[k]] != "":
                lines.append("| %s | %s | %s |" % (k, r[ix[k]], units[ix[k]]))
        for k, i in ix.items():
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 100:
                    lines.append("| %s | %s | samples |" % (k, r[i]))
        if "dram__bytes_read.sum" in ix:
            def to_bytes(key):
                v, u = float(r[ix[key]]), units[ix[key]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            tr = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            lines.append("| **traffic = dram read + write** | %.0f | byte |" % tr)
        lines.append("")
    # top stall sites of the first kernel (source page)
    src = ncu_csv(rep, "source")
    hi = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if hi:
        h = src[hi[0]]
        sx = {k: i for i, k in enumerate(h)}
        seg = []
        for r in src[hi[0] + 1:]:
            if r and r[0] == "Kernel Name":
                break
            if len(r) == len(h):
                seg.append(r)
        tot = sum(int(r[sx["# Samples"]]) for r in seg) or 1
        lines += ["## top stall sites (first captured launch, %d SASS instructions, %d samples)" % (len(seg), tot), "",
                  "| samples | % | SASS | long_sb | short_sb | wait |", "|---|---|---|---|---|---|"]
        for r in sorted(seg, key=lambda r: -int(r[sx["# Samples"]]))[:25]:
            n = int(r[sx["# Samples"]])
            lines.append("| %d | %.1f | `%s` | %s | %s | %s |" % (n, 100.0 * n / tot, r[sx["Source"]].strip()[:60],
                                                                  r[sx["stall_long_sb"]], r[sx["stall_short_sb"]], r[sx["stall_wait"]]))
        lines.append("")
    open(dst, "w").write("\n".join(lines))
    return data, ix, units


def summarise_launches(path, dst):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {k: i for i, k in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[ix["Metric Value"]])
        except ValueError:
            continue
        a = agg.setdefault(r[ix["Kernel Name"]].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1
    lines = ["# %s" % os.path.basename(path), "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` over kernel-by-kernel passes of the bench workload "
             "(tools/prof_target.py).  Per-launch times are cold-cache and serialised: compare SHARES.", "",
             "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, a in agg.items():
        lines.append("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))
    open(dst, "w").write("\n".join(lines) + "\n")


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    traffic = {}
    tpath = os.path.join(dst, "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for f in sorted(os.listdir(src)):
        p = os.path.join(src, f)
        if f.endswith(".ncu-rep"):
            data, ix, units = summarise_report(p, os.path.join(dst, "%s_%s.md" % (tag, f[:-8])))
            for cfg in ("cfg2", "cfg5", "cfg3"):
                if cfg in f and "gl_iter" in f and data:
                    r = data[0]
                    mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
                    tr = sum(float(r[ix[k]]) * mul.get(units[ix[k]].lower(), 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    traffic[cfg] = tr
        elif f.startswith("launches") and f.endswith(".csv"):
            summarise_launches(p, os.path.join(dst, "%s_%s.md" % (tag, f[:-4])))
            shutil.copy(p, os.path.join(dst, "%s_%s" % (tag, f)))
        elif f.endswith(".json") or f in ("pytest_gpu.log", "smoke.log", "gpu.txt"):
            shutil.copy(p, os.path.join(dst, "%s_%s" % (tag, f)))
        elif f.startswith("sanitizer_") and f.endswith(".log"):     # keep the verdict lines, not the whole log
            lines = [ln for ln in open(p, errors="replace").read().splitlines() if "SUMMARY" in ln or "sanitize target" in ln]
            with open(os.path.join(dst, "%s_%s" % (tag, f)), "w") as out:
                out.write("\n".join(lines) + "\n")
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("profiles/ updated from", src)


if __name__ == "__main__":
    main()
