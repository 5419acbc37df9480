#!/usr/bin/env python
"""Turns the ncu artefacts gpurun brought back (gpurun_out/) into the small text summaries kept
under profiles/:  python profiles/summarize.py <tag>"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEEP = (
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
    "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
)


def launches(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, rows = r, rows[i + 1:]
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")) * scale.get(r[ui], 1))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none",
           "# cold-cache, serialised launches: compare SHARES, not absolutes",
           f"# total {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches",
           "share%  total_ms  launches  avg_us  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{100 * sum(v) / tot:6.2f} {sum(v) / 1e6:9.3f} {len(v):9d} {sum(v) / len(v) / 1e3:9.1f}  {k[:110]}")
    return "\n".join(out) + "\n"


def full(name):
    rep = os.path.join(OUT, name + ".ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none ({name}), selected raw metrics per captured launch"]
    for v in rows[2:]:
        out.append("## " + v[h.index("Kernel Name")][:120])
        for a, b, c in zip(h, u, v):
            if any(a.startswith(k) for k in KEEP):
                out.append(f"{a:95s} {c:>20s} {b}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    tag = sys.argv[1]
    dst = os.path.join(ROOT, "profiles")
    if os.path.exists(os.path.join(OUT, f"launches_{tag}.csv")):
        open(os.path.join(dst, f"{tag}_launches.txt"), "w").write(launches(tag))
    for kind in ("ensemble", "cloud", "population"):
        name = f"prof_{kind}_{tag}"
        if os.path.exists(os.path.join(OUT, name + ".ncu-rep")):
            open(os.path.join(dst, f"{tag}_{kind}_ncu.txt"), "w").write(full(name))
    print("written to", dst)
