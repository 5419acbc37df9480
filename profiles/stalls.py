#!/usr/bin/env python
"""Warp-stall sampling of a captured kernel by reason and by opcode, from the SASS page of an
.ncu-rep:  python profiles/stalls.py gpurun_out/prof_ensemble_r01g.ncu-rep > profiles/<tag>_stalls.txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print("#", rows[0][1][:110])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
tot, opc, opn, n = collections.Counter(), collections.Counter(), collections.Counter(), 0
for r in data:
    try:
        s = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    n += s
    src = r[ix["Source"]].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    opc[op] += s
    opn[op] += int(r[ix["Instructions Executed"]] or 0)
    for c in stall_cols:
        try:
            tot[c] += int(r[ix[c]])
        except ValueError:
            pass
print(f"# {n} warp-stall samples (ncu --set full, --clock-control none)")
print("stall reason            share%")
for k, v in tot.most_common(12):
    print(f"  {k:22s} {100 * v / n:5.1f}")
print("opcode      samples%   warp-instructions executed")
for k, v in opc.most_common(18):
    print(f"  {k:10s} {100 * v / n:6.1f}   {opn[k]:.3e}")
