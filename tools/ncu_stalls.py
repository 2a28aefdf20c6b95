#!/usr/bin/env python
"""Summarise the source page of an ncu capture: stall reasons over all samples, the instructions that collect the most
samples, and samples per source line.  python tools/ncu_stalls.py capture.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hdr_i], [r for r in rows[hdr_i + 1:] if len(r) > 10]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = defaultdict(int)
nsamp = ninst = 0
for r in data:
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    nsamp += int(r[ix["# Samples"]] or 0)
    ninst += int(r[ix["Instructions Executed"]] or 0)
print(rows[0][1] if len(rows[0]) > 1 else "")
print(f"samples {nsamp}  warp instructions {ninst:.4g}")
for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {s:26s} {100 * v / nsamp:5.1f} %")
print("top instructions by samples: samples, executions, SASS, two largest stall reasons")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n_top]:
    st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  {int(r[ix['# Samples']]):8d} {int(r[ix['Instructions Executed']]):11d}  {r[ix['Source']].strip()[:64]:64s} {st}")
