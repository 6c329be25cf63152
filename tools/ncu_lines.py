#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals from an .ncu-rep (ncu --page source --print-source cuda,sass).

    python tools/ncu_lines.py report.ncu-rep [min_pct]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
lines = []   # (file, line, src, inst, thread_inst, samples)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_tinst = hdr.index("Thread Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[0] == "-" or not r[0].isdigit():
        continue   # SASS rows under a CUDA line
    try:
        lines.append((cur_file, int(r[0]), r[1], int(r[i_inst]), int(r[i_tinst]), int(r[i_samp])))
    except (ValueError, IndexError):
        pass
tot = sum(l[3] for l in lines) or 1
tots = sum(l[5] for l in lines) or 1
print(f"total warp instructions {tot}, samples {tots}")
for f, ln, src, inst, tinst, samp in lines:
    if 100.0 * inst / tot >= min_pct or 100.0 * samp / tots >= min_pct:
        print(f"{f}:{ln:4d} inst {100.0*inst/tot:5.1f}%  thr/inst {tinst/max(inst,1):5.1f}  samples {100.0*samp/tots:5.1f}%  | {src.strip()[:110]}")
