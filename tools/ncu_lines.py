#!/usr/bin/env python3
"""Per-CUDA-source-line instruction / stall attribution from an .ncu-rep (needs -lineinfo).
usage: python tools/ncu_lines.py prof.ncu-rep [min_pct]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None
hd = None
recs = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 3 and r[0] == "Line No":
        hd = r
        continue
    if hd is None or len(r) < len(hd) - 5:
        continue
    if r[2] != "-":
        continue            # SASS row
    try:
        n = int(r[hd.index("Instructions Executed")])
        st = int(r[hd.index("Warp Stall Sampling (All Samples)")] or 0)
    except ValueError:
        continue
    recs.append((fname, int(r[0]), n, st, r[1]))
tot = sum(x[2] for x in recs)
tst = sum(x[3] for x in recs)
print("total warp instr %d, stall samples %d" % (tot, tst))
for f, ln, n, st, src in recs:
    if 100.0 * n / tot >= min_pct or 100.0 * st / max(1, tst) >= min_pct:
        print("%-22s %4d  inst %5.2f%%  stall %5.2f%%  %s" % (f, ln, 100.0 * n / tot, 100.0 * st / max(1, tst), src.strip()[:100]))
