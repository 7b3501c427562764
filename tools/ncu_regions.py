#!/usr/bin/env python
"""Share of executed warp instructions / stall samples per source-line range of a kernel.
usage: tools/ncu_regions.py report.ncu-rep kernel_regex file.cu name:lo-hi [name:lo-hi ...]"""
import csv, subprocess, sys, io
rep, kern, src = sys.argv[1:4]
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""; tot = {}; ti = ts = 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[2] != "-": continue
    ix = {h: k for k, h in enumerate(hdr)}
    inst = float(r[ix["Instructions Executed"]] or 0); samp = float(r[ix["# Samples"]] or 0)
    ti += inst; ts += samp
    name = "other:" + fname
    if fname == src:
        ln = int(r[0])
        for n, lo, hi in regions:
            if lo <= ln <= hi: name = n; break
    t = tot.setdefault(name, [0.0, 0.0]); t[0] += inst; t[1] += samp
print(f"total warp-inst {ti:.0f} samples {ts:.0f}")
for n, (i, s) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {n:28s} inst {i/ti*100:5.1f}%  samples {s/ts*100:5.1f}%")
