#!/usr/bin/env python
"""Top source lines of a kernel from an ncu report: share of executed instructions and of stall samples.
usage: tools/ncu_lines.py report.ncu-rep kernel_regex [launch_skip]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[2] != "-": continue                      # SASS rows
    ix = {h: k for k, h in enumerate(hdr)}
    inst = float(r[ix["Instructions Executed"]] or 0); samp = float(r[ix["# Samples"]] or 0)
    tinst = float(r[ix["Thread Instructions Executed"]] or 0)
    agg.append((fname, r[0], r[1].strip(), inst, samp, tinst))
ti = sum(a[3] for a in agg) or 1; ts = sum(a[4] for a in agg) or 1
print(f"total warp-inst {ti:.0f}  samples {ts:.0f}")
for a in sorted(agg, key=lambda a: -(a[3] / ti + a[4] / ts))[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print(f"{a[0][:14]:14s}:{a[1]:>4s} inst {a[3]/ti*100:5.1f}% samp {a[4]/ts*100:5.1f}% thr/inst {a[5]/max(a[3],1):4.1f} | {a[2][:105]}")
