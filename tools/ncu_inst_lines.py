#!/usr/bin/env python
"""Source lines of a kernel ranked by executed warp instructions (ncu report with --import-source on).
usage: tools/ncu_inst_lines.py report.ncu-rep kernel_regex [launch_skip] [top_n]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""; agg = {}; src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[2] != "-": continue
    ix = {h: k for k, h in enumerate(hdr)}
    key = (fname, int(r[0]))
    agg[key] = agg.get(key, 0) + float(r[ix["Instructions Executed"]] or 0)
    src[key] = r[1].strip()
tot = sum(agg.values()) or 1
print(f"total warp instructions {tot:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v / tot * 100:5.1f}% {k[0][:14]:14s}:{k[1]:4d} | {src[k][:110]}")
