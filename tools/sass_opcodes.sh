#!/bin/bash
# SASS opcode evidence of the in-tree library (profiles/rNN_sass_opcodes.txt): TMA bulk copy, mbarrier, cp.async, warp
# reductions, programmatic dependent launch; and that no tensor-core / library kernels are involved.
cd "$(dirname "$0")/.."
so=bonnie-32_b200/libb32raster.so
echo "# cuobjdump -sass $so  ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
echo "# kernels (sm_100a):"
cuobjdump -sass $so | grep -o "Function : [^ ]*" | sed 's/Function : //' | c++filt | sed 's/(.*//' | sort | uniq -c | sort -rn
echo
echo "# opcode counts over all kernels (selected families)"
cuobjdump -sass $so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sort | uniq -c | sort -rn > /tmp/ops.txt
grep -E "UBLKCP|SYNCS|LDGSTS|LDGDEPBAR|DEPBAR|REDUX|CREDUX|ACQBULK|PREEXIT|ATOMS|ATOMG|RED\.|VOTE|SHFL|MATCH|BAR\.|HMMA|UTC|TCGEN|WGMMA|IMMA|UTMA|FFMA|FMUL|FADD|MUFU" /tmp/ops.txt
echo
echo "# per kernel: the instructions that prove TMA / mbarrier / cp.async / warp reductions / PDL"
cuobjdump -sass $so | python3 -c '
import re, sys, subprocess, collections
fn = None; k = collections.Counter()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and re.match(r"UBLKCP|SYNCS\.|LDGSTS|ACQBULK|PREEXIT|REDUX|CREDUX", m.group(1)): k[(fn, m.group(1))] += 1
names = sorted({f for f, _ in k})
dem = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")))
for (f, op), n in sorted(k.items()):
    print("%5d  %-34s %s" % (n, op, re.sub(r"\(.*", "", dem[f])))
'
echo
echo "# -fmad=false: no user-level multiply-add is fused (the reference never fuses).  The FFMA that remain belong to the IEEE-exact"
echo "# division / square-root sequences (-prec-div/-prec-sqrt: MUFU.RCP or MUFU.RSQ seed + FFMA Newton steps, ~4 FFMA per division):"
echo "#   FFMA total:" $(awk '$2 ~ /^FFMA/ {n+=$1} END {print n}' /tmp/ops.txt) "  MUFU.RCP:" $(awk '$2=="MUFU.RCP" {print $1}' /tmp/ops.txt) "  MUFU.RSQ:" $(awk '$2=="MUFU.RSQ" {print $1}' /tmp/ops.txt)
