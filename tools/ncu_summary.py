#!/usr/bin/env python
"""Key metrics per kernel launch from an ncu report. usage: tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]
stalls = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("=====", r[ix["Kernel Name"]][:60])
    for w in want:
        if w in ix:
            print(f"  {w:70s} {r[ix[w]]:>16s} {units[ix[w]]}")
    st = sorted(((float(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:6]
    for v, h in st:
        print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:8.2f} warps/issue")
