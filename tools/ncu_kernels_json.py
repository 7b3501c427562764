#!/usr/bin/env python
"""Per-kernel summary (JSON) of one `ncu --set full` capture; bench.py reads roofline.traffic from it.
usage: tools/ncu_kernels_json.py report.ncu-rep > profiles/rNN_ncu_kernels.json"""
import csv, io, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, name, scale_unit=True):
    v = float(r[ix[name]] or 0); u = units[ix[name]]
    if scale_unit:
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
    return v
res = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    res[name] = {"gpu_time_us": round(val(r, "gpu__time_duration.sum"), 3),
                 "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                 "registers": val(r, "launch__registers_per_thread"),
                 "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                 "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "warp_inst": val(r, "smsp__inst_executed.sum")}
json.dump(res, sys.stdout, indent=1)
