#!/usr/bin/env python
"""profiles/rNN_ncu_kernels.json from two `ncu --set full` captures of tools/ncu_c4.py: one with ncu's default cache control
(caches flushed before every kernel) and one with --cache-control none (steady state: what a frame finds in L2).
usage: tools/ncu_kernels_json.py flushed.ncu-rep steady.ncu-rep > profiles/r02_ncu_kernels.json
The enqueued frames' launches are used (fill shape OpSparse, the clear folded into k_setup)."""
import csv, io, json, subprocess, sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        key = "k_setup" if "k_setup" in name.split("(")[0] else ("k_fill_opaque_dense" if "OpCfg<512" in name else "k_fill_opaque")
        f = lambda m: float(r[ix[m]].replace(",", "")) if m in ix and r[ix[m]] not in ("", "n/a") else None
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = f("dram__bytes_read.sum") * scale[rows[1][ix["dram__bytes_read.sum"]]]
        wr = f("dram__bytes_write.sum") * scale[rows[1][ix["dram__bytes_write.sum"]]]
        res[key] = {"time_us": f("gpu__time_duration.sum"), "dram_bytes": rd + wr, "warp_instructions": f("smsp__inst_executed.sum"),
                    "registers": f("launch__registers_per_thread"), "grid": f("launch__grid_size"), "block": f("launch__block_size"),
                    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "dram_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}      # the last launch of each kind wins
    return res


fl, stdy = load(sys.argv[1]), load(sys.argv[2])
out = {}
for k in fl:
    out[k] = dict(fl[k])
    out[k]["dram_bytes_flushed"] = out[k].pop("dram_bytes")
    out[k]["dram_bytes_steady"] = stdy.get(k, {}).get("dram_bytes")
    out[k]["time_us_steady"] = stdy.get(k, {}).get("time_us")
out["_how"] = "ncu --set full --clock-control none [--cache-control none] -k regex:k_setup|k_fill_opaque python tools/ncu_c4.py; C4 frame, 320x240"
out.pop("k_fill_opaque_dense", None) if False else None
json.dump(out, sys.stdout, indent=1)
