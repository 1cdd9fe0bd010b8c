#!/usr/bin/env python
"""Summarise the `ncu --set full` captures of tools/gpu_profile.sh into profiles/<tag>_summary.json and the raw CSVs.

usage: tools/ncu_summary.py <tag> [sets_per_launch]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
sets = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
out = {"command": "python bench.py --sets %d --steps 1 --warmup 3 --no-cpu under `ncu --set full --clock-control none "
                  "--import-source on -k regex:<kernel> -s 3 -c 1` (tools/gpu_profile.sh %s), after the same command "
                  "exited 0 without ncu" % (sets, tag), "sets_per_launch": sets}
for name, rep in (("spectrum_kernel", "prof_spectrum_%s.ncu-rep" % tag), ("nthcomp_kernel", "prof_nthcomp_%s.ncu-rep" % tag),
                  ("hist_kernel", "prof_hist_%s.ncu-rep" % tag), ("nthcomp_interp_kernel", "prof_interp_%s.ncu-rep" % tag)):
    if not os.path.exists(os.path.join(ROOT, "gpurun_out", rep)):
        continue
    path = os.path.join(ROOT, "gpurun_out", rep)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    short = "interp" if "interp" in name else name.split("_")[0]
    open(os.path.join(ROOT, "profiles", "%s_%s_raw.csv" % (tag, short)), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                d[h] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                pass
    out[name] = d


def gbytes(m):
    f = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[m["unit"]]
    return m["value"] * f


sp = out["spectrum_kernel"]
tot = gbytes(sp["dram__bytes_read.sum"]) + gbytes(sp["dram__bytes_write.sum"])
out["spectrum_dram_bytes_per_launch"] = tot
out["spectrum_dram_bytes_per_set"] = tot / sets
json.dump(out, open(os.path.join(ROOT, "profiles", "%s_summary.json" % tag), "w"), indent=1)
print("spectrum: %.2f ms, %.0f B/set DRAM; fp64 pipe %.1f %%, issue %.1f %%" % (
    sp["gpu__time_duration.sum"]["value"], tot / sets,
    sp["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"],
    sp["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"]))
nt = out["nthcomp_kernel"]
print("nthcomp: %.2f ms, read %.2f %s write %.2f %s; issue %.1f %%" % (
    nt["gpu__time_duration.sum"]["value"], nt["dram__bytes_read.sum"]["value"], nt["dram__bytes_read.sum"]["unit"],
    nt["dram__bytes_write.sum"]["value"], nt["dram__bytes_write.sum"]["unit"],
    nt["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"]))
