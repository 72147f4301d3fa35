#!/usr/bin/env python
"""ncu --page raw --csv export (one kernel launch) -> compact JSON summary.  usage: summarize_ncu.py <raw.csv> <out.json> <note>"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "smsp__warps_eligible.avg.per_cycle_active"]
out = {"Kernel Name": vals[hdr.index("Kernel Name")], "capture": sys.argv[3]}
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        out[k] = f"{vals[i]} {units[i]}".strip()
st = {h[33:]: float(vals[hdr.index(h)]) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h}
tot = sum(st.values()) or 1.0
out["stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot > 0.01}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out)[:400])
