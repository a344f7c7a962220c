"""Summarise an .ncu-rep into the handful of numbers the roofline discussion needs.
usage: python scripts/ncu_summary.py <report.ncu-rep> [kernel-name-substring]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "dram__bytes.sum.per_second", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if flt and flt not in d.get("Kernel Name", ""):
        continue
    s = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k in hdr:
        base = k.split(".", 2)[-1] if k.count(".") >= 2 and k.split(".")[0].isupper() else k
        for want in KEYS:
            if k.endswith(want):
                u = units[hdr.index(k)]
                s[want] = f"{d[k]} {u}".strip()
    out.append(s)
print(json.dumps(out, indent=1))
