#!/usr/bin/env python
"""Key metrics of an ncu --set full capture -> JSON under profiles/ (same layout as round 1's summaries).
python tools/ncu_summary.py capture.ncu-rep out.json "command line of the capture" """
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]
rep, out, command = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
ix = {h: i for i, h in enumerate(hdr)}
metrics = {m: [vals[ix[m]], units[ix[m]]] for m in WANT if m in ix}
res = {"command": command, "kernel": vals[ix["Kernel Name"]] if "Kernel Name" in ix else "", "metrics": metrics}
try:
    a, e = float(metrics["sm__cycles_active.avg"][0]), float(metrics["sm__cycles_elapsed.avg"][0])
    res["derived"] = {"sm_active_of_elapsed": a / e,
                      "dram_bytes_per_launch": (float(metrics["dram__bytes_read.sum"][0]) + float(metrics["dram__bytes_write.sum"][0]))
                      * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(metrics["dram__bytes_read.sum"][1], 1.0)}
except (KeyError, ValueError):
    pass
with open(out, "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res.get("derived", {})), len(metrics), "metrics ->", out)
