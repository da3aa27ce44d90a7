#!/usr/bin/env python
"""Compact per-launch table of the metrics that matter from an `ncu --set full` report (raw page).

  python tools/ncu_raw_summary.py REPORT.ncu-rep > profiles/NAME.csv
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
cols = [(w, h.index(w)) for w in WANT if w in h]
kn = h.index("Kernel Name")
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + [f"{n} [{units[i]}]" for n, i in cols])
for r in rows[2:]:
    name = r[kn].split("(")[0].replace("void ", "").replace("b200ssl::", "")
    w.writerow([name] + [r[i] for _, i in cols])
