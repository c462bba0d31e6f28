#!/usr/bin/env python
"""Summarise an .ncu-rep: per captured kernel the headline metrics, stall reasons and (optionally) hot source lines."""
import csv, subprocess, sys, io, collections, signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)   # piping into head must not end in a traceback
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:70])
    for k in keys:
        if k in d: print(f"   {k} = {d[k]}  [{rows[1][hdr.index(k)]}]")
    st = [(float(v), h) for h, v in d.items() if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
    for v, h in sorted(st, reverse=True)[:7]:
        print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.2f}")
