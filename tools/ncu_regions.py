#!/usr/bin/env python
"""Instruction-count breakdown of a kernel by code region (runs of consecutive SASS instructions with the same
execution count).  usage: ncu_regions.py rep kernel-regex [min-share]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if 'Source' in r][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:   # first captured launch only (each launch has its own table)
    if r == hdr: break
    if len(r) == len(hdr): data.append(r)
tot = sum(int(r[col['Instructions Executed']]) for r in data)
print('total warp-instructions', tot, 'static', len(data))
run, runs = None, []
for r in data:
    ex = int(r[col['Instructions Executed']])
    if run and abs(ex - run[0]) <= 0.02 * max(run[0], 1):
        run[1] += 1; run[2] += ex; run[4] += int(r[col['# Samples']])
    else:
        if run: runs.append(run)
        run = [ex, 1, ex, r[col['Source']].strip()[:60], int(r[col['# Samples']])]
runs.append(run)
for ex, n, s, src, smp in runs:
    if s > tot * min_share:
        print(f'{ex:9d} x {n:4d} instrs = {s/1e6:7.1f}M ({100*s/tot:4.1f}%) samples {smp:5d}  first: {src}')
