#!/usr/bin/env python
"""Per-instruction stall hot spots from an .ncu-rep (source page).  usage: ncu_hotspots.py rep [kernel-regex] [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kre] if kre else [])
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, stdout=subprocess.PIPE, text=True).stdout)))
hi = [i for i, r in enumerate(rows) if 'Source' in r][0]
hdr = rows[hi]; data = []
for r in rows[hi + 1:]:   # first captured launch only (each launch has its own table)
    if r == hdr: break
    if len(r) == len(hdr): data.append(r)
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
for r in data:
    for h in reasons: tot[h] += int(r[col[h]] or 0) if (r[col[h]] or '0').isdigit() else 0
allc = sum(tot.values())
print("static instrs", len(data), "samples", allc)
print("by reason:", ", ".join(f"{h[6:]} {100*c/allc:.1f}%" for h, c in tot.most_common(9)))
for r in sorted(data, key=lambda r: -int(r[col['# Samples']]))[:top]:
    rs = sorted(((int(r[col[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{int(r[col['# Samples']]):6d} exec {int(r[col['Instructions Executed']]):9d}  {r[col['Source']].strip()[:70]:70s} {rs}")
