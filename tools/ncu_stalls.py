#!/usr/bin/env python
"""Stall-reason totals and the hottest SASS lines of one kernel of an ncu report (source page).
usage: ncu_stalls.py report.ncu-rep launch_index [top_n]"""
import csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2]); top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h = r[1]; rows = []
for x in r[2:]:
    if len(x) != len(h) or x[0] == 'Address':
        if x and x[0] in ('Address', 'Kernel Name') and rows: break
        continue
    rows.append(x)
iS = h.index('# Samples'); isrc = h.index('Source')
cols = [c for c in h if c.startswith('stall_') and '(' not in c]
idx_ = [h.index(c) for c in cols]
f = lambda v: int(v or 0)
print(r[0][1][:100]); print('total samples', sum(f(x[iS]) for x in rows), 'lines', len(rows))
print({c[6:]: s for c, s in ((c, sum(f(x[i]) for x in rows)) for c, i in zip(cols, idx_)) if s})
top = sorted(range(len(rows)), key=lambda i: -f(rows[i][iS]))[:top_n]
for i in sorted(top):
    x = rows[i]; print(i, x[iS], x[isrc][:70], {c[6:]: x[j] for c, j in zip(cols, idx_) if f(x[j]) > 0})
