#!/usr/bin/env python
"""Samples per CUDA source line (file:line) of one kernel of an ncu report, with the stall reasons of each line.
usage: ncu_srclines.py report.ncu-rep launch_index [top_n]"""
import csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2]); top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, agg, total, fn = None, None, {}, 0, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[2] != "-": continue        # keep the per-source-line rows (Address == '-')
    iS = hdr.index("# Samples")
    n = int(r[iS] or 0)
    if n == 0: continue
    st = {c[6:]: int(r[i] or 0) for i, c in enumerate(hdr) if c.startswith("stall_") and "(" not in c and int(r[i] or 0) > 0}
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, {}, r[1]])
    a[0] += n
    for k, v in st.items(): a[1][k] = a[1].get(k, 0) + v
    total += n
print(fn[:110] if fn else "?"); print("total samples", total)
for (f, ln), (n, st, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{n:6d} {100.0*n/total:5.1f}%  {f}:{ln:<5d} {src.strip()[:90]}  {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
