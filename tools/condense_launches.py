#!/usr/bin/env python
"""Condense an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`) into
id,kernel,grid,block,duration_ns: torch-internal kernel names cut to 60 characters, ours (namespace b200f::) kept up
to the argument list.  Usage: condense_launches.py launches.csv out.csv "<the ncu command line>" """
import csv
import sys


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    lines = [ln for ln in open(src, errors="replace") if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, data = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        if cmd:
            f.write(f"# {cmd}\n")
        f.write("# condensed: torch-internal kernels (at::...) cut to 60 characters, ours (namespace b200f::) up to the argument list\n")
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in data:
            if r[col["Metric Name"]] != "gpu__time_duration.sum":
                continue
            name = r[col["Kernel Name"]].replace("void ", "")
            ours = ("umma::", "rowops::", "gallery::", "head_simt::", "simt::", "b200f::")
            if name.startswith(ours):                         # ncu prints our kernels without the outer b200f:: namespace
                name = name.split("(")[0].replace("b200f::", "")
            else:
                name = name[:60]
            val = float(r[col["Metric Value"]].replace(",", ""))
            unit = r[col["Metric Unit"]]
            ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
            f.write(f'{r[col["ID"]]},"{name}","{r[col["Grid Size"]]}","{r[col["Block Size"]]}",{int(round(ns))}\n')


if __name__ == "__main__":
    main()
