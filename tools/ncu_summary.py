#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the few per-kernel numbers the
design discussion uses, as a markdown table + a JSON file.  Usage: ncu_summary.py report.ncu-rep out_prefix"""
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_read"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_size", "cluster"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in data:
        name = r[col["Kernel Name"]]
        rec = {"kernel": name.split("(")[0].replace("void ", "").replace("b200f::", "")[:70]}
        for k, short in KEYS:
            if k in col:
                v = r[col[k]]
                try:
                    rec[short] = float(v)
                except ValueError:
                    rec[short] = v
                rec[short + "_unit"] = units[col[k]]
        recs.append(rec)
    json.dump(recs, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n\n")
        f.write("| kernel | time | tensor active % | issue active % | DRAM read | DRAM write | L2->SM read | L2 hit % | regs | grid | cluster |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in recs:
            g = lambda k: (f"{r[k]:.4g} {r.get(k + '_unit', '')}".strip() if isinstance(r.get(k), float) else str(r.get(k, "")))
            f.write(f"| `{r['kernel']}` | {g('time')} | {g('tensor_active_pct')} | {g('issue_active_pct')} | {g('dram_read')} | "
                    f"{g('dram_write')} | {g('l2_to_sm_read')} | {g('l2_hit_pct')} | {g('regs')} | {g('grid')} | {g('cluster')} |\n")
    print("wrote", out + ".md", out + ".json")


if __name__ == "__main__":
    main()
