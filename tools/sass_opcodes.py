#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what the shipped library runs on (tcgen05 MMA, TMEM loads, TMA
loads / stores / prefetches, mbarrier traffic), from `cuobjdump -sass` of the in-tree libb200face.so.
Usage: python tools/sass_opcodes.py [lib.so] > profiles/rNN_sass_opcodes.txt     (needs no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "facerecognition-multiarchitecture-pipeline_b200", "libb200face.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "ELECT", "MUFU.EX2", "STG.E.ENL2.256", "LDG", "STG", "ATOMG", "REDG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
counts, cur, order = {}, None, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); order.append(cur); continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} -- instruction counts per kernel (static), sm_100a")
print("# UTCHMMA = tcgen05.mma (the count includes its .2CTA form, listed again on its own; likewise UTMALDG), LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAPF = TMA tensor load / store / L2 prefetch, UTCBAR = tcgen05.commit")
tot = collections.Counter()
for f in order:
    c = counts[f]
    if not any(c[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF")):
        continue
    name = demangle(f)
    i = name.rfind(">(")
    name = (name[:i + 1] if i > 0 else name.split("(")[0]).replace("void ", "").replace("b200f::", "").replace("(bool)", "")
    print(f"{name[:160]}")
    print("    " + "  ".join(f"{o}={c[o]}" for o in OPS if c[o]) + f"  total={c['_total']}")
    tot.update({o: c[o] for o in OPS})
print("# library totals (tensor-engine kernels only): " + "  ".join(f"{o}={tot[o]}" for o in OPS if tot[o]))
print(f"# kernels in the library: {len(order)}; with tcgen05 / TMA instructions: {sum(1 for f in order if any(counts[f][o] for o in ('UTCHMMA','LDTM','UTMALDG')))}")
