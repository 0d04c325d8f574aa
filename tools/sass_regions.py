#!/usr/bin/env python
"""Static SASS size of a kernel by source region (function of dsb_seedcore.h / file): nvdisasm --print-line-info output.
usage: sass_regions.py disasm.txt kernel_name [src.h]"""
import re, sys
dis, kern = sys.argv[1], sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 else "desamba_b200/csrc/dsb_seedcore.h"
base = src.split("/")[-1]
starts = []
for i, l in enumerate(open(src), 1):
    m = re.match(r"^(?:SC_HDN?|static|__device__|SC_HD)\b.*?\b(\w+)\s*\(", l)
    if m and not l.startswith("\t"): starts.append((i, m.group(1)))
def region(f, ln):
    if f != base: return f
    name = "?"
    for s, n in starts:
        if s <= ln: name = n
        else: break
    return name
inside = False; cur = ("?", 0); agg = {}; total = 0
for l in open(dis):
    if l.startswith(".text."): inside = kern in l
    if ".section" in l and ".text." in l: inside = kern in l
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        k = region(*cur); agg[k] = agg.get(k, 0) + 1; total += 1
print(f"# {kern}: {total} SASS instructions = {total * 16 / 1024:.0f} KB")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:40]: print(f"{v:6d}  {k}")
