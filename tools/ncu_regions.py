#!/usr/bin/env python
"""Warp instructions / stall samples / active lanes of an ncu capture aggregated over source-line REGIONS (functions of
dsb_seedcore.h by default).  usage: ncu_regions.py x.ncu-rep [file.h]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; src = sys.argv[2] if len(sys.argv) > 2 else "desamba_b200/csrc/dsb_seedcore.h"
base = src.split("/")[-1]
# function start lines
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
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 9 or cur is None or r[2] != "-": continue
    try: ln = int(r[0]); samp = int(float(r[6] or 0)); inst = int(float(r[7] or 0)); thr = int(float(r[8] or 0))
    except ValueError: continue
    k = region(cur, ln); a = agg.setdefault(k, [0, 0, 0]); a[0] += samp; a[1] += inst; a[2] += thr
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
print(f"# {rep}: {tot} stall samples, {toti} warp instructions")
print("# stall share | instruction share | active lanes | region")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] * 1000 < toti: continue
    print(f"{100*v[0]/tot:5.1f}% {100*v[1]/toti:5.1f}% {v[2]/max(v[1],1):5.1f}  {k}")
