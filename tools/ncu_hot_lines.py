#!/usr/bin/env python
"""Hottest CUDA source lines of an ncu capture (needs -lineinfo + --import-source on).  usage: ncu_hot_lines.py x.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 9 or cur is None or r[2] != "-": continue
    try: ln = int(r[0]); samp = int(float(r[6] or 0)); inst = int(float(r[7] or 0)); thr = int(float(r[8] or 0))
    except ValueError: continue
    if samp or inst: agg[(cur, ln)] = (samp, inst, thr, r[1].strip()[:120])
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
print(f"# {rep}: {tot} stall samples, {toti} warp instructions, {sum(v[2] for v in agg.values())/toti:.1f} active lanes per instruction")
print("# share of stall samples | share of warp instructions | active lanes | file:line | source")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{100*v[0]/tot:5.1f}% {100*v[1]/toti:5.1f}% {v[2]/max(v[1],1):5.1f}  {k[0]}:{k[1]}  {v[3]}")
