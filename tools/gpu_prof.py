#!/usr/bin/env python
"""Developer tool: per-read / per-phase device time of k_classify on a synthetic batch (gpurun).  usage: gpu_prof.py [n_reads]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import oracle_binding as ob
import desamba_b200 as dsb
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ob.ensure_demo_index()
_, seqs = bench.make_batch(ob, n, 0, 0, "/tmp")
cat, offs = ob.pack(seqs)
ix = dsb.Index(ob.DEMO_IDX, 0); ctx = dsb.Context(ix, warps_per_sm=int(os.environ.get("PROF_WARPS", "0")) or None)
ctx.upload(cat, offs)
for _ in range(2):
    ctx.run(10**6); ctx.sync()
print("kernel ms", dict(zip(dsb.KERNEL_NAMES, [round(x, 2) for x in ctx.kernel_ms()])))
res = ctx.download()
P = ctx.profile().astype(np.float64) * 1024 / 1.965e6     # ms at 1965 MHz
names = ["fast", "chain", "slow", "match", "middle", "right", "left", "total"]
print("sum over reads (warp-ms):", {k: round(float(P[:, i].sum()), 1) for i, k in enumerate(names)})
print("mean per read (ms):", {k: round(float(P[:, i].mean()), 3) for i, k in enumerate(names)})
order = np.argsort(-P[:, 7])
print("worst reads:")
for i in order[:int(os.environ.get("PROF_WORST", "12"))]:
    print(f"  read {i} len={len(seqs[i])} n_anc={res.rr['n_anchor'][i]} n_hit={res.rr['n_hit'][i]} fast={res.rr['fast_classify'][i]} " + " ".join(f"{k}={P[i, j]:.1f}" for j, k in enumerate(names)))
for j, k in enumerate(names[:7]):
    i = int(np.argmax(P[:, j]))
    print(f"max {k}: read {i} len={len(seqs[i])} n_anc={res.rr['n_anchor'][i]} " + " ".join(f"{kk}={P[i, jj]:.1f}" for jj, kk in enumerate(names)))
q = np.percentile(P[:, 7], [50, 90, 99, 99.9, 100])
print("total ms percentiles 50/90/99/99.9/100:", [round(float(x), 2) for x in q])
lens = np.array([len(s) for s in seqs])
print("ideal kernel ms if perfectly balanced over", 148 * 16, "warps:", round(float(P[:, 7].sum()) / (148 * 16), 2))
