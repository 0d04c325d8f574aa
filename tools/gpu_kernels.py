#!/usr/bin/env python
"""Developer tool (gpurun): the 11 kernel times and the work-list sizes of one bench batch run alone (no other batch in flight).
usage: gpu_kernels.py [reads] [workload long|short]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_binding as ob, bench, desamba_b200 as dsb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
wl = sys.argv[2] if len(sys.argv) > 2 else "long"
ob.ensure_demo_index()
os.makedirs("/tmp/dsb_bench", exist_ok=True)
_, seqs = bench.make_batch(ob, n, 0, 0, "/tmp/dsb_bench", wl)
cat, offs = ob.pack(seqs)
ix = dsb.Index(ob.DEMO_IDX, 0); ctx = dsb.Context(ix)
ctx.upload(cat, offs)
for _ in range(3):
    ctx.run(10**6); ctx.sync()
ms = np.zeros(11)
for _ in range(3):
    ctx.run(10**6); ctx.sync(); ms += ctx.kernel_ms()
ms /= 3
print(f"{n} {wl} reads, {int(offs[-1])/1e6:.0f} Mbases, DSB_HEAVY_BLOCKS={os.environ.get('DSB_HEAVY_BLOCKS','-')}: total {ms.sum():.1f} ms")
print("  " + "  ".join(f"{k} {v:.2f}" for k, v in zip(dsb.KERNEL_NAMES, ms)))
print("  work:", ctx.work())
