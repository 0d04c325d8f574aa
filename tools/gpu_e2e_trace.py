#!/usr/bin/env python
"""Developer tool: per-step wall-clock trace of concurrent dsb_classify_batch calls (one thread per context)."""
import os, sys, time, threading
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import oracle_binding as ob
import desamba_b200 as dsb
from desamba_b200.api import PinnedBuffer
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
NC = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
pageable = len(sys.argv) > 4 and sys.argv[4] == "pageable"      # inputs in ordinary memory: the staged H2D path the driver uses
quiet = len(sys.argv) > 5
ob.ensure_demo_index()
os.makedirs("/tmp/dsb_bench", exist_ok=True)
batches = []
for b in range(2):
    _, seqs = bench.make_batch(ob, n, 0, b, "/tmp/dsb_bench")
    cat, offs = ob.pack(seqs)
    pc, po = PinnedBuffer(len(cat)), PinnedBuffer(len(offs) * 8)
    pc.array[:] = cat; po.array.view(np.uint64)[:] = offs
    if pageable: batches.append((np.array(cat, copy=True), np.array(offs, dtype=np.uint64, copy=True), None, None, len(seqs)))
    else: batches.append((pc.array, po.array.view(np.uint64), pc, po, len(seqs)))
ix = dsb.Index(ob.DEMO_IDX, 0)
ctxs = [dsb.Context(ix) for _ in range(NC)]
outs = []
for t in range(NC):
    rr = PinnedBuffer(n * dsb.RR_DTYPE.itemsize); h = PinnedBuffer(24 * n * dsb.HIT_DTYPE.itemsize)
    outs.append((rr.array.view(dsb.RR_DTYPE), h.array.view(dsb.HIT_DTYPE), rr, h))
for t in range(NC):
    for b in range(2):
        ctxs[t].classify_into(batches[b][0], batches[b][1], outs[t][0], outs[t][1], 10**6)
log = []
go = threading.Event()
def worker(t):
    go.wait()
    for k in range(t, steps, NC):
        b = batches[k % 2]
        t0 = time.perf_counter()
        ctxs[t].classify_into(b[0], b[1], outs[t][0], outs[t][1], 10**6)
        t1 = time.perf_counter()
        log.append((t, k, k % 2, t0, t1, ctxs[t].kernel_ms(), ctxs[t].timeline(ctxs[0])))
ths = [threading.Thread(target=worker, args=(t,)) for t in range(NC)]
for th in ths: th.start()
time.sleep(0.2); ctxs[0].mark(0); ctxs[0].sync(); T0 = time.perf_counter(); go.set()
for th in ths: th.join()
T1 = time.perf_counter()
# device time line: which contexts have a kernel running, in 5-ms bins (letters = kernel group of dsb_batch_kernel_ms, '.' = the
# stream holds the batch but no kernel runs: upload or waiting for the host, ' ' = nothing)
if len(sys.argv) > 6:
    end = max(x[6][12] for x in log); nb = int(end / 5) + 1
    rows = [[" "] * nb for _ in range(NC)]
    for t, k, b, t0, t1, ms, tl in log:
        for i in range(int(max(tl[0], 0) / 5), min(nb, int(tl[1] / 5) + 1)): rows[t][i] = "."
        for g in range(11):
            if tl[2 + g] - tl[1 + g] < 0.05: continue
            for i in range(int(tl[1 + g] / 5), min(nb, int(tl[2 + g] / 5) + 1)): rows[t][i] = "PISCscscXHF"[g]
    for t in range(NC): print("ctx %d |%s|" % (t, "".join(rows[t])))
    busy = sum(1 for i in range(nb) if any(rows[t][i] not in " ." for t in range(NC)))
    print("5-ms bins with a kernel running: %d of %d" % (busy, nb))
for t, k, b, t0, t1, ms, tl in sorted(log, key=lambda x: x[3]):
    if quiet: break
    print(f"thread {t} step {k} batch {b}: start {1e3*(t0-T0):7.1f} ms  end {1e3*(t1-T0):7.1f} ms  wall {1e3*(t1-t0):6.1f}  kernels {sum(ms):6.1f} ms  {[round(x,1) for x in ms]}")
print("total", round(1e3 * (T1 - T0), 1), "ms for", steps, "steps of", n, "reads,", NC, "contexts,", "pageable" if pageable else "pinned", "inputs:", round(1e3 * (T1 - T0) / steps, 1), "ms per step;", {k: round(v, 3) for k, v in ctxs[0].host_seconds().items()})
