"""Developer tool (gpurun): the very-long-read case of tests/test_gpu_parity.py with per-read error codes printed."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, oracle_binding as ob, desamba_b200 as dsb
ob.build(); idx = ob.ensure_demo_index()
ix = dsb.Index(idx, 0); orc = ob.Oracle(idx)
_, seqs, _ = ob.read_fastq(ob.sim_set("long10", "long", 300, 0.10, 20261020), 120)
reads = [b"".join(seqs[0:20]), b"".join(seqs[20:70]), b"".join(seqs[70:120])[:400000], seqs[3]]
cat, offs = ob.pack(reads)
rr_o, hits_o, mx_o = orc.classify(cat, offs)
print("oracle n_anchor", rr_o["n_anchor"], "n_hit", rr_o["n_hit"], "len", [len(r) for r in reads])
for kw in ({}, {"max_anchors": 65536, "max_matches": 65536}):
    ctx = dsb.Context(ix, **kw)
    try:
        res = ctx.classify(cat, offs)
        bad = ob.compare_results(res.rr, res.hits, rr_o, hits_o, None, max_report=10)
        print(kw, "ok, mismatches:", len(bad), bad[:5], "kernel ms", [round(x, 2) for x in ctx.kernel_ms()])
    except dsb.api.DsbError as e:
        print(kw, "ERROR", e, "per-read error codes:", getattr(e, "result", None) and e.result.rr["error"], "n_anchor", getattr(e, "result", None) and e.result.rr["n_anchor"])
    ctx.close()
