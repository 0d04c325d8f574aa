#!/usr/bin/env python
"""Developer tool (runs on the GPU box via gpurun): stage-by-stage parity of the CUDA path against the oracle.
usage: python tools/gpu_check.py [set ...]   sets: demo long10 long30 short1 mixed   -> gpurun_out/gpu_check.log"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_binding as ob
import desamba_b200 as dsb

SETS = {"demo": None, "long10": ("long", 300, 0.10, 20261020), "long30": ("long", 300, 0.30, 20261021),
        "short1": ("short", 5000, 0.01, 20261022), "mixed": ("mixed", (150, 1500), 0, 20261024)}


def main():
    want = sys.argv[1:] or ["demo", "long10", "long30", "short1", "mixed"]
    os.makedirs("gpurun_out", exist_ok=True)
    log = open("gpurun_out/gpu_check.log", "w")
    def say(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True); log.write(s + "\n"); log.flush()
    ob.ensure_demo_index()
    t0 = time.time()
    ix = dsb.Index(ob.DEMO_IDX, 0)
    say(f"index loaded in {time.time()-t0:.1f}s, HBM {ix.hbm_bytes/1e6:.0f} MB, l_ek {ix.l_ek}")
    ctx = dsb.Context(ix)
    orc = ob.Oracle(ob.DEMO_IDX)
    ok_all = True
    for name in want:
        path = ob.DEMO_FQ if name == "demo" else ob.sim_set(name, *SETS[name])
        names, seqs, _ = ob.read_fastq(path)
        cat, offs = ob.pack(seqs)
        orc.counters(reset=True)
        t = time.time(); rr_o, hits_o, mx_o = orc.classify(cat, offs); t_o = time.time() - t
        cnt_o = orc.counters()
        t = time.time(); res = None
        try:
            res = ctx.classify(cat, offs)
        except dsb.DsbError as e:
            say(f"[{name}] GPU error: {e}")
            if e.code != -5: ok_all = False; continue
            res = ctx.download() if False else None
        t_g = time.time() - t
        if res is None: ok_all = False; continue
        # stage 1: seeds
        n_seed_bad = 0
        for i in range(min(len(seqs), 400)):
            for s in (0, 1):
                sg, tg = ctx.seeds(i, s); so, to = orc.seeds(seqs[i], s)
                if len(seqs[i]) < 40: continue
                if tg != to or sg.tobytes() != so.tobytes():
                    n_seed_bad += 1
                    if n_seed_bad <= 3:
                        say(f"[{name}] seeds differ read {i} strand {s}: gpu n={len(sg)} ts={tg} oracle n={len(so)} ts={to}")
                        for k in range(min(len(sg), len(so))):
                            if sg[k].tobytes() != so[k].tobytes(): say(f"    first diff at {k}: gpu={sg[k]} oracle={so[k]}"); break
        bad = ob.compare_results(res.rr, res.hits, rr_o, hits_o, names, max_report=8)
        n_bad_total = len(ob.compare_results(res.rr, res.hits, rr_o, hits_o, names, max_report=10**9))
        cnt_g = ctx.counters()
        say(f"[{name}] reads={len(seqs)} bases={len(cat)} oracle {t_o:.2f}s gpu(e2e) {t_g:.3f}s kernels_ms={['%.2f' % x for x in ctx.kernel_ms()]} sum={sum(ctx.kernel_ms()):.2f} "
            f"seed_mismatch={n_seed_bad} read_mismatch={n_bad_total} max_read_l gpu={res.max_read_l} oracle={mx_o}")
        say(f"    counters gpu: " + " ".join(f"{k}={cnt_g[k]}" for k in ("n_bit0", "n_bit1", "n_prefix", "n_occ", "n_locate", "n_getref", "n_getref_bytes")))
        say(f"    counters orc: " + " ".join(f"{k}={cnt_o[k]}" for k in ("n_bit0", "n_bit1", "n_prefix", "n_occ", "n_locate", "n_getref", "n_getref_bytes")))
        for b in bad: say("    " + b)
        if n_seed_bad or n_bad_total or res.max_read_l != mx_o: ok_all = False
    say("ALL OK" if ok_all else "MISMATCHES")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
