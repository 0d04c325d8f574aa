#!/usr/bin/env bash
# Developer tool (gpurun): BASELINE.json configs[3] at reduced scale -- synthetic multi-genome reference (15 "species" x 4
# "strains" x 5 Mbp = 300 Mbp, SURVEY.md 8d cfg 4 recipe), index built ON THE BOX by the unmodified reference binaries in
# oracle/_ref (the 1.8 GB index cannot travel), then parity of the CUDA path against the oracle on samples and a bench run
# with the FM index / SA / bit tables resident in HBM but larger than L2.
set -uo pipefail
cd "$(dirname "$0")/.."
W=/tmp/big; mkdir -p $W gpurun_out
NS=${1:-15}
t0=$(date +%s)
python tools/gen_synth_ref.py $W/syn.fa $NS 4 5000000 20261023
bash oracle/build_index.sh $W/syn.fa $W/idx > $W/build.log 2>&1 || { tail -5 $W/build.log; exit 1; }
t1=$(date +%s); echo "reference + index built in $((t1 - t0)) s"; ls -la $W/idx
python - <<PY 2>&1 | tee gpurun_out/big_index_parity.log
import sys, time; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, oracle_binding as ob, desamba_b200 as dsb
IDX, FA = "$W/idx", "$W/syn.fa"
ix = dsb.Index(IDX, 0); ctx = dsb.Context(ix); orc = ob.Oracle(IDX)
print(f"index in HBM: {ix.hbm_bytes/1e9:.2f} GB, l_ek {ix.l_ek}")
for name, mode, n, err, seed in (("big_long10", "long", 1500, 0.10, 31), ("big_long30", "long", 1000, 0.30, 32), ("big_short", "short", 20000, 0.01, 33)):
    import subprocess, os
    p = f"$W/{name}.fq"
    subprocess.run([ob.SIMREADS, mode, FA, str(n), str(err), str(seed), p], check=True)
    names, seqs, _ = ob.read_fastq(p)
    cat, offs = ob.pack(seqs)
    orc.counters(reset=True)
    t = time.time(); rr_o, hits_o, mx_o = orc.classify(cat, offs); t_o = time.time() - t
    cnt_o = orc.counters()
    res = ctx.classify(cat, offs)
    bad = ob.compare_results(res.rr, res.hits, rr_o, hits_o, names, max_report=10**9)
    cnt_g = ctx.counters()
    same_cnt = all(cnt_g[k] == cnt_o[k] for k in ("n_bit0", "n_bit1", "n_prefix", "n_occ", "n_locate", "n_getref", "n_getref_bytes"))
    print(f"[{name}] reads={len(seqs)} hits={int(rr_o['n_hit'].sum())} secondaries={int((hits_o['primary']==2).sum())} oracle {t_o:.1f}s mismatching reads={len(bad)} counters_equal={same_cnt} kernels_ms={[round(x,2) for x in ctx.kernel_ms()]}")
    for b in bad[:5]: print("   ", b)
PY
python bench.py --index-dir $W/idx --fasta $W/syn.fa --reads-per-step 16384 --steps 6 --cpu-sample 4096 > gpurun_out/bench_syn300.json 2> gpurun_out/bench_syn300.err
tail -c 2500 gpurun_out/bench_syn300.json; tail -3 gpurun_out/bench_syn300.err
echo "total $(( $(date +%s) - t0 )) s"
