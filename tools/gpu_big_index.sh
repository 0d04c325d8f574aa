#!/usr/bin/env bash
# Developer tool (gpurun): BASELINE.json configs[3] / [4] at the scale the GPU-minute budget allows -- a synthetic multi-genome
# reference (NS "species" x 4 "strains" x 5 Mbp, SURVEY.md 8d cfg 4 recipe; NS = 20 -> 400 Mbp, ~276 M distinct 31-mers -> the
# builder picks the l_ek = 17 / 31-bit class, FM index and SA of ~0.6 GB each: resident in HBM, far beyond L2), index built ON THE
# BOX by the unmodified reference binaries of oracle/_ref (it cannot travel: ~3 GB), then
#   1. parity of the CUDA path against the oracle on long-10 %, long-30 % and short samples, all algorithmic counters included
#   2. configs[4]: mixed long / short reads, DES_FULL, driver text against `deSAMBA_zero classify -t 1 -f DES_FULL` run here
#   3. bench.py on this index (roofline of the seeding kernel with the FM index in HBM, driver leg, CPU reference)
#   4. one ncu --set full capture of k_seed on this index
# usage: tools/gpu_big_index.sh [NS=20] [TAG=big]
set -uo pipefail
cd "$(dirname "$0")/.."
NS=${1:-20}; TAG=${2:-big}
W=/tmp/big; mkdir -p $W gpurun_out
t0=$(date +%s)
python tools/gen_synth_ref.py $W/syn.fa $NS 4 5000000 20261023
bash oracle/build_index.sh $W/syn.fa $W/index > $W/build.log 2>&1 || { tail -5 $W/build.log; exit 1; }
t1=$(date +%s); echo "reference ($(( $(stat -c %s $W/syn.fa) / 1000000 )) MB) + index built in $((t1 - t0)) s"; ls -la $W/index
grep -E "l_e_kmer|kmer" $W/build.log | tail -3
python - <<PY 2>&1 | tee gpurun_out/${TAG}_parity.log
import sys, time, subprocess, os; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, oracle_binding as ob, desamba_b200 as dsb
IDX, FA = "$W/index", "$W/syn.fa"
ix = dsb.Index(IDX, 0); ctx = dsb.Context(ix); orc = ob.Oracle(IDX)
print(f"index in HBM: {ix.hbm_bytes/1e9:.2f} GB, l_ek {ix.l_ek}, {len(ix.ref_names)} reference sequences")
for name, mode, n, err, seed in (("big_long10", "long", 1500, 0.10, 31), ("big_long30", "long", 1000, 0.30, 32), ("big_short", "short", 20000, 0.01, 33)):
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
    keys = ("n_bit0", "n_bit1", "n_prefix", "n_occ", "n_locate", "n_getref", "n_getref_bytes")
    same_cnt = all(cnt_g[k] == cnt_o[k] for k in keys)
    print(f"[{name}] reads={len(seqs)} hits={int(rr_o['n_hit'].sum())} secondaries={int((hits_o['primary']==2).sum())} oracle {t_o:.1f}s "
          f"mismatching reads={len(bad)} counters_equal={same_cnt} retries={ctx.retries()} kernels_ms={[round(x,2) for x in ctx.kernel_ms()]}")
    for b in bad[:5]: print("   ", b)
# configs[4]: mixed long / short, DES_FULL, the unmodified reference (zero-init build, -t 1) as the judge of the driver's text
p = "$W/big_mixed.fq"
subprocess.run([ob.SIMREADS, "mixed", FA, "400", "4000", "34", p], check=True)
ref = os.path.join(ob.REF_DIR, "deSAMBA_zero")
t = time.time(); subprocess.run([ref, "classify", "-t", "1", "-f", "DES_FULL", "-o", "$W/mixed.ref", IDX, p], check=True, capture_output=True); t_r = time.time() - t
drv = os.path.join("desamba_b200", "bin", "deSAMBA-b200")
t = time.time(); subprocess.run([drv, "classify", "-f", "DES_FULL", "-B", "997", "-o", "$W/mixed.gpu", IDX, p], check=True, capture_output=True); t_g = time.time() - t
a, b = open("$W/mixed.ref", "rb").read(), open("$W/mixed.gpu", "rb").read()
ra, rb = a.split(b"\n\n"), b.split(b"\n\n")
diff = sum(1 for x, y in zip(ra, rb) if x != y) + abs(len(ra) - len(rb))
print(f"[big_mixed DES_FULL] 4400 reads: reference -t 1 {t_r:.1f} s, driver {t_g:.1f} s (index load included), identical text: {a == b}, differing read blocks: {diff}, SEC lines {a.count(b' SEC ')}")
PY
python bench.py --index-dir $W/index --fasta $W/syn.fa --reads-per-step 32768 --steps 8 --warmup 3 --cpu-sample 4096 --driver-files 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
k = d["roofline"]["kernels"]
print("bench on the big index: value %.0f Mbases/s, e2e %.0f, step %.1f ms (sequential %.1f)" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["ms_per_step_sequential"]))
print("  " + "  ".join(f"{n} {v['ms_per_step']:.2f}" for n, v in k.items()))
r = d["roofline"]; print("  k_seed achieved %.0f GB/s (8d bytes) / %.0f (own layout), frac of HBM peak %.3f, of the 128-B gather rate %.3f" % (r["achieved"], r["achieved_own_layout"], r["frac"], r["frac_of_random_gather"] or 0))
print("  driver:", d.get("driver")); print("  cpu:", d.get("cpu_baseline"))
PY
tail -2 gpurun_out/${TAG}_bench.err
timeout 500 ncu --set full --import-source on --clock-control none -k regex:'^k_seed$' --launch-count 1 -o gpurun_out/${TAG}_k_seed -f python bench.py --index-dir $W/index --fasta $W/syn.fa --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-driver --in-flight 1 --reads-per-step 16384 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_k_seed.ncu-rep gpurun_out/${TAG}_k_seed_raw.txt > /dev/null 2>&1
python tools/ncu_regions.py gpurun_out/${TAG}_k_seed.ncu-rep > gpurun_out/${TAG}_k_seed_regions.txt 2>&1
grep -E "dram__bytes|duration|lts__t_sector_hit|thread_inst_executed_per|issue_active|long_scoreboard_per|no_instruction_per" gpurun_out/${TAG}_k_seed_raw.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:'^k_encode_probe$' --launch-count 1 -o gpurun_out/${TAG}_k_encode_probe -f python bench.py --index-dir $W/index --fasta $W/syn.fa --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-driver --in-flight 1 --reads-per-step 16384 > gpurun_out/${TAG}_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_k_encode_probe.ncu-rep gpurun_out/${TAG}_k_encode_probe_raw.txt > /dev/null 2>&1
grep -E "dram__bytes|duration|dram__throughput" gpurun_out/${TAG}_k_encode_probe_raw.txt
echo "total $(( $(date +%s) - t0 )) s"
