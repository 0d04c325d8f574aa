#!/usr/bin/env bash
# Developer tool (gpurun): wall-clock of the C driver `deSAMBA-b200 classify` vs the reference `deSAMBA classify -t <cores>`
# on the same FASTQ (FASTQ parse + classify + text output), and a byte comparison of the SAM files.
set -uo pipefail
cd "$(dirname "$0")/.."
N=${1:-32768}
python - <<PY
import sys; sys.path.insert(0, "tests")
import oracle_binding as ob, bench, os
ob.ensure_demo_index()
os.makedirs("/tmp/dsb_bench", exist_ok=True)
paths, seqs = bench.make_batch(ob, $N, 0, 0, "/tmp/dsb_bench")
bench.write_fastq("/tmp/dsb_bench/driver.fq", seqs)
print("reads", len(seqs), "bases", sum(map(len, seqs)))
PY
IDX=oracle/_ref/demo/idx; FQ=/tmp/dsb_bench/driver.fq
C=$(nproc)
now() { date +%s.%N; }
for i in 1 2; do t0=$(now); oracle/_ref/deSAMBA_stock classify -t $C -f SAM $IDX $FQ -o /tmp/dsb_bench/ref.sam 2>&1 | grep -E "processed"; t1=$(now); echo "reference -t $C: wall $(echo "$t1 - $t0" | bc) s"; done
for i in 1 2; do t0=$(now); desamba_b200/bin/deSAMBA-b200 classify -f SAM $IDX $FQ -o /tmp/dsb_bench/gpu.sam 2>&1 | grep -E "processed|GPUs|error"; t1=$(now); echo "deSAMBA-b200: wall $(echo "$t1 - $t0" | bc) s (includes index load to HBM)"; done
oracle/_ref/deSAMBA_zero classify -t 1 -f SAM $IDX $FQ -o /tmp/dsb_bench/zero.sam 2>/dev/null
cmp /tmp/dsb_bench/gpu.sam /tmp/dsb_bench/zero.sam && echo "SAM identical to the parity oracle (zero-init reference, -t 1)"
echo "lines differing from stock -t $C: $(diff /tmp/dsb_bench/gpu.sam /tmp/dsb_bench/ref.sam | grep -c '^<') of $(wc -l < /tmp/dsb_bench/gpu.sam)"
