#!/usr/bin/env bash
# Developer tool (gpurun): wall-clock of the C driver `deSAMBA-b200 classify` vs the reference `deSAMBA classify -t <cores>`
# on the same FASTQ (FASTQ parse + classify + text output), and a byte comparison of the SAM files.
# usage: gpu_driver_bench.sh [reads=262144] [ref_reads=32768]
set -uo pipefail
cd "$(dirname "$0")/.."
N=${1:-262144}; NREF=${2:-32768}
python - <<PY
import sys; sys.path.insert(0, "tests")
import oracle_binding as ob
ob.build(); ob.ensure_demo_index()
PY
D=/tmp/dsb_bench; mkdir -p $D
FA=oracle/_ref/demo/viral-gs.fa; IDX=oracle/_ref/demo/idx; SIM=desamba_b200/bin/simreads
$SIM long $FA $((N / 2)) 0.10 20261030 $D/a.fq; $SIM long $FA $((N / 2)) 0.30 20261031 $D/b.fq
cat $D/a.fq $D/b.fq > $D/driver.fq; rm -f $D/a.fq $D/b.fq
head -n $((NREF * 4)) $D/driver.fq > $D/driver_ref.fq
ls -la $D/driver.fq $D/driver_ref.fq
C=$(nproc)
now() { date +%s.%N; }
el() { python -c "print(f'{$2 - $1:.2f}')"; }
for i in 1 2; do t0=$(now); oracle/_ref/deSAMBA_stock classify -t $C -f SAM $IDX $D/driver_ref.fq -o $D/ref.sam 2>&1 | grep -E "processed"; t1=$(now); echo "reference -t $C on the first $NREF reads: wall $(el $t0 $t1) s"; done
for P in 0 8; do for i in 1 2; do t0=$(now); desamba_b200/bin/deSAMBA-b200 classify -f SAM -P $P $IDX $D/driver.fq -o $D/gpu.sam 2>&1 | grep -E "processed|GPUs|error"; t1=$(now); echo "deSAMBA-b200 -P $P on $N reads: wall $(el $t0 $t1) s (includes index load to HBM)"; done; done
desamba_b200/bin/deSAMBA-b200 classify -f SAM $IDX $D/driver_ref.fq -o $D/gpu_ref.sam 2>/dev/null
oracle/_ref/deSAMBA_zero classify -t 1 -f SAM $IDX $D/driver_ref.fq -o $D/zero.sam 2>/dev/null
cmp $D/gpu_ref.sam $D/zero.sam && echo "SAM of the first $NREF reads identical to the parity oracle (zero-init reference, -t 1)"
echo "lines differing from stock -t $C: $(diff $D/gpu_ref.sam $D/ref.sam | grep -c '^<') of $(wc -l < $D/gpu_ref.sam)"
