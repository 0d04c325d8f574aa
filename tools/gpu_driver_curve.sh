#!/usr/bin/env bash
# Developer tool (gpurun --gpus N): the driver's `-g k` curve, k = 1, 2, 4, 8 <= N, on 16 k copies of the bench step's FASTQ in
# /dev/shm (its own "processed in" interval and host timers), and the host pipeline alone (DSB_HOST_ONLY=1) on the same input.
# usage: tools/gpu_driver_curve.sh N
set -uo pipefail
cd "$(dirname "$0")/.."
N=${1:-2}
echo "cores: $(nproc)"
python - <<'PY'
import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_binding as ob, bench
ob.ensure_demo_index()
os.makedirs("/tmp/dsb_bench", exist_ok=True)
fq = "/dev/shm/dsb_step.fq"
if not os.path.exists(fq):
    _, seqs = bench.make_batch(ob, 65536, 0, 0, "/tmp/dsb_bench")
    bench.write_fastq(fq, seqs)
PY
IDX=oracle/_ref/demo/idx
for k in 1 2 4 8; do
	[ $k -le $N ] || continue
	FILES=$(for i in $(seq $((16 * k))); do echo -n "/dev/shm/dsb_step.fq "; done)
	DSB_VERBOSE=1 desamba_b200/bin/deSAMBA-b200 classify -g $k -f SAM -o /dev/shm/dsb_out.sam $IDX $FILES 2> /tmp/drv.err
	echo "== driver -g $k, $((16 * k)) step files ($((8 * k)) Gbases): $(grep -E 'sequences processed' /tmp/drv.err)"
	grep -E "host time|GPU calls:|GPUs:" /tmp/drv.err | sed 's/^/     /'
	md5sum /dev/shm/dsb_out.sam | cut -c1-12
done
FILES=$(for i in $(seq $((16 * N))); do echo -n "/dev/shm/dsb_step.fq "; done)
DSB_HOST_ONLY=1 DSB_VERBOSE=1 desamba_b200/bin/deSAMBA-b200 classify -f SAM -o /dev/shm/dsb_out.sam $IDX $FILES 2> /tmp/drv.err
echo "== host pipeline only, $((16 * N)) step files: $(grep -E 'sequences processed' /tmp/drv.err)"
grep -E "host time" /tmp/drv.err | sed 's/^/     /'
