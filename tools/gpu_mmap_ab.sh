#!/usr/bin/env bash
# Developer tool (gpurun): block input of the FASTQ reader by pread against mmap (DSB_FQ_MMAP), host pipeline alone and with one GPU.
cd "$(dirname "$0")/.."
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
FILES=$(for i in $(seq 32); do echo -n "/dev/shm/dsb_step.fq "; done)
echo "cores $(nproc)"
for rep in 1 2; do for m in 0 1; do
  echo -n "host only mmap=$m: "; DSB_FQ_MMAP=$m DSB_HOST_ONLY=1 DSB_VERBOSE=1 desamba_b200/bin/deSAMBA-b200 classify -f SAM -o /dev/shm/o.sam x $FILES 2>&1 | grep -E "host time" | cut -c30-120
done; done
for rep in 1 2; do for m in 0 1; do
  echo -n "-g 1 mmap=$m: "; DSB_FQ_MMAP=$m DSB_VERBOSE=1 desamba_b200/bin/deSAMBA-b200 classify -g 1 -f SAM -o /dev/shm/o.sam $IDX $FILES 2>&1 | grep -E "sequences processed|host time" | cut -c1-150 | tr '\n' ' '; echo; md5sum /dev/shm/o.sam | cut -c1-12
done; done
