#!/usr/bin/env bash
# Developer tool (gpurun): BASELINE configs[2] at its stated size through the product -- 20 M x 150 bp reads (20 copies of a 1 M-read
# simulated FASTQ in /dev/shm) through `deSAMBA-b200 classify -g 1 -f SAM`, and the unmodified reference (-t <cores>) on one copy.
set -uo pipefail
cd "$(dirname "$0")/.."
python - <<'PY'
import os, sys, subprocess
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_binding as ob
ob.ensure_demo_index()
p = "/dev/shm/dsb_short1m.fq"
if not os.path.exists(p):
    subprocess.run([ob.SIMREADS, "short", ob.DEMO_FA, "1000000", "0.01", "20261031", p], check=True)
print("1 M x 150 bp FASTQ:", os.path.getsize(p) >> 20, "MiB")
PY
IDX=oracle/_ref/demo/idx
FILES=$(for i in $(seq 20); do echo -n "/dev/shm/dsb_short1m.fq "; done)
DSB_VERBOSE=1 desamba_b200/bin/deSAMBA-b200 classify -g 1 -f SAM -o /dev/shm/o.sam $IDX $FILES 2> /tmp/drv.err
echo "== driver, 20 M x 150 bp: $(grep -E 'sequences processed' /tmp/drv.err)"
grep -E "host time|GPU calls:| at +[0-9.]+ s" /tmp/drv.err | sed 's/^/     /'
grep -c -v "^@" /dev/shm/o.sam | sed 's/^/     SAM lines: /'
awk '$2 != 4' /dev/shm/o.sam | wc -l | sed 's/^/     classified lines: /'
oracle/_ref/deSAMBA_stock classify -t $(nproc) -f SAM -o /dev/shm/o_ref.sam $IDX /dev/shm/dsb_short1m.fq 2> /tmp/ref.err
echo "== reference -t $(nproc), 1 M x 150 bp: $(grep -E 'sequences processed|processed in' /tmp/ref.err | tail -1)"
rm -f /dev/shm/o.sam /dev/shm/o_ref.sam /dev/shm/o_head.sam
