#!/usr/bin/env bash
# Developer tool (gpurun): `deSAMBA-b200 classify` on N copies of the bench step's FASTQ (in /dev/shm) under several settings:
# the driver's own "processed in" interval and where its host time goes.   usage: tools/gpu_driver_ab.sh [copies] [gpus]
set -uo pipefail
cd "$(dirname "$0")/.."
N=${1:-12}; G=${2:-1}
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
print("step FASTQ:", os.path.getsize(fq) >> 20, "MiB")
PY
IDX=oracle/_ref/demo/idx
FILES=$(for i in $(seq $N); do echo -n "/dev/shm/dsb_step.fq "; done)
run() { # label env... -- args
	local label=$1; shift
	local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
	env DSB_VERBOSE=1 "${envs[@]}" desamba_b200/bin/deSAMBA-b200 classify -g $G "$@" -f SAM -o /dev/shm/dsb_out.sam $IDX $FILES 2> /tmp/drv.err
	echo "== $label: $(grep -E 'sequences processed' /tmp/drv.err)"
	grep -E "host time|batch buffers|index:" /tmp/drv.err | sed 's/^/     /'
}
run "default (-c 3 -M 512)" X=1 --
run "-c 4" X=1 -- -c 4
run "-c 4 -P 12" X=1 -- -c 4 -P 12
run "-c 6 -P 12" X=1 -- -c 6 -P 12
run "-c 4 -P 12 -M 256" X=1 -- -c 4 -P 12 -M 256
md5sum /dev/shm/dsb_out.sam | cut -c1-12
