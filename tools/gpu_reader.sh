#!/usr/bin/env bash
# Developer tool (gpurun): where the host time of `deSAMBA-b200 classify` goes -- cores of the box, the driver under several
# reader thread counts (-P), and the host pipeline's own ceiling (DSB_HOST_ONLY=1: no GPU calls per batch).
# usage: tools/gpu_reader.sh [copies] [gpus] ["P values"]
set -uo pipefail
cd "$(dirname "$0")/.."
N=${1:-16}; G=${2:-1}; PS=${3:-"8 16 32 48"}
echo "cores: $(nproc)  $(lscpu | grep -E 'Model name|Socket|NUMA node\(s\)' | tr -s ' ' | tr '\n' ';')"
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
	local t0=$(date +%s%N)
	env DSB_VERBOSE=1 "${envs[@]}" desamba_b200/bin/deSAMBA-b200 classify -g $G "$@" -f SAM -o /dev/shm/dsb_out.sam $IDX $FILES 2> /tmp/drv.err
	echo "== $label: $(grep -E 'sequences processed' /tmp/drv.err)  wall $(( ($(date +%s%N) - t0) / 1000000 )) ms"
	grep -E "host time|GPU calls:| at +[0-9.]+ s" /tmp/drv.err | sed 's/^/     /'
}
run "default" X=1 --
md5sum /dev/shm/dsb_out.sam | cut -c1-12
for p in $PS; do run "-P $p" X=1 -- -P $p; done
