#!/usr/bin/env bash
# Developer tool (gpurun --gpus N): the multi-GPU paths -- driver `-g k` for k = 1..N on 16 k copies of the bench step's FASTQ
# (its own "processed in" interval, index load + device-to-device copies), the 2-GPU driver test, and bench.py under torchrun.
# usage: tools/gpu_multi.sh N [TAG]
set -uo pipefail
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-multi}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; nvidia-smi topo -m 2>/dev/null | head -12
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or clone" 2>&1 | tail -2
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
	grep -E "host time|GPU calls:|GPUs:| at +[0-9.]+ s" /tmp/drv.err | sed 's/^/     /'
	md5sum /dev/shm/dsb_out.sam | cut -c1-12
done
if [ $N -gt 1 ]; then
	python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err
	python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_$N.json')); print('bench N=$N: value %.0f e2e %.0f Mbases/s' % (d['value'], d['e2e']['value'])); print(' driver:', {k: d['driver'].get(k) for k in ('gpus','processed_s','wall_s','index_load_s','index_clone_s','value')})"
	tail -2 gpurun_out/${TAG}_bench_$N.err
fi
