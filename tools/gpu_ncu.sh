#!/usr/bin/env bash
# Developer tool (gpurun): one ncu --set full capture of a kernel of the bench workload + the summaries kept under profiles/.
# usage: tools/gpu_ncu.sh TAG KERNEL [launch-skip] [reads]
set -uo pipefail
cd "$(dirname "$0")/.."
TAG=$1; K=$2; SKIP=${3:-0}; READS=${4:-16384}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"^$K\$" --launch-skip $SKIP --launch-count 1 -o gpurun_out/${TAG}_$K -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-driver --in-flight 1 --reads-per-step $READS > gpurun_out/${TAG}_${K}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_$K.ncu-rep gpurun_out/${TAG}_${K}_raw.txt > /dev/null 2>&1
python tools/ncu_hot_lines.py gpurun_out/${TAG}_$K.ncu-rep 80 > gpurun_out/${TAG}_${K}_hot_lines.txt 2>&1
for f in dsb_classify.cuh dsb_seedcore.h dsb_seed.cuh dsb_batch.cu; do python tools/ncu_regions.py gpurun_out/${TAG}_$K.ncu-rep desamba_b200/csrc/$f > gpurun_out/${TAG}_${K}_regions_$f.txt 2>&1; done
head -2 gpurun_out/${TAG}_${K}_hot_lines.txt
