#!/usr/bin/env bash
# Developer tool (gpurun): k_seed time of the bench workload under several state-vote policies (DSB_SEED_POLICY=policy,min),
# then one ncu --set full capture of k_seed with the default.   usage: tools/gpu_seed_ab.sh TAG "3,3 3,2 2,4 ..."
set -uo pipefail
cd "$(dirname "$0")/.."
TAG=${1:-ab}; shift
mkdir -p gpurun_out
for pol in ${1:-"3,3"}; do
	DSB_SEED_POLICY=$pol timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${pol/,/_}.json 2> gpurun_out/${TAG}_bench.err
	python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${pol/,/_}.json"))
k = d["roofline"]["kernels"]
print("policy $pol: value %.0f Mbases/s  step %.1f ms (sequential %.1f)  " % (d["value"], d["ms_per_step"], d["ms_per_step_sequential"]) + "  ".join(f"{n} {v['ms_per_step']:.2f}" for n, v in k.items()))
PY
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_seed --launch-count 1 -o gpurun_out/${TAG}_k_seed -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --in-flight 1 --reads-per-step 16384 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_k_seed.ncu-rep gpurun_out/${TAG}_k_seed_raw.txt > /dev/null 2>&1
python tools/ncu_hot_lines.py gpurun_out/${TAG}_k_seed.ncu-rep 60 > gpurun_out/${TAG}_k_seed_hot_lines.txt 2>&1
python tools/ncu_regions.py gpurun_out/${TAG}_k_seed.ncu-rep > gpurun_out/${TAG}_k_seed_regions.txt 2>&1
head -3 gpurun_out/${TAG}_k_seed_hot_lines.txt
