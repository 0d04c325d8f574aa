#!/usr/bin/env bash
# Developer tool (gpurun): kernel times (one batch alone) and the bench value of library variants built under build_variants/
# (lib_<name>.so, e.g. other -DSEED_WARPS_PER_SM / -DCLASSIFY_WARPS_PER_SM) against the default build.
# usage: tools/gpu_variants.sh "name1 name2 ..."
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default $1; do
	if [ $v = default ]; then unset DSB_LIB; else export DSB_LIB=$PWD/build_variants/lib_$v.so; fi
	echo "== $v"
	timeout 300 python tools/gpu_kernels.py 65536 2>&1 | head -2
	timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-driver > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
	python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_$v.json")); print("  bench: value %.0f Mbases/s, step %.1f ms" % (d["value"], d["ms_per_step"]))
except Exception as e: print("  bench failed", e)
PY
done
