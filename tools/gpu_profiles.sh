#!/usr/bin/env bash
# Developer tool (gpurun): the end-of-round evidence under profiles/<TAG>/ -- the default bench line, the launch list of the
# same command (gpu__time_duration.sum), one ncu --set full capture of ALL launches of one full step (per-kernel summaries +
# traffic.json for bench.py's roofline.traffic), and single-kernel captures with per-line / per-function breakdowns.
# usage: tools/gpu_profiles.sh TAG
set -uo pipefail
cd "$(dirname "$0")/.."
TAG=${1:-r2}; O=gpurun_out/$TAG; mkdir -p $O
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json')); print('bench: value %.0f e2e %.0f driver %.0f Mbases/s, reference %.0f' % (d['value'], d['e2e']['value'], (d.get('driver') or {}).get('value', 0), (d.get('cpu_baseline') or {}).get('value', 0)))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-driver --in-flight 1 > $O/launches_bench.log 2>&1
python - <<PY
import csv, collections
tot = collections.Counter(); n = collections.Counter()
for r in csv.reader(open("$O/launches.csv")):
    if len(r) > 5 and r[-1].replace(".", "").replace(",", "").isdigit() and "gpu__time_duration" in r[-3]:
        k = r[4].split("(")[0]; v = float(r[-1].replace(",", "")); u = r[-2]
        tot[k] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1}.get(u, 1e-6); n[k] += 1
s = sum(tot.values()) or 1
print("launch list shares:", "  ".join(f"{k} {100*v/s:.1f}% ({n[k]}x)" for k, v in tot.most_common(9)))
PY
# one full step at the default size: skip the 2 index-preparation kernels + 3 warm-up steps (11 launches each), capture 11
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_(encode|islands|seed|chain|score|finalize)' --launch-skip 33 --launch-count 11 -o $O/step_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-driver --in-flight 1 > $O/step_full_ncu.log 2>&1
python tools/ncu_traffic.py $O/step_full.ncu-rep $O/traffic.json 65536 "long|viral-gs|65536" > /dev/null 2>&1
python tools/ncu_kernels.py $O/step_full.ncu-rep $O/kernels_summary.txt > /dev/null 2>&1
python -c "
import json; t=json.load(open('$O/traffic.json')); print('DRAM per step:', '  '.join('%s %.2f GB (%.1f ms)' % (k, v['dram_bytes']/1e9, v['ncu_duration_ms']) for k, v in t.items() if isinstance(v, dict)))"
rm -f $O/step_full.ncu-rep                       # (55 MB; gpurun brings back at most 64 MiB: the summaries above are what is kept)
for K in k_seed k_score k_chain; do
	tools/gpu_ncu.sh $TAG/one $K 0 16384 > /dev/null 2>&1
	[ $K = k_seed ] || rm -f $O/one_$K.ncu-rep
done
du -sh $O; ls $O | head -40
