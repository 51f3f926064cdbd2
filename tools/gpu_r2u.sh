#!/bin/bash
# rank-sort unique_rows: tests, bench, kernel time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== unique tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -k "unique or filter_coarse or select_anchor" > gpurun_out/u_uniq.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/u_uniq.log
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/u_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/u_tests.log
echo "=== bench 20"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-modes tf32 > gpurun_out/u_bench_20.json 2> gpurun_out/u_bench_20.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"unique|select_anchor|proposals" -s 8 -c 16 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline --e2e-modes tf32 > gpurun_out/u_ncu_launch.log 2>&1; echo "rc=$?"
grep -E "unique" gpurun_out/u_launches.csv | awk -F'","' '{print substr($5,1,30), $NF}' | head -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/u_bench_20.json').read().strip().splitlines()[-1])
print('value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'], d['config']['step_ms_quantiles'])
print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
