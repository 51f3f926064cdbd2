#!/bin/bash
# Round-2 run K: unique_rows with warp-local bitonic sub-steps; full tests; bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/k_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/k_tests.log
echo "=== bench 20"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/k_bench_20.json 2> gpurun_out/k_bench_20.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"unique_rows|select_anchor|proposals|nc_" -s 20 -c 40 --csv --log-file gpurun_out/k_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/k_ncu_launch.log 2>&1; echo "rc=$?"
grep -E "unique_rows|select_anchor" gpurun_out/k_launches.csv | awk -F'","' '{print $5, $NF}' | head -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/k_bench_20.json').read().strip().splitlines()[-1])
print('value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])
print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
r=d['roofline']; print({k:r[k] for k in ('gap_ms_per_step','kernel_event_sum_ms_per_step')}); print(d['config'].get('step_ms_quantiles'))
PY
