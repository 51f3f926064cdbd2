#!/bin/bash
# Round-2 run N: unique_rows with compile-time shared addressing; full tests + bench + launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/n_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/n_tests.log
echo "=== bench 20"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n_bench_20.json 2> gpurun_out/n_bench_20.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"unique_rows|select_anchor|proposals" -s 8 -c 16 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/n_ncu_launch.log 2>&1; echo "rc=$?"
grep -E "unique_rows" gpurun_out/n_launches.csv | awk -F'","' '{print $NF}' | head -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n_bench_20.json').read().strip().splitlines()[-1])
print('value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])
print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
