#!/bin/bash
# tests + repeated driver-shaped bench runs (pinned slabs, event pool)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/s_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s_tests.log
for i in 1 2 3 4 5 6 7 8; do
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --backbone-fp32 > gpurun_out/s_bench_$i.json 2> gpurun_out/s_bench_$i.err; echo "run $i rc=$?"
done
python - <<'PY'
import json
for i in range(1,9):
    try:
        d=json.loads(open(f'gpurun_out/s_bench_{i}.json').read().strip().splitlines()[-1])
        print(i, 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config']['step_ms_quantiles'], d['clocks'].get('samples'), d['clocks'].get('slowest_nvml_query_ms'))
    except Exception as e:
        print(i, 'unreadable', e)
PY
