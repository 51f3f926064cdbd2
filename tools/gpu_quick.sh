#!/bin/bash
# quick check: parity tests + one bench line (+ optional extra args for bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -x > gpurun_out/q_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/q_tests.log
echo "=== bench"; timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "rc=$?"; tail -3 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])
print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
print(d['roofline']); print(d['clocks'])
PY
