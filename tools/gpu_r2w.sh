#!/bin/bash
# validation of the final defaults: full GPU tests, smoke, driver-shaped bench, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/w_smoke.log
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/w_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/w_tests.log
echo "=== bench 20/5"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/w_bench_20.json 2> gpurun_out/w_bench_20.err; echo "rc=$?"
echo "=== bench default"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('w_bench_20','w_bench'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['config']['step_ms_quantiles'])
    print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
