#!/bin/bash
# conv1: deferred half-1 MMAs (epilogue of half 0 overlaps the next tile) -- bit-identity tests + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x -k "fused_gather or end_to_end_vs_oracle_benchmark or full_size or golden_train or config1 or refine_vs_oracle or determin" > gpurun_out/z_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/z_tests.log
echo "=== bench 20"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-modes tf32 > gpurun_out/z_bench_20.json 2> gpurun_out/z_bench_20.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z_bench_20.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))
print({k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
PY
