#!/bin/bash
# CTA-pair GEMM bring-up: bit-identity test, then bench with/without the pair kernel on the conv launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pair test"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 200 -x -k "cta_pair" 2>&1 | tail -15
for gp in "$@"; do
  echo "=== bench gemm_pair=$gp"
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --gemm-pair $gp > gpurun_out/pair_$gp.json 2> gpurun_out/pair_$gp.err; echo "rc=$?"; tail -3 gpurun_out/pair_$gp.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/pair_$gp.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))
print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
done
