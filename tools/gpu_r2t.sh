#!/bin/bash
# overlap-mode backbone: test + A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== test"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -k "backbone" > gpurun_out/t_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_tests.log
for ov in 1 0 1 0; do
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-modes tf32,fp16 --e2e-overlap $ov > gpurun_out/t_bench_ov$ov.json 2> gpurun_out/t_bench_ov$ov.err; echo "overlap $ov rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/t_bench_ov$ov.json').read().strip().splitlines()[-1])
print('overlap', $ov, 'hot', round(d['value'],1), 'e2e tf32', round(d['e2e']['value'],1), 'fp16', round(d['e2e'].get('fp16_channels_last_backbone_value',0),1), 'traffic', d['roofline']['traffic'])
PY
done
