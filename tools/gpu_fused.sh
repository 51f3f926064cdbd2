#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 600 -k "fused or tc11 or tc31 or band31 or end_to_end" -x > gpurun_out/fu_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/fu_tests.log
for f in 0 1 2; do
python - <<PY
import subprocess, json, sys
sys.path.insert(0,'.')
import bench
PY
done
for f in 0 1; do
echo "=== bench fuse_gather=$f"
P2P_FUSE=$f timeout 600 python - <<'PY'
import os, sys, json, subprocess
f=os.environ['P2P_FUSE']
r=subprocess.run([sys.executable,'bench.py','--steps','30','--warmup','3','--no-cpu-baseline','--fuse-gather',f],capture_output=True,text=True)
try:
    d=json.loads(r.stdout.strip().splitlines()[-1])
    print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))
    print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
except Exception as e:
    print('ERR', e, r.stderr[-1500:])
PY
done
