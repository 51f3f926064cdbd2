#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "=== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "rc=$?"; tail -3 gpurun_out/n${N}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['clocks'])
PY
