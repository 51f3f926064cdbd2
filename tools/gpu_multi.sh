#!/bin/bash
# multi-GPU validation: bench at N=1 and N=$1 (torchrun, NCCL), plus the new config tests on GPU 0
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_gpus.txt
echo "=== config tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 900 -k "config1 or config3" > gpurun_out/m_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/m_tests.log
echo "=== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_1.json 2> gpurun_out/m_bench_1.err; echo "rc=$?"; tail -2 gpurun_out/m_bench_1.err
echo "=== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/m_bench_$N.json 2> gpurun_out/m_bench_$N.err; echo "rc=$?"; tail -5 gpurun_out/m_bench_$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads(open(f'gpurun_out/m_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['clocks'])
    except Exception as e: print(n, 'ERR', e)
PY
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err; echo "rc=$?"; tail -c 600 gpurun_out/m_bench_ref.json
