#!/bin/bash
# Multi-GPU evidence: weak scaling (one pair per rank per step) and BASELINE configs[4] (512 pairs sharded, strong scaling
# with the cross-rank bit-equality check).  usage: gpu_r2_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "=== weak N=$N"; timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 --e2e-modes tf32 > gpurun_out/m_bench_${N}gpu.json 2> gpurun_out/m_bench_${N}gpu.err; echo "rc=$?"; tail -2 gpurun_out/m_bench_${N}gpu.err
echo "=== strong 512 pairs N=$N"; timeout 1200 $RUN bench.py --gpus $N --pairs 512 --warmup 3 > gpurun_out/m_bench_${N}gpu_512pairs.json 2> gpurun_out/m_bench_${N}gpu_512pairs.err; echo "rc=$?"; tail -2 gpurun_out/m_bench_${N}gpu_512pairs.err
[ "$N" -le 2 ] && echo "=== 1 GPU, same box, same flags" && timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/m_bench_1gpu_samebox.json 2> gpurun_out/m_bench_1gpu_samebox.err; echo "rc=$?"
python - <<PY
import json
for f in ('m_bench_${N}gpu','m_bench_${N}gpu_512pairs','m_bench_1gpu_samebox'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],2), 'n_gpus', d['n_gpus'], 'scaling', d['scaling'], 'steps', d['steps'], 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'cross', d.get('cross_rank_check'))
PY
