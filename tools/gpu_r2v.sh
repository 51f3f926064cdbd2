#!/bin/bash
# A/B: NC layer 2 with two CTAs per SM; window map with streaming stores
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== nc tests (2 CTAs)"; P2P_OPTIONS=nc_l2_mode=8 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -k "neigh_consensus_tensor or large_shapes or coarse_stages or fused_gather" > gpurun_out/v_nc.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/v_nc.log
for m in 0 8 0 8; do
  P2P_OPTIONS=nc_l2_mode=$m timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-modes tf32 > gpurun_out/v_bench_$m.json 2> gpurun_out/v_bench_$m.err; echo "mode $m rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/v_bench_$m.json').read().strip().splitlines()[-1])
print('l2_mode', $m, 'hot', round(d['value'],1), 'nc', round(d['kernels']['nc']['ms_per_launch'],4), 'prep', round(d['kernels']['prep']['ms_per_launch'],4), 'conv1', round(d['kernels']['conv1_mid']['ms_per_launch'],4))
PY
done
echo "=== ncu launches (2 CTAs)"; P2P_OPTIONS=nc_l2_mode=8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nc_|window_map|feature_prep" -s 12 -c 24 --csv --log-file gpurun_out/v_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline --e2e-modes tf32 > gpurun_out/v_ncu_launch.log 2>&1; echo "rc=$?"
grep -E "nc_l2|window_map|nc_l1" gpurun_out/v_launches.csv | awk -F'","' '{print substr($5,1,28), $NF}' | head -9
