#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/gpu_quick.sh
echo "=== band stats"; timeout 600 python tools/band_stats.py 2>&1 | tail -16
echo "=== ncu fused"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_conv1_fused -s 2 -c 2 -o gpurun_out/r6_prof_fused -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r6_ncu_fused.log 2>&1; echo "rc=$?"
