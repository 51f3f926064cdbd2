#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_quick.sh
echo "=== ncu nc/fc/gather"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nc_layer|fc_parse|patch_gather|flag_risky" -s 9 -c 9 -o gpurun_out/r3_prof_other -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_ncu_other.log 2>&1; echo "rc=$?"
