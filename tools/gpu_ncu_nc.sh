#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nc_layer" -s 2 -c 2 -o gpurun_out/r4_prof_nc -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_ncu_nc.log 2>&1; echo "rc=$?"
