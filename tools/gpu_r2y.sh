#!/bin/bash
# ncu --set full of the umma_gemm launches of one pair (conv2 1-pass, band convs 3-pass, FC, correlation), final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 12 -c 12 -o gpurun_out/y_prof_umma -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-modes tf32 > gpurun_out/y_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/y_ncu.log
