#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== nc debug uniform"; timeout 600 python tools/nc_debug.py uniform > gpurun_out/h_ncdbg_uniform.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/h_ncdbg_uniform.log
echo "=== nc debug consensus"; timeout 600 python tools/nc_debug.py consensus > gpurun_out/h_ncdbg_consensus.log 2>&1; echo "rc=$?"; grep "^dims" gpurun_out/h_ncdbg_consensus.log
