#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do
echo "=== stall probe sampler $i"; timeout 600 python tools/stall_probe.py 600 sampler > gpurun_out/q_stall_sampler$i.log 2>&1; echo "rc=$?"; grep -E "^ms/step|^sampler|^step|quantiles" gpurun_out/q_stall_sampler$i.log | cut -c1-400 | head -8
done
echo "=== stall probe nosampler"; timeout 600 python tools/stall_probe.py 600 > gpurun_out/q_stall_nogc.log 2>&1; echo "rc=$?"; grep -E "^ms/step|^step|quantiles" gpurun_out/q_stall_nogc.log | cut -c1-400 | head -6
