#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/gpu_quick.sh
echo "=== band stats"; timeout 600 python tools/band_stats.py 2>&1 | tail -12
