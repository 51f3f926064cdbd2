#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/r2_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_tests.log
echo "=== bench default"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "rc=$?"; tail -c 2500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
for v in "--seg-len 1" "--seg-len 9" "--seg-len 0" "--mid-passes 1" "--corr-passes 3"; do
  n=$(echo $v | tr -d ' -'); echo "=== bench $v"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline $v > gpurun_out/r2_bench_$n.json 2> gpurun_out/r2_bench_$n.err; echo "rc=$?"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_$n.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})" 2>&1 | tail -3
done
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full umma"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 4 -c 4 -o gpurun_out/r2_prof_umma -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_umma.log 2>&1; echo "rc=$?"
echo "=== ncu full others"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nc_layer|corr_pool|patch_gather|fc_parse" -s 5 -c 7 -o gpurun_out/r2_prof_other -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_other.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -15
