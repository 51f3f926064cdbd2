#!/bin/bash
# Round profile: bench line + ncu launch list + ncu --set full of the main kernels (1 GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_layer|patch_gather|fc_parse|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|nchw_to_nhwc|nsq_rgb|flag_risky|delta'
echo "=== bench"; timeout 900 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/p_bench.json
echo "=== bench ref"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err; echo "rc=$?"; tail -c 900 gpurun_out/p_bench_ref.json
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 33 -c 66 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k "$KREG" -s 33 -c 33 -o gpurun_out/p_prof_all -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/p_*
