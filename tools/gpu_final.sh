#!/bin/bash
# Final round-1 evidence: full GPU test suite, smoke, bench (default flags), reference arm, ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_layer|patch_gather|fc_parse|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|nchw_to_nhwc|nsq_rgb|flag_risky|delta'
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/f_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/f_tests.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/f_smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "rc=$?"; tail -2 gpurun_out/f_bench.err
if [ -z "$SKIP_REF" ]; then echo "=== bench ref"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "rc=$?"; fi
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 66 -c 99 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full gemm"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_ -s 14 -c 7 -o gpurun_out/f_prof_umma -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_umma.log 2>&1; echo "rc=$?"
if [ -z "$SKIP_OTHER" ]; then echo "=== ncu full nc+gather"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"nc_layer|patch_gather|fc_parse" -s 16 -c 8 -o gpurun_out/f_prof_other -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_other.log 2>&1; echo "rc=$?"; fi
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])
print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()}); print(d['clocks']); print(d['cpu_baseline'])
PY
