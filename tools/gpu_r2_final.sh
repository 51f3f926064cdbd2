#!/bin/bash
# Round-2 evidence run on one B200: smoke, GPU tests, bench lines (driver-shaped 20/5, default 100/5 with the CPU
# baseline, reference arm, BASELINE configs[3]), ncu launch list, ncu --set full captures of the dominant kernels.
cd "$(dirname "$0")/.."
P=${1:-p}
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|select_anchor|feature_prep|window_map|flag_risky|delta|absmax'
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${P}_smoke.log
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/${P}_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${P}_tests.log
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err; echo "rc=$?"
echo "=== bench 20/5"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${P}_bench_20.json 2> gpurun_out/${P}_bench_20.err; echo "rc=$?"
echo "=== bench default"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "rc=$?"
echo "=== bench configs[3] 1024x768 ptmax 1000"; timeout 900 python bench.py --height 768 --width 1024 --ptmax 1000 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${P}_bench_cfg3.json 2> gpurun_out/${P}_bench_cfg3.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 70 -c 140 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 8 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full gemm"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"umma_conv1_tma|umma_gemm_kernel<1, 0, 2|umma_gemm_kernel<3, 1, 1|umma_gemm_kernel<3, 1, 2" -s 4 -c 6 -o gpurun_out/${P}_prof_gemm -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_ncu_gemm.log 2>&1; echo "rc=$?"
echo "=== ncu full nc + coarse"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"nc_|corr|l2norm|mutual|proposals|unique_rows|window_map|feature_prep" -s 13 -c 13 -o gpurun_out/${P}_prof_coarse -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_ncu_coarse.log 2>&1; echo "rc=$?"
python - <<PY
import json
for f in ('${P}_bench_ref','${P}_bench_20','${P}_bench','${P}_bench_cfg3'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],4), 'ms/step', round(d['ms_per_step'],3), 'e2e', d['e2e'].get('value'), 'launches', d.get('gpu_launches'))
    if 'kernels' in d and d['kernels']: print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    if d.get('roofline'): print({k:d['roofline'][k] for k in ('kernel','achieved','peak','frac','gap_ms_per_step')})
    print(d.get('clocks')); print(d.get('cpu_baseline'))
PY
