#!/bin/bash
# Round-2 run E: tests, bench default (NC v4, depth-3 pipeline), reference arm, configs[3] line, ncu lists.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|select_anchor|feature_prep|window_map|flag_risky|delta|absmax'
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/e_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/e_tests.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/e_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/e_smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "rc=$?"; tail -2 gpurun_out/e_bench.err
echo "=== bench driver-shaped"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_20.json 2> gpurun_out/e_bench_20.err; echo "rc=$?"
echo "=== bench depth 2"; timeout 900 python bench.py --steps 40 --depth 2 --no-cpu-baseline > gpurun_out/e_bench_d2.json 2> gpurun_out/e_bench_d2.err; echo "rc=$?"
echo "=== bench config3"; timeout 900 python bench.py --height 768 --width 1024 --ptmax 1000 --steps 20 --no-cpu-baseline > gpurun_out/e_bench_cfg3.json 2> gpurun_out/e_bench_cfg3.err; echo "rc=$?"; tail -2 gpurun_out/e_bench_cfg3.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 60 -c 120 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full nc"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"nc_l1|nc_l2|nc_combine|window_map" -s 4 -c 4 -o gpurun_out/e_prof_nc -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu_nc.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ('e_bench','e_bench_20','e_bench_d2','e_bench_cfg3','e_bench_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', d['e2e'], 'launches', d['gpu_launches'])
    if 'kernels' in d:
        print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
        r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','gap_ms_per_step','kernel_event_sum_ms_per_step','band_rows_fraction')}); print(d['clocks']); print(d.get('refine_only'))
PY
