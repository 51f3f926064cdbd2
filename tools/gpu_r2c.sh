#!/bin/bash
# Round-2 run C: tests (tie-aware), bench default (window-map conv1 + rewritten NC kernels), A/B, ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|select_anchor|feature_prep|window_map|flag_risky|delta|absmax'
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/c_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/c_tests.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/c_smoke.log
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "rc=$?"; tail -2 gpurun_out/c_bench.err
echo "=== bench fg1"; timeout 900 python bench.py --steps 40 --no-cpu-baseline --fuse-gather 1 > gpurun_out/c_bench_fg1.json 2> gpurun_out/c_bench_fg1.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 60 -c 120 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full nc"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"nc_l1|nc_l2|nc_combine|window_map|unique_rows" -s 5 -c 5 -o gpurun_out/c_prof_nc -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_nc.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ('c_bench','c_bench_fg1'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
    print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','gap_ms_per_step','kernel_event_sum_ms_per_step','band_rows_fraction')}); print(d['clocks']); print(d.get('refine_only'))
PY
