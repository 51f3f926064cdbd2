#!/bin/bash
# Round-2 run B: full tests with the tensor-core NC as default, A/B benches (fuse_gather 1 vs 3), ncu of the new kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|select_anchor|feature_prep|window_map|flag_risky|delta|absmax'
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/b_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/b_tests.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/b_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/b_smoke.log
echo "=== bench fg1"; timeout 900 python bench.py --steps 40 --no-cpu-baseline > gpurun_out/b_bench_fg1.json 2> gpurun_out/b_bench_fg1.err; echo "rc=$?"; tail -2 gpurun_out/b_bench_fg1.err
echo "=== bench fg3"; timeout 900 python bench.py --steps 40 --no-cpu-baseline --fuse-gather 3 > gpurun_out/b_bench_fg3.json 2> gpurun_out/b_bench_fg3.err; echo "rc=$?"; tail -2 gpurun_out/b_bench_fg3.err
echo "=== bench fg3 band35"; timeout 900 python bench.py --steps 40 --no-cpu-baseline --fuse-gather 3 --mid-band 35 > gpurun_out/b_bench_fg3_b35.json 2> gpurun_out/b_bench_fg3_b35.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 60 -c 120 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full nc"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"nc_umma|nc_combine" -s 3 -c 3 -o gpurun_out/b_prof_nc -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_nc.log 2>&1; echo "rc=$?"
echo "=== ncu full conv1 tma"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"umma_conv1_tma|window_map" -s 2 -c 3 -o gpurun_out/b_prof_conv1tma -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --fuse-gather 3 > gpurun_out/b_ncu_conv1tma.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ('b_bench_fg1','b_bench_fg3','b_bench_fg3_b35'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
    print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','gap_ms_per_step','kernel_event_sum_ms_per_step','band_rows_fraction')}); print(d['clocks']); print(d.get('refine_only'))
PY
