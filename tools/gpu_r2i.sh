#!/bin/bash
# Round-2 run I: NC v6 (bulk-copy staged layer 1, shifted-window layer 2), bench with the profiler pass split off.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|select_anchor|feature_prep|window_map|flag_risky|delta|absmax'
echo "=== nc unit"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 -k "neigh_consensus or large_shapes or coarse_stages" > gpurun_out/i_nc.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/i_nc.log
cat gpurun_out/parity_nc_layer2_mode1.json gpurun_out/parity_nc_layer2_mode2.json 2>/dev/null | tr -d '\n ' | head -c 3000; echo
echo "=== bench 20"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench_20.json 2> gpurun_out/i_bench_20.err; echo "rc=$?"
echo "=== bench 20 mode2"; P2P_OPTIONS=nc_l2_mode=2 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench_20_m2.json 2> gpurun_out/i_bench_20_m2.err; echo "rc=$?"
echo "=== bench 100"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "rc=$?"
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/i_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/i_tests.log
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 60 -c 120 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/i_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full nc"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"nc_pad|nc_l1|nc_l2|nc_combine" -s 4 -c 4 -o gpurun_out/i_prof_nc -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/i_ncu_nc.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ('i_bench_20','i_bench_20_m2','i_bench'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', d['e2e'], 'launches', d['gpu_launches'])
    print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','gap_ms_per_step','kernel_event_sum_ms_per_step','band_rows_fraction')}); print(d['clocks']); print(d.get('refine_only')); print(d['config'].get('step_ms_quantiles'))
PY
