#!/bin/bash
# Round-2 run F: NC v5 + probes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|select_anchor|feature_prep|window_map|flag_risky|delta|absmax'
echo "=== nc unit"; timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 -k "neigh_consensus or large_shapes or coarse_stages" > gpurun_out/f_nc.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/f_nc.log
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/f_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/f_tests.log
echo "=== determinism probe"; timeout 900 python tools/determinism_probe.py > gpurun_out/f_det.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/f_det.log
echo "=== gap probe"; timeout 900 python tools/gap_probe.py > gpurun_out/f_gap.log 2>&1; echo "rc=$?"; tail -26 gpurun_out/f_gap.log
echo "=== bench 20"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_20.json 2> gpurun_out/f_bench_20.err; echo "rc=$?"
echo "=== bench 100"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 60 -c 120 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full nc"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"nc_l1|nc_l2|nc_combine" -s 3 -c 3 -o gpurun_out/f_prof_nc -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_nc.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ('f_bench_20','f_bench'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', d['e2e'], 'launches', d['gpu_launches'])
    print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','gap_ms_per_step','kernel_event_sum_ms_per_step','band_rows_fraction')}); print(d['clocks']); print(d.get('refine_only'))
PY
