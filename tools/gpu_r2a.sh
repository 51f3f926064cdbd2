#!/bin/bash
# Round-2 baseline on the representative workload: tests, bench (new + legacy workload), band stats, ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_rows|select_anchor|nchw_to_nhwc|nsq_rgb|flag_risky|delta'
echo "=== nc unit"; timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 -k "neigh_consensus" > gpurun_out/a_nc.log 2>&1; NCRC=$?; echo "rc=$NCRC"; tail -12 gpurun_out/a_nc.log
if [ "$NCRC" != "0" ]; then echo "!!! tensor-core NC failed its unit test: everything below runs with nc_impl=0"; export P2P_OPTIONS="nc_impl=0"; fi
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/a_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/a_tests.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/a_smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "rc=$?"; tail -2 gpurun_out/a_bench.err
echo "=== bench nc_impl 0"; timeout 600 python bench.py --nc-impl 0 --steps 30 --no-cpu-baseline > gpurun_out/a_bench_nc0.json 2> gpurun_out/a_bench_nc0.err; echo "rc=$?"; tail -2 gpurun_out/a_bench_nc0.err
echo "=== bench legacy"; timeout 600 python bench.py --legacy-workload --steps 30 --no-cpu-baseline > gpurun_out/a_bench_legacy.json 2> gpurun_out/a_bench_legacy.err; echo "rc=$?"; tail -2 gpurun_out/a_bench_legacy.err
echo "=== band stats"; timeout 900 python tools/band_stats.py > gpurun_out/a_band.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_band.log
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 60 -c 120 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_launch.log 2>&1; echo "rc=$?"
echo "=== ncu full conv1 fused"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_conv1_fused -s 2 -c 2 -o gpurun_out/a_prof_conv1 -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_conv1.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ('a_bench','a_bench_nc0','a_bench_legacy'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', d['e2e'], 'launches', d['gpu_launches'])
    print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    print(d['roofline']); print(d['clocks']); print(d['config'].get('distinct_proposals_first_pairs')); print(d.get('refine_only'))
PY
