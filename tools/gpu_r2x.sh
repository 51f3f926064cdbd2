#!/bin/bash
# validation after the proposals_dir1 change + refreshed launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREG='regex:umma_|nc_|patch_gather|fc_parse|fc3_parse|pooled_split|corr_pool|l2norm|mutual_apply|rowcolmax|proposals|unique_r|select_anchor|feature_prep|window_map|flag_risky|delta|absmax'
echo "=== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 1200 > gpurun_out/x_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/x_tests.log
echo "=== bench 20/5"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/x_bench_20.json 2> gpurun_out/x_bench_20.err; echo "rc=$?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 70 -c 140 --csv --log-file gpurun_out/x_launches.csv python bench.py --steps 8 --warmup 1 --no-cpu-baseline --e2e-modes tf32 > gpurun_out/x_ncu_launch.log 2>&1; echo "rc=$?"
python - <<'PY'
import json, csv, collections
d=json.loads(open('gpurun_out/x_bench_20.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['config']['step_ms_quantiles'])
print({k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
rows=list(csv.reader(open('gpurun_out/x_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
dd=collections.defaultdict(list)
for r in rows[start+2:]:
    if len(r)>vi:
        try: dd[r[ki].split('(')[0][:50]].append(float(r[vi].replace(',','')))
        except: pass
for k,v in sorted(dd.items(), key=lambda kv:-sum(kv[1])): print(f'{k:52s} n={len(v):3d} avg={sum(v)/len(v)/1000:8.1f} us')
PY
