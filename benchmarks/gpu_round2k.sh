#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -3 gpurun_out/${T}_gpu_tests.log
timeout 900 python benchmarks/configs.py --only cfg4 --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1; echo "configs rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_configs.json'));print([(x['triangles'], round(x['bvh_build_ms'],2), round(x['rays_per_sec_batched']/1e9,1)) for x in d['cfg4']['sweep']])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${T}_build_launches.csv \
   python bench.py --steps 4 --warmup 3 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r03s_build_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split('(')[0][-40:]
    agg.setdefault(k,[]).append(float(r[-1]))
tot=0
for k,v in agg.items():
    if 'trace' in k or 'sum' in k or 'Entries' in k or 'rng' in k or 'prep' in k: continue
    print(f"{k:42s} n={len(v):3d} total={sum(v)/1e3:8.1f}us"); tot+=sum(v)
print('build kernels total us', tot/1e3)
PY
