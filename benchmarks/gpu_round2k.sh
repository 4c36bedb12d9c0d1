#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -3 gpurun_out/${T}_gpu_tests.log
timeout 900 python benchmarks/configs.py --only cfg4 --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1; echo "configs rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_configs.json'));print([(x['triangles'], round(x['bvh_build_ms'],2), round(x['rays_per_sec_batched']/1e9,1)) for x in d['cfg4']['sweep']])"
timeout 900 python benchmarks/configs.py --only cfg4 --out gpurun_out/${T}_configs2.json > gpurun_out/${T}_configs2.log 2>&1
python -c "import json;d=json.load(open('gpurun_out/${T}_configs2.json'));print([(x['triangles'], round(x['bvh_build_ms'],2), round(x['rays_per_sec_batched']/1e9,1)) for x in d['cfg4']['sweep']])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
