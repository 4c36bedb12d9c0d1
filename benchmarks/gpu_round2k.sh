#!/bin/bash
# Final-tree check of this milestone: all GPU tests, bench, cfg5 single GPU.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -4 gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.log; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_bench.json'));print(d['value']/1e9, d['e2e']['value']/1e9, {k:(round(v['rays_per_sec']/1e9,2), round(v['e2e_rays_per_sec']/1e9,2)) for k,v in d['modes'].items() if isinstance(v,dict)})"
timeout 900 python benchmarks/pose_batch.py --poses 100000 --samples 64 --native --chunk 2048 --mode fused > gpurun_out/${T}_pose_batch_100k_native_1gpu.json 2> gpurun_out/${T}_pose_batch.log; cut -c1-420 gpurun_out/${T}_pose_batch_100k_native_1gpu.json
timeout 900 python benchmarks/configs.py --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1; echo "configs rc=$?"
