#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03l}
mkdir -p gpurun_out
timeout 900 python benchmarks/pose_batch.py --poses 100000 --samples 64 --native --chunk 2048 --mode fused > gpurun_out/${T}_pose_batch_100k_native_1gpu.json 2> gpurun_out/${T}_pose_batch.log; cut -c1-600 gpurun_out/${T}_pose_batch_100k_native_1gpu.json
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_sharding.py -x -q -m gpu 2>&1 | tail -3
