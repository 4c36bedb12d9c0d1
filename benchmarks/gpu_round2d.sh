#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r02d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all rc=$?"
tail -14 gpurun_out/${T}_gpu_tests.log
timeout 900 python benchmarks/mode_matrix.py --lists 0,1,2 --out gpurun_out/${T}_mode_matrix.json > gpurun_out/${T}_mode_matrix.log 2>&1; echo "matrix rc=$?"
grep grays_device gpurun_out/${T}_mode_matrix.log | cut -c1-60,150-330
timeout 1200 python benchmarks/configs.py --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.out 2> gpurun_out/${T}_configs.log; echo "configs rc=$?"
tail -5 gpurun_out/${T}_configs.log
timeout 600 python benchmarks/speed_test.py > gpurun_out/${T}_speed_test_protocol.json 2> gpurun_out/${T}_speed_test_protocol.txt; echo "speedtest rc=$?"
timeout 600 python benchmarks/pose_batch.py --poses 2048 --samples 64 --native --chunk 256 > gpurun_out/${T}_pose_batch_native_1gpu.json 2> gpurun_out/${T}_pose_batch_native_1gpu.log; echo "native1 rc=$?"; cat gpurun_out/${T}_pose_batch_native_1gpu.json
timeout 600 python benchmarks/pose_batch.py --poses 2048 --samples 64 > gpurun_out/${T}_pose_batch_1gpu.json 2>> gpurun_out/${T}_pose_batch_native_1gpu.log; cat gpurun_out/${T}_pose_batch_1gpu.json
