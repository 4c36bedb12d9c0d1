#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 -s > gpurun_out/${T}_gpu_tests_full.log 2>&1; echo "all gpu tests rc=$?"
grep -v "^\[PyEye\]\|^WARNING\|^ERROR: Unable" gpurun_out/${T}_gpu_tests_full.log | tail -16 > gpurun_out/${T}_gpu_tests.log; cat gpurun_out/${T}_gpu_tests.log | cut -c1-200
cat gpurun_out/viewer_screenshot_ieee_vs_fast_math.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
