#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_outputs.py -x -q -m gpu -s -k "fast_math_mode or viewer_screenshot" > gpurun_out/${T}_fast_tests_full.log 2>&1; echo "rc=$?"
grep -v "^\[PyEye\]\|^WARNING\|^ERROR: Unable" gpurun_out/${T}_fast_tests_full.log | tail -12 | cut -c1-260
cat gpurun_out/reference_frames_ieee_vs_fast_math.txt
