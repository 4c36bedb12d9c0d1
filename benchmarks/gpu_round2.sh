#!/bin/bash
# One gpurun call: new-mode tests, mode matrix, whole GPU suite, one ncu capture of the fused packet kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02b_smi.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_modes.py -x -q -m gpu > gpurun_out/r02b_modes_tests.log 2>&1; echo "modes rc=$?"
tail -5 gpurun_out/r02b_modes_tests.log
timeout 900 python benchmarks/mode_matrix.py --out gpurun_out/r02b_mode_matrix.json > gpurun_out/r02b_mode_matrix.log 2>&1; echo "matrix rc=$?"
grep grays_device gpurun_out/r02b_mode_matrix.log | cut -c1-260
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_gpu_tests.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/r02b_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.log; echo "bench rc=$?"
cut -c1-600 gpurun_out/r02b_bench.json
CR_REDUCE=fused timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traceCompound -s 1 -c 1 -f -o gpurun_out/r02b_k1_fused_lists \
   python bench.py --steps 34 --warmup 3 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/r02b_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | head -30
