#!/bin/bash
# Wavefront queue + phase-switched walk: parity tests, sweeps (default build and the CR_INLINE_PHASED variant), ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r02j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py -x -q -m gpu --durations=8 > gpurun_out/${T}_modes_tests.log 2>&1; echo "modes tests rc=$?"
tail -5 gpurun_out/${T}_modes_tests.log
CR_LIB_PATH=$PWD/compound-ray_b200/lib/variants/libEyeRenderer3_phased.so timeout 900 python -m pytest tests/test_gpu_modes.py -x -q -m gpu > gpurun_out/${T}_modes_tests_phased.log 2>&1; echo "modes tests (phased inline) rc=$?"
tail -3 gpurun_out/${T}_modes_tests_phased.log
timeout 900 python benchmarks/wavefront_sweep.py --refills 12,16,24 --node-lanes 6,8,12,16 --out gpurun_out/${T}_wavefront.json > gpurun_out/${T}_wavefront.log 2>&1; echo "sweep rc=$?"
CR_LIB_PATH=$PWD/compound-ray_b200/lib/variants/libEyeRenderer3_phased.so timeout 900 python benchmarks/wavefront_sweep.py --modes 1:0 --refills 16 --node-lanes 8,12 --inline-lanes 1,6,8,12 --out gpurun_out/${T}_wavefront_phased.json > gpurun_out/${T}_wavefront_phased.log 2>&1; echo "sweep (phased inline) rc=$?"
for K in k_traceQueue k_traceCompound; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -s 1 -c 1 -f -o gpurun_out/${T}_${K}_full \
     python bench.py --mode fused --steps 20 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_ncu_${K}.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la gpurun_out | grep ${T}
