#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03u}
mkdir -p gpurun_out
V=$PWD/compound-ray_b200/lib/variants
CR_LIB_PATH=$V/libEyeRenderer3_cpasync.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py -x -q -m gpu 2>&1 | tail -2
run() { name=$1; shift
  env "$@" timeout 600 python benchmarks/wavefront_sweep.py --modes 1:0 --refills 12 --node-lanes 8 --out gpurun_out/${T}_ab_${name}.json > gpurun_out/${T}_ab_${name}.log 2>&1; echo "$name rc=$?"; }
run base CR_X=1
run cpasync CR_LIB_PATH=$V/libEyeRenderer3_cpasync.so
run base2 CR_X=1
run cpasync2 CR_LIB_PATH=$V/libEyeRenderer3_cpasync.so
