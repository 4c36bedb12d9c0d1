#!/bin/bash
# K0b on the cone axis's node copy; frontier depth cap sweep.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r02q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${T}_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python benchmarks/wavefront_sweep.py --modes 1:0 --refills 12 --node-lanes 8 --inline-lanes ${NL:-16} --out gpurun_out/${T}_ab_${name}.json > gpurun_out/${T}_ab_${name}.log 2>&1; echo "$name rc=$?"
}
run default CR_X=1
run lev20 CR_ENTRY_MAX_LEVELS=20
run lev16 CR_ENTRY_MAX_LEVELS=16
run lev13 CR_ENTRY_MAX_LEVELS=13
run lev10 CR_ENTRY_MAX_LEVELS=10
run default2 CR_X=1
ls gpurun_out | grep ${T}
