#!/bin/bash
# SM-affine hand-out of the trace kernel's units: all GPU tests with the new default, then in-run A/B against the previous build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04d}
mkdir -p gpurun_out
V=$PWD/compound-ray_b200/lib/variants
run() { name=$1; shift
  env CR_SM_AFFINE_VERBOSE=1 "$@" timeout 600 python benchmarks/wavefront_sweep.py --modes 1:0 --refills 12 --node-lanes 8 --out gpurun_out/${T}_ab_${name}.json > gpurun_out/${T}_ab_${name}.log 2>&1; echo "$name rc=$?"
  grep "SM ids" gpurun_out/${T}_ab_${name}.log
  grep -E '"what": "(batch, no lists, no queue|per frame, no lists)"' gpurun_out/${T}_ab_${name}.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('   ', d['what'], {k: round(v, 4) for k, v in d.items() if k in ('grays_device', 'ms_per_frame', 'per_frame_wall_ms', 'trace_ms', 'frontier_ms', 'grays_per_frame_abi')}, d.get('rows_equal_first_variant'))"; }
for i in 1 2; do
run base$i CR_LIB_PATH=$V/libEyeRenderer3_base.so
run affine$i CR_SM_AFFINE=1
run slotlive$i CR_LIB_PATH=$V/libEyeRenderer3_slotlive.so
done
