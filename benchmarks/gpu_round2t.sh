#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04h}
mkdir -p gpurun_out
V=$PWD/compound-ray_b200/lib/variants
run() { name=$1; shift
  env "$@" timeout 600 python benchmarks/wavefront_sweep.py --modes 1:0 --refills 12 --node-lanes 8 --out gpurun_out/${T}_ab_${name}.json > gpurun_out/${T}_ab_${name}.log 2>&1; echo "$name rc=$?"
  grep -E '"what": "(per frame, no lists)"' gpurun_out/${T}_ab_${name}.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('   ', d['what'], {k: round(v, 4) for k, v in d.items() if k in ('grays_device', 'ms_per_frame', 'per_frame_wall_ms', 'trace_ms', 'frontier_ms', 'reduce_ms', 'grays_per_frame_abi')}, d.get('rows_equal_first_variant'))"; }
run nofinish CR_FUSED_FINISH=0
run nofence CR_LIB_PATH=$V/libEyeRenderer3_nofence.so
run nohostcopy CR_LIB_PATH=$V/libEyeRenderer3_nohostcopy.so
