#!/bin/bash
# frontier pass v2 with branch-free decisions: parity + fuzz + frontier_ms
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python compound-ray_b200/tools/frontier_fuzz.py --configs 200 --seed 13 2>&1 | tail -1
run() { name=$1; shift
  env "$@" timeout 600 python benchmarks/wavefront_sweep.py --modes 1:0 --refills 12 --node-lanes 8 --out gpurun_out/${T}_ab_${name}.json > gpurun_out/${T}_ab_${name}.log 2>&1; echo "$name rc=$?"
  grep -E '"what": "(batch, no lists, no queue|per frame, no lists)"' gpurun_out/${T}_ab_${name}.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('   ', d['what'], {k: round(v, 4) for k, v in d.items() if k in ('grays_device', 'ms_per_frame', 'per_frame_wall_ms', 'trace_ms', 'frontier_ms', 'reduce_ms', 'grays_per_frame_abi')}, d.get('rows_equal_first_variant'))"; }
run a CR_X=1
run b CR_X=1
