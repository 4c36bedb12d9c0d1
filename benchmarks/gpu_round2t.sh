#!/bin/bash
# pair mode of the SM-affine hand-out: parity (forced on everywhere), then A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04m}
mkdir -p gpurun_out
CR_SM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_bench_size.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do for pr in 1 0; do
env CR_SM_PAIR=$pr timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-modes 2>/dev/null > gpurun_out/${T}_bench_pair${pr}_$i.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_pair${pr}_$i.json')); print('bench pair=$pr: value %.2f e2e %.2f (%.4f ms)' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step']))"
done; done
