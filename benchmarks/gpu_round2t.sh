#!/bin/bash
# pre-draw for large single frames: parity (focused), then bench A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_bench_size.py -x -q -m gpu 2>&1 | tail -15 | cut -c1-200
for pre in 1 0; do
env CR_PREDRAW=$pre timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-modes 2>/dev/null > gpurun_out/${T}_bench_predraw${pre}.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_predraw${pre}.json')); print('bench predraw=$pre: value %.2f e2e %.2f (%.4f ms) launches %s' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['gpu_launches']))"
done
