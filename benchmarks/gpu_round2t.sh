#!/bin/bash
# programmatic dependent launch (frontier pass -> trace -> reduction) + frontier pass v2 without the near/far swap: parity, A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python compound-ray_b200/tools/frontier_fuzz.py --configs 200 --seed 12 2>&1 | tail -1
for i in 1 2; do for pdl in 1 0; do
env CR_PDL=$pdl timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-modes 2>/dev/null > gpurun_out/${T}_bench_pdl${pdl}_$i.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_pdl${pdl}_$i.json')); print('bench pdl=$pdl: value %.2f e2e %.2f (%.4f ms) launches %s' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['gpu_launches']))"
done; done
