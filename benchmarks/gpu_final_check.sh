#!/bin/bash
# final tree: all GPU tests, smoke, the driver's bench invocation (N = 1)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -8 gpurun_out/${T}_gpu_tests.log | cut -c1-160
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/${T}_bench_steps20.json 2> gpurun_out/${T}_bench.log; echo "bench20 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_steps20.json')); r=d['roofline']
print('value %.2f e2e %.2f (%.4f ms) issue frac %.3f traffic %s dram_frac %.4f clocks %s launches %s' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step'], r['frac'], r['traffic'], r['hbm']['dram_frac'], d['clocks'], d['gpu_launches']))
print({k:(round(v['rays_per_sec']/1e9,2),round(v['e2e_rays_per_sec']/1e9,2)) for k,v in d['modes'].items() if isinstance(v,dict)})"
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
