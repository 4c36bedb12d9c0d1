#!/bin/bash
# node loads with an L2 evict-last policy: parity (focused), A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04u}
mkdir -p gpurun_out
V=$PWD/compound-ray_b200/lib/variants
CR_LIB_PATH=$V/libEyeRenderer3_evictlast.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do for v in default evictlast; do
case $v in default) E="CR_X=1";; evictlast) E="CR_LIB_PATH=$V/libEyeRenderer3_evictlast.so";; esac
env $E timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-modes 2>/dev/null > gpurun_out/${T}_bench_${v}_$i.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_${v}_$i.json')); print('bench $v: value %.2f e2e %.2f (%.4f ms)' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step']))"
done; done
