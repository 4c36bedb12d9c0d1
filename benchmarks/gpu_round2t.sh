#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04k}
mkdir -p gpurun_out
V=$PWD/compound-ray_b200/lib/variants
CR_LIB_PATH=$V/libEyeRenderer3_latewait.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do for v in default latewait; do
case $v in default) E="CR_X=1";; latewait) E="CR_LIB_PATH=$V/libEyeRenderer3_latewait.so";; esac
env $E timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-modes 2>/dev/null > gpurun_out/${T}_bench_${v}_$i.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_${v}_$i.json')); print('bench $v: value %.2f e2e %.2f (%.4f ms)' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step']))"
done; done
timeout 600 python benchmarks/speed_test.py > gpurun_out/${T}_speed_test_protocol.txt 2>&1; echo "speed test rc=$?"
grep -h "^ *S=" gpurun_out/${T}_speed_test_protocol.txt | cut -c1-72
