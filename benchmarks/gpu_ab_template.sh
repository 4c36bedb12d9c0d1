#!/bin/bash
# Template of the in-run A/B calls of the last session (one gpurun call = one box): focused parity with the variant library, then
# bench.py lines of the default build and the variant in alternation.  Variants: `make -C compound-ray_b200 variant NAME=x DEFS="-D..."`
# (compile-time switches) or environment switches.  usage: VARIANT=name bash benchmarks/gpu_round2t.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-ab}
NAME=${VARIANT:-base}
mkdir -p gpurun_out
V=$PWD/compound-ray_b200/lib/variants
CR_LIB_PATH=$V/libEyeRenderer3_${NAME}.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do for v in default $NAME; do
case $v in default) E="CR_X=1";; *) E="CR_LIB_PATH=$V/libEyeRenderer3_${NAME}.so";; esac
env $E timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-modes 2>/dev/null > gpurun_out/${T}_bench_${v}_$i.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_${v}_$i.json')); print('bench $v: value %.2f e2e %.2f (%.4f ms)' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step']))"
done; done
