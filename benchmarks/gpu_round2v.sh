#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04f}
V=$PWD/compound-ray_b200/lib/variants
for v in new base new_noaffine; do
  case $v in new) E="CR_X=1";; base) E="CR_LIB_PATH=$V/libEyeRenderer3_base.so";; new_noaffine) E="CR_SM_AFFINE=0";; esac
  env $E timeout 600 python benchmarks/speed_test.py --samples 1,4,8,32,1000 --frames 300 > gpurun_out/${T}_speed_${v}.txt 2>&1; echo "$v rc=$?"
  grep -h "^ *S=" gpurun_out/${T}_speed_${v}.txt | cut -c1-70
done
