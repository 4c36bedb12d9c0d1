#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_modes.py -x -q -m gpu -k "read_ahead or frame_groups or sm_affine" 2>&1 | tail -2
for v in default noaffine default noaffine; do
  case $v in default*) E="CR_X=1";; nopdl) E="CR_PDL=0";; noaffine) E="CR_SM_AFFINE=0";; esac
  echo "== $v"; env $E timeout 300 python benchmarks/speed_probe.py 8,1000,2000,3200 500 2>&1 | grep '^{' | cut -c1-260
done
