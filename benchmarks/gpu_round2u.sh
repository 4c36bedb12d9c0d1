#!/bin/bash
# ncu --set full of the batched and the per-frame trace kernel (SM-affine hand-out)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04b}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_traceCompound -s 1 -c 1 -f -o gpurun_out/${T}_k1_batched_full \
   python bench.py --mode fused --steps 20 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_ncu_full_batched.log 2>&1; echo "ncu full batched rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_traceCompound -s 8 -c 1 -f -o gpurun_out/${T}_k1_perframe_full \
   python bench.py --mode fused --steps 20 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_ncu_full_perframe.log 2>&1; echo "ncu full per-frame rc=$?"
