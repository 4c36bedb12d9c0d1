#!/bin/bash
# same box: 1 rank on GPU 0, 1 rank on GPU 1, 2 ranks (driver's launch), 2 ranks without the SM-affine hand-out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04q}
p() { python -c "
import json,sys; d=json.load(open('$1')); print('$2', round(d['value']/1e9,2), round(d['ms_per_step'],4), [round(x,2) for x in d['batch_ms']], 'e2e', round(d['e2e']['value']/1e9,2), d.get('per_rank'), d['clocks'])"; }
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/${T}_1gpu_dev0.json 2>/dev/null; p gpurun_out/${T}_1gpu_dev0.json dev0
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/${T}_1gpu_dev1.json 2>/dev/null; p gpurun_out/${T}_1gpu_dev1.json dev1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_2gpu.json 2>/dev/null; p gpurun_out/${T}_2gpu.json 2gpu
CR_SM_AFFINE=0 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_2gpu_noaffine.json 2>/dev/null; p gpurun_out/${T}_2gpu_noaffine.json 2gpu_noaffine
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
