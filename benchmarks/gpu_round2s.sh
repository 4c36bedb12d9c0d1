#!/bin/bash
# 8-GPU box, final round-2 tree: the driver's torchrun launch of bench.py at N = 8, 4, 2, 1, reference arm, cfg5 as one job.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03g}
mkdir -p gpurun_out
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.log; echo "bench$N rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/${T}_bench_${N}gpu.json'));print('$N GPUs:', d['value']/1e9, d['ms_per_step'], d['batch_ms'], d['e2e']['value']/1e9, d.get('per_rank'))"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.log
python -c "import json;d=json.load(open('gpurun_out/${T}_bench_1gpu.json'));print('1 GPU :', d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530"
timeout 600 $TR bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref_8gpu.json 2> gpurun_out/${T}_bench_ref_8gpu.log
python -c "import json;d=json.load(open('gpurun_out/${T}_bench_ref_8gpu.json'));print('reference arm under torchrun x8:', d['value'], d['cpu_baseline']['cores'])"
timeout 900 $TR benchmarks/pose_batch.py --poses 100000 --samples 64 --native --chunk 2048 --mode fused > gpurun_out/${T}_pose_batch_100k_native_8gpu.json 2> gpurun_out/${T}_pose_batch_8gpu.log; cut -c1-420 gpurun_out/${T}_pose_batch_100k_native_8gpu.json
timeout 900 python benchmarks/pose_batch.py --poses 100000 --samples 64 --native --chunk 2048 --mode fused > gpurun_out/${T}_pose_batch_100k_native_1gpu.json 2>> gpurun_out/${T}_pose_batch_8gpu.log; cut -c1-420 gpurun_out/${T}_pose_batch_100k_native_1gpu.json
