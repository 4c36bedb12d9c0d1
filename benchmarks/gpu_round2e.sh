#!/bin/bash
# 2-GPU call: the driver's torchrun launch of bench.py (both arms), and the library's own NCCL data plane.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r02e}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref_2gpu.json 2> gpurun_out/${T}_bench_ref_2gpu.log; echo "ref rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_bench_ref_2gpu.json'));print('reference arm under torchrun:', d['value'], d['cpu_baseline']['cores'])"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.log; echo "bench2 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_bench_2gpu.json'));print('2 GPUs:', d['value']/1e9, d['ms_per_step'], d['batch_ms'], d['e2e']['value']/1e9)"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/${T}_bench_1gpu.json 2>> gpurun_out/${T}_bench_2gpu.log
python -c "import json;d=json.load(open('gpurun_out/${T}_bench_1gpu.json'));print('1 GPU :', d['value']/1e9, d['ms_per_step'])"
timeout 600 $TR benchmarks/pose_batch.py --poses 2048 --samples 64 --native --chunk 256 > gpurun_out/${T}_pose_batch_native_2gpu.json 2> gpurun_out/${T}_pose_batch_native_2gpu.log; echo "native2 rc=$?"; cut -c1-400 gpurun_out/${T}_pose_batch_native_2gpu.json
timeout 600 $TR benchmarks/pose_batch.py --poses 2047 --samples 64 --native --chunk 100 > gpurun_out/${T}_pose_batch_native_2gpu_odd.json 2>> gpurun_out/${T}_pose_batch_native_2gpu.log; echo "native2 odd rc=$?"; cut -c1-400 gpurun_out/${T}_pose_batch_native_2gpu_odd.json
timeout 600 python benchmarks/pose_batch.py --poses 2047 --samples 64 > gpurun_out/${T}_pose_batch_1gpu_odd.json 2>> gpurun_out/${T}_pose_batch_native_2gpu.log; cut -c1-400 gpurun_out/${T}_pose_batch_1gpu_odd.json
timeout 600 $TR benchmarks/pose_batch.py --poses 2048 --samples 64 --chunk 256 > gpurun_out/${T}_pose_batch_torch_2gpu.json 2>> gpurun_out/${T}_pose_batch_native_2gpu.log; cut -c1-400 gpurun_out/${T}_pose_batch_torch_2gpu.json
# the library's data plane without torch in the process: two plain processes, unique id through a file
rm -f /tmp/cr_nccl_id
(RANK=1 WORLD_SIZE=2 LOCAL_RANK=1 timeout 300 python benchmarks/pose_batch.py --poses 2048 --samples 64 --native --chunk 256 --id-file /tmp/cr_nccl_id > /dev/null 2> gpurun_out/${T}_pose_batch_notorch_rank1.log &)
RANK=0 WORLD_SIZE=2 LOCAL_RANK=0 timeout 300 python benchmarks/pose_batch.py --poses 2048 --samples 64 --native --chunk 256 --id-file /tmp/cr_nccl_id > gpurun_out/${T}_pose_batch_notorch_2gpu.json 2> gpurun_out/${T}_pose_batch_notorch_rank0.log; echo "notorch rc=$?"; cut -c1-400 gpurun_out/${T}_pose_batch_notorch_2gpu.json
sleep 2
timeout 900 $TR benchmarks/pose_batch.py --poses 100000 --samples 64 --native --chunk 4096 --mode fused > gpurun_out/${T}_pose_batch_100k_native_2gpu.json 2>> gpurun_out/${T}_pose_batch_native_2gpu.log; cut -c1-400 gpurun_out/${T}_pose_batch_100k_native_2gpu.json
