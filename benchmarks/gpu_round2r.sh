#!/bin/bash
# Evidence refresh for the final round-2 tree: all GPU tests, ncu captures for profiles/k1_traffic.json, full ncu of the batched,
# per-frame and frame-group kernels, bench lines, launch list, other configs, speed-test protocol, small-batch A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=10 -s > gpurun_out/${T}_gpu_tests_full.log 2>&1; echo "all gpu tests rc=$?"
grep -v "^\[PyEye\]\|^WARNING\|^ERROR: Unable" gpurun_out/${T}_gpu_tests_full.log | tail -18 > gpurun_out/${T}_gpu_tests.log; cat gpurun_out/${T}_gpu_tests.log | cut -c1-200
cat gpurun_out/viewer_screenshot_ieee_vs_fast_math.txt
M="dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum"
ARGS=""
for spec in fused:20 fused:40 ordered:20 ordered:34 fused_fast:20 fused_fast:40; do
  mode=${spec%%:*}; F=${spec##*:}
  timeout 600 ncu --clock-control none --metrics $M -k regex:k_traceCompound -s 1 -c 1 -f -o gpurun_out/${T}_k1_${mode}_F${F} \
     python bench.py --mode $mode --steps $F --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_ncu_${mode}_F${F}.log 2>&1
  echo "ncu $spec rc=$?"
  ARGS="$ARGS ${mode}:${F}:gpurun_out/${T}_k1_${mode}_F${F}.ncu-rep"
done
python profiles/make_k1_traffic.py $ARGS > gpurun_out/${T}_k1_traffic.txt 2>&1; cat gpurun_out/${T}_k1_traffic.txt
cp profiles/k1_traffic.json gpurun_out/${T}_k1_traffic.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_traceCompound -s 1 -c 1 -f -o gpurun_out/${T}_k1_batched_full \
   python bench.py --mode fused --steps 20 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_ncu_full_batched.log 2>&1; echo "ncu full batched rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_traceCompound -s 8 -c 1 -f -o gpurun_out/${T}_k1_perframe_full \
   python bench.py --mode fused --steps 20 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_ncu_full_perframe.log 2>&1; echo "ncu full per-frame rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_traceCompound -s 4 -c 1 -f -o gpurun_out/${T}_k1_groups_cfg5_full \
   python benchmarks/pose_batch.py --poses 4096 --samples 64 --native --chunk 2048 --mode fused > gpurun_out/${T}_ncu_full_groups.log 2>&1; echo "ncu full frame groups rc=$?"
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.log; echo "bench rc=$?"
cut -c1-300 gpurun_out/${T}_bench.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_steps20.json 2>> gpurun_out/${T}_bench.log; echo "bench20 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.log; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
   python bench.py --steps 20 --warmup 3 --repeats 2 --no-cpu-baseline --no-modes > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 python benchmarks/configs.py --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1; echo "configs rc=$?"
timeout 600 python benchmarks/speed_test.py > gpurun_out/${T}_speed_test_protocol.txt 2>&1; echo "speed test rc=$?"
grep -h "^ *S=" gpurun_out/${T}_speed_test_protocol.txt | cut -c1-100
timeout 600 python benchmarks/small_batch_ab.py 8192 > gpurun_out/${T}_small_batch_ab.log 2>&1; grep "^{" gpurun_out/${T}_small_batch_ab.log | cut -c1-230
timeout 900 python benchmarks/pose_batch.py --poses 100000 --samples 64 --native --chunk 2048 --mode fused > gpurun_out/${T}_pose_batch_100k_native_1gpu.json 2> gpurun_out/${T}_pose_batch.log; cut -c1-420 gpurun_out/${T}_pose_batch_100k_native_1gpu.json
ls gpurun_out | grep ${T} | wc -l
