#!/bin/bash
# Evidence refresh for the final round-2 tree: all GPU tests, ncu captures for profiles/k1_traffic.json, full ncu of the batched and
# per-frame trace kernels, bench lines, launch list, configs, speed-test protocol.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -14 gpurun_out/${T}_gpu_tests.log | cut -c1-160
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
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
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.log; echo "bench rc=$?"
cut -c1-200 gpurun_out/${T}_bench.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_steps20.json 2>> gpurun_out/${T}_bench.log; echo "bench20 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
   python bench.py --steps 20 --warmup 3 --repeats 2 --no-cpu-baseline --no-modes > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 python benchmarks/configs.py --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1; echo "configs rc=$?"
timeout 600 python benchmarks/speed_test.py > gpurun_out/${T}_speed_test_protocol.txt 2>&1; echo "speed test rc=$?"
grep -h "^ *S=" gpurun_out/${T}_speed_test_protocol.txt | cut -c1-100
