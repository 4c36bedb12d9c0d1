#!/bin/bash
# One gpurun call of round 2: GPU tests, mode matrix, ncu captures for profiles/k1_traffic.json, bench line, launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r02c}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_scripts.py tests/test_gpu_bench_size.py tests/test_gpu_modes.py -q -m gpu --durations=12 > gpurun_out/${T}_new_tests.log 2>&1; echo "new tests rc=$?"
tail -25 gpurun_out/${T}_new_tests.log
timeout 900 python benchmarks/mode_matrix.py --out gpurun_out/${T}_mode_matrix.json > gpurun_out/${T}_mode_matrix.log 2>&1; echo "matrix rc=$?"
grep grays_device gpurun_out/${T}_mode_matrix.log | cut -c1-230
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
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.log; echo "bench rc=$?"
cut -c1-300 gpurun_out/${T}_bench.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_steps20.json 2>> gpurun_out/${T}_bench.log; echo "bench20 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
   python bench.py --steps 20 --warmup 3 --repeats 2 --no-cpu-baseline --no-modes > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/${T}_gpu_tests.log
ls -la gpurun_out | tail -30
