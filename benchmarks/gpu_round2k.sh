#!/bin/bash
# Overlapped frontier pass: tests, A/B in one process, bench; racecheck again after the traceList fix.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r03c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -4 gpurun_out/${T}_gpu_tests.log
run() { name=$1; shift
  env "$@" timeout 600 python benchmarks/wavefront_sweep.py --modes 1:0 --refills 12 --node-lanes 8 --out gpurun_out/${T}_ab_${name}.json > gpurun_out/${T}_ab_${name}.log 2>&1; echo "$name rc=$?"; }
run overlap CR_X=1
run nooverlap CR_OVERLAP_FRONTIER=0
run overlap2 CR_X=1
run nooverlap2 CR_OVERLAP_FRONTIER=0
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.log; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_bench.json'));print(d['value']/1e9, d['e2e']['value']/1e9, {k:(round(v['rays_per_sec']/1e9,2), round(v['e2e_rays_per_sec']/1e9,2)) for k,v in d['modes'].items() if isinstance(v,dict)})"
CR_OVERLAP_FRONTIER=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/${T}_bench_nooverlap.json 2>> gpurun_out/${T}_bench.log
python -c "import json;d=json.load(open('gpurun_out/${T}_bench_nooverlap.json'));print('no overlap:', d['value']/1e9, d['e2e']['value']/1e9)"
timeout 1500 compute-sanitizer --tool racecheck python compound-ray_b200/tools/sanitize_run.py > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -h "RACECHECK SUMMARY\|tour done" gpurun_out/${T}_sanitizer_racecheck.log | tail -3
