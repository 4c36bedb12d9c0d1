#!/bin/bash
# Read-ahead for standing cameras: tests + speed-test protocol + bench sanity.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r02t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -12 gpurun_out/${T}_gpu_tests.log
timeout 600 python benchmarks/speed_test.py > gpurun_out/${T}_speed_test_protocol.txt 2>&1; echo "speed test rc=$?"
CR_READ_AHEAD=0 timeout 600 python benchmarks/speed_test.py --samples 1,8,32,128,1000,3200 --frames 300 > gpurun_out/${T}_speed_test_noreadahead.txt 2>&1; echo "speed test (no read-ahead) rc=$?"
grep -h "^ *S=" gpurun_out/${T}_speed_test_protocol.txt gpurun_out/${T}_speed_test_noreadahead.txt | cut -c1-100
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.log; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_bench.json'));print(d['value']/1e9, d['e2e']['value']/1e9)"
