#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CR_SANITIZE_SECTION=affine
for tool in memcheck racecheck initcheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool python compound-ray_b200/tools/sanitize_run.py > gpurun_out/r04r_sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tour" gpurun_out/r04r_sanitize_$tool.log | tail -3
done
timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-modes 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value']/1e9,2), round(d['e2e']['value']/1e9,2), d['clocks'])"
