#!/bin/bash
# ncu --set full of the per-frame frontier pass (second form, branch-free decisions)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r04v}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_buildEntries -s 9 -c 2 -f -o gpurun_out/${T}_k0b_full \
   python bench.py --mode fused --steps 20 --repeats 1 --no-cpu-baseline --no-modes > gpurun_out/${T}_ncu_k0b.log 2>&1; echo "ncu k0b rc=$?"
ncu -i gpurun_out/${T}_k0b_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:40], r[h.index('launch__grid_size')], r[h.index('gpu__time_duration.sum')], r[h.index('smsp__inst_executed.sum')])"
