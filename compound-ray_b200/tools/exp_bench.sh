# usage: bash compound-ray_b200/tools/exp_bench.sh "ENV=.. ENV2=.." ...   (one bench run per argument)
run() { echo -n "$1 :: "; env $1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('Grays/s %.3f  e2e %.3f  nodes/ray %.2f (root %.2f) tris/ray %.2f  bvh_ms %.1f' % (d['value']/1e9, d['e2e']['value']/1e9, r['nodes_per_ray'], r['nodes_per_ray_from_root'], r['tris_per_ray'], d['bvh_build_ms']))"; }
if [ $# -eq 0 ]; then run "CR_X=0"; else for a in "$@"; do run "$a"; done; fi
