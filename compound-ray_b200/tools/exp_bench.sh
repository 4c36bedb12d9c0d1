run() { echo -n "$1 :: "; env $1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('Grays/s %.3f  e2e %.3f  nodes/ray %.2f tris/ray %.2f  bvh_ms %.1f' % (d['value']/1e9, d['e2e']['value']/1e9, r['nodes_per_ray'], r['tris_per_ray'], d['bvh_build_ms']))"; }
run "CR_X=0"
run "CR_MORTON_PER_AXIS=1"
run "CR_LEAF_SIZE=1"
run "CR_LEAF_SIZE=2"
run "CR_LEAF_SIZE=8"
run "CR_LIB_PATH=/root/repo/compound-ray_b200/lib/libEyeRenderer3_mb10.so"
run "CR_LIB_PATH=/root/repo/compound-ray_b200/lib/libEyeRenderer3_mb12.so"
