#!/bin/bash
# same-box A/B of two builds of the library: tools/ab/lib_head.so vs tools/ab/lib_new.so
L=lyft-3d-object-detection_b200/liblyftvoxel_b200.so
cp $L /tmp/lib_keep.so
for round in 1 2; do
  for v in head new; do
    cp tools/ab/lib_$v.so $L
    python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v', d['ms_per_step'],d['stages']['bev']['ms_per_step'],d['stages']['pillarize']['ms_per_step'], d['stages']['scatter']['ms_per_step'], d['pillar_path_with_fused_pfn']['ms_per_step'])"
  done
done
cp /tmp/lib_keep.so $L
