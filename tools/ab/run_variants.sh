#!/bin/bash
# same-box A/B of several builds of the library: tools/ab/lib_<name>.so, names given as arguments.
# A name of the form name@ENV=VAL runs lib_<name>.so with that environment variable set.
L=lyft-3d-object-detection_b200/liblyftvoxel_b200.so
cp $L /tmp/lib_keep.so
for round in 1 2; do
  for spec in "$@"; do
    v=${spec%%@*}; e=""
    if [[ "$spec" == *@* ]]; then e=${spec#*@}; fi
    cp tools/ab/lib_$v.so $L
    env $e python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-verify --no-eager-ref 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('$spec', 'step', d['ms_per_step'],'bev',d['stages']['bev']['ms_per_step'],'pillarize',d['stages']['pillarize']['ms_per_step'], 'scatter', d['stages']['scatter']['ms_per_step'], 'pfn', d['pillar_path_with_fused_pfn']['ms_per_step'])"
  done
done
cp /tmp/lib_keep.so $L
