#!/bin/bash
# Builds liblyftvoxel_b200.so in-tree for sm_100a (B200) only.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
OBJS=""
for f in lv_api lv_bev lv_draw lv_voxel lv_filter lv_pillar lv_ingest lv_png lv_half lv_pfn_train; do
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ lv_common.cuh -nt $f.o ] || [ lv_decorate.cuh -nt $f.o ] || [ ../../include/lyft_voxel.h -nt $f.o ]; then
    $NVCC $FLAGS ${LV_PTXAS_V:+-Xptxas -v} -c $f.cu -o $f.o &
  fi
  OBJS="$OBJS $f.o"
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../liblyftvoxel_b200.so $OBJS -cudart static
echo "built $(cd .. && pwd)/liblyftvoxel_b200.so"
