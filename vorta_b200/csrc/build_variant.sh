#!/bin/sh
# Perf-experiment builds of the library with extra -D flags, next to the product library:
#   sh build_variant.sh timeline -DVB_TIMELINE      ->  ../lib/exp/libvb_timeline.so   (use with VB_LIB_PATH=...)
set -e
cd "$(dirname "$0")"
name=$1; shift
mkdir -p ../lib/exp /tmp/vb_variant_$name
for f in vb_attn vb_kernels vb_block vb_api; do
  /usr/local/cuda/bin/nvcc "$@" -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -c $f.cu -o /tmp/vb_variant_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/exp/libvb_$name.so /tmp/vb_variant_$name/*.o -lcudart
echo built ../lib/exp/libvb_$name.so
