#!/bin/bash
# Build libemap_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
# Translation units compile in parallel (mlp_tc.cu with its 32 kernel instantiations dominates).
set -e
cd "$(dirname "$0")"
SRC=emap_b200/csrc
OUT=emap_b200/lib
mkdir -p $OUT build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
objs=""
pids=""
names=""
for f in cabi pack mlp_tc mlp_rg mlp_rev mlp_dw rays mlp_bwd extract rendering; do
  [ -f $SRC/$f.cu ] || continue
  if [ ! -f build/$f.o ] || [ $SRC/$f.cu -nt build/$f.o ] || [ $SRC/common.cuh -nt build/$f.o ] || [ $SRC/mlp_dev.cuh -nt build/$f.o ] || [ $SRC/host.h -nt build/$f.o ] || [ include/emap_b200.h -nt build/$f.o ]; then
    extra=""
    [ $f = rays ] && extra="-fmad=false"
    [ $f = extract ] && extra="-fmad=false"
    echo "nvcc $f.cu"
    ( $NVCC $FLAGS $extra -c $SRC/$f.cu -o build/$f.o.tmp 2> build/$f.log && mv build/$f.o.tmp build/$f.o ) &
    pids="$pids $!"
    names="$names $f"
  fi
  objs="$objs build/$f.o"
done
i=0
set -- $names
for p in $pids; do
  i=$((i+1))
  n=$(eval echo \${$i})
  if ! wait $p; then echo "nvcc $n.cu FAILED"; cat build/$n.log; rm -f build/$n.o.tmp; exit 1; fi
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libemap_b200.so $objs
echo "built $OUT/libemap_b200.so"
