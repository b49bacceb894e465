#!/bin/bash
# Build libmmunet_b200.so (sm_100a only) in-tree.  Usage: mm-unet_b200/csrc/build.sh [-j]
set -e
cd "$(dirname "$0")"
OUT=../mmunet_b200/libmmunet_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -diag-suppress 128 -Xcompiler -fPIC"
mkdir -p build
pids=()
for f in capi selective_scan_fwd selective_scan_bwd causal_conv1d scan_order snake_sample group_norm; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ scan_tiles.cuh -nt build/$f.o ] || [ scan3.cuh -nt build/$f.o ] || [ scan3_fwd.cuh -nt build/$f.o ] || [ scan3_bwd.cuh -nt build/$f.o ] || [ scan4.cuh -nt build/$f.o ] || [ scan4_bwd.cuh -nt build/$f.o ] || [ ../../include/mmunet_b200.h -nt build/$f.o ]; then
    nvcc $FLAGS -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT build/*.o
echo "built $OUT"
