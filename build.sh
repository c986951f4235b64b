#!/bin/bash
# Build libvinet_b200.so (sm_100a only) in-tree: the .so travels to the GPU box with the snapshot.
set -e
cd "$(dirname "$0")"
SRC=vinet_b200/csrc
OUT=vinet_b200/libvinet_b200.so
if [ "$1" = "--clean" ]; then rm -rf build "$OUT"; fi      # full recompile (VINET_CLEAN_BUILD=1 python -c "import __graft_entry__ as g; g.build()")
mkdir -p build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$SRC --expt-relaxed-constexpr ${VINET_NVCC_EXTRA}"
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ $SRC/common.cuh -nt "$o" ] || [ $SRC/gather.cuh -nt "$o" ] || [ $SRC/tc_ptx.cuh -nt "$o" ] || [ include/vinet_b200.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/*.o -lcudart
echo "built $OUT"
