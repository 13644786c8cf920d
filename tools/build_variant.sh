#!/bin/bash
# Builds a variant of libganffn.so into gan_ffn_b200/<name> without touching the regular objects.
# usage: tools/build_variant.sh libganffn_trace.so -DGANFFN_TC_TRACE [more nvcc flags]
set -e
out=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/gan_ffn_b200/csrc
obj=/tmp/ganffn_variant_$(basename "$out" .so)
mkdir -p "$obj"
ARCH="-gencode arch=compute_100a,code=sm_100a"
pids=()
for f in capi gemm_simt gemm_tc attention attention_mma rowwise losses head net graph; do
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr "$@" \
    -c "$csrc/$f.cu" -o "$obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc $ARCH -shared -o "$root/gan_ffn_b200/$out" "$obj"/*.o -lcudart
echo "built gan_ffn_b200/$out"
