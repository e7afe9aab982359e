#!/bin/bash
# Build libb200gen.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../lib"
mkdir -p "$out" "$here/.obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
pids=()
for f in core igemm norm elementwise vq attention_small flash_attn decode; do
  if [ ! -f "$here/.obj/$f.o" ] || [ "$here/$f.cu" -nt "$here/.obj/$f.o" ] || [ "$here/common.cuh" -nt "$here/.obj/$f.o" ] || [ "$here/../../include/b200gen.h" -nt "$here/.obj/$f.o" ]; then
    $NVCC $FLAGS -c "$here/$f.cu" -o "$here/.obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -shared -o "$out/libb200gen.so" "$here"/.obj/*.o -Xcompiler -fPIC
echo "built $out/libb200gen.so"
