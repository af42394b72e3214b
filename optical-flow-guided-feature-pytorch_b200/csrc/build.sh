#!/usr/bin/env bash
# Build liboffk.so for sm_100a in-tree (the .so travels to the GPU box with the repo snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../liboffk.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v
       -I"$ROOT/include" -I"$HERE")
OBJS=()
for f in offk_api offk_gemm_simt offk_gemm_tc offk_stencil offk_head; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/$f.o" 2> "$HERE/$f.ptxas.log" || { cat "$HERE/$f.ptxas.log" >&2; exit 1; }
  OBJS+=("$HERE/$f.o")
done
"$NVCC" -shared -o "$OUT" "${OBJS[@]}" -gencode arch=compute_100a,code=sm_100a -lcudart
echo "built $OUT"
