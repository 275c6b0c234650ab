#!/bin/bash
# Build libvican_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libvican_b200.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared "$HERE/vican_b200.cu" -o "$OUT" "$@"
echo "built $OUT"
