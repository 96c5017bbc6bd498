#!/bin/sh
# Builds libraysect_b200.so for sm_100a in-tree (source_b200/libraysect_b200.so).
# -fmad=false: the reference's x86-64 build contains no FMA, and parity of t / hit ids requires
# that products and sums round separately.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${RSB_OUT:-$HERE/../libraysect_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -lineinfo -fmad=false \
    -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-O2,-ffp-contract=off -shared \
    ${RSB_NVCC_EXTRA} \
    -o "$OUT" "$HERE/raysect_b200.cu" "$HERE/kdtree_host.cpp" "$HERE/scene_pack.cpp"
echo "built $OUT"
