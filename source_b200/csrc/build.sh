#!/bin/sh
# Builds libraysect_b200.so for sm_100a in-tree (source_b200/libraysect_b200.so).
# -fmad=false: the reference's x86-64 build contains no FMA, and parity of t / hit ids requires
# that products and sums round separately.
# The CUDA translation unit (3 minutes of ptxas) is kept as an object file next to the output and rebuilt only
# when a .cu/.cuh/.h it includes (or the flags) changed; the two host-only .cpp files are compiled every time.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${RSB_OUT:-$HERE/../libraysect_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OBJDIR="${RSB_OBJDIR:-$HERE/../../build/obj}"
mkdir -p "$OBJDIR"
TAG="$(printf '%s' "${RSB_NVCC_EXTRA}" | cksum | cut -d' ' -f1)"
CUOBJ="$OBJDIR/raysect_b200_$TAG.o"
FLAGS="-std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2,-ffp-contract=off,-pthread ${RSB_NVCC_EXTRA}"
stale=0
[ -f "$CUOBJ" ] || stale=1
for f in "$HERE"/*.cu "$HERE"/*.cuh "$HERE"/*.h "$HERE"/../../include/*.h "$HERE/build.sh"; do
    [ "$stale" = 1 ] && break
    [ "$f" -nt "$CUOBJ" ] && stale=1
done
if [ "$stale" = 1 ]; then
    "$NVCC" $FLAGS -c -o "$CUOBJ.tmp" "$HERE/raysect_b200.cu"
    mv "$CUOBJ.tmp" "$CUOBJ"
fi
"$NVCC" $FLAGS -c -o "$OBJDIR/kdtree_host_$TAG.o" "$HERE/kdtree_host.cpp"
"$NVCC" $FLAGS -c -o "$OBJDIR/scene_pack_$TAG.o" "$HERE/scene_pack.cpp"
"$NVCC" -shared -Xcompiler -pthread -gencode arch=compute_100a,code=sm_100a -o "$OUT.tmp" "$CUOBJ" "$OBJDIR/kdtree_host_$TAG.o" "$OBJDIR/scene_pack_$TAG.o"
mv "$OUT.tmp" "$OUT"
echo "built $OUT"
