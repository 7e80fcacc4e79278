#!/usr/bin/env bash
# Build libcge.so (the C-ABI library, include/cge.h) for sm_100a, in-tree.
#   -fmad=false            : no FMA contraction anywhere in device code (bit parity with the reference)
#   -prec-div/-prec-sqrt   : IEEE division and square root (nvcc defaults, stated explicitly)
#   -Xcompiler -ffp-contract=off : same guarantee for the host-side triangle precompute
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
PKG=$(cd "$HERE/.." && pwd)
REPO=$(cd "$PKG/.." && pwd)
EXTRA=${CGE_NVCC_EXTRA:-}
OUT=${CGE_OUT:-$PKG/libcge.so}
LOG=${OUT%.so}_ptxas.log   # registers / spills / shared memory per kernel (-Xptxas -v)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 ${CGE_LINEINFO--lineinfo} -std=c++17 \
    -fmad=false -prec-div=true -prec-sqrt=true \
    -Xcompiler -fPIC,-ffp-contract=off,-O2,-Wall,-Wno-unused-function,-pthread \
    -Xptxas -v $EXTRA \
    -I"$REPO/include" -I"$HERE" \
    -shared -o "$OUT" "$HERE/cge_api.cu" "$HERE/bvh_sah_gpu.cu" "$HERE/bvh_build.cpp" "$HERE/bvh_sah.cpp" -ldl 2> "$LOG" || { cat "$LOG" >&2; exit 1; }
grep -E "error|warning" "$LOG" | grep -v "ptxas info" | head -20 >&2 || true
echo "built $OUT"
