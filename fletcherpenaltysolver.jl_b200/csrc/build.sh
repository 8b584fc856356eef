#!/usr/bin/env bash
# Builds libfpsb200.so in-tree for sm_100a (B200). No JIT cache, no torch dependency.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${FPSB_OUT:-$HERE/../libfpsb200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function $ARCH"
mkdir -p "$HERE/build"
# Krylov kernels: no FMA contraction so the scalar recurrences round like the CPU reference
"$NVCC" $COMMON -fmad=false ${FPSB_DEFS:-} ${PTXAS_V:+-Xptxas -v} -c "$HERE/fpsb_krylov.cu" -o "$HERE/build/fpsb_krylov.o"
"$NVCC" $COMMON ${FPSB_DEFS:-} ${PTXAS_V:+-Xptxas -v} -c "$HERE/fpsb_ldlt.cu" -o "$HERE/build/fpsb_ldlt.o"
"$NVCC" $COMMON -c "$HERE/fpsb_api.cu" -o "$HERE/build/fpsb_api.o"
EXTRA=""
for f in "$HERE"/fpsb_symbolic.cpp "$HERE"/fpsb_batch.cu "$HERE"/fpsb_fpnlp.cu; do
  if [ -f "$f" ]; then
    o="$HERE/build/$(basename "${f%.*}").o"
    # the batch kernel mirrors the reference's scalar LDL' operation by operation: no FMA contraction
    FM=""; case "$f" in *fpsb_batch.cu|*fpsb_fpnlp.cu) FM="-fmad=false";; esac
    "$NVCC" $COMMON $FM -c "$f" -o "$o"
    EXTRA="$EXTRA $o"
  fi
done
"$NVCC" $ARCH -shared -o "$OUT" -ldl "$HERE/build/fpsb_krylov.o" "$HERE/build/fpsb_ldlt.o" "$HERE/build/fpsb_api.o" $EXTRA -lcudart
echo "built $OUT"
