#!/usr/bin/env bash
# Builds the reference's own device hot path (ComputeKingKernel + Submatrix + KingResult) for sm_100a.
#
# The reference (cuking.cu) cannot be built whole here: it needs Abseil C++, google-cloud-cpp and nlohmann-json,
# none of which are installed, and its Run() only accepts gs:// URIs.  Its hot path, however, is
# self-contained: /root/reference/cuking.cu:100-314.  This recipe slices those lines OUT OF THE REFERENCE WHERE IT
# LIES into the git-ignored oracle/_ref/ (never into tracked source) and compiles them behind oracle/ref_harness.cu.
# Outputs: oracle/_ref/cuking_ref_extract.cuh (generated), oracle/_ref/libcuking_ref.so.
# The .so travels to the GPU box with gpurun (oracle/_ref/ is git-ignored, not gpurun-ignored).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CUKING_REFERENCE:-/root/reference}/cuking.cu"
OUT="$HERE/_ref"
if [[ ! -f "$REF" ]]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT/libcuking_ref.so if any" >&2
  exit 0
fi
mkdir -p "$OUT"
# from "// Custom deleter for RAII-style CUDA-managed array." up to (not including) "// Atomically clears a bit"
awk '/^\/\/ Custom deleter for RAII-style CUDA-managed array\./{on=1} /^\/\/ Atomically clears a bit in a bit set\./{on=0} on' \
  "$REF" > "$OUT/cuking_ref_extract.cuh"
grep -q "__global__ void ComputeKingKernel" "$OUT/cuking_ref_extract.cuh"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Wno-deprecated-declarations \
  -Xcompiler -fPIC -shared -I "$OUT" -o "$OUT/libcuking_ref.so" "$HERE/ref_harness.cu"
echo "build_ref: built $OUT/libcuking_ref.so"
