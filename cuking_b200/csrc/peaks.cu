// Live measurement of the integer-pipe peaks the pairwise kernel's roofline is quoted against (the
// bounding resource is the POPC issue rate, which MEASURED_PEAKS.json does not contain).  Same method as
// tools/int_pipe_peaks.cu: 8 independent dependent chains per thread, 2 CTAs x 256 threads per SM, best of 5.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "internal.cuh"

namespace ck {
namespace {

constexpr int kChains = 8;

template <int kOp>
__global__ void __launch_bounds__(256) pipe_peak_kernel(uint32_t *out, int iters) {
  uint32_t a[kChains];
  const uint32_t b = threadIdx.x * 2654435761u + blockIdx.x, c = b ^ 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < kChains; ++i) a[i] = b + i * 0x01010101u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        if (kOp == 0) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
        if (kOp == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kOp>
cudaError_t measure(ck_ctx *ctx, uint32_t *d_out, double *lane_ops_per_s) {
  const int blocks = ctx->num_sms * 2, threads = 256, iters = 4096;
  cudaStream_t s = ctx->stream;
  float best = 1e30f;
  for (int rep = 0; rep < 7; ++rep) {
    cudaEventRecord(ctx->ev[0], s);
    pipe_peak_kernel<kOp><<<blocks, threads, 0, s>>>(d_out, iters);
    cudaEventRecord(ctx->ev[1], s);
    cudaError_t e = cudaEventSynchronize(ctx->ev[1]);
    if (e != cudaSuccess) return e;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    if (rep >= 2) best = std::min(best, ms);
  }
  *lane_ops_per_s = double(iters) * 4 * kChains * threads * blocks / (double(best) * 1e-3);
  return cudaGetLastError();
}

}  // namespace
}  // namespace ck

extern "C" int ck_measure_int_peaks(ck_ctx *ctx, double *popc_lane_ops_per_s, double *lop3_lane_ops_per_s) {
  using namespace ck;
  if (!ctx || !popc_lane_ops_per_s || !lop3_lane_ops_per_s) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  uint32_t *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_out, size_t(ctx->num_sms) * 2 * 256 * 4);
  if (e == cudaSuccess) e = measure<0>(ctx, d_out, popc_lane_ops_per_s);
  if (e == cudaSuccess) e = measure<1>(ctx, d_out, lop3_lane_ops_per_s);
  if (d_out) cudaFree(d_out);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) return fail_cuda(e, "ck_measure_int_peaks", __FILE__, __LINE__);
  return CK_OK;
}
