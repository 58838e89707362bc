// Live measurement of the integer-pipe peaks the pairwise kernel's roofline is quoted against (the
// bounding resource is the POPC issue rate, which MEASURED_PEAKS.json does not contain).  Same method as
// tools/int_pipe_peaks.cu: 8 independent dependent chains per thread, 2 CTAs x 256 threads per SM, best of 5.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "internal.cuh"
#include "king_common.cuh"
#include "umma_common.cuh"

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

// ---- dense kind::mxf4 tensor rate: the denominator of the mxf4 pairwise kernel's roofline ------------------------------
// Two issuer lanes per CTA, each streaming tcgen05.mma.kind::mxf4 (M = 128, N = 208, K = 64, A from TMEM, B from shared
// memory, unit block scales) into its own accumulator, one CTA per SM, operands resident - nothing but the tensor pipe
// can limit it.  This is the best configuration tools/umma_mxf4_probe.cu found on B200 (profiles/r01_mxf4_probe.txt).
constexpr uint32_t kRateN = 208, kRateSteps = 20000, kRateKBytes = 128, kRateLBO = 128, kRateSBO = (kRateKBytes / 16) * 128;
constexpr uint32_t kRateColA = 416, kRateColSF = 480;

// genotype-like operand words: eight E2M1 nibbles drawn from {0, 0.5, +1, -1} by a hash of the index
__device__ __forceinline__ uint32_t rate_operand(uint32_t i, uint32_t random) {
  if (!random) return 0x22222222u;  // all +1
  uint32_t w = 0, x = i * 0x9e3779b9u + 0x7f4a7c15u;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12;
    const uint32_t c = (x >> 7) & 3u;
    w |= (c == 0 ? 0x0u : c == 1 ? 0x1u : c == 2 ? 0x2u : 0xAu) << (4 * q);
  }
  return w;
}

__global__ void __launch_bounds__(128) fp4_rate_kernel(uint32_t steps, uint32_t random_operands) {
  __shared__ __align__(1024) uint8_t smem[256 * kRateKBytes];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 2);
    mbar_fence_init();
  }
  for (uint32_t e = tid; e < 256 * kRateKBytes / 4; e += blockDim.x) reinterpret_cast<uint32_t *>(smem)[e] = rate_operand(e, random_operands);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t lane_base = tmem_base + ((warp * 32u) << 16);
  uint32_t v[8];
#pragma unroll
  for (uint32_t q = 0; q < 8; ++q) v[q] = 0x7f7f7f7fu;
  for (uint32_t c = 0; c < 32; c += 8) tmem_store8(lane_base + kRateColSF + c, v);
  for (uint32_t c = 0; c < 64; c += 8) {
#pragma unroll
    for (uint32_t q = 0; q < 8; ++q) v[q] = rate_operand(0x10000u + tid * 64u + c + q, random_operands);
    tmem_store8(lane_base + kRateColA + c, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  if ((tid & 31) == 0 && warp < 2) {
    const uint32_t idesc = (1u << 7) | (1u << 10) | ((kRateN >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    const uint32_t d = tmem_base + warp * kRateN, sf = tmem_base + kRateColSF;
    for (uint32_t st = 0; st < steps; ++st) {
      const uint32_t ks = st & 3;
      const uint64_t db = umma_smem_desc(smem_u32(smem) + ks * 2 * kRateLBO, kRateLBO, kRateSBO);
      const uint32_t acc = st > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%5], p;\n\t"
          "}\n" ::"r"(d),
          "r"(tmem_base + kRateColA + ks * 8 + warp * 8), "l"(db), "r"(idesc), "r"(acc), "r"(sf)
          : "memory");
    }
    umma_commit_arrive(&bar);
  }
  __syncwarp();
  mbar_wait_suspend(&bar, 0);
  tcgen05_after_sync();
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

}  // namespace
}  // namespace ck

extern "C" int ck_measure_fp4_peak(ck_ctx *ctx, double *ops_per_s) {
  using namespace ck;
  if (!ctx || !ops_per_s) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK_CUDA(cudaEventRecord(ctx->ev[0], s));
    fp4_rate_kernel<<<ctx->num_sms, 128, 0, s>>>(kRateSteps, 0u);
    CK_CUDA(cudaGetLastError());
    CK_CUDA(cudaEventRecord(ctx->ev[1], s));
    CK_CUDA(cudaEventSynchronize(ctx->ev[1]));
    if (rep >= 1) best = std::min(best, elapsed_ms(ctx->ev[0], ctx->ev[1]));
  }
  // 2 issuers x 128 x 208 x 64 MACs per step and SM; 2 ops per MAC
  *ops_per_s = 2.0 * 2.0 * 128.0 * kRateN * 64.0 * double(kRateSteps) * ctx->num_sms / (double(best) * 1e-3);
  return CK_OK;
}

extern "C" int ck_measure_fp4_peak_sustained(ck_ctx *ctx, double seconds, double *ops_per_s) {
  using namespace ck;
  if (!ctx || !ops_per_s || !(seconds > 0.0) || seconds > 30.0) return fail(CK_ERR_INVALID_ARGUMENT, "bad argument");
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  // one launch takes about 2.3 ms; run them back to back for `seconds` and time the second half, when the board has
  // settled at whatever clock its power limit allows under a saturated tensor pipe
  // operands are genotype-like random E2M1 values here: the tensor core's power depends on how its operand bits toggle
  CK_CUDA(cudaEventRecord(ctx->ev[0], s));
  fp4_rate_kernel<<<ctx->num_sms, 128, 0, s>>>(kRateSteps, 1u);
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  CK_CUDA(cudaEventSynchronize(ctx->ev[1]));
  const double one_ms = std::max(0.1, double(elapsed_ms(ctx->ev[0], ctx->ev[1])));
  const int launches = std::max(8, int(seconds * 1e3 / one_ms)), half = launches / 2;
  for (int i = 0; i < launches; ++i) {
    if (i == half) CK_CUDA(cudaEventRecord(ctx->ev[0], s));
    fp4_rate_kernel<<<ctx->num_sms, 128, 0, s>>>(kRateSteps, 1u);
  }
  CK_CUDA(cudaGetLastError());
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  CK_CUDA(cudaEventSynchronize(ctx->ev[1]));
  const double ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
  *ops_per_s = 2.0 * 2.0 * 128.0 * kRateN * 64.0 * double(kRateSteps) * ctx->num_sms * double(launches - half) / (ms * 1e-3);
  return CK_OK;
}

extern "C" int ck_measure_int_peaks(ck_ctx *ctx, double *popc_lane_ops_per_s, double *lop3_lane_ops_per_s) {
  using namespace ck;
  if (!ctx || !popc_lane_ops_per_s || !lop3_lane_ops_per_s) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  uint32_t *d_out = nullptr;
  cudaError_t e = dev_alloc(ctx, reinterpret_cast<void **>(&d_out), size_t(ctx->num_sms) * 2 * 256 * 4);
  if (e == cudaSuccess) e = measure<0>(ctx, d_out, popc_lane_ops_per_s);
  if (e == cudaSuccess) e = measure<1>(ctx, d_out, lop3_lane_ops_per_s);
  if (d_out) cudaFree(d_out);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) return fail_cuda(e, "ck_measure_int_peaks", __FILE__, __LINE__);
  return CK_OK;
}
