// Device helpers shared by the pairwise kernels: mbarrier / bulk-copy PTX wrappers, the kinship expression and the
// warp-aggregated result append.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.cuh"

namespace ck {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Same, but each probe may suspend the thread in hardware for up to `hint_ns` (it still wakes as soon as the phase
// completes): a long wait costs a handful of probes instead of one every ~30 clocks — issue slots and power that the
// producer warps sharing the scheduler can use.
__device__ __forceinline__ void mbar_wait_suspend(uint64_t *bar, uint32_t parity, uint32_t hint_ns = 20000u) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
  }
}
// The same operations by 32-bit shared address.  Hot loops compute the address of barrier i as base + 8 i once;
// going through a generic pointer makes the compiler rebuild the shared window (S2UR SR_CgaCtaId + LEA) at every use.
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_suspend_addr(uint32_t bar, uint32_t parity, uint32_t hint_ns = 20000u) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
  }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); completion is signalled on the mbarrier.
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// kinship exactly as the reference evaluates it (cuking.cu:286-294; SASS order FADD(bh,bh), FFMA(opp,-4,.),
// FADD, FADD, FMUL(min,4), IEEE division, FADD 0.5).  All integer->float conversions round to nearest.
__device__ __forceinline__ float kinship(uint32_t het_i, uint32_t het_j, uint32_t both_het, uint32_t opp) {
  const uint32_t min_hets = het_i < het_j ? het_i : het_j;
  const float bh = __uint2float_rn(both_het);
  float num = __fmaf_rn(__uint2float_rn(opp), -4.f, __fadd_rn(bh, bh));
  num = __fsub_rn(num, __uint2float_rn(het_i));
  num = __fsub_rn(num, __uint2float_rn(het_j));
  const float den = __fmul_rn(4.f, __uint2float_rn(min_hets));
  return __fadd_rn(0.5f, __fdiv_rn(num, den));
}

// Rows per band of the tensor-core kernels' tile enumeration (band_tiles.cu: kBandRowTiles x kBandTileRows).
constexpr uint32_t kDenseBandRows = 1024;
static_assert(kDenseBandRows == kBandRowTiles * kBandTileRows, "band height");

// Offset of pair (li, lj) - row / column index inside the sub-matrix - among the pairs of its band, in the sorted
// (row-major) order: a triangular row li holds the columns (li, num_cols), a rectangular one all of them.
__host__ __device__ inline unsigned long long dense_offset_in_band(uint32_t li, uint32_t lj, uint32_t num_cols, bool triangular) {
  const unsigned long long b0 = (li / kDenseBandRows) * kDenseBandRows, r = li - b0;
  if (!triangular) return r * num_cols + lj;
  return r * (num_cols - 1ull) - (b0 * r + r * (r - 1ull) / 2ull) + (lj - li - 1ull);
}
// pairs of the band that starts at row b0
__host__ __device__ inline unsigned long long dense_band_pairs(uint32_t b0, uint32_t num_rows, uint32_t num_cols, bool triangular) {
  const unsigned long long rows = (num_rows - b0 < kDenseBandRows) ? num_rows - b0 : kDenseBandRows;
  if (!triangular) return rows * num_cols;
  return rows * (num_cols - 1ull) - (b0 * rows + rows * (rows - 1ull) / 2ull);
}

// Threshold + result store of one pair per lane (cuking.cu:297-312).  Must be called by all 32 lanes of a warp.
//   pair : this lane holds an i < j pair of the sub-matrix;  maybe : it may pass the threshold (a cheap screen said so)
// Sparse mode: one atomicAdd per warp and call instead of one per retained pair (warp-aggregated append); the 64-bit
// counter cannot wrap, and the host turns counter > max_results into the reference's overflow error (:747-751).
// Dense mode (p.dense_band_base): the record goes straight to its slot in the sorted output.
__device__ __forceinline__ void emit_pair(const KingLaunch &p, bool pair, bool maybe, uint32_t gi, uint32_t gj, float kin,
                                          uint32_t opp, uint32_t conc, uint32_t both_het, uint32_t shared) {
  const uint32_t lane = threadIdx.x & 31;
  const bool emit = pair && maybe && (kin > p.kin_threshold);  // strict; NaN / -inf never pass (cuking.cu:297)
  if (p.dense_band_base != nullptr) {
    const uint32_t holes = __ballot_sync(0xffffffffu, pair && !emit);
    if (holes != 0 && int(lane) == __ffs(holes) - 1) atomicAdd(p.holes, (unsigned long long)__popc(holes));
    if (pair) {
      const uint32_t li = gi - p.row_global0, lj = gj - p.col_global0;
      const unsigned long long slot = p.dense_band_base[li / kDenseBandRows] + dense_offset_in_band(li, lj, p.num_cols, p.triangular != 0);
      uint2 *dst = reinterpret_cast<uint2 *>(p.results + slot);  // our own 256-byte aligned buffer: 24-byte records are 8-aligned
      if (emit) {
        const uint32_t ibs2 = conc + both_het;
        dst[0] = make_uint2(gi, gj);
        dst[1] = make_uint2(__float_as_uint(kin), opp);
        dst[2] = make_uint2(shared - opp - ibs2, ibs2);
      } else {
        dst[0] = make_uint2(0xffffffffu, 0xffffffffu);
      }
    }
    return;
  }
  const uint32_t ballot = __ballot_sync(0xffffffffu, emit);
  if (ballot == 0) return;
  const int leader = __ffs(ballot) - 1;
  unsigned long long base = 0;
  if (int(lane) == leader) base = atomicAdd(p.counter, (unsigned long long)__popc(ballot));  // :299, once per warp
  base = __shfl_sync(0xffffffffu, base, leader);
  if (emit) {
    const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
    if (slot < p.max_results) {  // :300
      ck_result res;
      res.sample_i = gi;
      res.sample_j = gj;
      res.kin = kin;
      res.ibs0 = opp;                           // :305
      res.ibs2 = conc + both_het;               // :306
      res.ibs1 = shared - res.ibs0 - res.ibs2;  // :307
      p.results[slot] = res;
    }
  }
}

}  // namespace ck
