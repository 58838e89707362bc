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

// Threshold + warp-aggregated append of one pair per lane (cuking.cu:297-312).  Must be called by all 32 lanes of a
// warp.  One atomicAdd per warp and call instead of one per retained pair; the 64-bit counter cannot wrap, and the
// host turns counter > max_results into the reference's overflow error (:747-751).
__device__ __forceinline__ void emit_pair(const KingLaunch &p, bool valid, uint32_t gi, uint32_t gj, float kin,
                                          uint32_t opp, uint32_t conc, uint32_t both_het, uint32_t shared) {
  const uint32_t lane = threadIdx.x & 31;
  const bool emit = valid && (kin > p.kin_threshold);  // strict; NaN / -inf never pass (cuking.cu:297)
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
