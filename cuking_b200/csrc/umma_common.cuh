// tcgen05 / TMEM PTX wrappers and the tile rasterisation shared by the two tensor-core pairwise kernels
// (king_umma_kernel.cu: kind::i8, king_fp4_kernel.cu: kind::mxf4).  sm_100a only.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "king_common.cuh"

namespace ck {

// K-major, no-swizzle canonical operand tile: 8-row x 16-byte core matrices, K-adjacent core matrices LBO bytes
// apart, 8-row groups SBO bytes apart; descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((smem_addr >> 4) & 0x3fff) | (uint64_t((lbo_bytes >> 4) & 0x3fff) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3fff) << 32) | (uint64_t(1) << 46);
}
// arrives on the mbarrier once every tcgen05.mma this thread has issued so far has completed
__device__ __forceinline__ void umma_commit_arrive(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_arrive_addr(uint32_t bar) {  // by 32-bit shared address
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_load16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_load4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_store8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 16-byte shared-memory store by 32-bit shared address.  Through a generic pointer the compiler emitted ST.E.128 and
// split some of the stores into 32- and 64-bit pieces (22 % extra wavefronts, profiles/r01_king_fp4_cfg2_ncu.txt).
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tcgen05_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t elected;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(elected));
  return elected;
}

// Linear tile index -> (row tile, column tile) of the band enumeration (band_tiles.cu).
__device__ __forceinline__ void band_decode(const BandTiles &tiles, unsigned long long t, uint32_t &ti, uint32_t &tj) {
  uint32_t lo = 0, hi = tiles.num_bands;  // largest b with band_prefix[b] <= t
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (tiles.band_prefix[mid] <= t) lo = mid; else hi = mid;
  }
  const uint32_t band_rows = band_rows_padded(tiles.num_row_tiles - lo * kBandRowTiles);
  const uint32_t q = uint32_t(t - tiles.band_prefix[lo]);
  ti = lo * kBandRowTiles + q % band_rows;
  tj = tiles.band_first_col[lo] + q / band_rows;
}

}  // namespace ck
