// Pairwise KING on the 5th-generation tensor cores: tcgen05.mma.kind::i8 over indicator vectors (sm_100a only).
//
// The six counters of ComputeKingKernel (/root/reference/cuking.cu:214-240) are bilinear forms of per-sample indicator
// vectors over the sites.  With h = [het], y = [hom-ref or hom-alt], x = [hom-alt] - [hom-ref] (all 0 where the
// genotype is missing):
//     x_i.x_j = conc - opp     y_i.y_j = conc + opp     h_i.h_j = both_het
//     h_i.y_j = (i het, j hom) y_i.h_j = (i hom, j het)
//     het_i = hh + hy,  het_j = hh + yh,  shared = yy + yh + hy + hh,  conc = (yy + xx)/2,  opp = (yy - xx)/2
// i.e. five int8 GEMMs with exact s32 accumulation (counts <= num_sites < 2^31), issued as three MMAs per 32 sites:
//     D_xx[128 x 80]        += x_i  . x_j^T
//     D_y [128 x (80|80)]   += y_i  . [y_j ; h_j]^T        -> (yy | yh)
//     D_h [128 x (80|80)]   += h_i  . [y_j ; h_j]^T        -> (hy | hh)
// (h and y are stored as -1 instead of +1: every product pairs two of them, so the sign cancels.)
// One CTA owns a 128 (rows = TMEM lanes) x 80 (columns) tile of sample pairs: 400 TMEM columns of accumulators plus
// a 4-deep ring of A operands (3 x 8 columns per 32-site step) = 496 of the SM's 512 columns.
//
// The int8 operands are never stored in HBM (8x the packed genotypes; the kernel would turn L2-bound).  They are
// expanded on the fly from the 4-bit genotype codes (csrc/layout.cuh): a code nibble is used directly as a PRMT byte
// selector into an 8-byte value table, so 4 genotypes become 4 operand bytes in ONE instruction per operand.
//   * warps 0-7   A operands: one thread per row sample, the two groups of four warps take alternate 32-site steps,
//                 and write straight into TMEM with tcgen05.st (thread = TMEM lane = row) — no shared-memory traffic;
//   * warps 8-12  B operands: two threads per column sample write K-major no-swizzle canonical shared-memory tiles,
//                 four 32-site steps per stage so that the proxy fence and barrier round trip are amortised;
//   * one lane of each of warps 13-15 issues one of the three MMAs (A from TMEM, B from shared memory) and releases
//                 the A slot / B stage with tcgen05.commit.
// Measured on B200 (tools/umma_i8_probe.cu, profiles/r01_umma_probe.txt): kind::i8 peaks at 8192 MAC/clk/SM; one
// issuing thread sustains only one M=128 MMA per 100-160 clk whatever N is, three issuers reach the peak — hence three
// issuers.  The first version of this kernel (both operands in shared memory, bit planes expanded with shifts) ran at
// 81 % of the shared-memory pipe and 50 % tensor-pipe activity (profiles/r01_king_umma_v1_ncu.txt).
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "internal.cuh"
#include "king_common.cuh"
#include "umma_common.cuh"

namespace ck {

namespace {

constexpr uint32_t kUM = 128, kUN = 80;            // tile rows (A operand, TMEM lanes) x tile columns (B operand)
constexpr uint32_t kASlots = 4;                    // A-operand ring in TMEM: one 32-site step per slot
constexpr uint32_t kAStageSteps = 2;               // steps per A stage; the two groups of A warps own one stage each
constexpr uint32_t kBStageSteps = 4;               // 32-site steps per shared-memory B stage
constexpr uint32_t kBStages = 3;
constexpr uint32_t kLBO = 128;                     // bytes between K-adjacent 8x16-byte core matrices
constexpr uint32_t kSBO = kBStageSteps * 2 * kLBO; // bytes between 8-row groups (a stage holds 32*kBStageSteps K bytes)
constexpr uint32_t kBTile = (kUN / 8) * kSBO;      // one B operand plane of one stage
constexpr uint32_t kBStageBytes = 3 * kBTile;
constexpr size_t kUmmaSmem = size_t(kBStages) * kBStageBytes + 1024;  // + alignment slack
constexpr uint32_t kUThreads = 512;
constexpr uint32_t kAWarps = 8, kBWarps = (2 * kUN) / 32, kExpanderWarps = kAWarps + kBWarps;  // 8 + 5
constexpr uint32_t kIssuers = 3;                   // warps 13, 14, 15: x, y and h MMAs
constexpr uint32_t kAPrefetch = 2, kBPrefetch = 2; // register prefetch depth in stages
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColXX = 0, kColY = kUN, kColH = 3 * kUN;  // accumulators: xx | (yy|yh) | (hy|hh)
constexpr uint32_t kColA = 5 * kUN;                // A ring: slot s at kColA + 24 s: x, y, h (8 columns each)
static_assert(kBWarps * 32 == 2 * kUN, "two threads per column sample must fill whole warps");
static_assert(kUM == kBandTileRows && kUN == kBandTileCols, "band enumeration tile shape");
static_assert(kColA + 24 * kASlots <= kTmemCols, "TMEM budget");
static_assert(kASlots == 2 * kAStageSteps && kBStageSteps == 2 * kAStageSteps, "stage geometry");
static_assert(kChunkWords % (2 * kAStageSteps * kAPrefetch) == 0 && kChunkWords % (kBStageSteps * kBPrefetch) == 0, "loop unrolling");

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  // K-major, no swizzle: ((8,n),2):((16 B, SBO), LBO); version 1 (Blackwell)
  return uint64_t((smem_addr >> 4) & 0x3fff) | (uint64_t((kLBO >> 4) & 0x3fff) << 16) |
         (uint64_t((kSBO >> 4) & 0x3fff) << 32) | (uint64_t(1) << 46);
}
__host__ __device__ constexpr uint32_t make_idesc_i8(uint32_t M, uint32_t N, bool a_signed, bool b_signed) {
  return (2u << 4) /* D = s32 */ | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

#ifdef CK_UMMA_PROFILE
__device__ unsigned long long g_umma_prof[16];
#define PROF_T() clock64()
#define PROF_ADD(slot, dt) do { if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_umma_prof[slot], (unsigned long long)(dt)); } while (0)
#else
#define PROF_T() 0ull
#define PROF_ADD(slot, dt) do { (void)(dt); } while (0)
#endif

// One code word = 8 genotypes (nibbles: 1 het, 2 hom-alt, 4 hom-ref, 0 missing) -> 8 bytes of each operand row.
// PRMT picks, per output byte, the table byte indexed by the nibble: tables hold the operand value of each code.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ void expand_codes(uint32_t z, uint32_t &x_lo, uint32_t &x_hi, uint32_t &y_lo, uint32_t &y_hi,
                                             uint32_t &h_lo, uint32_t &h_hi) {
  const uint32_t zh = z >> 16;  // PRMT reads its four selector nibbles from bits 0-15
  //            code:  0     1     2     3 | 4     (5-7 unused)
  // x = alt - ref     0     0    +1     . | -1
  // y = -[hom]        0     0    -1     . | -1
  // h = -[het]        0    -1     0     . |  0
  x_lo = prmt(0x00010000u, 0x000000ffu, z);
  x_hi = prmt(0x00010000u, 0x000000ffu, zh);
  y_lo = prmt(0x00ff0000u, 0x000000ffu, z);
  y_hi = prmt(0x00ff0000u, 0x000000ffu, zh);
  h_lo = prmt(0x0000ff00u, 0x00000000u, z);
  h_hi = prmt(0x0000ff00u, 0x00000000u, zh);
}

__global__ void __launch_bounds__(kUThreads, 1) king_umma_kernel(const KingLaunch p, const BandTiles tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[2], empty_a[2], full_b[kBStages], empty_b[kBStages], acc_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which tile: band order (band_tiles.cu) ----
  uint32_t ti, tj;
  band_decode(tiles, p.tile_begin + blockIdx.x, ti, tj);
  const uint32_t row0 = ti * kUM, col0 = tj * kUN;           // offsets inside the sub-matrix
  if (row0 >= p.num_rows) return;  // phantom row tile that pads an odd last band (band_tiles.cu)
  const uint32_t rows_here = min(kUM, p.num_rows - row0), cols_here = min(kUN, p.num_cols - col0);
  const uint32_t i0 = p.row_global0 + row0, j0 = p.col_global0 + col0;
  if (j0 + cols_here - 1 <= i0) return;  // no i < j pair in this tile (below the diagonal): whole CTA leaves

  if (warp == kExpanderWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(&full_a[s], kAWarps / 2);  // the four warps of the group that owns the stage
      mbar_init(&empty_a[s], kIssuers);
    }
    for (uint32_t s = 0; s < kBStages; ++s) {
      mbar_init(&full_b[s], kBWarps);
      mbar_init(&empty_b[s], kIssuers);
    }
    mbar_init(&acc_bar, kIssuers);
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t num_steps = p.words;  // one 32-site word per step; p.words is a multiple of kChunkWords = 16

  if (warp < kAWarps) {
    // ===== A expanders: one thread per row; group g (4 warps) owns A stage g = TMEM slots 2g, 2g+1 and fills it with
    // the steps {4n + 2g, 4n + 2g + 1}, n = 0, 1, ...  =====
    const uint32_t group = warp >> 2, srow = (warp & 3) * 32 + lane;
    const uint32_t slot = p.row_slot0 + row0 + srow;
    const uint32_t blk = slot / kTileSamples, ln = slot % kTileSamples;
    const bool in_range = srow < rows_here;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.codes) + size_t(blk) * p.words * kTileSamples + ln;
    const uint32_t ta = tmem_base + ((uint32_t(warp & 3) * 32u) << 16) + kColA + group * (kAStageSteps * 24);
    const uint32_t num_fills = num_steps / kASlots;  // fills of this group's stage
    uint4 z[kAPrefetch][kAStageSteps];
    auto load_fill = [&](uint32_t n, uint4 (&dst)[kAStageSteps]) {
#pragma unroll
      for (uint32_t q = 0; q < kAStageSteps; ++q)
        dst[q] = (in_range && n < num_fills) ? __ldg(src + size_t(n * kASlots + group * kAStageSteps + q) * kTileSamples)
                                             : make_uint4(0, 0, 0, 0);
    };
#pragma unroll
    for (uint32_t u = 0; u < kAPrefetch; ++u) load_fill(u, z[u]);
    for (uint32_t n0 = 0; n0 < num_fills; n0 += kAPrefetch) {
#pragma unroll
      for (uint32_t u = 0; u < kAPrefetch; ++u) {
        const uint32_t n = n0 + u;
        uint32_t x[kAStageSteps][8], y[kAStageSteps][8], h[kAStageSteps][8];
        const unsigned long long p0 = PROF_T();
#pragma unroll
        for (uint32_t q = 0; q < kAStageSteps; ++q) {
          expand_codes(z[u][q].x, x[q][0], x[q][1], y[q][0], y[q][1], h[q][0], h[q][1]);
          expand_codes(z[u][q].y, x[q][2], x[q][3], y[q][2], y[q][3], h[q][2], h[q][3]);
          expand_codes(z[u][q].z, x[q][4], x[q][5], y[q][4], y[q][5], h[q][4], h[q][5]);
          expand_codes(z[u][q].w, x[q][6], x[q][7], y[q][6], y[q][7], h[q][6], h[q][7]);
          // keep the expansion ahead of the barrier wait below (the compiler otherwise sinks it onto the refill's critical path)
          asm volatile("" ::"r"(x[q][0]), "r"(x[q][1]), "r"(x[q][2]), "r"(x[q][3]), "r"(x[q][4]), "r"(x[q][5]), "r"(x[q][6]), "r"(x[q][7]));
          asm volatile("" ::"r"(y[q][0]), "r"(y[q][1]), "r"(y[q][2]), "r"(y[q][3]), "r"(y[q][4]), "r"(y[q][5]), "r"(y[q][6]), "r"(y[q][7]));
          asm volatile("" ::"r"(h[q][0]), "r"(h[q][1]), "r"(h[q][2]), "r"(h[q][3]), "r"(h[q][4]), "r"(h[q][5]), "r"(h[q][6]), "r"(h[q][7]));
        }
        load_fill(n + kAPrefetch, z[u]);  // refill the registers just consumed
        const unsigned long long p1 = PROF_T();
        if (n > 0) mbar_wait_suspend(&empty_a[group], (n - 1) & 1u);  // the MMAs that read the previous fill have completed
        __syncwarp();  // tcgen05.st is warp-collective; the polling loop may leave the lanes diverged
        const unsigned long long p2 = PROF_T();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (uint32_t q = 0; q < kAStageSteps; ++q) {
          tmem_st8(ta + q * 24, x[q]);
          tmem_st8(ta + q * 24 + 8, y[q]);
          tmem_st8(ta + q * 24 + 16, h[q]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          // The issuers wait on full_a only.  Group 0's fill n starts B stage n, so it is group 0 that makes sure the
          // B operands are in shared memory before it announces the A stage (acquire on full_b, release on full_a).
          if (group == 0) mbar_wait_suspend(&full_b[n % kBStages], (n / kBStages) & 1u);
          mbar_arrive(&full_a[group]);
        }
        const unsigned long long p3 = PROF_T();
        PROF_ADD(0, p1 - p0);
        PROF_ADD(1, p2 - p1);
        PROF_ADD(2, p3 - p2);
        PROF_ADD(3, 1);
      }
    }
  } else if (warp < kExpanderWarps) {
    // ================= B expanders: two threads per column sample, kBStageSteps steps per stage =================
    const uint32_t idx = tid - kAWarps * 32;
    const uint32_t half = idx / kUN, srow = idx % kUN;  // half: K bytes 16*half .. 16*half+15 of every step
    const uint32_t slot = p.col_slot0 + col0 + srow;
    const uint32_t blk = slot / kTileSamples, ln = slot % kTileSamples;
    const bool in_range = srow < cols_here;
    const uint2 *src = reinterpret_cast<const uint2 *>(p.codes) + (size_t(blk) * p.words * kTileSamples + ln) * 2 + half;
    const uint32_t b_off = (srow >> 3) * kSBO + (srow & 7) * 16 + half * kLBO;
    const uint32_t num_bstages = num_steps / kBStageSteps;
    const uint32_t smem_base = smem_u32(smem);
    uint2 z[kBPrefetch][kBStageSteps];
    auto load_stage = [&](uint32_t m, uint2 (&dst)[kBStageSteps]) {
#pragma unroll
      for (uint32_t q = 0; q < kBStageSteps; ++q)
        dst[q] = (in_range && m < num_bstages) ? __ldg(src + size_t(m * kBStageSteps + q) * kTileSamples * 2) : make_uint2(0, 0);
    };
#pragma unroll
    for (uint32_t u = 0; u < kBPrefetch; ++u) load_stage(u, z[u]);
    for (uint32_t m0 = 0; m0 < num_bstages; m0 += kBPrefetch) {
#pragma unroll
      for (uint32_t u = 0; u < kBPrefetch; ++u) {
        const uint32_t m = m0 + u;
        const uint32_t s = m % kBStages, fill = m / kBStages;
        uint32_t x[kBStageSteps][4], y[kBStageSteps][4], h[kBStageSteps][4];
        const unsigned long long p0 = PROF_T();
#pragma unroll
        for (uint32_t q = 0; q < kBStageSteps; ++q) {
          expand_codes(z[u][q].x, x[q][0], x[q][1], y[q][0], y[q][1], h[q][0], h[q][1]);
          expand_codes(z[u][q].y, x[q][2], x[q][3], y[q][2], y[q][3], h[q][2], h[q][3]);
        }
        load_stage(m + kBPrefetch, z[u]);
        const unsigned long long p1 = PROF_T();
        if (fill > 0) mbar_wait_suspend(&empty_b[s], (fill - 1) & 1u);  // the MMAs that read this stage have completed
        const unsigned long long p2 = PROF_T();
        const uint32_t stage = smem_base + s * kBStageBytes + b_off;
#pragma unroll
        for (uint32_t q = 0; q < kBStageSteps; ++q) {
          sts128(stage + q * 2 * kLBO, x[q][0], x[q][1], x[q][2], x[q][3]);
          sts128(stage + kBTile + q * 2 * kLBO, y[q][0], y[q][1], y[q][2], y[q][3]);
          sts128(stage + 2 * kBTile + q * 2 * kLBO, h[q][0], h[q][1], h[q][2], h[q][3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (tensor core)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_b[s]);
        const unsigned long long p3 = PROF_T();
        PROF_ADD(4, p1 - p0);
        PROF_ADD(5, p2 - p1);
        PROF_ADD(6, p3 - p2);
        PROF_ADD(7, 1);
      }
    }
  } else {
    // ================= MMA issuers: warps 13 (x.x), 14 (y.[y;h]), 15 (h.[y;h]) =================
    // The whole warp runs the loop (warp-uniform control flow keeps the descriptor arithmetic in the uniform
    // datapath); one elected lane issues the MMAs and commits.
    const uint32_t which = warp - kExpanderWarps;
    const uint32_t idesc = which == 0 ? make_idesc_i8(kUM, kUN, true, true) : make_idesc_i8(kUM, 2 * kUN, true, true);
    const uint32_t d_addr = tmem_base + (which == 0 ? kColXX : which == 1 ? kColY : kColH);
    const uint32_t a_addr = tmem_base + kColA + which * 8;
    const uint64_t b_desc0 = make_smem_desc(smem_u32(smem) + (which == 0 ? 0u : kBTile));  // x tile, or stacked [y ; h]
    uint32_t elected;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(elected));
    const uint32_t num_astages = num_steps / kAStageSteps;
    for (uint32_t ma = 0; ma < num_astages; ++ma) {
      const uint32_t g = ma & 1u, mb = ma >> 1, sb = mb % kBStages;  // A stage g, B stage mb (4 steps = 2 A stages)
      const unsigned long long q0 = PROF_T();
      mbar_wait_suspend(&full_a[g], (ma >> 1) & 1u);  // covers the B stage too (see the A expanders)
      const unsigned long long q1 = PROF_T();
      PROF_ADD(8 + which, q1 - q0);
      PROF_ADD(11, which == 0 ? 1 : 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elected) {
#pragma unroll
        for (uint32_t q = 0; q < kAStageSteps; ++q) {
          const uint32_t b_bytes = sb * kBStageBytes + (g * kAStageSteps + q) * 2 * kLBO;
          umma_i8_ts(d_addr, a_addr + (g * kAStageSteps + q) * 24, b_desc0 + uint64_t(b_bytes >> 4), idesc,
                     (ma > 0 || q > 0) ? 1u : 0u);
        }
        umma_commit(&empty_a[g]);                 // arrives when this thread's MMAs so far have completed
        if (g == 1) umma_commit(&empty_b[sb]);    // second half of the B stage done
      }
      __syncwarp();
    }
    if (elected) umma_commit(&acc_bar);  // this issuer's accumulator is final
  }

  // ================= epilogue: all 16 warps; thread = row (TMEM lane quadrant warp % 4), column chunks by warp / 4 ====
  {
    __syncwarp();
    mbar_wait_suspend(&acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t quad = warp & 3, group = warp >> 2;
    const uint32_t r = quad * 32 + lane;
    const uint32_t gi = i0 + r;
    const uint32_t lane_base = tmem_base + ((quad * 32u) << 16);
    for (uint32_t c0 = group * 16; c0 < kUN; c0 += 64) {
      uint32_t xx[16], yy[16], yh[16], hy[16], hh[16];
      tmem_ld16(lane_base + kColXX + c0, xx);
      tmem_ld16(lane_base + kColY + c0, yy);
      tmem_ld16(lane_base + kColY + kUN + c0, yh);
      tmem_ld16(lane_base + kColH + c0, hy);
      tmem_ld16(lane_base + kColH + kUN + c0, hh);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (uint32_t q = 0; q < 16; ++q) {
        const uint32_t c = c0 + q, gj = j0 + c;
        const uint32_t both_het = hh[q];
        const uint32_t het_i = hh[q] + hy[q];                 // i het where j is defined
        const uint32_t het_j = hh[q] + yh[q];                 // j het where i is defined
        const uint32_t shared = yy[q] + yh[q] + hy[q] + hh[q];
        const uint32_t conc = uint32_t(int32_t(yy[q]) + int32_t(xx[q])) >> 1;
        const uint32_t opp = uint32_t(int32_t(yy[q]) - int32_t(xx[q])) >> 1;
        const bool in_tile = r < rows_here && c < cols_here;
        const float kin = kinship(het_i, het_j, both_het, opp);
        if (p.dump_counts != nullptr && in_tile) {
          const size_t idx = size_t(row0 + r) * p.num_cols + (col0 + c);
          ck_counts out;
          out.het_i = het_i; out.het_j = het_j; out.both_het = both_het;
          out.opposing_hom = opp; out.concordant_hom = conc; out.shared_sites = shared;
          p.dump_counts[idx] = out;
          p.dump_kin[idx] = kin;
        }
        emit_pair(p, in_tile && gi < gj, true, gi, gj, kin, opp, conc, both_het, shared);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  __syncwarp();
  if (warp == kExpanderWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

}  // namespace

#ifdef CK_UMMA_PROFILE
extern "C" void ck_debug_umma_prof(unsigned long long *out) {
  cudaMemcpyFromSymbol(out, g_umma_prof, sizeof(g_umma_prof));
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_umma_prof, z, sizeof(z));
}
#endif

uint64_t king_umma_num_tiles(const KingLaunch &k) { return band_num_tiles(k, kUN); }

cudaError_t launch_king_umma(const KingLaunch &k, uint32_t total_blocks, ck_ctx *ctx, cudaStream_t s, uint32_t *launches) {
  (void)total_blocks;  // every row / column a tile reads lies inside the shard's allocated blocks
  static std::atomic<uint64_t> configured{0};  // one bit per device
  if (cudaError_t e = optin_dynamic_smem(king_umma_kernel, kUmmaSmem, configured); e != cudaSuccess) return e;
  if (k.tile_end <= k.tile_begin) return cudaSuccess;
  BandTiles tiles{};
  cudaError_t e = band_prepare(k, kUN, ctx, s, nullptr, &tiles);
  if (e != cudaSuccess) return e;
  constexpr uint64_t kMaxGrid = 1ull << 30;
  for (uint64_t t = k.tile_begin; e == cudaSuccess && t < k.tile_end; t += kMaxGrid) {
    KingLaunch part = k;
    part.tile_begin = t;
    part.tile_end = (t + kMaxGrid < k.tile_end) ? t + kMaxGrid : k.tile_end;
    king_umma_kernel<<<unsigned(part.tile_end - part.tile_begin), kUThreads, kUmmaSmem, s>>>(part, tiles);
    if (launches) ++*launches;
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace ck
