// Pairwise KING on the 5th-generation tensor cores at the FP4 rate: tcgen05.mma.kind::mxf4.block_scale over E2M1
// indicator vectors with all block scales = 2^0 and fp32 accumulation (sm_100a only).  Variant 3, the default.
//
// Same algebra as king_umma_kernel.cu (the six counters of ComputeKingKernel, /root/reference/cuking.cu:214-240, as
// five bilinear forms of per-sample indicator vectors):
//     x = [hom-alt] - [hom-ref]   y = [hom]   h = [het] / 2          (all 0 where the genotype is missing)
//     D_xx = x_i.x_j = conc - opp            D_y = y_i.[y_j ; h_j] = (conc + opp | (i hom, j het) / 2)
//     D_h = h_i.[y_j ; h_j] = ((i het, j hom) / 2 | both_het / 4)
// kind::mxf4 multiplies 64 sites per instruction — twice kind::i8 — and the operands are 4 bits wide, so operand
// expansion, TMEM stores and shared-memory traffic per site all halve as well.  Exactness: every operand is 0, 0.5 or
// +-1 in E2M1, every product a multiple of 1/4, every partial sum a multiple of 1/4 below 2^24 / 4: exactly representable
// in the fp32 accumulator.  tools/umma_mxf4_probe.cu measured the tensor core's accumulation to be exact for counts up
// to 2^23 on this pool's B200s (profiles/r01_mxf4_probe.txt); capi.cu routes cohorts with more than 2^23 sites to the
// int8 kernel (exact to 2^31) instead.
//
// Operands are expanded on the fly from 4-bit genotype codes (layout.cuh) chosen so that the expansion is ONE logic
// instruction per operand per 8 genotypes: code = 1 het, 2 hom-alt, 0xA hom-ref, 0 missing, i.e. E2M1(0.5), E2M1(+1),
// E2M1(-1), 0, and   x = z & 0xAAAAAAAA   y = z & 0x22222222   h = z & 0x11111111.
//   * warps 0-7   A operands: one thread per row sample; the two groups of four warps take alternate A stages (two 64-site
//                 steps each) and write straight into a 4-slot TMEM ring with tcgen05.st (thread = TMEM lane = row);
//                 a slot is announced as soon as it is stored, a stage is released by one commit per issuer;
//   * warps 8-12  B operands: two threads per column sample (32 sites each per step) write K-major no-swizzle canonical
//                 shared-memory tiles, several steps per stage so the proxy fence and barrier round trip are amortised;
//   * one lane of each of warps 13-15 issues one of the three MMAs per step (A from TMEM, B from shared memory) and
//                 releases the A slot / B stage with tcgen05.commit;
//   * epilogue    all 16 warps read the fp32 accumulators with tcgen05.ld, convert to the exact integer counts, compute
//                 kinship in the reference's fp32 order and append through the warp-aggregated atomic.
// TMEM: 400 accumulator columns + 4 x 24 A columns + 16 scale-factor columns (all 0x7F = 2^0; every byte is the same,
// so the scale-factor layout is immaterial) = 512.
// Tiles are enumerated in bands of 8 row tiles (band_tiles.cu), column-major inside a band, so that the ~148 tiles in flight share
// 8 row blocks and ~19 column blocks and the genotype codes are served from L2.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "internal.cuh"
#include "king_common.cuh"
#include "umma_common.cuh"

namespace ck {

namespace {

#ifndef CK_FP4_TILE_N  // tile shape experiments: -DCK_FP4_TILE_N=64 -DCK_FP4_SLOTS=6 (profiles/r01_fp4_tuning.md)
#define CK_FP4_TILE_N 80
#define CK_FP4_SLOTS 4
#endif
constexpr uint32_t kFM = 128, kFN = CK_FP4_TILE_N;  // tile rows (A operand, TMEM lanes) x tile columns (B operand)
constexpr uint32_t kFSlots = CK_FP4_SLOTS;  // A ring in TMEM: one 64-site step per slot; AS slots form one A stage (barrier pair)
constexpr uint32_t kFGroups = 2;            // groups of four A warps; group g fills the A stages with stage % 2 == g
constexpr uint32_t kFSub = 4;               // B expanders work in sub-stages of 4 steps (register prefetch unit)
constexpr uint32_t kFLBO = 128;             // bytes between K-adjacent 8x16-byte core matrices
constexpr uint32_t kFThreads = 512;
constexpr uint32_t kFAWarps = 8, kFBWarps = (2 * kFN) / 32, kFExpWarps = kFAWarps + kFBWarps;  // 8 + 5
constexpr uint32_t kFIssuers = 3;           // warps 13, 14, 15: x, y and h MMAs
constexpr uint32_t kFAPrefetchSteps = 4;    // A register prefetch depth in steps of the group (= 8 steps ahead)
constexpr uint32_t kFBPrefetch = 2;         // B register prefetch depth in sub-stages (= 8 steps ahead)
constexpr uint32_t kFColXX = 0, kFColY = kFN, kFColH = 3 * kFN;  // accumulators: xx | (yy|yh) | (hy|hh)
constexpr uint32_t kFColA = 5 * kFN;        // A ring: slot s at kFColA + 24 s: x, y, h (8 columns = 64 E2M1 each)
constexpr uint32_t kFColSF = kFColA + 24 * kFSlots;  // 16 columns of scale factors
constexpr uint32_t kFTmemCols = 512;
static_assert(kFBWarps * 32 == 2 * kFN, "two threads per column sample must fill whole warps");
static_assert(kFColSF + 16 <= kFTmemCols, "TMEM budget");
static_assert(kFM == kBandTileRows && (kFN == kBandTileCols || CK_FP4_TILE_N != 80), "band enumeration tile shape");
static_assert(kChunkWords % (2 * kFGroups * kFAPrefetchSteps) == 0 && kChunkWords % (2 * kFSub * kFBPrefetch) == 0, "loop unrolling");

template <uint32_t AS, uint32_t BS, uint32_t NS>
struct Fp4Geo {
  static_assert(BS % kFSub == 0 && BS % AS == 0 && kFSlots % AS == 0 && (AS == 1 || AS == 2), "stage geometry");
  static constexpr uint32_t kAStages = kFSlots / AS;        // A stages in the TMEM ring
  static constexpr uint32_t kSBO = BS * 2 * kFLBO;          // a stage holds 32 K-bytes (64 sites) per step
  static constexpr uint32_t kTile = (kFN / 8) * kSBO;       // one B operand plane of one stage
  static constexpr uint32_t kStageBytes = 3 * kTile;
  static constexpr size_t kSmem = size_t(NS) * kStageBytes + 1024;  // + alignment slack
};

// block-scaled instruction descriptor: E2M1 x E2M1 (format code 1 under kind::mxf4), UE8M0 scales, both K-major,
// dense K = 64, scale-factor ids 0
__host__ __device__ constexpr uint32_t make_idesc_mxf4(uint32_t M, uint32_t N) {
  return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T, fp32 accumulation, block scales read from TMEM
__device__ __forceinline__ void umma_mxf4_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t tmem_sf,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%5], p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(tmem_sf)
      : "memory");
}

#ifdef CK_UMMA_PROFILE
__device__ unsigned long long g_fp4_prof[16];
#define FPROF_T() clock64()
#define FPROF_ADD(slot, dt) do { if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_fp4_prof[slot], (unsigned long long)(dt)); } while (0)
#else
#define FPROF_T() 0ull
#define FPROF_ADD(slot, dt) do { (void)(dt); } while (0)
#endif

// Pins eight values in registers at this point of the instruction stream: without it the compiler sinks the operand
// expansion below the barrier wait that follows, i.e. onto the critical path of the A-slot refill.
__device__ __forceinline__ void pin8(const uint32_t (&v)[8]) {
  asm volatile("" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void expand_fp4(uint32_t z, uint32_t &x, uint32_t &y, uint32_t &h) {
  x = z & 0xAAAAAAAAu;  // +1 hom-alt (0x2), -1 hom-ref (0xA)
  y = z & 0x22222222u;  // 1 hom
  h = z & 0x11111111u;  // 0.5 het
}

template <uint32_t AS, uint32_t BS, uint32_t NS>
__global__ void __launch_bounds__(kFThreads, 1) king_fp4_kernel(const KingLaunch p, const BandTiles tiles) {
  using G = Fp4Geo<AS, BS, NS>;
  constexpr uint32_t kAStages = G::kAStages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kFSlots], empty_a[kFSlots], full_b[NS], empty_b[NS], acc_bar;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.tile_flags != nullptr && p.tile_flags[blockIdx.x] == 0) return;  // behind the screen kernel: only the tiles it flagged

  // ---- which tile: band order (band_tiles.cu) ----
  uint32_t ti, tj;
  band_decode(tiles, p.tile_begin + blockIdx.x, ti, tj);
  const uint32_t row0 = ti * kFM, col0 = tj * kFN;  // offsets inside the sub-matrix
  if (row0 >= p.num_rows) return;  // phantom row tile that pads an odd last band (band_tiles.cu)
  const uint32_t rows_here = min(kFM, p.num_rows - row0), cols_here = min(kFN, p.num_cols - col0);
  const uint32_t i0 = p.row_global0 + row0, j0 = p.col_global0 + col0;
  if (j0 + cols_here - 1 <= i0) return;  // no i < j pair in this tile (below the diagonal): whole CTA leaves

  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (warp == kFExpWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kFTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < kFSlots; ++s) mbar_init(&full_a[s], kFAWarps / kFGroups);  // per slot: the four warps of the filling group
    for (uint32_t s = 0; s < kAStages; ++s) mbar_init(&empty_a[s], kFIssuers);             // per stage: one commit per issuer
    for (uint32_t s = 0; s < NS; ++s) {
      mbar_init(&full_b[s], kFBWarps);
      mbar_init(&empty_b[s], kFIssuers);
    }
    mbar_init(&acc_bar, kFIssuers);
    mbar_fence_init();
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t num_steps = p.words / 2;  // one 64-site step = two 32-site code words; p.words is a multiple of 16
  const unsigned long long t_start = FPROF_T();

  if (warp < kFAWarps) {
    // ===== A expanders: one thread per row; group g expands the steps {2n + g} into TMEM slot step % 4.  A slot is
    // refilled as soon as the three MMAs that read it have completed, while the other three keep the tensor pipe busy.
    const uint32_t group = warp >> 2, srow = (warp & 3) * 32 + lane;
    // rows beyond the tile's edge re-read its first row (always allocated): their pairs are masked in the epilogue, and
    // the loads stay unconditional
    const uint32_t slot = p.row_slot0 + row0 + (srow < rows_here ? srow : 0u);
    const uint32_t blk = slot / kTileSamples, ln = slot % kTileSamples;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.codes) + size_t(blk) * p.words * kTileSamples + ln;
    const uint32_t lane_base = tmem_base + ((uint32_t(warp & 3) * 32u) << 16);
    const uint32_t ta = lane_base + kFColA;
    if (group == 0) {  // scale factors: 2^0 everywhere; ordered before the first MMA by this group's first full_a arrive
      uint32_t one[8];
#pragma unroll
      for (uint32_t q = 0; q < 8; ++q) one[q] = 0x7f7f7f7fu;
      tmem_store8(lane_base + kFColSF, one);
      tmem_store8(lane_base + kFColSF + 8, one);
    }
    // item n of group g = A stage number kFGroups * n + g = steps [AS * (kFGroups n + g), + AS): a group fills whole
    // stages (measured faster than both groups filling half of every stage: each warp then sits on every refill's
    // critical path, profiles/r01_fp4_tuning.md)
    constexpr uint32_t kPrefetch = kFAPrefetchSteps / AS;  // items
    const uint32_t num_items = num_steps / (AS * kFGroups);
    uint4 z[kPrefetch][AS][2];
    auto load_item = [&](uint32_t n, uint4 (&dst)[AS][2]) {
      n = min(n, num_items - 1);  // the prefetch beyond the last item re-reads it
#pragma unroll
      for (uint32_t a = 0; a < AS; ++a) {
        const uint4 *s0 = src + size_t((n * kFGroups + group) * AS + a) * (2 * kTileSamples);
        dst[a][0] = __ldg(s0);
        dst[a][1] = __ldg(s0 + kTileSamples);
      }
    };
#pragma unroll
    for (uint32_t u = 0; u < kPrefetch; ++u) load_item(u, z[u]);
    for (uint32_t n0 = 0; n0 < num_items; n0 += kPrefetch) {
#pragma unroll
      for (uint32_t u = 0; u < kPrefetch; ++u) {
        const uint32_t stage_no = (n0 + u) * kFGroups + group, astage = stage_no % kAStages;
        uint32_t x[AS][8], y[AS][8], h[AS][8];
        const unsigned long long p0 = FPROF_T();
#pragma unroll
        for (uint32_t a = 0; a < AS; ++a) {
          expand_fp4(z[u][a][0].x, x[a][0], y[a][0], h[a][0]);
          expand_fp4(z[u][a][0].y, x[a][1], y[a][1], h[a][1]);
          expand_fp4(z[u][a][0].z, x[a][2], y[a][2], h[a][2]);
          expand_fp4(z[u][a][0].w, x[a][3], y[a][3], h[a][3]);
          expand_fp4(z[u][a][1].x, x[a][4], y[a][4], h[a][4]);
          expand_fp4(z[u][a][1].y, x[a][5], y[a][5], h[a][5]);
          expand_fp4(z[u][a][1].z, x[a][6], y[a][6], h[a][6]);
          expand_fp4(z[u][a][1].w, x[a][7], y[a][7], h[a][7]);
          pin8(x[a]);
          pin8(y[a]);
          pin8(h[a]);
        }
        load_item(n0 + u + kPrefetch, z[u]);  // refill the registers just consumed
        const unsigned long long p1 = FPROF_T();
        if (stage_no >= kAStages) mbar_wait_suspend(&empty_a[astage], ((stage_no / kAStages) - 1) & 1u);  // previous readers done
        __syncwarp();  // tcgen05.st is warp-collective; the polling loop may leave the lanes diverged
        const unsigned long long p2 = FPROF_T();
        tcgen05_after_sync();
        // a slot is announced as soon as it is stored (the issuers start on it while the next one is written); the
        // stage is released as a whole (one commit per issuer per stage: commits are the issuers' expensive instruction)
#pragma unroll
        for (uint32_t a = 0; a < AS; ++a) {
          const uint32_t aslot = astage * AS + a;
          tmem_store8(ta + aslot * 24, x[a]);
          tmem_store8(ta + aslot * 24 + 8, y[a]);
          tmem_store8(ta + aslot * 24 + 16, h[a]);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tcgen05_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_a[aslot]);
        }
        const unsigned long long p3 = FPROF_T();
        FPROF_ADD(0, p1 - p0);
        FPROF_ADD(1, p2 - p1);
        FPROF_ADD(2, p3 - p2);
        FPROF_ADD(3, 1);
      }
    }
  } else if (warp < kFExpWarps) {
    // ===== B expanders: two threads per column sample (32 sites of every step each), BS steps per stage =====
    const uint32_t idx = tid - kFAWarps * 32;
    const uint32_t half = idx / kFN, srow = idx % kFN;  // half: K bytes 16*half .. 16*half+15 of every step
    const uint32_t slot = p.col_slot0 + col0 + (srow < cols_here ? srow : 0u);  // see the A expanders
    const uint32_t blk = slot / kTileSamples, ln = slot % kTileSamples;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.codes) + (size_t(blk) * p.words + half) * kTileSamples + ln;
    const uint32_t b_off = (srow >> 3) * G::kSBO + (srow & 7) * 16 + half * kFLBO;
    const uint32_t num_subs = num_steps / kFSub;
    const uint32_t smem_base = smem_u32(smem);
    constexpr uint32_t kSubsPerStage = BS / kFSub;
    uint4 z[kFBPrefetch][kFSub];
    auto load_sub = [&](uint32_t m, uint4 (&dst)[kFSub]) {
#pragma unroll
      for (uint32_t q = 0; q < kFSub; ++q) dst[q] = __ldg(src + size_t(min(m, num_subs - 1) * kFSub + q) * (2 * kTileSamples));
    };
#pragma unroll
    for (uint32_t u = 0; u < kFBPrefetch; ++u) load_sub(u, z[u]);
    for (uint32_t m0 = 0; m0 < num_subs; m0 += kFBPrefetch) {
#pragma unroll
      for (uint32_t u = 0; u < kFBPrefetch; ++u) {
        const uint32_t m = m0 + u;
        const uint32_t st = m / kSubsPerStage, sub = m % kSubsPerStage;  // stage counter, sub-stage inside it
        const uint32_t s = st % NS, fill = st / NS;
        uint32_t x[kFSub][4], y[kFSub][4], h[kFSub][4];
        const unsigned long long p0 = FPROF_T();
#pragma unroll
        for (uint32_t q = 0; q < kFSub; ++q) {
          expand_fp4(z[u][q].x, x[q][0], y[q][0], h[q][0]);
          expand_fp4(z[u][q].y, x[q][1], y[q][1], h[q][1]);
          expand_fp4(z[u][q].z, x[q][2], y[q][2], h[q][2]);
          expand_fp4(z[u][q].w, x[q][3], y[q][3], h[q][3]);
        }
        load_sub(m + kFBPrefetch, z[u]);
        const unsigned long long p1 = FPROF_T();
        if (sub == 0 && fill > 0) mbar_wait_suspend(&empty_b[s], (fill - 1) & 1u);  // the MMAs that read this stage have completed
        const unsigned long long p2 = FPROF_T();
        const uint32_t stage = smem_base + s * G::kStageBytes + b_off + sub * kFSub * 2 * kFLBO;
#pragma unroll
        for (uint32_t q = 0; q < kFSub; ++q) {
          sts128(stage + q * 2 * kFLBO, x[q][0], x[q][1], x[q][2], x[q][3]);
          sts128(stage + G::kTile + q * 2 * kFLBO, y[q][0], y[q][1], y[q][2], y[q][3]);
          sts128(stage + 2 * G::kTile + q * 2 * kFLBO, h[q][0], h[q][1], h[q][2], h[q][3]);
        }
        if (sub == kSubsPerStage - 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (tensor core)
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_b[s]);
        }
        const unsigned long long p3 = FPROF_T();
        FPROF_ADD(4, p1 - p0);
        FPROF_ADD(5, p2 - p1);
        FPROF_ADD(6, p3 - p2);
        FPROF_ADD(7, 1);
      }
    }
  } else {
    // ===== MMA issuers: warps 13 (x.x), 14 (y.[y;h]), 15 (h.[y;h]).  The whole warp runs the loop (warp-uniform control
    // flow keeps the descriptor arithmetic in the uniform datapath); one elected lane issues the MMA and the commits.
    const uint32_t which = warp - kFExpWarps;
    if (which < kFIssuers) {
    const uint32_t idesc = which == 0 ? make_idesc_mxf4(kFM, kFN) : make_idesc_mxf4(kFM, 2 * kFN);
    const uint32_t d_addr = tmem_base + (which == 0 ? kFColXX : which == 1 ? kFColY : kFColH);
    const uint32_t a_addr = tmem_base + kFColA + which * 8;
    const uint32_t sf_addr = tmem_base + kFColSF;
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem) + (which == 0 ? 0u : G::kTile), kFLBO, G::kSBO);  // x, or stacked [y ; h]
    const uint32_t elected = elect_one();
    for (uint32_t step = 0; step < num_steps; step += AS) {
      const uint32_t stage_no = step / AS, astage = stage_no % kAStages, mb = step / BS, sb = mb % NS, q = step % BS;
      const unsigned long long q0 = FPROF_T();
      if (q == 0) mbar_wait_suspend(&full_b[sb], (mb / NS) & 1u);
      unsigned long long waited = FPROF_T() - q0;
#pragma unroll
      for (uint32_t a = 0; a < AS; ++a) {
        const uint32_t aslot = astage * AS + a;
        const unsigned long long w0 = FPROF_T();
        mbar_wait_suspend(&full_a[aslot], (stage_no / kAStages) & 1u);
        waited += FPROF_T() - w0;
        tcgen05_after_sync();
        if (elected) {
          const uint32_t b_bytes = sb * G::kStageBytes + (q + a) * 2 * kFLBO;
          umma_mxf4_ts(d_addr, a_addr + aslot * 24, b_desc0 + uint64_t(b_bytes >> 4), idesc, sf_addr, (step + a) > 0 ? 1u : 0u);
        }
      }
      if (elected) {
        umma_commit_arrive(&empty_a[astage]);                       // arrives when this thread's MMAs so far have completed
        if (q + AS == BS) umma_commit_arrive(&empty_b[sb]);         // last steps of the B stage
      }
      __syncwarp();
      const unsigned long long q1 = FPROF_T();
      FPROF_ADD(8 + which, waited);
      if (which == 1) FPROF_ADD(11, q1 - q0 - waited);
    }
    if (elected) umma_commit_arrive(&acc_bar);  // this issuer's accumulator is final
    }
  }

  // ===== epilogue: all 16 warps; thread = row (TMEM lane quadrant warp % 4), 20 columns per warp group (16 + 4) =====
  {
    __syncwarp();
    mbar_wait_suspend(&acc_bar, 0);
    tcgen05_after_sync();
    const unsigned long long t_main = FPROF_T();
    if (tid == 0) { FPROF_ADD(12, t_main - t_start); FPROF_ADD(13, num_steps); }
    const uint32_t quad = warp & 3, group = warp >> 2;
    const uint32_t r = quad * 32 + lane;
    const uint32_t gi = i0 + r;
    const uint32_t lane_base = tmem_base + ((quad * 32u) << 16);
    // Cheap conservative screen before the exact kinship, in fp32 on the raw accumulators (no conversions).  With
    // every count at most 2^23 the accumulators are exact and the two expressions below are exact or off by one ulp
    // (relative 2^-23, far inside the margin whenever they matter: |num| >= 2^23 implies den >= 2^23 / |thr - 0.5|):
    //     num = 2 both_het - 4 opp - het_i - het_j = 2 (D_xx - D_yy - D_hy - D_yh)          (cuking.cu:289-292)
    //     den = 4 min(het_i, het_j)                = 8 (2 D_hh + min(D_hy, D_yh))           (cuking.cu:293)
    // and kin = fl(0.5 + fl(num / den)) differs from the real value by < 2^-22 relative, so a pair with
    // num <= (thr - 0.5 - margin) den cannot pass the strict threshold test and skips the conversions and the IEEE
    // division.  den == 0 implies num <= 0 (both_het <= min_hets), i.e. -inf / NaN, which the reference never emits.
    const float screen4 = 4.f * (p.kin_threshold - 0.5f - (1e-3f + 1e-5f * fabsf(p.kin_threshold)));
    const bool dump = p.dump_counts != nullptr, dense = p.dense_band_base != nullptr;
    auto finish = [&](uint32_t c, uint32_t xx, uint32_t yy, uint32_t yh, uint32_t hy, uint32_t hh) {
      const uint32_t gj = j0 + c;
      const float fxx = __uint_as_float(xx), fyy = __uint_as_float(yy), fyh = __uint_as_float(yh), fhy = __uint_as_float(hy),
                  fhh = __uint_as_float(hh);
      const float half_num = (fxx - fyy) - (fhy + fyh);
      const float eighth_den = fmaf(2.f, fhh, fminf(fhy, fyh));
      const bool in_tile = r < rows_here && c < cols_here;
      const bool pair = in_tile && gi < gj;
      const bool cand = pair && half_num > screen4 * eighth_den;
      if (__ballot_sync(0xffffffffu, cand || dump || (dense && pair)) == 0) return;  // dense output: every pair owns a slot
      // the accumulators hold exact multiples of 1/4 (see the header): scale back to integer counts
      const int32_t n_xx = __float2int_rn(fxx);                    // conc - opp (signed)
      const uint32_t n_yy = uint32_t(__float2int_rn(fyy));         // conc + opp
      const uint32_t n_yh = uint32_t(__float2int_rn(2.f * fyh));   // i hom, j het
      const uint32_t n_hy = uint32_t(__float2int_rn(2.f * fhy));   // i het, j hom
      const uint32_t both_het = uint32_t(__float2int_rn(4.f * fhh));
      const uint32_t het_i = both_het + n_hy;  // i het where j is defined
      const uint32_t het_j = both_het + n_yh;  // j het where i is defined
      const uint32_t opp = uint32_t(int32_t(n_yy) - n_xx) >> 1;
      const uint32_t shared = n_yy + n_yh + n_hy + both_het;
      const uint32_t conc = uint32_t(int32_t(n_yy) + n_xx) >> 1;
      const float kin = kinship(het_i, het_j, both_het, opp);
      if (dump && in_tile) {
        const size_t idx = size_t(row0 + r) * p.num_cols + (col0 + c);
        ck_counts out;
        out.het_i = het_i; out.het_j = het_j; out.both_het = both_het;
        out.opposing_hom = opp; out.concordant_hom = conc; out.shared_sites = shared;
        p.dump_counts[idx] = out;
        p.dump_kin[idx] = kin;
      }
      emit_pair(p, pair, cand, gi, gj, kin, opp, conc, both_het, shared);
    };
    constexpr uint32_t kColsPerGroup = kFN / 4;  // 20
    static_assert(kColsPerGroup == 20 || kColsPerGroup == 16, "epilogue column split");
    const uint32_t c0 = group * kColsPerGroup;
    {
      uint32_t xx[16], yy[16], yh[16], hy[16], hh[16];
      tmem_load16(lane_base + kFColXX + c0, xx);
      tmem_load16(lane_base + kFColY + c0, yy);
      tmem_load16(lane_base + kFColY + kFN + c0, yh);
      tmem_load16(lane_base + kFColH + c0, hy);
      tmem_load16(lane_base + kFColH + kFN + c0, hh);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (uint32_t q = 0; q < 16; ++q) finish(c0 + q, xx[q], yy[q], yh[q], hy[q], hh[q]);
    }
    if constexpr (kColsPerGroup > 16) {
      uint32_t xx[4], yy[4], yh[4], hy[4], hh[4];
      tmem_load4(lane_base + kFColXX + c0 + 16, xx);
      tmem_load4(lane_base + kFColY + c0 + 16, yy);
      tmem_load4(lane_base + kFColY + kFN + c0 + 16, yh);
      tmem_load4(lane_base + kFColH + c0 + 16, hy);
      tmem_load4(lane_base + kFColH + kFN + c0 + 16, hh);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (uint32_t q = 0; q < 4; ++q) finish(c0 + 16 + q, xx[q], yy[q], yh[q], hy[q], hh[q]);
    }
    if (tid == 0) FPROF_ADD(14, FPROF_T() - t_main);
  }
  tcgen05_before_sync();
  __syncthreads();
  __syncwarp();
  if (warp == kFExpWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kFTmemCols));
}

// ---- host side ------------------------------------------------------------------------------------------------------

struct Fp4Config { uint32_t as, bs, ns; };
Fp4Config fp4_config() {  // stage geometry; CUKING_FP4_STAGE = "<steps per A stage>x<steps per B stage>x<B stages>" is a tuning knob
  static Fp4Config cfg = [] {
    Fp4Config c{2, 8, 3};
    if (const char *v = getenv("CUKING_FP4_STAGE")) {
      unsigned a = 0, b = 0, n = 0;
      if (sscanf(v, "%ux%ux%u", &a, &b, &n) == 3 && (a == 1 || a == 2) && ((b == 4 && n == 4) || (b == 8 && n == 3))) c = Fp4Config{a, b, n};
    }
    return c;
  }();
  return cfg;
}

template <uint32_t AS, uint32_t BS, uint32_t NS>
cudaError_t launch_cfg(const KingLaunch &part, const BandTiles &tiles, cudaStream_t s) {
  static std::atomic<uint64_t> configured{0};  // one bit per device
  if (cudaError_t e = optin_dynamic_smem(king_fp4_kernel<AS, BS, NS>, Fp4Geo<AS, BS, NS>::kSmem, configured); e != cudaSuccess) return e;
  king_fp4_kernel<AS, BS, NS><<<unsigned(part.tile_end - part.tile_begin), kFThreads, Fp4Geo<AS, BS, NS>::kSmem, s>>>(part, tiles);
  return cudaGetLastError();
}

}  // namespace

#ifdef CK_UMMA_PROFILE
extern "C" void ck_debug_fp4_prof(unsigned long long *out) {
  cudaMemcpyFromSymbol(out, g_fp4_prof, sizeof(g_fp4_prof));
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_fp4_prof, z, sizeof(z));
}
#endif

uint64_t king_fp4_num_tiles(const KingLaunch &k) { return band_num_tiles(k, kFN); }

cudaError_t king_fp4_prepare(const KingLaunch &k, ck_ctx *ctx, cudaStream_t s, std::vector<uint64_t> *band_prefix) {
  return band_prepare(k, kFN, ctx, s, band_prefix, nullptr);
}

cudaError_t launch_king_fp4(const KingLaunch &k, uint32_t total_blocks, ck_ctx *ctx, cudaStream_t s, uint32_t *launches) {
  (void)total_blocks;  // every row / column a tile reads lies inside the shard's allocated blocks
  if (k.tile_end <= k.tile_begin) return cudaSuccess;
  BandTiles tiles{};
  cudaError_t e = band_prepare(k, kFN, ctx, s, nullptr, &tiles);
  if (e != cudaSuccess) return e;
  const Fp4Config cfg = fp4_config();
  constexpr uint64_t kMaxGrid = 1ull << 30;
  for (uint64_t t = k.tile_begin; e == cudaSuccess && t < k.tile_end; t += kMaxGrid) {
    KingLaunch part = k;
    part.tile_begin = t;
    part.tile_end = (t + kMaxGrid < k.tile_end) ? t + kMaxGrid : k.tile_end;
    if (cfg.as == 1 && cfg.bs == 4) e = launch_cfg<1, 4, 4>(part, tiles, s);
    else if (cfg.as == 1) e = launch_cfg<1, 8, 3>(part, tiles, s);
    else if (cfg.bs == 4) e = launch_cfg<2, 4, 4>(part, tiles, s);
    else e = launch_cfg<2, 8, 3>(part, tiles, s);
    if (launches) ++*launches;
  }
  return e;
}

}  // namespace ck
