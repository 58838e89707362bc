// Three-product screen in front of the mxf4 pairwise kernel (variant 5; sm_100a only).
//
// A pair passes the threshold only if   kin = 0.5 - D / (4 min(het_i, het_j)) > thr   (cuking.cu:289-297), where
//     D = sum over the sites both samples are called at of (g_i - g_j)^2 = het_i + het_j - 2 both_het + 4 opp
// is the squared genotype distance.  With the E2M1 indicator vectors of king_fp4_kernel.cu (x = +-1 hom-alt / hom-ref,
// y = 1 hom, h = 0.5 het) and w = y + h (one LOP3: z & 0x33333333),
//     D = 2 (y_i.w_j + h_i.y_j - x_i.x_j)
// THREE exact products of N = 80 instead of the five (N = 80 + 160 + 160) the counters need: 127 instead of 212 tensor
// clocks per 64 sites, 240 instead of 400 accumulator columns - so the A ring in TMEM is 8 slots deep instead of 4 and the
// refill latency that holds the five-product kernel at 88 % tensor activity is off the critical path.  min(het_i, het_j)
// over the jointly called sites is at most the minimum of the two samples' het counts over ALL sites (sample_totals_kernel),
// so   D < 4 (0.5 - thr) min(Het_i, Het_j)   is a necessary condition, exact in fp32 up to a margin.  For unrelated
// pairs kin is near 0 and the bound is off by about missing_rate / 2: at the thresholds cuKING is run with the screen
// rejects everything but the related pairs and their immediate neighbourhood.
// The kernel computes nothing else: a tile that holds at least one candidate pair sets its byte in p.tile_flags, and
// launch_king_screen runs king_fp4_kernel over the same tile range behind it - its CTAs leave at once where the byte is
// 0 - which produces the records of the flagged tiles exactly as variant 3 does.  Every record therefore comes out of the
// five-product kernel and its reference-order epilogue: same pairs, same bits.
// Dense output (negative thresholds keep every pair) and the count dump bypass the screen (dispatch_king).
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

#ifndef CK_SCREEN_SLOTS
#define CK_SCREEN_SLOTS 8
#endif
constexpr uint32_t kFM = 128, kFN = 80;     // tile rows (A operand, TMEM lanes) x tile columns (B operand): the mxf4 kernel's tiles
constexpr uint32_t kFSlots = CK_SCREEN_SLOTS;  // A ring in TMEM: one 64-site step per slot; AS slots form one A stage (barrier pair)
constexpr uint32_t kFGroups = 2;            // groups of four A warps; group g fills the A stages with stage % 2 == g
constexpr uint32_t kFSub = 4;               // B expanders work in sub-stages of 4 steps (register prefetch unit)
constexpr uint32_t kFLBO = 128;             // bytes between K-adjacent 8x16-byte core matrices
constexpr uint32_t kFThreads = 512;
constexpr uint32_t kFAWarps = 8, kFBWarps = (2 * kFN) / 32, kFExpWarps = kFAWarps + kFBWarps;  // 8 + 5
constexpr uint32_t kFIssuers = 3;           // warps 13, 14, 15: x, y and h MMAs
constexpr uint32_t kFAPrefetchSteps = 4;    // A register prefetch depth in steps of the group (= 8 steps ahead)
constexpr uint32_t kFBPrefetch = 2;         // B register prefetch depth in sub-stages (= 8 steps ahead)
constexpr uint32_t kFColXX = 0, kFColYW = kFN, kFColHY = 2 * kFN;  // accumulators: x.x | y.w | h.y
constexpr uint32_t kFColA = 3 * kFN;        // A ring: slot s at kFColA + 24 s: x, y, h (8 columns = 64 E2M1 each)
constexpr uint32_t kFColSF = kFColA + 24 * kFSlots;  // 16 columns of scale factors
constexpr uint32_t kFTmemCols = 512;
static_assert(kFBWarps * 32 == 2 * kFN, "two threads per column sample must fill whole warps");
static_assert(kFColSF + 16 <= kFTmemCols, "TMEM budget");
static_assert(kFM == kBandTileRows && kFN == kBandTileCols, "band enumeration tile shape");
static_assert(kChunkWords % (2 * kFGroups * kFAPrefetchSteps) == 0 && kChunkWords % (2 * kFSub * kFBPrefetch) == 0, "loop unrolling");

template <uint32_t AS, uint32_t BS, uint32_t NS>
struct Fp4Geo {
  static_assert(BS % kFSub == 0 && BS % AS == 0 && kFSlots % AS == 0 && (AS == 1 || AS == 2 || AS == 4), "stage geometry");
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
__device__ unsigned long long g_screen_prof[16];
#define FPROF_T() clock64()
#define FPROF_ADD(slot, dt) do { if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_screen_prof[slot], (unsigned long long)(dt)); } while (0)
#else
#define FPROF_T() 0ull
#define FPROF_ADD(slot, dt) do { (void)(dt); } while (0)
#endif

// Pins eight values in registers at this point of the instruction stream: without it the compiler sinks the operand
// expansion below the barrier wait that follows, i.e. onto the critical path of the A-slot refill.
__device__ __forceinline__ void pin8(const uint32_t (&v)[8]) {
  asm volatile("" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
#ifndef CK_S3_PF_STEPS  // L2 prefetch of the column operand this many steps ahead (see king_screen1_kernel.cu).  Off: this kernel
                        // runs at the board's power cap, where the prefetch costs more than it saves (cfg2: 622 vs 611 ms with 24)
#define CK_S3_PF_STEPS 0
#endif
constexpr uint32_t kL2PrefetchSteps = CK_S3_PF_STEPS;
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void expand_fp4(uint32_t z, uint32_t &x, uint32_t &y, uint32_t &h) {  // A operands
  x = z & 0xAAAAAAAAu;  // +1 hom-alt (0x2), -1 hom-ref (0xA)
  y = z & 0x22222222u;  // 1 hom
  h = z & 0x11111111u;  // 0.5 het
}
__device__ __forceinline__ void expand_b(uint32_t z, uint32_t &x, uint32_t &y, uint32_t &w) {  // B operands
  x = z & 0xAAAAAAAAu;
  y = z & 0x22222222u;
  w = z & 0x33333333u;  // 1 hom, 0.5 het: y + h
}

template <uint32_t AS, uint32_t BS, uint32_t NS>
__global__ void __launch_bounds__(kFThreads, 1) king_screen_kernel(const KingLaunch p, const BandTiles tiles) {
  using G = Fp4Geo<AS, BS, NS>;
  constexpr uint32_t kAStages = G::kAStages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kFSlots], empty_a[kFSlots], full_b[NS], empty_b[NS], acc_bar;  // full_a: one per STAGE is used
  __shared__ uint32_t tmem_base_smem;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

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
    // super-stage = kFGroups consecutive A stages (one per group): ONE full / empty barrier pair, so that an issuer pays one
    // wait and one commit per kFGroups * AS steps
    for (uint32_t s = 0; s < kAStages / kFGroups; ++s) mbar_init(&full_a[s], kFAWarps);  // all eight A warps
    for (uint32_t s = 0; s < kAStages / kFGroups; ++s) mbar_init(&empty_a[s], kFIssuers);  // one commit per issuer
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
        constexpr uint32_t kSuper = kAStages / kFGroups;  // super-stages in the ring
        const uint32_t super_no = n0 + u, ss = super_no % kSuper;
        if (super_no >= kSuper) mbar_wait_suspend(&empty_a[ss], ((super_no / kSuper) - 1) & 1u);  // previous readers done
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
        }
        // one announcement per stage: with an 8-slot ring the issuers need not start on a half-written stage, and their loop -
        // a barrier wait, the MMAs, a commit - is what bounds a step (profiles/r02_screen_kernels.md)
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[ss]);
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
      for (uint32_t q = 0; q < kFSub; ++q) {
        const uint32_t step = min(m, num_subs - 1) * kFSub + q;
        const uint4 *s0 = src + size_t(step) * (2 * kTileSamples);
        dst[q] = __ldg(s0);
        if (kL2PrefetchSteps != 0 && step + kL2PrefetchSteps < num_steps) prefetch_l2(s0 + size_t(kL2PrefetchSteps) * (2 * kTileSamples));
      }
    };
#pragma unroll
    for (uint32_t u = 0; u < kFBPrefetch; ++u) load_sub(u, z[u]);
    for (uint32_t m0 = 0; m0 < num_subs; m0 += kFBPrefetch) {
#pragma unroll
      for (uint32_t u = 0; u < kFBPrefetch; ++u) {
        const uint32_t m = m0 + u;
        const uint32_t st = m / kSubsPerStage, sub = m % kSubsPerStage;  // stage counter, sub-stage inside it
        const uint32_t s = st % NS, fill = st / NS;
        uint32_t x[kFSub][4], y[kFSub][4], h[kFSub][4];  // h holds w = y + h on the B side
        const unsigned long long p0 = FPROF_T();
#pragma unroll
        for (uint32_t q = 0; q < kFSub; ++q) {
          expand_b(z[u][q].x, x[q][0], y[q][0], h[q][0]);
          expand_b(z[u][q].y, x[q][1], y[q][1], h[q][1]);
          expand_b(z[u][q].z, x[q][2], y[q][2], h[q][2]);
          expand_b(z[u][q].w, x[q][3], y[q][3], h[q][3]);
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
    // ===== MMA issuers: warps 13 (x.x), 14 (y.w), 15 (h.y), N = 80 each.  The whole warp runs the loop (warp-uniform
    // control flow keeps the descriptor arithmetic in the uniform datapath); one elected lane issues the MMA and the commits.
    const uint32_t which = warp - kFExpWarps;
    if (which < kFIssuers) {
    const uint32_t idesc = make_idesc_mxf4(kFM, kFN);
    const uint32_t d_addr = tmem_base + (which == 0 ? kFColXX : which == 1 ? kFColYW : kFColHY);
    const uint32_t a_addr = tmem_base + kFColA + which * 8;
    const uint32_t sf_addr = tmem_base + kFColSF;
    // B planes of a stage: x | y | w.  x.x reads x, y.w reads w, h.y reads y.
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem) + (which == 0 ? 0u : which == 1 ? 2u * G::kTile : G::kTile), kFLBO, G::kSBO);
    const uint32_t elected = elect_one();
    constexpr uint32_t kSuperSteps = AS * kFGroups, kSuper = kAStages / kFGroups;
    static_assert(BS % kSuperSteps == 0, "a super-stage must not straddle B stages");
    for (uint32_t step = 0; step < num_steps; step += kSuperSteps) {
      const uint32_t super_no = step / kSuperSteps, ss = super_no % kSuper, mb = step / BS, sb = mb % NS, q = step % BS;
      const unsigned long long q0 = FPROF_T();
      if (q == 0) mbar_wait_suspend(&full_b[sb], (mb / NS) & 1u);
      mbar_wait_suspend(&full_a[ss], (super_no / kSuper) & 1u);
      const unsigned long long waited = FPROF_T() - q0;
      tcgen05_after_sync();
      if (elected) {
#pragma unroll
        for (uint32_t a = 0; a < kSuperSteps; ++a) {  // slot of step (step + a) = ss * kSuperSteps + a: group a / AS, its step a % AS
          const uint32_t b_bytes = sb * G::kStageBytes + (q + a) * 2 * kFLBO;
          umma_mxf4_ts(d_addr, a_addr + (ss * kSuperSteps + a) * 24, b_desc0 + uint64_t(b_bytes >> 4), idesc, sf_addr, (step + a) > 0 ? 1u : 0u);
        }
        umma_commit_arrive(&empty_a[ss]);                            // arrives when this thread's MMAs so far have completed
        if (q + kSuperSteps == BS) umma_commit_arrive(&empty_b[sb]);  // last steps of the B stage
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
  bool any = false;
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
    // kin > thr needs D < 4 (0.5 - thr) min(het_i, het_j) <= 4 (0.5 - thr) min(Het_i, Het_j) with the het counts over all
    // sites on the right.  The accumulators are exact multiples of 1/4 (header of king_fp4_kernel.cu), so half_d below is
    // exact or off by an ulp; the reference's fp32 kin differs from the real value by a few ulps: a relative margin of
    // 1e-4 plus one site on the bound covers both.  thr >= 0.5 (bound <= 0) leaves no candidate, like the reference.
    const float bound2 = 2.f * (0.5f - p.kin_threshold) * 1.0001f;  // on D / 2
    const uint32_t row_slot = p.row_slot0 + row0 + (r < rows_here ? r : 0u);
    const float het_i = __uint2float_rn(__ldg(p.sample_totals + row_slot).x);
    auto screen = [&](uint32_t c, uint32_t xx, uint32_t yw, uint32_t hy) {
      const uint32_t gj = j0 + c;
      const bool pair = r < rows_here && c < cols_here && gi < gj;
      const float het_j = __uint2float_rn(__ldg(p.sample_totals + p.col_slot0 + col0 + (c < cols_here ? c : 0u)).x);  // warp-uniform address
      const float half_d = (__uint_as_float(yw) + __uint_as_float(hy)) - __uint_as_float(xx);
      any = any || (pair && half_d < fmaf(bound2, fminf(het_i, het_j), 1.f));
    };
    constexpr uint32_t kColsPerGroup = kFN / 4;  // 20
    const uint32_t c0 = group * kColsPerGroup;
    {
      uint32_t xx[16], yw[16], hy[16];
      tmem_load16(lane_base + kFColXX + c0, xx);
      tmem_load16(lane_base + kFColYW + c0, yw);
      tmem_load16(lane_base + kFColHY + c0, hy);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (uint32_t q = 0; q < 16; ++q) screen(c0 + q, xx[q], yw[q], hy[q]);
    }
    {
      uint32_t xx[4], yw[4], hy[4];
      tmem_load4(lane_base + kFColXX + c0 + 16, xx);
      tmem_load4(lane_base + kFColYW + c0 + 16, yw);
      tmem_load4(lane_base + kFColHY + c0 + 16, hy);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (uint32_t q = 0; q < 4; ++q) screen(c0 + 16 + q, xx[q], yw[q], hy[q]);
    }
    if (tid == 0) FPROF_ADD(14, FPROF_T() - t_main);
  }
  tcgen05_before_sync();
  if (__syncthreads_or(any ? 1 : 0) && tid == 0) {
    p.tile_flags[blockIdx.x] = 1;
    atomicAdd(p.flagged_counter, 1ull);
  }
  __syncwarp();
  if (warp == kFExpWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kFTmemCols));
}

// ---- host side ------------------------------------------------------------------------------------------------------

template <uint32_t AS, uint32_t BS, uint32_t NS>
cudaError_t launch_cfg(const KingLaunch &part, const BandTiles &tiles, cudaStream_t s) {
  static std::atomic<uint64_t> configured{0};  // one bit per device
  if (cudaError_t e = optin_dynamic_smem(king_screen_kernel<AS, BS, NS>, Fp4Geo<AS, BS, NS>::kSmem, configured); e != cudaSuccess) return e;
  king_screen_kernel<AS, BS, NS><<<unsigned(part.tile_end - part.tile_begin), kFThreads, Fp4Geo<AS, BS, NS>::kSmem, s>>>(part, tiles);
  return cudaGetLastError();
}

}  // namespace

#ifdef CK_UMMA_PROFILE
extern "C" void ck_debug_screen_prof(unsigned long long *out) {
  cudaMemcpyFromSymbol(out, g_screen_prof, sizeof(g_screen_prof));
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_screen_prof, z, sizeof(z));
}
#endif

// Screen kernel over the tile range, then the mxf4 kernel over the same range restricted to the flagged tiles.  Ranges are
// cut at 2^24 tiles so that the flag bytes stay a small grow-only scratch of the ctx.
cudaError_t launch_king_screen(const KingLaunch &k, uint32_t total_blocks, ck_ctx *ctx, cudaStream_t s, uint32_t *launches) {
  if (k.tile_end <= k.tile_begin) return cudaSuccess;
  BandTiles tiles{};
  cudaError_t e = band_prepare(k, kFN, ctx, s, nullptr, &tiles);
  if (e != cudaSuccess) return e;
  constexpr uint64_t kMaxGrid = 1ull << 24;
  const size_t need = size_t(std::min<uint64_t>(k.tile_end - k.tile_begin, kMaxGrid));
  if (ctx->tile_flags_bytes < need) {
    if (ctx->tile_flags) {
      if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;  // an earlier launch may still read the old buffer
      cudaFree(ctx->tile_flags);
      ctx->tile_flags = nullptr;
      ctx->tile_flags_bytes = 0;
    }
    const size_t want = std::max<size_t>(need + need / 2, size_t(1) << 20);
    if ((e = dev_alloc(ctx, reinterpret_cast<void **>(&ctx->tile_flags), want)) != cudaSuccess) return e;
    ctx->tile_flags_bytes = want;
  }
  for (uint64_t t = k.tile_begin; e == cudaSuccess && t < k.tile_end; t += kMaxGrid) {
    KingLaunch part = k;
    part.tile_begin = t;
    part.tile_end = (t + kMaxGrid < k.tile_end) ? t + kMaxGrid : k.tile_end;
    part.tile_flags = ctx->tile_flags;
    part.flagged_counter = ctx->d_screen_flagged;
    ctx->screen_tiles += part.tile_end - part.tile_begin;
    ctx->screen_level_used = part.screen_level == 1 ? 1 : 3;
    if ((e = cudaMemsetAsync(ctx->tile_flags, 0, size_t(part.tile_end - part.tile_begin), s)) != cudaSuccess) return e;
    e = part.screen_level == 1 ? launch_king_screen1(part, tiles, s) : launch_cfg<2, 8, 3>(part, tiles, s);
    if (launches) ++*launches;
    if (e == cudaSuccess) e = launch_king_fp4(part, total_blocks, ctx, s, launches);  // reads part.tile_flags
  }
  return e;
}

}  // namespace ck
