// One-product screen in front of the pairwise kernels (first level of variant 5; sm_100a only).
//
// king_screen_kernel.cu bounds kin through the exact squared genotype distance D = 2 (w_i.w_j - h_i.h_j - x_i.x_j) - three
// products.  Two of them only count sites: with y = [hom] and the per-sample totals Y (hom sites), Het, Def = Y + Het over
// the S sites of the cohort,
//     D = sum_both y_i + sum_both y_j - 2 x_i.x_j,      sum_both y_i >= max(0, Y_i - (S - Def_j))
// (a hom site of i is lost from the joint count only where j is missing), so
//     D / 2 >= L = (max(0, Y_i + Def_j - S) + max(0, Y_j + Def_i - S)) / 2 - x_i.x_j
// needs ONE exact product, x_i.x_j = concordant - opposing homozygotes.  kin > thr implies
// D / 2 < 2 (0.5 - thr) min(het_i, het_j) <= 2 (0.5 - thr) min(Het_i, Het_j), hence L below that bound: a necessary
// condition whose slack is the two samples' missing counts - a few per cent of D at the call rates cuKING's inputs have.
// dispatch (king_screen_kernel.cu: launch_king_screen) uses this level when the cohort's call rate predicts that
// unrelated pairs stay clear of the threshold, the three-product screen otherwise.
//
// With one 80-column accumulator per column tile there is TMEM for a 128 x 320 tile (four column tiles of the mxf4 kernel's
// enumeration side by side: 0.34 instead of 0.63 bytes of genotype codes per pair from L2, which is what bounds these
// kernels) and a 16-slot A ring.  Roles as in king_fp4_kernel.cu: warps 0-7 expand the row samples' x into TMEM, warps
// 8-12 the column samples' x into shared memory (each thread four columns), one lane each of warps 13 and 14 issues the N = 160
// MMAs of two column tiles for a four-step A stage behind one barrier wait; all 16 warps screen the accumulators.  The CTA of
// the first column tile of a group of four takes the following ones along (as far as they are inside the launch range); a
// tile whose left-hand neighbour lies outside the range leads the rest of its group.  Output: one flag byte per tile of the launch, as in king_screen_kernel.cu.
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.cuh"
#include "king_common.cuh"
#include "umma_common.cuh"

namespace ck {

namespace {

constexpr uint32_t kSM = 128, kSN = 80;       // tile rows x columns of ONE column tile (the mxf4 kernel's tile)
constexpr uint32_t kSTiles = 4;                // column tiles a CTA screens side by side
constexpr uint32_t kSWide = kSTiles * kSN;     // 320 columns
constexpr uint32_t kSSlots = 16;               // A ring in TMEM: one 64-site step (8 columns of x) per slot
constexpr uint32_t kSAS = 4;                   // steps per A stage: ONE barrier pair, wait and commit per stage - the issuing
                                               // lane's loop (wait ~75 clk, MMA ~55, commit ~100) is what bounds a step otherwise
constexpr uint32_t kSAStages = kSSlots / kSAS;
constexpr uint32_t kSGroups = 2;               // groups of four A warps; group g fills the A stages with stage % 2 == g
constexpr uint32_t kSBS = 8, kSNS = 2;         // steps per B stage, B stages
constexpr uint32_t kSSub = 2;                  // B expanders work in sub-stages of 2 steps (register prefetch unit)
constexpr uint32_t kSLBO = 128;                // bytes between K-adjacent 8x16-byte core matrices
constexpr uint32_t kSSBO = kSBS * 2 * kSLBO;   // bytes between 8-row groups: a stage holds 32 K-bytes (64 sites) per step
constexpr uint32_t kSStageBytes = (kSWide / 8) * kSSBO;  // 80 KB
constexpr size_t kSSmem = size_t(kSNS) * kSStageBytes + 1024;
constexpr uint32_t kSThreads = 512;
constexpr uint32_t kSAWarps = 8, kSBWarps = (2 * kSN) / 32, kSExpWarps = kSAWarps + kSBWarps;  // 8 + 5
constexpr uint32_t kSIssuers = 2;              // warps 13, 14: column tiles {0, 1} and {2, 3} (N = 160 each, own accumulator columns)
constexpr uint32_t kSAPrefetchItems = 1;       // A register prefetch depth in items (stages) of the group
constexpr uint32_t kSBPrefetch = 2;            // B register prefetch depth in sub-stages
constexpr uint32_t kSColAcc = 0, kSColA = kSWide, kSColSF = kSColA + 8 * kSSlots;
constexpr uint32_t kSTmemCols = 512;
static_assert(kSColSF + 16 <= kSTmemCols, "TMEM budget");
static_assert(kSM == kBandTileRows && kSN == kBandTileCols, "band enumeration tile shape");
static_assert(kChunkWords % (2 * kSAS * kSGroups * kSAPrefetchItems) == 0 && kChunkWords % (2 * kSSub * kSBPrefetch) == 0 &&
                  kChunkWords % (2 * kSBS) == 0 && kSBS % kSAS == 0 && kSBS % kSSub == 0 && kSTiles == 2 * kSIssuers, "loop unrolling");

__host__ __device__ constexpr uint32_t make_idesc_mxf4(uint32_t M, uint32_t N) {  // see king_fp4_kernel.cu
  return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}
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
__device__ __forceinline__ void pin8(const uint32_t (&v)[8]) {  // see king_fp4_kernel.cu
  asm volatile("" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
constexpr uint32_t kXMask = 0xAAAAAAAAu;  // x = z & mask: +1 hom-alt (0x2), -1 hom-ref (0xA)
// The register prefetch covers ~4 steps; a quarter of the loads miss L2 (the band's working set is larger than L2) and then
// take longer than that: ncu shows the B warps parked on the long scoreboard.  An L2 prefetch kPrefetchSteps ahead costs no
// register and turns those misses into hits (DRAM runs at 15 % of its bandwidth).
#ifndef CK_S1_PF_STEPS  // tuning: -DCK_S1_PF_STEPS=n -DCK_S1_PF_A=0/1 (profiles/r02_screen_kernels.md)
#define CK_S1_PF_STEPS 24
#define CK_S1_PF_A 0
#endif
constexpr uint32_t kL2PrefetchSteps = CK_S1_PF_STEPS;
constexpr bool kL2PrefetchA = CK_S1_PF_A != 0;  // the row operand too: measured slower (281 vs 272 ms on cfg2; its expanders have slack and the board is at its power cap)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__global__ void __launch_bounds__(kSThreads, 1) king_screen1_kernel(const KingLaunch p, const BandTiles tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kSAStages], empty_a[kSAStages], full_b[kSNS], empty_b[kSNS], acc_bar;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which tiles: band order (band_tiles.cu); this CTA's tile and, for an even column of the band, its right neighbour ----
  const unsigned long long t = p.tile_begin + blockIdx.x;
  uint32_t band = 0;
  {
    uint32_t hi = tiles.num_bands;  // largest b with band_prefix[b] <= t
    while (hi - band > 1) {
      const uint32_t mid = (band + hi) >> 1;
      if (tiles.band_prefix[mid] <= t) band = mid; else hi = mid;
    }
  }
  const uint32_t band_rows = band_rows_padded(tiles.num_row_tiles - band * kBandRowTiles);
  const uint32_t q = uint32_t(t - tiles.band_prefix[band]);
  const uint32_t ti = band * kBandRowTiles + q % band_rows, col_rel = q / band_rows;
  const uint32_t tj = tiles.band_first_col[band] + col_rel;
  // Groups of kSTiles adjacent column tiles of a band: the first member that lies inside the launch range leads (it is
  // member 0, or its left neighbour is outside the range) and takes the following members along as far as they exist and lie
  // inside the range; everybody else returns.
  const uint32_t member = col_rel % kSTiles;
  if (member != 0 && t - band_rows >= p.tile_begin) return;
  uint32_t ntiles = 1;
  while (member + ntiles < kSTiles && tj + ntiles < tiles.num_col_tiles && t + (unsigned long long)ntiles * band_rows < p.tile_end) ++ntiles;
  const uint32_t row0 = ti * kSM, col0 = tj * kSN;  // offsets inside the sub-matrix
  if (row0 >= p.num_rows) return;  // phantom row tile that pads an odd last band
  const uint32_t rows_here = min(kSM, p.num_rows - row0), cols_here = min(ntiles * kSN, p.num_cols - col0);
  const uint32_t i0 = p.row_global0 + row0, j0 = p.col_global0 + col0;
  if (j0 + cols_here - 1 <= i0) return;  // no i < j pair in these tiles: their flags stay 0

  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (warp == kSExpWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kSTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < kSAStages; ++s) mbar_init(&full_a[s], kSAWarps / kSGroups);  // the four warps of the filling group
    for (uint32_t s = 0; s < kSAStages; ++s) mbar_init(&empty_a[s], kSIssuers);          // one commit per issuer
    for (uint32_t s = 0; s < kSNS; ++s) {
      mbar_init(&full_b[s], kSBWarps);
      mbar_init(&empty_b[s], kSIssuers);
    }
    mbar_init(&acc_bar, kSIssuers);
    mbar_fence_init();
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t num_steps = p.words / 2;  // one 64-site step = two 32-site code words; p.words is a multiple of 16

  if (warp < kSAWarps) {
    // ===== A expanders: one thread per row; group g expands the stages {2n + g} (two steps each) into the TMEM ring =====
    const uint32_t group = warp >> 2, srow = (warp & 3) * 32 + lane;
    const uint32_t slot = p.row_slot0 + row0 + (srow < rows_here ? srow : 0u);  // rows beyond the edge re-read the first row
    const uint32_t blk = slot / kTileSamples, ln = slot % kTileSamples;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.codes) + size_t(blk) * p.words * kTileSamples + ln;
    const uint32_t lane_base = tmem_base + ((uint32_t(warp & 3) * 32u) << 16);
    const uint32_t ta = lane_base + kSColA;
    if (group == 0) {  // scale factors: 2^0 everywhere; ordered before the first MMA by this group's first full_a arrive
      uint32_t one[8];
#pragma unroll
      for (uint32_t k = 0; k < 8; ++k) one[k] = 0x7f7f7f7fu;
      tmem_store8(lane_base + kSColSF, one);
      tmem_store8(lane_base + kSColSF + 8, one);
    }
    const uint32_t num_items = num_steps / (kSAS * kSGroups);
    uint4 z[kSAPrefetchItems][kSAS][2];
    auto load_item = [&](uint32_t n, uint4 (&dst)[kSAS][2]) {
      n = min(n, num_items - 1);  // the prefetch beyond the last item re-reads it
#pragma unroll
      for (uint32_t a = 0; a < kSAS; ++a) {
        const uint32_t step = (n * kSGroups + group) * kSAS + a;
        const uint4 *s0 = src + size_t(step) * (2 * kTileSamples);
        dst[a][0] = __ldg(s0);
        dst[a][1] = __ldg(s0 + kTileSamples);
        if (kL2PrefetchA && step + kL2PrefetchSteps < num_steps) {
          prefetch_l2(s0 + size_t(kL2PrefetchSteps) * (2 * kTileSamples));
          prefetch_l2(s0 + size_t(kL2PrefetchSteps) * (2 * kTileSamples) + kTileSamples);
        }
      }
    };
#pragma unroll
    for (uint32_t u = 0; u < kSAPrefetchItems; ++u) load_item(u, z[u]);
    for (uint32_t n0 = 0; n0 < num_items; n0 += kSAPrefetchItems) {
#pragma unroll
      for (uint32_t u = 0; u < kSAPrefetchItems; ++u) {
        const uint32_t stage_no = (n0 + u) * kSGroups + group, astage = stage_no % kSAStages;
        uint32_t x[kSAS][8];
#pragma unroll
        for (uint32_t a = 0; a < kSAS; ++a) {
          x[a][0] = z[u][a][0].x & kXMask; x[a][1] = z[u][a][0].y & kXMask; x[a][2] = z[u][a][0].z & kXMask; x[a][3] = z[u][a][0].w & kXMask;
          x[a][4] = z[u][a][1].x & kXMask; x[a][5] = z[u][a][1].y & kXMask; x[a][6] = z[u][a][1].z & kXMask; x[a][7] = z[u][a][1].w & kXMask;
          pin8(x[a]);
        }
        load_item(n0 + u + kSAPrefetchItems, z[u]);  // refill the registers just consumed
        if (stage_no >= kSAStages) mbar_wait_suspend(&empty_a[astage], ((stage_no / kSAStages) - 1) & 1u);  // previous readers done
        __syncwarp();  // tcgen05.st is warp-collective; the polling loop may leave the lanes diverged
        tcgen05_after_sync();
#pragma unroll
        for (uint32_t a = 0; a < kSAS; ++a) tmem_store8(ta + (astage * kSAS + a) * 8, x[a]);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[astage]);
      }
    }
  } else if (warp < kSExpWarps) {
    // ===== B expanders: two threads per column quadruple (srow + 80 c), 32 sites of every step each, kSBS steps per stage =====
    const uint32_t idx = tid - kSAWarps * 32;
    const uint32_t half = idx / kSN, srow = idx % kSN;  // half: K bytes 16*half .. 16*half+15 of every step
    const uint4 *codes4 = reinterpret_cast<const uint4 *>(p.codes);
    uint32_t src[kSTiles];  // element offsets into codes4 (the buffer holds far fewer than 2^32 uint4)
#pragma unroll
    for (uint32_t c = 0; c < kSTiles; ++c) {
      const uint32_t col = srow + c * kSN;
      const uint32_t slot = p.col_slot0 + col0 + (col < cols_here ? col : 0u);  // columns beyond the edge re-read the first one
      const uint32_t blk = slot / kTileSamples, ln = slot % kTileSamples;
      src[c] = (blk * p.words + half) * kTileSamples + ln;
    }
    const uint32_t b_off0 = (srow >> 3) * kSSBO + (srow & 7) * 16 + half * kSLBO;  // column srow + 80 c: + c * (80 / 8) * kSSBO
    const uint32_t num_subs = num_steps / kSSub;
    const uint32_t smem_base = smem_u32(smem);
    constexpr uint32_t kSubsPerStage = kSBS / kSSub;
    uint4 z[kSBPrefetch][kSTiles][kSSub];
    auto load_sub = [&](uint32_t m, uint4 (&dst)[kSTiles][kSSub]) {
#pragma unroll
      for (uint32_t c = 0; c < kSTiles; ++c)
#pragma unroll
        for (uint32_t k = 0; k < kSSub; ++k) {
          const uint32_t step = min(m, num_subs - 1) * kSSub + k;
          const uint4 *s0 = codes4 + src[c] + size_t(step) * (2 * kTileSamples);
          dst[c][k] = __ldg(s0);
          if (step + kL2PrefetchSteps < num_steps) prefetch_l2(s0 + size_t(kL2PrefetchSteps) * (2 * kTileSamples));
        }
    };
#pragma unroll
    for (uint32_t u = 0; u < kSBPrefetch; ++u) load_sub(u, z[u]);
    for (uint32_t m0 = 0; m0 < num_subs; m0 += kSBPrefetch) {
#pragma unroll
      for (uint32_t u = 0; u < kSBPrefetch; ++u) {
        const uint32_t m = m0 + u;
        const uint32_t st = m / kSubsPerStage, sub = m % kSubsPerStage;  // stage counter, sub-stage inside it
        const uint32_t s = st % kSNS, fill = st / kSNS;
        if (sub == 0 && fill > 0) mbar_wait_suspend(&empty_b[s], (fill - 1) & 1u);  // the MMAs that read this stage have completed
        const uint32_t stage = smem_base + s * kSStageBytes + sub * kSSub * 2 * kSLBO;
#pragma unroll
        for (uint32_t c = 0; c < kSTiles; ++c)
#pragma unroll
          for (uint32_t k = 0; k < kSSub; ++k)
            sts128(stage + b_off0 + c * (kSN / 8) * kSSBO + k * 2 * kSLBO, z[u][c][k].x & kXMask, z[u][c][k].y & kXMask, z[u][c][k].z & kXMask, z[u][c][k].w & kXMask);
        load_sub(m + kSBPrefetch, z[u]);  // refill the registers just consumed
        if (sub == kSubsPerStage - 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (tensor core)
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_b[s]);
        }
      }
    }
  } else if (warp < kSExpWarps + kSIssuers) {
    // ===== MMA issuers: warp 13 the column tiles 0 and 1, warp 14 the tiles 2 and 3 - one N = 160 (80 where only one of the two
    // exists) MMA per step each, into their own accumulator columns; the whole warp runs the loop, one lane issues.  An issuer
    // without tiles keeps step with the barriers (its commits complete at once) so that the phase counts stay aligned.
    const uint32_t which = warp - kSExpWarps;
    const uint32_t my_tiles = ntiles > 2 * which ? min(2u, ntiles - 2 * which) : 0u;
    const uint32_t idesc = make_idesc_mxf4(kSM, my_tiles == 2 ? 2 * kSN : kSN);
    const uint32_t d_addr = tmem_base + kSColAcc + which * 2 * kSN, a_addr = tmem_base + kSColA, sf_addr = tmem_base + kSColSF;
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem) + which * (2 * kSN / 8) * kSSBO, kSLBO, kSSBO);
    const uint32_t elected = elect_one();
    for (uint32_t step = 0; step < num_steps; step += kSAS) {
      const uint32_t stage_no = step / kSAS, astage = stage_no % kSAStages, mb = step / kSBS, sb = mb % kSNS, qb = step % kSBS;
      if (qb == 0) mbar_wait_suspend(&full_b[sb], (mb / kSNS) & 1u);
      mbar_wait_suspend(&full_a[astage], (stage_no / kSAStages) & 1u);
      tcgen05_after_sync();
      if (elected) {
        if (my_tiles != 0) {
#pragma unroll
          for (uint32_t a = 0; a < kSAS; ++a) {
            const uint32_t b_bytes = sb * kSStageBytes + (qb + a) * 2 * kSLBO;
            umma_mxf4_ts(d_addr, a_addr + (astage * kSAS + a) * 8, b_desc0 + uint64_t(b_bytes >> 4), idesc, sf_addr, (step + a) > 0 ? 1u : 0u);
          }
        }
        umma_commit_arrive(&empty_a[astage]);                     // arrives when this thread's MMAs so far have completed
        if (qb + kSAS == kSBS) umma_commit_arrive(&empty_b[sb]);  // last steps of the B stage
      }
      __syncwarp();
    }
    if (elected) umma_commit_arrive(&acc_bar);  // this issuer's accumulator columns are final
  }

  // ===== epilogue: all 16 warps; thread = row (TMEM lane quadrant warp % 4), 20 columns of each column tile per warp group =====
  uint32_t any_mask = 0;  // bit k: column tile k holds a candidate seen by this thread
  {
    __syncwarp();
    mbar_wait_suspend(&acc_bar, 0);
    tcgen05_after_sync();
    const uint32_t quad = warp & 3, group = warp >> 2;
    const uint32_t r = quad * 32 + lane;
    const uint32_t gi = i0 + r;
    const uint32_t lane_base = tmem_base + ((quad * 32u) << 16);
    // see the header: candidate iff L < 2 (0.5 - thr) min(Het_i, Het_j); margins as in king_screen_kernel.cu
    const float bound2 = 2.f * (0.5f - p.kin_threshold) * 1.0001f;
    const float sites = __uint2float_rn(p.num_sites);
    const uint2 tot_i = __ldg(p.sample_totals + p.row_slot0 + row0 + (r < rows_here ? r : 0u));  // (het, hom)
    const float het_i = __uint2float_rn(tot_i.x), hom_i = __uint2float_rn(tot_i.y), def_i = __uint2float_rn(tot_i.x + tot_i.y);
    auto screen = [&](uint32_t c, uint32_t xx) -> bool {
      const uint32_t gj = j0 + c;
      const bool pair = r < rows_here && c < cols_here && gi < gj;
      const uint2 tot_j = __ldg(p.sample_totals + p.col_slot0 + col0 + (c < cols_here ? c : 0u));  // warp-uniform address
      const float het_j = __uint2float_rn(tot_j.x), hom_j = __uint2float_rn(tot_j.y), def_j = __uint2float_rn(tot_j.x + tot_j.y);
      // all quantities are integers below 2^24 (num_sites <= 2^23): exact in fp32
      const float lower = 0.5f * (fmaxf(0.f, hom_i + def_j - sites) + fmaxf(0.f, hom_j + def_i - sites)) - __uint_as_float(xx);
      return pair && lower < fmaf(bound2, fminf(het_i, het_j), 1.f);
    };
    constexpr uint32_t kColsPerGroup = kSN / 4;  // 20
    const uint32_t c0 = group * kColsPerGroup;
    for (uint32_t tile = 0; tile < ntiles; ++tile) {  // ntiles is uniform over the CTA
      const uint32_t cb = tile * kSN + c0;
      bool any = false;
      {
        uint32_t xx[16];
        tmem_load16(lane_base + kSColAcc + cb, xx);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (uint32_t k = 0; k < 16; ++k) any = screen(cb + k, xx[k]) || any;
      }
      {
        uint32_t xx[4];
        tmem_load4(lane_base + kSColAcc + cb + 16, xx);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) any = screen(cb + 16 + k, xx[k]) || any;
      }
      if (any) any_mask |= 1u << tile;
    }
  }
  tcgen05_before_sync();
  uint32_t flagged = 0;
#pragma unroll
  for (uint32_t tile = 0; tile < kSTiles; ++tile)  // one round per column tile (every thread takes part in every round)
    if (__syncthreads_or(int((any_mask >> tile) & 1u))) flagged |= 1u << tile;
  if (tid == 0 && flagged != 0) {
    for (uint32_t tile = 0; tile < ntiles; ++tile)
      if ((flagged >> tile) & 1u) p.tile_flags[blockIdx.x + size_t(tile) * band_rows] = 1;
    atomicAdd(p.flagged_counter, (unsigned long long)__popc(flagged));
  }
  __syncwarp();
  if (warp == kSExpWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kSTmemCols));
}

}  // namespace

cudaError_t launch_king_screen1(const KingLaunch &part, const BandTiles &tiles, cudaStream_t s) {
  static std::atomic<uint64_t> configured{0};  // one bit per device
  if (cudaError_t e = optin_dynamic_smem(king_screen1_kernel, kSSmem, configured); e != cudaSuccess) return e;
  king_screen1_kernel<<<unsigned(part.tile_end - part.tile_begin), kSThreads, kSSmem, s>>>(part, tiles);
  return cudaGetLastError();
}

}  // namespace ck
