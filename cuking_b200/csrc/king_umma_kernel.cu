// Pairwise KING on the 5th-generation tensor cores: tcgen05.mma.kind::i8 over indicator vectors (sm_100a only).
//
// The six counters of ComputeKingKernel (/root/reference/cuking.cu:214-240) are bilinear forms of per-sample indicator
// vectors over the sites.  With h = [het], y = [hom-ref or hom-alt], x = [hom-alt] - [hom-ref] (all 0 where the
// genotype is missing):
//     x_i.x_j = conc - opp     y_i.y_j = conc + opp     h_i.h_j = both_het
//     h_i.y_j = (i het, j hom) y_i.h_j = (i hom, j het)
//     het_i = hh + hy,  het_j = hh + yh,  shared = yy + yh + hy + hh,  conc = (yy + xx)/2,  opp = (yy - xx)/2
// i.e. five int8 GEMMs with exact s32 accumulation (counts <= num_sites < 2^31), issued as three MMAs per 32 sites:
//     D_xx[128 x 96]        += x_i  . x_j^T
//     D_y [128 x (96|96)]   += y_i  . [y_j ; h_j]^T        -> (yy | yh)
//     D_h [128 x (96|96)]   += h_i  . [y_j ; h_j]^T        -> (hy | hh)
// which fills 480 of the 512 TMEM columns of the SM.  One CTA owns a 128 (rows, TMEM lanes) x 96 (columns) tile of
// sample pairs.  The int8 operands are never stored in HBM (they would be 8x the bit planes and the kernel would turn
// L2-bound): warps 0-6 expand the compute bit planes (H, D, A; csrc/layout.cuh) into K-major no-swizzle canonical
// shared-memory tiles each stage, one thread per sample, and hand them to the single MMA-issuing thread through
// mbarriers; tcgen05.commit releases a stage when the tensor core has consumed it.
//
// Measured on B200 (tools/umma_i8_probe.cu): kind::i8 peaks at 8192 MAC/clk/SM, but an M=128 MMA with both operands
// in shared memory takes >= ~91-112 clk whatever N is (A-operand read), so MMAs must be wide: N = 192 for the two
// stacked products; only the x.x product (N = 96) runs below peak.
//
// The site order inside a 32-site K step is permuted (site 8b+j of the word -> K byte 4j+b) identically for both
// operands, which leaves every dot product unchanged and makes the expansion 2 ALU ops per 4 bytes:
// (word >> j) & 0x01010101.
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "internal.cuh"
#include "king_common.cuh"

namespace ck {

namespace {

constexpr uint32_t kUM = 128, kUN = 96;            // tile rows (A operand, TMEM lanes) x tile columns (B operand)
constexpr uint32_t kStageWords = 2;                // 32-site words per pipeline stage
constexpr uint32_t kStageK = 32 * kStageWords;     // K bytes per stage
constexpr uint32_t kUStages = 4;
constexpr uint32_t kLBO = 128;                     // bytes between K-adjacent 8x16-byte core matrices
constexpr uint32_t kSBO = (kStageK / 16) * 128;    // bytes between 8-row groups
constexpr uint32_t kATile = (kUM / 8) * kSBO;      // one A operand plane of one stage
constexpr uint32_t kBTile = (kUN / 8) * kSBO;
constexpr uint32_t kStageBytes = 3 * kATile + 3 * kBTile;
constexpr size_t kUmmaSmem = size_t(kUStages) * kStageBytes + 1024;  // + alignment slack
constexpr uint32_t kUThreads = 256;                // warps 0-3: A expanders + epilogue, 4-6: B expanders, 7: MMA issuer
constexpr uint32_t kExpanderWarps = 7;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColXX = 0, kColY = kUN, kColH = 3 * kUN;  // accumulator column bases: xx | (yy|yh) | (hy|hh)

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  // K-major, no swizzle: ((8,n),2):((16 B, SBO), LBO); version 1 (Blackwell)
  return uint64_t((smem_addr >> 4) & 0x3fff) | (uint64_t((kLBO >> 4) & 0x3fff) << 16) |
         (uint64_t((kSBO >> 4) & 0x3fff) << 32) | (uint64_t(1) << 46);
}
__host__ __device__ constexpr uint32_t make_idesc_i8(uint32_t M, uint32_t N, bool a_signed, bool b_signed) {
  return (2u << 4) /* D = s32 */ | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
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

struct UmmaTiles {  // alive-tile enumeration (tiles with at least one i < j pair), built on the host per launch
  const unsigned long long *row_prefix;  // [num_row_tiles + 1] alive tiles before row tile t
  const uint32_t *first_col;             // [num_row_tiles] first alive column tile of row tile t
  uint32_t num_row_tiles, num_col_tiles;
  uint32_t total_blocks;                 // 64-sample plane blocks allocated (reads beyond are treated as missing)
};

// Expands one 32-site word of one sample into the three int8 operand rows (32 K bytes each) of the canonical tile.
//   out byte 4j+b  <-  site 8b+j
__device__ __forceinline__ void expand_store(uint32_t H, uint32_t D, uint32_t A, uint8_t *op_x, uint8_t *op_y,
                                             uint8_t *op_h, uint32_t row_off, uint32_t word_in_stage, uint32_t tile_bytes) {
  (void)tile_bytes;
  const uint32_t R = D & ~H & ~A;  // hom-ref
  uint32_t x[8], y[8], h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t hj = (H >> j) & 0x01010101u;
    const uint32_t aj = (A >> j) & 0x01010101u;
    const uint32_t rj = (R >> j) & 0x01010101u;
    h[j] = hj;
    y[j] = aj | rj;
    x[j] = rj * 255u + aj;  // bytes: +1 hom-alt, 0xFF = -1 hom-ref (disjoint, no inter-byte carry)
  }
  const uint32_t off = row_off + word_in_stage * 2 * kLBO;  // a word = 32 K bytes = two core matrices along K
  *reinterpret_cast<uint4 *>(op_x + off) = make_uint4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<uint4 *>(op_x + off + kLBO) = make_uint4(x[4], x[5], x[6], x[7]);
  *reinterpret_cast<uint4 *>(op_y + off) = make_uint4(y[0], y[1], y[2], y[3]);
  *reinterpret_cast<uint4 *>(op_y + off + kLBO) = make_uint4(y[4], y[5], y[6], y[7]);
  *reinterpret_cast<uint4 *>(op_h + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4 *>(op_h + off + kLBO) = make_uint4(h[4], h[5], h[6], h[7]);
}

__global__ void __launch_bounds__(kUThreads, 1) king_umma_kernel(const KingLaunch p, const UmmaTiles tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kUStages], empty_bar[kUStages], acc_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which tile ----
  const unsigned long long t = p.tile_begin + blockIdx.x;
  uint32_t lo = 0, hi = tiles.num_row_tiles;  // largest ti with row_prefix[ti] <= t
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (tiles.row_prefix[mid] <= t) lo = mid; else hi = mid;
  }
  const uint32_t ti = lo, tj = tiles.first_col[ti] + uint32_t(t - tiles.row_prefix[ti]);
  const uint32_t row0 = ti * kUM, col0 = tj * kUN;           // offsets inside the sub-matrix
  const uint32_t rows_here = min(kUM, p.num_rows - row0), cols_here = min(kUN, p.num_cols - col0);
  const uint32_t i0 = p.row_global0 + row0, j0 = p.col_global0 + col0;

  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < kUStages; ++s) {
      mbar_init(&full_bar[s], kExpanderWarps);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t num_stages_total = p.words / kStageWords;  // p.words is a multiple of 16

  if (warp < kExpanderWarps) {
    // ================= expanders: one thread per sample =================
    const bool is_a = warp < 4;
    const uint32_t srow = is_a ? tid : tid - 128;                                   // row of the A tile / of the B tile
    const uint32_t slot = is_a ? p.row_block0 * kTileSamples + row0 + srow : p.col_block0 * kTileSamples + col0 + srow;
    const uint32_t blk = slot / kTileSamples, ln = slot % kTileSamples;
    const bool in_range = blk < tiles.total_blocks && (is_a ? srow < rows_here : srow < cols_here);
    const uint32_t *src = p.compute + (size_t(blk) * p.words * kComputePlanes) * kTileSamples + ln;
    const uint32_t op_base = is_a ? 0u : 3 * kATile;
    const uint32_t tile_bytes = is_a ? kATile : kBTile;
    const uint32_t row_off = (srow >> 3) * kSBO + (srow & 7) * 16;

    auto load_stage = [&](uint32_t st, uint32_t (&w)[kStageWords][3]) {
#pragma unroll
      for (uint32_t q = 0; q < kStageWords; ++q)
#pragma unroll
        for (uint32_t pl = 0; pl < 3; ++pl)
          w[q][pl] = (in_range && st < num_stages_total)
                         ? __ldg(src + (size_t(st * kStageWords + q) * kComputePlanes + pl) * kTileSamples)
                         : 0u;  // out-of-range samples / stages: everything missing
    };
    uint32_t w0[kStageWords][3], w1[kStageWords][3], w2[kStageWords][3];
    load_stage(0, w0);
    load_stage(1, w1);
    for (uint32_t st = 0; st < num_stages_total; ++st) {
      load_stage(st + 2, w2);  // prefetch two stages ahead (covers L2/HBM latency)
      const uint32_t s = st % kUStages, fill = st / kUStages;
      if (fill > 0) mbar_wait(&empty_bar[s], (fill - 1) & 1u);  // the tensor core has consumed the previous fill
      uint8_t *stage = smem + size_t(s) * kStageBytes + op_base;
#pragma unroll
      for (uint32_t q = 0; q < kStageWords; ++q)
        expand_store(w0[q][kPlaneH], w0[q][kPlaneD], w0[q][kPlaneA], stage, stage + tile_bytes, stage + 2 * tile_bytes,
                     row_off, q, tile_bytes);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (tensor core)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
#pragma unroll
      for (uint32_t q = 0; q < kStageWords; ++q)
#pragma unroll
        for (uint32_t pl = 0; pl < 3; ++pl) {
          w0[q][pl] = w1[q][pl];
          w1[q][pl] = w2[q][pl];
        }
    }
  } else if (lane == 0) {
    // ================= MMA issuer: one thread =================
    constexpr uint32_t idesc_xx = make_idesc_i8(kUM, kUN, true, true);
    constexpr uint32_t idesc_yh = make_idesc_i8(kUM, 2 * kUN, false, false);
    const uint32_t smem_base = smem_u32(smem);
    for (uint32_t st = 0; st < num_stages_total; ++st) {
      const uint32_t s = st % kUStages, fill = st / kUStages;
      mbar_wait(&full_bar[s], fill & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_base + s * kStageBytes, b0 = a0 + 3 * kATile;
#pragma unroll
      for (uint32_t q = 0; q < kStageWords; ++q) {
        const uint32_t koff = q * 2 * kLBO;
        const uint32_t acc = (st > 0 || q > 0) ? 1u : 0u;
        umma_i8(tmem_base + kColXX, make_smem_desc(a0 + koff), make_smem_desc(b0 + koff), idesc_xx, acc);
        umma_i8(tmem_base + kColY, make_smem_desc(a0 + kATile + koff), make_smem_desc(b0 + kBTile + koff), idesc_yh, acc);
        umma_i8(tmem_base + kColH, make_smem_desc(a0 + 2 * kATile + koff), make_smem_desc(b0 + kBTile + koff), idesc_yh, acc);
      }
      umma_commit(&empty_bar[s]);  // arrives when the MMAs above have finished reading this stage
    }
    umma_commit(&acc_bar);  // all accumulators final
  }

  // ================= epilogue: warps 0-3, thread = row (TMEM lane) =================
  if (warp < 4) {
    mbar_wait(&acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t r = tid;
    const uint32_t gi = i0 + r;
    const uint32_t lane_base = tmem_base + ((warp * 32u) << 16);
    for (uint32_t c0 = 0; c0 < kUN; c0 += 16) {
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
        emit_pair(p, in_tile && gi < gj, gi, gj, kin, opp, conc, both_het, shared);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  __syncwarp();
  if (warp == 7) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

// ---- host side: alive-tile table ---------------------------------------------------------------------------------

struct TileTable {
  std::vector<unsigned long long> row_prefix;
  std::vector<uint32_t> first_col;
  uint32_t num_row_tiles = 0, num_col_tiles = 0;
};

TileTable build_tile_table(const KingLaunch &k) {
  TileTable tt;
  tt.num_row_tiles = ceil_div(k.num_rows, kUM);
  tt.num_col_tiles = ceil_div(k.num_cols, kUN);
  tt.row_prefix.assign(tt.num_row_tiles + 1, 0);
  tt.first_col.assign(std::max<uint32_t>(tt.num_row_tiles, 1), 0);
  for (uint32_t ti = 0; ti < tt.num_row_tiles; ++ti) {
    const uint64_t i_min = uint64_t(k.row_global0) + uint64_t(ti) * kUM;
    // first column tile whose largest j exceeds i_min:  col_global0 + min(96 tj + 95, num_cols - 1) > i_min
    uint32_t first = tt.num_col_tiles;
    if (uint64_t(k.col_global0) + k.num_cols - 1 > i_min) {
      if (i_min < k.col_global0) {
        first = 0;
      } else {
        const uint64_t need = i_min - k.col_global0 + 1;  // need local j_max >= need
        first = uint32_t(need <= kUN - 1 ? 0 : (need - (kUN - 1) + kUN - 1) / kUN);
        if (first >= tt.num_col_tiles) first = tt.num_col_tiles - 1;  // the last (ragged) tile holds num_cols - 1
      }
    }
    tt.first_col[ti] = first;
    tt.row_prefix[ti + 1] = tt.row_prefix[ti] + (tt.num_col_tiles - first);
  }
  return tt;
}

}  // namespace

uint64_t king_umma_num_tiles(const KingLaunch &k) {
  if (k.num_rows == 0 || k.num_cols == 0) return 0;
  return build_tile_table(k).row_prefix.back();
}

cudaError_t launch_king_umma(const KingLaunch &k, uint32_t total_blocks, cudaStream_t s, uint32_t *launches) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(king_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kUmmaSmem));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (k.tile_end <= k.tile_begin) return cudaSuccess;
  const TileTable tt = build_tile_table(k);
  unsigned long long *d_prefix = nullptr;
  uint32_t *d_first = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&d_prefix), tt.row_prefix.size() * 8, s);
  if (e != cudaSuccess) return e;
  e = cudaMallocAsync(reinterpret_cast<void **>(&d_first), tt.first_col.size() * 4, s);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(d_prefix, tt.row_prefix.data(), tt.row_prefix.size() * 8, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_first, tt.first_col.data(), tt.first_col.size() * 4, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the host vectors die with this frame
  UmmaTiles tiles{d_prefix, d_first, tt.num_row_tiles, tt.num_col_tiles, total_blocks};
  constexpr uint64_t kMaxGrid = 1ull << 30;
  for (uint64_t t = k.tile_begin; e == cudaSuccess && t < k.tile_end; t += kMaxGrid) {
    KingLaunch part = k;
    part.tile_begin = t;
    part.tile_end = (t + kMaxGrid < k.tile_end) ? t + kMaxGrid : k.tile_end;
    king_umma_kernel<<<unsigned(part.tile_end - part.tile_begin), kUThreads, kUmmaSmem, s>>>(part, tiles);
    if (launches) ++*launches;
    e = cudaGetLastError();
  }
  cudaFreeAsync(d_prefix, s);
  cudaFreeAsync(d_first, s);
  return e;
}

}  // namespace ck
