// Tiled pairwise KING kernel for sm_100a (LOP3 + POPC path).
//
// Replaces ComputeKingKernel, /root/reference/cuking.cu:191-314.  The reference runs one 128-thread block per
// sample pair and re-reads both samples' 2 x 25 KB from L2/HBM for every pair (:218-240), reduces six counters
// with shuffles + shared atomics (:244-282) and appends through one global atomic per pair (:297-312).
//
// Here one CTA owns a 64 x 64-sample tile.  The K loop walks the site words in chunks of kChunkWords; each chunk
// of the row block and of the column block is ONE contiguous 12 KB run of the tile-major compute planes
// (layout.cuh) that a single elected thread stages into shared memory with cp.async.bulk (TMA, completes on an
// mbarrier) through a kStages-deep ring.  Each of the 256 threads owns a 4 x 4 block of pairs and keeps its
// 16 x 5 counters in registers for the whole K loop — no cross-thread reduction exists.  Per 32 sites and pair it
// issues 5 POPC + 6 LOP3 (the reference: 12 POPC.32 + ~22 LOP3): only five counters are
// independent, shared_sites == opp + conc + het_i + het_j - both_het, so concordant_hom is recovered in the
// epilogue.  The CSA variant folds two site words per counter with a carry-save adder (2 LOP3) before counting,
// halving POPC — the quarter-rate pipe (16 lanes/clk/SM measured, profiles/r01_int_pipe_peaks.jsonl) — at the
// price of 5 more LOP3 per pair-word.
// Epilogue: kinship in the reference's fp32 operation order (:289-294), strict threshold (:297), warp-aggregated
// slot reservation (one atomic per warp and ballot instead of one per pair, :299).
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.cuh"
#include "king_common.cuh"

namespace ck {

namespace {

constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr uint32_t kChunkU32 = kChunkWords * kComputePlanes * kTileSamples;  // 3072 words = 12 KB per block-chunk
constexpr uint32_t kChunkBytes = kChunkU32 * 4;
constexpr size_t kSmemBytes = size_t(kStages) * 2 * kChunkBytes;             // 96 KB

// One LOP3 with an explicit truth table (a = 0xF0, b = 0xCC, c = 0xAA).  Written as PTX so that ptxas keeps the
// intended 3-input functions; left to itself the compiler re-associates the carry-save majority across the AND
// terms that feed it and emits ~60 % more LOP3s (18 instead of 11 per pair-word).
template <int kLut>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(kLut));
  return d;
}
constexpr int kLutXor3 = 0x96;      // a ^ b ^ c
constexpr int kLutMaj = 0xE8;       // majority(a, b, c)
constexpr int kLutAndNor = 0x10;    // a & ~(b | c)
constexpr int kLutAndXor = 0x60;    // a & (b ^ c)

// Linear tile index -> (row block, column block).  Triangular grids enumerate bj >= bi row-major.
__device__ __forceinline__ void tile_coords(uint64_t t, uint32_t nrb, uint32_t ncb, bool triangular, uint32_t *bi,
                                            uint32_t *bj) {
  if (!triangular) {
    *bi = uint32_t(t / ncb);
    *bj = uint32_t(t % ncb);
    return;
  }
  // row b starts at off(b) = b*n - b(b-1)/2 ; invert with a double sqrt and fix up
  const double n2 = 2.0 * double(nrb) + 1.0;
  double bd = (n2 - sqrt(n2 * n2 - 8.0 * double(t))) * 0.5;
  long long b = (long long)bd;
  if (b < 0) b = 0;
  if (b >= (long long)nrb) b = (long long)nrb - 1;
  auto off = [&](long long x) { return (unsigned long long)(x * (long long)nrb - x * (x - 1) / 2); };
  while (b > 0 && off(b) > t) --b;
  while (b + 1 < (long long)nrb && off(b + 1) <= t) ++b;
  *bi = uint32_t(b);
  *bj = uint32_t(b + (long long)(t - off(b)));
}

template <bool kCsa>
__global__ void __launch_bounds__(kThreads, 1) king_tile_kernel(const KingLaunch p) {
  extern __shared__ __align__(128) uint32_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kStages];

  const uint32_t tid = threadIdx.x;
  uint32_t bi, bj;
  tile_coords(p.tile_begin + blockIdx.x, p.num_row_blocks, p.num_col_blocks, p.triangular != 0, &bi, &bj);

  // global sample index range of this tile; leave early if no pair has i < j
  const uint32_t i0 = p.row_global0 + bi * kTileSamples;
  const uint32_t j0 = p.col_global0 + bj * kTileSamples;
  const uint32_t rows_here = min(kTileSamples, p.num_rows - bi * kTileSamples);
  const uint32_t cols_here = min(kTileSamples, p.num_cols - bj * kTileSamples);
  if (i0 >= j0 + cols_here - 1) return;  // every i >= every j (also covers cols_here == 1 on the diagonal)

  const bool same_block = (p.row_block0 + bi) == (p.col_block0 + bj);
  const uint32_t *row_src = p.compute + size_t(p.row_block0 + bi) * p.words * (kComputePlanes * kTileSamples);
  const uint32_t *col_src = p.compute + size_t(p.col_block0 + bj) * p.words * (kComputePlanes * kTileSamples);
  const uint32_t num_chunks = p.words / kChunkWords;
  const uint32_t stage_bytes = same_block ? kChunkBytes : 2 * kChunkBytes;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](uint32_t chunk) {
    const uint32_t s = chunk % kStages;
    uint32_t *dst = smem + size_t(s) * 2 * kChunkU32;
    mbar_expect_tx(&full_bar[s], stage_bytes);
    bulk_load(dst, row_src + size_t(chunk) * kChunkU32, kChunkBytes, &full_bar[s]);
    if (!same_block) bulk_load(dst + kChunkU32, col_src + size_t(chunk) * kChunkU32, kChunkBytes, &full_bar[s]);
  };
  if (tid == 0) {
    for (uint32_t c = 0; c < uint32_t(kStages - 1) && c < num_chunks; ++c) issue(c);
  }

  const uint32_t ty = tid >> 4, tx = tid & 15;  // thread owns rows 4ty..4ty+3, cols 4tx..4tx+3 of the tile
  uint32_t acc[4][4][5];
  uint32_t ones[kCsa ? 4 : 1][kCsa ? 4 : 1][5];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        acc[a][b][c] = 0;
        if constexpr (kCsa) ones[a][b][c] = 0;
      }

  for (uint32_t chunk = 0; chunk < num_chunks; ++chunk) {
    if (tid == 0 && chunk + (kStages - 1) < num_chunks) issue(chunk + (kStages - 1));
    const uint32_t s = chunk % kStages;
    mbar_wait(&full_bar[s], (chunk / kStages) & 1u);
    const uint32_t *rs = smem + size_t(s) * 2 * kChunkU32 + 4 * ty;
    const uint32_t *cs = smem + size_t(s) * 2 * kChunkU32 + (same_block ? 0u : kChunkU32) + 4 * tx;

    if constexpr (!kCsa) {
#pragma unroll 2
      for (uint32_t k = 0; k < kChunkWords; ++k) {
        const uint32_t o = k * (kComputePlanes * kTileSamples);
        const uint4 rH4 = *reinterpret_cast<const uint4 *>(rs + o + kPlaneH * kTileSamples);
        const uint4 rD4 = *reinterpret_cast<const uint4 *>(rs + o + kPlaneD * kTileSamples);
        const uint4 rA4 = *reinterpret_cast<const uint4 *>(rs + o + kPlaneA * kTileSamples);
        const uint4 cH4 = *reinterpret_cast<const uint4 *>(cs + o + kPlaneH * kTileSamples);
        const uint4 cD4 = *reinterpret_cast<const uint4 *>(cs + o + kPlaneD * kTileSamples);
        const uint4 cA4 = *reinterpret_cast<const uint4 *>(cs + o + kPlaneA * kTileSamples);
        const uint32_t rH[4] = {rH4.x, rH4.y, rH4.z, rH4.w}, rD[4] = {rD4.x, rD4.y, rD4.z, rD4.w},
                       rA[4] = {rA4.x, rA4.y, rA4.z, rA4.w};
        const uint32_t cH[4] = {cH4.x, cH4.y, cH4.z, cH4.w}, cD[4] = {cD4.x, cD4.y, cD4.z, cD4.w},
                       cA[4] = {cA4.x, cA4.y, cA4.z, cA4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t dd = rD[a] & cD[b];
            acc[a][b][0] += __popc(rH[a] & cH[b]);                            // both het
            acc[a][b][1] += __popc(rH[a] & cD[b]);                            // het_i where j defined
            acc[a][b][2] += __popc(rD[a] & cH[b]);                            // het_j where i defined
            acc[a][b][3] += __popc(dd);                                       // jointly defined
            acc[a][b][4] += __popc(lop3<kLutAndXor>(lop3<kLutAndNor>(dd, rH[a], cH[b]), rA[a], cA[b]));  // both hom, alleles differ
          }
      }
    } else {
#pragma unroll 1
      for (uint32_t k = 0; k < kChunkWords; k += 2) {
        const uint32_t o0 = k * (kComputePlanes * kTileSamples), o1 = o0 + kComputePlanes * kTileSamples;
        uint32_t rH[2][4], rD[2][4], rA[2][4], cH[2][4], cD[2][4], cA[2][4];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t o = w ? o1 : o0;
          const uint4 a0 = *reinterpret_cast<const uint4 *>(rs + o + kPlaneH * kTileSamples);
          const uint4 a1 = *reinterpret_cast<const uint4 *>(rs + o + kPlaneD * kTileSamples);
          const uint4 a2 = *reinterpret_cast<const uint4 *>(rs + o + kPlaneA * kTileSamples);
          const uint4 b0 = *reinterpret_cast<const uint4 *>(cs + o + kPlaneH * kTileSamples);
          const uint4 b1 = *reinterpret_cast<const uint4 *>(cs + o + kPlaneD * kTileSamples);
          const uint4 b2 = *reinterpret_cast<const uint4 *>(cs + o + kPlaneA * kTileSamples);
          rH[w][0] = a0.x; rH[w][1] = a0.y; rH[w][2] = a0.z; rH[w][3] = a0.w;
          rD[w][0] = a1.x; rD[w][1] = a1.y; rD[w][2] = a1.z; rD[w][3] = a1.w;
          rA[w][0] = a2.x; rA[w][1] = a2.y; rA[w][2] = a2.z; rA[w][3] = a2.w;
          cH[w][0] = b0.x; cH[w][1] = b0.y; cH[w][2] = b0.z; cH[w][3] = b0.w;
          cD[w][0] = b1.x; cD[w][1] = b1.y; cD[w][2] = b1.z; cD[w][3] = b1.w;
          cA[w][0] = b2.x; cA[w][1] = b2.y; cA[w][2] = b2.z; cA[w][3] = b2.w;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            uint32_t x[2][5];
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const uint32_t dd = rD[w][a] & cD[w][b];
              x[w][0] = rH[w][a] & cH[w][b];
              x[w][1] = rH[w][a] & cD[w][b];
              x[w][2] = rD[w][a] & cH[w][b];
              x[w][3] = dd;
              x[w][4] = lop3<kLutAndXor>(lop3<kLutAndNor>(dd, rH[w][a], cH[w][b]), rA[w][a], cA[w][b]);
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              // carry-save: ones + x0 + x1 = (ones ^ x0 ^ x1) + 2 * majority(ones, x0, x1)
              const uint32_t o = ones[a][b][c];
              ones[a][b][c] = lop3<kLutXor3>(o, x[0][c], x[1][c]);
              acc[a][b][c] += __popc(lop3<kLutMaj>(o, x[0][c], x[1][c]));
            }
          }
      }
    }
    __syncthreads();  // every thread is done with stage s before thread 0 refills it next iteration
  }

  // ---- epilogue: counters -> kinship -> threshold -> warp-aggregated append ----
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t both_het = acc[a][b][0], het_i = acc[a][b][1], het_j = acc[a][b][2], shared = acc[a][b][3],
               opp = acc[a][b][4];
      if constexpr (kCsa) {
        both_het = 2 * both_het + __popc(ones[a][b][0]);
        het_i = 2 * het_i + __popc(ones[a][b][1]);
        het_j = 2 * het_j + __popc(ones[a][b][2]);
        shared = 2 * shared + __popc(ones[a][b][3]);
        opp = 2 * opp + __popc(ones[a][b][4]);
      }
      // shared = opp + conc + het_i + het_j - both_het  (each jointly defined site is exactly one of: opposing hom,
      // concordant hom, het/het, het/hom, hom/het)
      const uint32_t conc = shared - opp - het_i - het_j + both_het;
      const uint32_t r = 4 * ty + a, c = 4 * tx + b;
      const uint32_t gi = i0 + r, gj = j0 + c;
      const bool valid = r < rows_here && c < cols_here && gi < gj;  // cuking.cu:199
      const float kin = kinship(het_i, het_j, both_het, opp);
      if (p.dump_counts != nullptr && r < rows_here && c < cols_here) {
        const size_t idx = size_t(bi * kTileSamples + r) * p.num_cols + (bj * kTileSamples + c);
        ck_counts out;
        out.het_i = het_i; out.het_j = het_j; out.both_het = both_het;
        out.opposing_hom = opp; out.concordant_hom = conc; out.shared_sites = shared;
        p.dump_counts[idx] = out;
        p.dump_kin[idx] = kin;
      }
      emit_pair(p, valid, true, gi, gj, kin, opp, conc, both_het, shared);
    }
  }
}

}  // namespace

uint64_t king_num_tiles(uint32_t num_row_blocks, uint32_t num_col_blocks, bool triangular) {
  if (triangular) return uint64_t(num_row_blocks) * (uint64_t(num_row_blocks) + 1) / 2;
  return uint64_t(num_row_blocks) * num_col_blocks;
}

cudaError_t launch_king(const KingLaunch &k, int variant, cudaStream_t s, uint32_t *launches) {
  static std::atomic<uint64_t> configured[2] = {{0}, {0}};  // one bit per device
  if (cudaError_t e = optin_dynamic_smem(king_tile_kernel<false>, kSmemBytes, configured[0]); e != cudaSuccess) return e;
  if (cudaError_t e = optin_dynamic_smem(king_tile_kernel<true>, kSmemBytes, configured[1]); e != cudaSuccess) return e;
  // CUDA grids are limited to 2^31-1 blocks in x; slice very large tile ranges into several launches.
  constexpr uint64_t kMaxGrid = 1ull << 30;
  for (uint64_t t = k.tile_begin; t < k.tile_end; t += kMaxGrid) {
    KingLaunch part = k;
    part.tile_begin = t;
    part.tile_end = (t + kMaxGrid < k.tile_end) ? t + kMaxGrid : k.tile_end;
    const unsigned grid = unsigned(part.tile_end - part.tile_begin);
    if (variant == 1)
      king_tile_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(part);
    else
      king_tile_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(part);
    if (launches) ++*launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace ck
