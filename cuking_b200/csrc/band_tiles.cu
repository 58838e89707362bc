// Tile enumeration shared by the two tensor-core pairwise kernels (king_fp4_kernel.cu, king_umma_kernel.cu): 128-row
// tiles grouped in bands of kBandRowTiles row tiles, column-major inside a band, so that the ~148 tiles in flight share
// 8 row blocks and ~19 column blocks (the genotype codes are then served from L2) and so that a band only pairs its
// rows with samples at or after them (which the streaming host-buffer seam relies on).  A band starts at the first
// column tile that holds an i < j pair for its FIRST row tile; the few tiles of its later row tiles that lie wholly
// below the diagonal are enumerated too and exit at once.  The linear tile index is the multi-GPU partition unit.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "internal.cuh"

namespace ck {

namespace {

struct BandTable {
  std::vector<unsigned long long> band_prefix;
  std::vector<uint32_t> band_first_col;
  uint32_t num_bands = 0, num_row_tiles = 0, num_col_tiles = 0;
};

// first column tile holding an i < j pair for row tile ti (num_col_tiles if there is none)
uint32_t first_alive_col(const KingLaunch &k, uint32_t tile_cols, uint32_t ti, uint32_t num_col_tiles) {
  const uint64_t i_min = uint64_t(k.row_global0) + uint64_t(ti) * kBandTileRows;
  if (uint64_t(k.col_global0) + k.num_cols - 1 <= i_min) return num_col_tiles;
  if (i_min < k.col_global0) return 0;
  const uint64_t need = i_min - k.col_global0 + 1;  // need a local column index >= need in the tile
  const uint32_t first = uint32_t(need / tile_cols);  // the tile that holds local column `need`
  return first < num_col_tiles ? first : num_col_tiles - 1;
}

BandTable build_band_table(const KingLaunch &k, uint32_t tile_cols) {
  BandTable bt;
  bt.num_row_tiles = ceil_div(k.num_rows, kBandTileRows);
  bt.num_col_tiles = ceil_div(k.num_cols, tile_cols);
  bt.num_bands = ceil_div(bt.num_row_tiles, kBandRowTiles);
  bt.band_prefix.assign(bt.num_bands + 1, 0);
  bt.band_first_col.assign(std::max<uint32_t>(bt.num_bands, 1), 0);
  for (uint32_t b = 0; b < bt.num_bands; ++b) {
    // an odd last band is padded with one phantom row tile (it exits at once) so that consecutive tiles (2m, 2m + 1) of a
    // band always share their column tile: the CTA-pair kernel runs them as one 256-row tile
    const uint32_t rows = band_rows_padded(bt.num_row_tiles - b * kBandRowTiles);
    const uint32_t first = first_alive_col(k, tile_cols, b * kBandRowTiles, bt.num_col_tiles);  // non-decreasing in the row tile
    bt.band_first_col[b] = first;
    bt.band_prefix[b + 1] = bt.band_prefix[b] + uint64_t(rows) * (bt.num_col_tiles - first);
  }
  return bt;
}

}  // namespace

uint64_t band_num_tiles(const KingLaunch &k, uint32_t tile_cols) {
  if (k.num_rows == 0 || k.num_cols == 0) return 0;
  return build_band_table(k, tile_cols).band_prefix.back();
}

cudaError_t band_prepare(const KingLaunch &k, uint32_t tile_cols, ck_ctx *ctx, cudaStream_t s, std::vector<uint64_t> *band_prefix,
                         BandTiles *tiles) {
  const uint64_t key[3] = {tile_cols, (uint64_t(k.num_rows) << 32) | k.num_cols, (uint64_t(k.row_global0) << 32) | k.col_global0};
  const uint32_t num_row_tiles = ceil_div(k.num_rows, kBandTileRows), num_bands = ceil_div(num_row_tiles, kBandRowTiles);
  const size_t first_bytes = size_t(std::max<uint32_t>(num_bands, 1)) * 4;
  auto fill_tiles = [&] {
    if (!tiles) return;
    tiles->band_prefix = static_cast<unsigned long long *>(ctx->tile_table);
    tiles->band_first_col = reinterpret_cast<uint32_t *>(static_cast<char *>(ctx->tile_table) + ctx->tile_table_bytes - first_bytes);
    tiles->num_bands = num_bands;
    tiles->num_row_tiles = num_row_tiles;
    tiles->num_col_tiles = ceil_div(k.num_cols, tile_cols);
  };
  const bool cached = ctx->tile_table != nullptr && ctx->tile_table_key[0] == key[0] && ctx->tile_table_key[1] == key[1] &&
                      ctx->tile_table_key[2] == key[2];
  if (cached && band_prefix == nullptr) {
    fill_tiles();
    return cudaSuccess;
  }
  const BandTable bt = build_band_table(k, tile_cols);
  if (band_prefix) band_prefix->assign(bt.band_prefix.begin(), bt.band_prefix.end());
  if (cached) {
    fill_tiles();
    return cudaSuccess;
  }
  const size_t prefix_bytes = (bt.band_prefix.size() * 8 + 255) & ~size_t(255);
  ctx->tile_table_key[0] = ~0ull;
  if (ctx->tile_table_bytes < prefix_bytes + first_bytes) {  // grow-only scratch owned by the ctx
    if (ctx->tile_table) cudaFree(ctx->tile_table);
    ctx->tile_table = nullptr;
    ctx->tile_table_bytes = 0;
    cudaError_t e = dev_alloc(ctx, &ctx->tile_table, prefix_bytes + first_bytes);
    if (e != cudaSuccess) return e;
    ctx->tile_table_bytes = prefix_bytes + first_bytes;
  }
  fill_tiles();
  auto *d_prefix = static_cast<unsigned long long *>(ctx->tile_table);
  auto *d_first = reinterpret_cast<uint32_t *>(static_cast<char *>(ctx->tile_table) + ctx->tile_table_bytes - first_bytes);
  cudaError_t e = cudaMemcpyAsync(d_prefix, bt.band_prefix.data(), bt.band_prefix.size() * 8, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_first, bt.band_first_col.data(), first_bytes, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the host vectors die with this frame
  if (e == cudaSuccess) {
    ctx->tile_table_key[0] = key[0];
    ctx->tile_table_key[1] = key[1];
    ctx->tile_table_key[2] = key[2];
  }
  return e;
}

}  // namespace ck
