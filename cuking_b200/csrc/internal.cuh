// Internal declarations shared by the translation units of libcuking_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cuking_b200.h"
#include "layout.cuh"
#include "shard_plan.h"

namespace ck {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int fail_cuda(cudaError_t e, const char *what, const char *file, int line);

#define CK_CUDA(expr)                                                      \
  do {                                                                     \
    cudaError_t ck_e_ = (expr);                                            \
    if (ck_e_ != cudaSuccess) return ::ck::fail_cuda(ck_e_, #expr, __FILE__, __LINE__); \
  } while (0)

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
};

}  // namespace ck

struct ck_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;       // own_stream or a caller-supplied one
  cudaStream_t copy_stream = nullptr;  // host staging copies overlap the pack kernel
  cudaEvent_t ev[2] = {nullptr, nullptr};
  ck_timings timings{};
  int num_sms = 0;
  // scratch owned by the ctx
  unsigned long long *d_counter = nullptr;  // [0] emitted-pair counter
  uint32_t *d_pack_err = nullptr;           // [0] invalid-genotype index+1 (min), [1] out-of-range index+1 (min)
  void *pinned[2] = {nullptr, nullptr};   // host staging for ck_pack_triples(on_device = 0)
  void *staging[2] = {nullptr, nullptr};  // device side of the same double buffer
  size_t pinned_bytes = 0;
  cudaEvent_t pinned_free[2] = {nullptr, nullptr};
  // synthetic triples scratch
  int64_t *syn_row = nullptr, *syn_col = nullptr;
  int32_t *syn_alt = nullptr;
  size_t syn_cap = 0;
  // emitted-pair buffer reused across ck_king* calls
  ck_result *result_buf = nullptr;
  size_t result_cap = 0;
  int king_variant = -1;  // -1 = library default
  // alive-tile table of the tcgen05 kernel (grow-only device scratch)
  void *tile_table = nullptr;
  size_t tile_table_bytes = 0;
  uint64_t tile_table_key[3] = {~0ull, 0, 0};  // (variant, rows/cols, global origins) of the table now on the device
  // grow-only scratch for the result sort (keys, indices, CUB temporaries, sorted records): no cudaMalloc/cudaFree
  // on the steady-state path
  void *sort_scratch = nullptr;
  size_t sort_scratch_bytes = 0;
  // small cache of released device buffers (planes, staging) so that a create / destroy cycle per call - the
  // host-buffer entry point ck_king_host_bitset - does not pay cudaMalloc / cudaFree of gigabytes every time
  static constexpr int kCacheSlots = 8;
  void *cache_ptr[kCacheSlots] = {};
  size_t cache_bytes[kCacheSlots] = {};
};

namespace ck {
cudaError_t ctx_alloc(ck_ctx *ctx, void **ptr, size_t bytes);  // exact-size reuse from the cache, else cudaMalloc
void ctx_release(ck_ctx *ctx, void *ptr, size_t bytes);        // into the cache (evicting the smallest entry) or cudaFree
}  // namespace ck


namespace ck {
struct KingStream;  // below
}

struct ck_planes {
  ck_ctx *ctx = nullptr;
  size_t raw_bytes = 0, compute_bytes = 0, codes_bytes = 0;  // sizes of the three buffers as allocated
  ck::SlotMap map{};
  uint32_t num_sites = 0;
  uint32_t words = 0;          // padded 32-bit words per plane (multiple of kChunkWords)
  uint32_t *raw = nullptr;     // [num_blocks][words][2][64]
  uint32_t *compute = nullptr; // [num_blocks][words][3][64]   (LOP3+POPC kernels)
  uint32_t *codes = nullptr;   // [num_blocks][words][64][4]   (tcgen05 kernel; allocated on first use)
  ck::KingStream *stream_state = nullptr;  // ck_king_stream_begin .. end
  bool compute_stale = true;   // raw changed since the last finalize into `compute`
  bool codes_stale = true;     // raw changed since the last finalize into `codes`
  int codes_kind = 0;          // which kernel variant `codes` was last derived for (2: int8 selectors, 3: E2M1 nibbles)
  size_t raw_words() const { return size_t(map.num_blocks) * words * ck::kRawPlanes * ck::kTileSamples; }
  size_t compute_words() const { return size_t(map.num_blocks) * words * ck::kComputePlanes * ck::kTileSamples; }
  size_t codes_words() const { return size_t(map.num_blocks) * words * ck::kTileSamples * 4; }
  void mark_stale() { compute_stale = codes_stale = true; }
};

namespace ck {

// ---- pack_kernels.cu ----
cudaError_t launch_fill_missing(uint32_t *raw, size_t num_words, cudaStream_t s);
cudaError_t launch_pack(const ck_planes &pl, const int64_t *row, const int64_t *col, const int32_t *alt, size_t n,
                        size_t index_base, uint32_t *d_err, cudaStream_t s);
cudaError_t launch_finalize(const ck_planes &pl, cudaStream_t s);
cudaError_t launch_finalize_codes(const ck_planes &pl, int kind, cudaStream_t s);
// the same for the 64-sample blocks [block0, block0 + num_blocks) only (pipelined host-buffer path)
cudaError_t launch_finalize_codes_range(const ck_planes &pl, int kind, uint32_t block0, uint32_t num_blocks, cudaStream_t s);
// d_rows points at the reference-layout row of slot ref_slot0 (a chunk of the bit set, or the whole of it with 0)
cudaError_t launch_import_ref_range(const ck_planes &pl, const uint64_t *d_rows, uint32_t ref_slot0, uint32_t block0,
                                    uint32_t num_blocks, cudaStream_t s);
cudaError_t launch_import_ref(const ck_planes &pl, const uint64_t *d_bit_set, cudaStream_t s);
cudaError_t launch_export_ref(const ck_planes &pl, uint64_t *d_bit_set, cudaStream_t s);
cudaError_t launch_synth_planes(const ck_planes &pl, uint64_t seed, uint32_t miss_thr, cudaStream_t s);
cudaError_t launch_synth_count(uint64_t seed, uint32_t miss_thr, uint32_t sample_begin, uint32_t sample_end,
                               uint32_t site_begin, uint32_t site_end, unsigned long long *d_site_counts,
                               cudaStream_t s);
cudaError_t launch_synth_emit(uint64_t seed, uint32_t miss_thr, uint32_t sample_begin, uint32_t sample_end,
                              uint32_t site_begin, uint32_t site_end, const unsigned long long *d_site_offsets,
                              int64_t *row, int64_t *col, int32_t *alt, cudaStream_t s);

// ---- king_kernel.cu ----
struct KingLaunch {
  const uint32_t *compute;  // compute planes (LOP3+POPC kernels)
  const uint32_t *codes;    // nibble-coded genotypes (tcgen05 kernel)
  uint32_t words;           // padded words per plane
  uint32_t row_block0, num_row_blocks;
  uint32_t col_block0, num_col_blocks;
  uint32_t row_global0, col_global0;  // global sample index of slot row_block0*64 / col_block0*64
  uint32_t num_rows, num_cols;
  uint32_t triangular;                // rows and columns are the same sample range
  uint64_t tile_begin, tile_end;
  float kin_threshold;
  uint64_t max_results;
  ck_result *results;                 // device
  unsigned long long *counter;        // device, emitted pairs
  ck_counts *dump_counts;             // optional dense [num_rows][num_cols] dump (parity hook), else nullptr
  float *dump_kin;
};
struct KingStream {  // state of one ck_king_stream_begin .. end session (also used by the pipelined host-buffer path)
  KingLaunch k{};
  std::vector<uint64_t> band_prefix;  // first linear tile of every band + total
  int variant = 3;                    // 3 = mxf4 kernel, 2 = int8 kernel (same 128 x 80 band-ordered tiles)
  uint32_t part_index = 0, num_parts = 1;
  uint32_t max_results = 0;
  uint32_t next_end = 0;              // rows [next_end, n) have been delivered
};
uint64_t king_num_tiles(uint32_t num_row_blocks, uint32_t num_col_blocks, bool triangular);
cudaError_t launch_king(const KingLaunch &k, int variant, cudaStream_t s, uint32_t *launches);

// ---- band_tiles.cu: tile enumeration shared by the two tensor-core kernels ----
constexpr uint32_t kBandTileRows = 128;  // tile rows (TMEM lanes)
constexpr uint32_t kBandRowTiles = 8;    // row tiles per band
constexpr uint32_t kBandTileCols = 80;   // tile columns of both tensor-core kernels (what the streaming seam assumes)
struct BandTiles {                       // device view of the band table (ctx scratch)
  const unsigned long long *band_prefix;  // [num_bands + 1] tiles before band b
  const uint32_t *band_first_col;         // [num_bands] first column tile enumerated in band b
  uint32_t num_bands, num_row_tiles, num_col_tiles;
};
uint64_t band_num_tiles(const KingLaunch &k, uint32_t tile_cols);
// Uploads the band table for this launch geometry unless it is already on the device (then: no copy, no stream
// synchronisation - the streaming seam launches many tile ranges back to back).  `band_prefix`, when not NULL, receives
// the first linear tile index of every band (+ the total); `tiles`, when not NULL, the device view.
cudaError_t band_prepare(const KingLaunch &k, uint32_t tile_cols, ck_ctx *ctx, cudaStream_t s, std::vector<uint64_t> *band_prefix,
                         BandTiles *tiles);

// ---- king_umma_kernel.cu (variant 2: tcgen05 int8 tensor-core formulation, 128 x 96 tiles) ----
uint64_t king_umma_num_tiles(const KingLaunch &k);
cudaError_t launch_king_umma(const KingLaunch &k, uint32_t total_blocks, ck_ctx *ctx, cudaStream_t s, uint32_t *launches);

// ---- king_fp4_kernel.cu (variant 3: tcgen05 kind::mxf4 formulation, 128 x 80 tiles, band-ordered tile enumeration) ----
uint64_t king_fp4_num_tiles(const KingLaunch &k);
cudaError_t launch_king_fp4(const KingLaunch &k, uint32_t total_blocks, ck_ctx *ctx, cudaStream_t s, uint32_t *launches);
// band table of this launch geometry for the mxf4 kernel's tile shape (band_prepare)
cudaError_t king_fp4_prepare(const KingLaunch &k, ck_ctx *ctx, cudaStream_t s, std::vector<uint64_t> *band_prefix);
constexpr uint32_t kFp4BandRows = kBandRowTiles * kBandTileRows;
constexpr uint32_t kFp4MaxSites = 1u << 23;  // exactness of the tensor core's fp32 accumulation was measured up to this count

}  // namespace ck
