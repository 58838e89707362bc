// Internal declarations shared by the translation units of libcuking_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <utility>
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

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (kernel, device) pair, and one
// process may drive several GPUs from several threads (bin/cuking --num_gpus / --all_shards): `done` holds one bit per
// device.  Two threads racing on the same device both set the attribute, which is harmless.
template <typename F>
cudaError_t optin_dynamic_smem(F func, size_t bytes, std::atomic<uint64_t> &done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (unsigned(dev) & 63u);
  if (dev < 64 && (done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  if (e == cudaSuccess && dev < 64) done.fetch_or(bit, std::memory_order_release);
  return e;
}

}  // namespace ck

struct ck_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;       // own_stream or a caller-supplied one
  cudaStream_t copy_stream = nullptr;  // host -> device staging copies overlap the kernels
  cudaStream_t d2h_stream = nullptr;   // device -> host result copies overlap the kernels (PCIe is full duplex)
  cudaEvent_t ev[2] = {nullptr, nullptr};
  ck_timings timings{};
  int num_sms = 0;
  // scratch owned by the ctx
  static constexpr uint32_t kHoleSlots = 1024;  // dense output: one below-threshold counter per output region
  unsigned long long *d_counter = nullptr;      // [0] emitted-pair counter, [1] spare, [2 ..] hole counters
  unsigned long long *h_holes = nullptr;        // pinned host mirror of the hole counters
  unsigned long long *d_pack_err = nullptr;     // [0] invalid-genotype index+1 (min), [1] out-of-range index+1 (min),
                                                // [2] dictionary index out of range (ck_pack_encoded), [3] spare
  // ck_pack_encoded may be called by several host threads at once (the decode threads of bin/cuking): every call takes a
  // free lane - its own stream, device staging buffer and error slots - so the calls overlap on the GPU instead of queueing
  // behind one host lock; lanes are created on demand and live as long as the ctx
  struct IngestLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t dep = nullptr, ev[2] = {nullptr, nullptr};
    void *staging = nullptr;  // grow-only device copy of one window's payloads, run tables and dictionaries
    size_t staging_bytes = 0;
    unsigned long long *d_err = nullptr;  // 4 slots like d_pack_err
    unsigned long long *h_err = nullptr;  // page-locked mirror
  };
  std::mutex lane_mu;
  std::vector<IngestLane *> lanes_free, lanes_all;
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
  int screen_level = 0;   // variant 5: 0 = chosen per evaluation from the cohort's call rate, 1 / 3 = forced (CUKING_SCREEN_LEVEL)
  // kind::mxf4 accumulation self-test (fp4_selftest.cu): 0 = not run yet, 1 = exact, -1 = inexact -> int8 kernel
  int fp4_state = 0;
  // alive-tile table of the tcgen05 kernel (grow-only device scratch)
  void *tile_table = nullptr;
  size_t tile_table_bytes = 0;
  uint8_t *tile_flags = nullptr;  // screen kernel: one byte per tile of a launch (grow-only)
  size_t tile_flags_bytes = 0;
  unsigned long long *d_screen_flagged = nullptr;  // tiles the screens have flagged on this ctx so far (ck_ctx_screen_stats)
  unsigned long long screen_tiles = 0;             // tiles they have been launched over
  int screen_level_used = 0;                       // screen of the last variant-5 launch: 1, 3, or 0 = none (mxf4 kernel alone)
  uint64_t tile_table_key[3] = {~0ull, 0, 0};  // (variant, rows/cols, global origins) of the table now on the device
  // dense output: first output slot of every band (device copy + the host vector the async upload reads)
  unsigned long long *dense_table = nullptr;
  size_t dense_table_entries = 0;
  std::vector<unsigned long long> dense_host;
  // grow-only scratch for the result sort (keys, indices, CUB temporaries, sorted records): no cudaMalloc/cudaFree
  // on the steady-state path
  void *sort_scratch = nullptr;
  size_t sort_scratch_bytes = 0;
  // page-locked double buffer for chunked result delivery (ck_king_view_sink)
  void *out_pinned[2] = {nullptr, nullptr};
  size_t out_pinned_records = 0;
  std::vector<cudaEvent_t> event_pool;  // timing-disabled events reused by the pipelined paths
  size_t events_used = 0;
  // small cache of released device buffers (planes, staging) so that a create / destroy cycle per call - the
  // host-buffer entry point ck_king_host_bitset - does not pay cudaMalloc / cudaFree of gigabytes every time
  static constexpr int kCacheSlots = 8;
  void *cache_ptr[kCacheSlots] = {};
  size_t cache_bytes[kCacheSlots] = {};
};

namespace ck {
cudaError_t ctx_alloc(ck_ctx *ctx, void **ptr, size_t bytes);  // exact-size reuse from the cache, else dev_alloc
void ctx_release(ck_ctx *ctx, void *ptr, size_t bytes);        // into the cache (evicting the smallest entry) or cudaFree
// cudaMalloc for everything the ctx owns: on cudaErrorMemoryAllocation the buffer cache is dropped and the allocation
// retried once, so idle cached gigabytes never turn into CK_ERR_OUT_OF_MEMORY
cudaError_t dev_alloc(ck_ctx *ctx, void **ptr, size_t bytes);
// a timing-disabled event from the ctx pool (valid until events_reset)
cudaError_t pool_event(ck_ctx *ctx, cudaEvent_t *out);
inline void events_reset(ck_ctx *ctx) { ctx->events_used = 0; }

struct DeviceGuard {  // every entry point runs on the ctx's device and restores the caller's
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    (void)cudaGetLastError();  // a stale non-sticky error of an earlier (successful) call must not be blamed on ours
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct DevBuf {  // RAII device allocation for per-call temporaries
  ck_ctx *ctx;
  void *p = nullptr;
  explicit DevBuf(ck_ctx *c) : ctx(c) {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return dev_alloc(ctx, &p, bytes ? bytes : 1); }
  template <typename T>
  T *as() const { return static_cast<T *>(p); }
};

inline float elapsed_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}
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
  // the codes buffer is followed by (het, hom) site counts per plane slot (uint2; the screen kernels bound kinship with
  // them) and by the two sums over all slots (2 x uint64; the host picks the screen level from the cohort's call rate)
  size_t codes_alloc_words() const { return codes_words() + size_t(map.num_blocks) * ck::kTileSamples * 2 + 4; }
  uint2 *sample_totals() const { return codes ? reinterpret_cast<uint2 *>(codes + codes_words()) : nullptr; }
  unsigned long long *totals_sums() const {
    return codes ? reinterpret_cast<unsigned long long *>(codes + codes_words() + size_t(map.num_blocks) * ck::kTileSamples * 2) : nullptr;
  }
  // kinship an unrelated pair's one-product bound (king_screen1_kernel.cu) is expected to reach, from the cohort's call rate
  // and het rate; < 0: not known (planes filled piecewise by the streaming seam) -> the three-product screen
  float screen1_floor = -1.f;
  void mark_stale() { compute_stale = codes_stale = true; }
};

namespace ck {

// ---- pack_kernels.cu ----
cudaError_t launch_fill_missing(uint32_t *raw, size_t num_words, cudaStream_t s);
cudaError_t launch_pack(const ck_planes &pl, const int64_t *row, const int64_t *col, const int32_t *alt, size_t n,
                        size_t index_base, unsigned long long *d_err, cudaStream_t s);
cudaError_t launch_pack_narrow(const ck_planes &pl, const uint32_t *row, const uint32_t *col, const uint8_t *alt, size_t n,
                               size_t index_base, unsigned long long *d_err, cudaStream_t s);
// device view of one ck_encoded_column (page_decode.cu): payload words (8-byte aligned, padded by 8 bytes), runs + sentinel
struct EncodedColumnDev {
  const uint32_t *words;
  const ck_run *runs;
  const void *dict;
  uint32_t num_runs, dict_len, width, skip;
};
cudaError_t launch_decode_pack(const ck_planes &pl, const EncodedColumnDev (&cols)[3], uint32_t num_rows, unsigned long long *d_err,
                               cudaStream_t s);
cudaError_t launch_finalize(const ck_planes &pl, cudaStream_t s);
cudaError_t launch_finalize_codes(const ck_planes &pl, int kind, cudaStream_t s);
// the same for the 64-sample blocks [block0, block0 + num_blocks) only (pipelined host-buffer path)
cudaError_t launch_finalize_codes_range(const ck_planes &pl, int kind, uint32_t block0, uint32_t num_blocks, cudaStream_t s,
                                        bool add_to_sums = false);
// call rate and het rate of `samples` samples (their (het, hom) sums) -> ck_planes::screen1_floor
float screen1_floor_from_sums(const unsigned long long sums[2], double samples, double sites);
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
  uint32_t row_slot0, col_slot0;      // plane slot of the first row / column sample (tensor-core kernels: any slot, so
                                      // that a shard can be a VIEW into the planes of a larger sample set)
  uint32_t row_block0, num_row_blocks;  // the same in 64-sample blocks (LOP3+POPC kernels: whole blocks only)
  uint32_t col_block0, num_col_blocks;
  uint32_t row_global0, col_global0;  // global sample index of the first row / column
  uint32_t num_rows, num_cols;
  uint32_t triangular;                // rows and columns are the same sample range
  uint64_t tile_begin, tile_end;
  float kin_threshold;
  uint64_t max_results;
  ck_result *results;                 // device
  unsigned long long *counter;        // device, emitted pairs
  ck_counts *dump_counts;             // optional dense [num_rows][num_cols] dump (parity hook), else nullptr
  float *dump_kin;
  // Dense output (tensor-core kernels): when not NULL, the record of pair (i, j) goes to the slot it has in the sorted
  // output - dense_band_base[band of i] + its closed-form offset inside the band - so neither the append counter nor a
  // sort is needed; a pair at or below the threshold leaves a hole (sample_i = 0xffffffff) and bumps *holes.
  const unsigned long long *dense_band_base;
  unsigned long long *holes;
  // Screen kernels (king_screen_kernel.cu, king_screen1_kernel.cu; variant 5): site counts of every plane slot, and one byte per tile
  // of the launch - written by the screen kernel (1 = the tile holds a pair that may pass the threshold), read by the mxf4
  // kernel launched behind it over the same tile range, whose CTAs leave at once where the byte is 0.
  const uint2 *sample_totals;  // (het, hom) site counts per plane slot
  uint8_t *tile_flags;
  unsigned long long *flagged_counter;  // += 1 per flagged tile (statistics)
  uint32_t num_sites;          // real (unpadded) sites
  int screen_level;            // 1: one-product screen first (king_screen1_kernel.cu), else the three-product screen
};
// Where the records of one evaluation go (king_api.cu).  Sparse: appended through the atomic counter, sorted afterwards.
// Dense: every pair has its slot in the sorted output; `regions` lists the output ranges in launch order with the event
// that marks them complete, so that their device -> host copies overlap the kernels still running.
struct OutRegion {
  unsigned long long offset, count;  // records
  cudaEvent_t ready;
  uint32_t holes_slot;               // index into ck_ctx::d_counter + 2
};
struct ResultPlan {
  bool dense = false;
  unsigned long long part_pairs = 0;  // i < j pairs of this part (= the dense buffer's size)
  std::vector<OutRegion> regions;
};
struct KingStream {  // state of one ck_king_stream_begin .. end session (also used by the pipelined host-buffer path)
  KingLaunch k{};
  std::vector<uint64_t> band_prefix;  // first linear tile of every band + total
  int variant = 3;                    // 3 = mxf4 kernel, 2 = int8 kernel (same 128 x 80 band-ordered tiles)
  uint32_t part_index = 0, num_parts = 1;
  uint32_t max_results = 0;
  uint32_t next_end = 0;              // rows [next_end, n) have been delivered
  ResultPlan plan;
  std::vector<std::pair<void *, size_t>> staged;  // device copies of host-delivered rows (ctx cache buffers)
};
void stream_discard(ck_planes *pl);  // king_api.cu: abandons an open stream session
// capi.cu
int planes_variant(const ck_planes *pl);   // the variant that runs on these planes (runs the mxf4 self-test on first use)
int ensure_compute(ck_planes *pl);   // derives what that variant reads from the raw planes
// fp4_selftest.cu: adversarial accumulation patterns through tcgen05.mma kind::mxf4 on this GPU; *exact = 1 when every
// checked accumulator equals the integer arithmetic
int fp4_selftest(ck_ctx *ctx, int *exact, std::string *detail);
uint64_t king_num_tiles(uint32_t num_row_blocks, uint32_t num_col_blocks, bool triangular);
cudaError_t launch_king(const KingLaunch &k, int variant, cudaStream_t s, uint32_t *launches);

// ---- band_tiles.cu: tile enumeration shared by the two tensor-core kernels ----
constexpr uint32_t kBandTileRows = 128;  // tile rows (TMEM lanes)
constexpr uint32_t kBandRowTiles = 8;    // row tiles per band
__host__ __device__ inline uint32_t band_rows_padded(uint32_t remaining_row_tiles) {  // row tiles enumerated in a band
  const uint32_t r = remaining_row_tiles < kBandRowTiles ? remaining_row_tiles : kBandRowTiles;
  return (r + 1u) & ~1u;
}
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
// king_fp4_pair_kernel.cu (variant 4: the same on CTA pairs, tcgen05 cta_group::2, 256 x 80 tiles sharing the B operand)
cudaError_t launch_king_fp4_pair(const KingLaunch &k, uint32_t total_blocks, ck_ctx *ctx, cudaStream_t s, uint32_t *launches);
uint64_t king_fp4_pair_num_tiles(const KingLaunch &k);
constexpr uint32_t kPairTileCols = 64;  // 256 x 64 pair tiles
// band table of this launch geometry for the mxf4 kernel's tile shape (band_prepare)
cudaError_t king_fp4_prepare(const KingLaunch &k, ck_ctx *ctx, cudaStream_t s, std::vector<uint64_t> *band_prefix);
// king_screen_kernel.cu (variant 5): three-product screen of every tile + the mxf4 kernel on the tiles it flags
cudaError_t launch_king_screen(const KingLaunch &k, uint32_t total_blocks, ck_ctx *ctx, cudaStream_t s, uint32_t *launches);
cudaError_t launch_king_screen1(const KingLaunch &part, const BandTiles &tiles, cudaStream_t s);  // king_screen1_kernel.cu
constexpr uint32_t kFp4BandRows = kBandRowTiles * kBandTileRows;
constexpr uint32_t kFp4MaxSites = 1u << 23;  // exactness of the tensor core's fp32 accumulation was measured up to this count

}  // namespace ck
