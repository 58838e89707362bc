// C ABI of libcuking_b200.so (include/cuking_b200.h), second half: the pairwise entry points - shard views, multi-GPU
// parts, result compaction / dense output, sort, copy-out, the streaming and host-buffer seams.
// Reference seam: /root/reference/cuking.cu:713-765 (result buffer, launch, overflow check, sort).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <new>
#include <string>
#include <vector>

#include "internal.cuh"
#include "king_common.cuh"

namespace ck {
namespace {

__global__ void make_sort_keys_kernel(const ck_result *res, size_t n, unsigned long long *keys, uint32_t *idx) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = (static_cast<unsigned long long>(res[i].sample_i) << 32) | res[i].sample_j;
    idx[i] = uint32_t(i);
  }
}
__global__ void gather_results_kernel(const ck_result *in, const uint32_t *idx, size_t n, ck_result *out) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

int ensure_result_buf(ck_ctx *ctx, size_t records) {
  if (ctx->result_cap >= records) return CK_OK;
  if (ctx->result_buf) cudaFree(ctx->result_buf);
  ctx->result_buf = nullptr;
  ctx->result_cap = 0;
  CK_CUDA(dev_alloc(ctx, reinterpret_cast<void **>(&ctx->result_buf), std::max<size_t>(records, 1) * sizeof(ck_result)));
  ctx->result_cap = records;
  return CK_OK;
}

int ensure_sort_scratch(ck_ctx *ctx, size_t bytes) {
  if (ctx->sort_scratch_bytes >= bytes) return CK_OK;
  if (ctx->sort_scratch) cudaFree(ctx->sort_scratch);
  ctx->sort_scratch = nullptr;
  ctx->sort_scratch_bytes = 0;
  const size_t want = bytes + bytes / 4;  // head-room so that slightly larger result sets do not reallocate
  CK_CUDA(dev_alloc(ctx, &ctx->sort_scratch, want));
  ctx->sort_scratch_bytes = want;
  return CK_OK;
}

int ensure_out_pinned(ck_ctx *ctx, size_t records) {
  if (ctx->out_pinned_records >= records) return CK_OK;
  for (void *&p : ctx->out_pinned) {
    if (p) cudaFreeHost(p);
    p = nullptr;
  }
  ctx->out_pinned_records = 0;
  for (void *&p : ctx->out_pinned) CK_CUDA(cudaHostAlloc(&p, records * sizeof(ck_result), cudaHostAllocDefault));
  ctx->out_pinned_records = records;
  return CK_OK;
}

// Sorts n device records by (sample_i, sample_j) — the pair is unique, so this equals the reference's
// (sample_i, sample_j, kin) order (cuking.cu:761-765).  `out` (device) receives the sorted records; when it is NULL
// they stay in the ctx scratch and *sorted points at them.
int sort_results(ck_ctx *ctx, const ck_result *in, size_t n, ck_result *out, const ck_result **sorted) {
  cudaStream_t s = ctx->stream;
  auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t tmp_bytes = 0;
  const long long items = static_cast<long long>(n);  // 64-bit item count: max_results may exceed INT_MAX
  {  // size query only (no work is launched); any valid device address serves as the pointer arguments
    auto *k64 = reinterpret_cast<unsigned long long *>(ctx->d_counter);
    auto *v32 = reinterpret_cast<uint32_t *>(ctx->d_counter);
    CK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k64, k64, v32, v32, items, 0, 64, s));
  }
  const size_t keys_b = align(n * 8), idx_b = align(n * 4), tmp_b = align(tmp_bytes), rec_b = out ? 0 : align(n * sizeof(ck_result));
  int rc = ensure_sort_scratch(ctx, 2 * keys_b + 2 * idx_b + tmp_b + rec_b);
  if (rc != CK_OK) return rc;
  char *base = static_cast<char *>(ctx->sort_scratch);
  auto *keys_a = reinterpret_cast<unsigned long long *>(base);
  auto *keys_o = reinterpret_cast<unsigned long long *>(base + keys_b);
  auto *idx_a = reinterpret_cast<uint32_t *>(base + 2 * keys_b);
  auto *idx_o = reinterpret_cast<uint32_t *>(base + 2 * keys_b + idx_b);
  void *tmp = base + 2 * keys_b + 2 * idx_b;
  ck_result *dst = out ? out : reinterpret_cast<ck_result *>(base + 2 * keys_b + 2 * idx_b + tmp_b);
  const unsigned grid = unsigned((n + 255) / 256);
  make_sort_keys_kernel<<<grid, 256, 0, s>>>(in, n, keys_a, idx_a);
  CK_CUDA(cudaGetLastError());
  CK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_o, idx_a, idx_o, items, 0, 64, s));
  gather_results_kernel<<<grid, 256, 0, s>>>(in, idx_o, n, dst);
  CK_CUDA(cudaGetLastError());
  if (sorted) *sorted = dst;
  return CK_OK;
}

// tile width the band table of a tensor-core variant is built for
uint32_t variant_tile_cols(int variant) { return variant == 4 ? kPairTileCols : kBandTileCols; }

uint64_t variant_num_tiles(int variant, const KingLaunch &k) {
  if (variant == 4) return king_fp4_pair_num_tiles(k);
  if (variant == 3 || variant == 5) return king_fp4_num_tiles(k);
  if (variant == 2) return king_umma_num_tiles(k);
  return king_num_tiles(k.num_row_blocks, k.num_col_blocks, k.triangular != 0);
}

cudaError_t dispatch_king(int variant, const ck_planes *pl, const KingLaunch &k, cudaStream_t s, uint32_t *launches) {
  // the screen needs nothing but sparse records out of the tile: dense output and the count dump take the mxf4 kernel
  if (variant == 5) {
    pl->ctx->screen_level_used = 0;
    if (k.dense_band_base || k.dump_counts || !k.sample_totals) return launch_king_fp4(k, pl->map.num_blocks, pl->ctx, s, launches);
    // Which screen.  Under the three-product bound an unrelated pair reaches a kinship of about r / 2 (r = the cohort's
    // missing rate), under the one-product bound (king_screen1_kernel.cu) about r + r^2 S / (2 het) (ck_planes::
    // screen1_floor, from the cohort's totals).  A screen is used when that leaves unrelated pairs clear of the threshold
    // (0.02 / 0.015 of kinship; without totals - planes filled piecewise - the three-product screen from 0.03 on); below,
    // nearly every tile would go on to the exact kernel anyway, which then runs alone.  Mispredicting only costs time.
    // CUKING_SCREEN_LEVEL = 1 / 3 (read when the ctx is created) forces a screen.
    KingLaunch ks = k;
    const float floor1 = pl->screen1_floor, floor3 = floor1 >= 0.f ? 0.5f * floor1 : 0.015f;
    int level = 0;  // 0 = no screen
    if (pl->ctx->screen_level != 0) level = pl->ctx->screen_level;
    else if (floor1 >= 0.f && k.kin_threshold - floor1 > 0.02f) level = 1;
    else if (k.kin_threshold - floor3 > 0.015f) level = 3;
    // king_screen1_kernel indexes the codes with 32-bit element offsets (registers): cohorts beyond 2^32 uint4 of codes (64 GiB)
    // take the three-product screen
    if (level == 1 && pl->codes_words() / 4 >= (size_t(1) << 32)) level = 3;
    if (level == 0) return launch_king_fp4(k, pl->map.num_blocks, pl->ctx, s, launches);
    ks.screen_level = level;
    return launch_king_screen(ks, pl->map.num_blocks, pl->ctx, s, launches);
  }
  if (variant == 4) return launch_king_fp4_pair(k, pl->map.num_blocks, pl->ctx, s, launches);
  if (variant == 3) return launch_king_fp4(k, pl->map.num_blocks, pl->ctx, s, launches);
  if (variant == 2) return launch_king_umma(k, pl->map.num_blocks, pl->ctx, s, launches);
  return launch_king(k, variant, s, launches);
}

// Launch geometry of the planes' own sub-matrix (view == NULL) or of a VIEW: any sub-matrix of the sample range the
// planes hold.  A cohort packed once can then serve every shard of a --split_factor run (cuking.cu:129-152 builds one
// bit set per shard process instead).  Views read samples at arbitrary plane slots: tensor-core kernels only.
int view_launch(const ck_planes *pl, const ck_submatrix *view, int variant, KingLaunch *out) {
  const ck_submatrix &sm = pl->map.sm;
  KingLaunch k{};
  k.compute = pl->compute;
  k.codes = pl->codes;
  k.sample_totals = pl->sample_totals();
  k.num_sites = pl->num_sites;
  k.words = pl->words;
  if (view == nullptr) {
    k.row_slot0 = 0;
    k.col_slot0 = pl->map.col_slot0;
    k.row_global0 = sm.i_begin;
    k.col_global0 = sm.j_begin;
    k.num_rows = sm_rows(sm);
    k.num_cols = sm_cols(sm);
    k.triangular = sm_diagonal(sm) ? 1u : 0u;
  } else {
    if (!sm_diagonal(sm)) return fail(CK_ERR_INVALID_ARGUMENT, "views need planes over one contiguous sample range (a diagonal sub-matrix)");
    if (view->i_end < view->i_begin || view->j_end < view->j_begin) return fail(CK_ERR_INVALID_ARGUMENT, "inverted sample range");
    if (view->i_begin < sm.i_begin || view->i_end > sm.i_end || view->j_begin < sm.i_begin || view->j_end > sm.i_end)
      return fail(CK_ERR_INVALID_ARGUMENT, "view outside the sample range of the planes");
    const bool diag = view->i_begin == view->j_begin && view->i_end == view->j_end;
    if (!diag && sm_rows(*view) > 0 && sm_cols(*view) > 0 && view->i_end > view->j_begin)
      return fail(CK_ERR_INVALID_ARGUMENT, "rows and columns of a view must be identical, or disjoint with the rows first");
    if (variant < 2) return fail(CK_ERR_INVALID_ARGUMENT, "views need a tensor-core kernel variant (2 or 3)");
    k.row_slot0 = view->i_begin - sm.i_begin;
    k.col_slot0 = view->j_begin - sm.i_begin;
    k.row_global0 = view->i_begin;
    k.col_global0 = view->j_begin;
    k.num_rows = sm_rows(*view);
    k.num_cols = sm_cols(*view);
    k.triangular = diag ? 1u : 0u;
  }
  k.row_block0 = k.row_slot0 / kTileSamples;
  k.col_block0 = k.col_slot0 / kTileSamples;
  k.num_row_blocks = ceil_div(k.num_rows, kTileSamples);
  k.num_col_blocks = ceil_div(k.num_cols, kTileSamples);
  *out = k;
  return CK_OK;
}

// Owner of band b when the pairs are split into num_parts parts: bands are dealt in snake order (0 .. P-1, P-1 .. 0, ...):
// the tile count of a band falls linearly with its index, so every pair (g, 2P-1-g) of a group carries the same work.
uint32_t band_owner(uint32_t band, uint32_t num_parts) {
  const uint32_t g = band % (2 * num_parts);
  return g < num_parts ? g : 2 * num_parts - 1 - g;
}

// Dense output is chosen when the caller's buffer has room for EVERY pair of the part and the threshold is negative
// (BASELINE configs[4]: --kin_threshold -1): then most pairs are kept, the append counter is a serialisation point and
// the sort touches tens of gigabytes - while the sorted position of a pair is a closed form.  CUKING_DENSE=0 / 1
// overrides the threshold rule (never / whenever it fits).
bool want_dense(float kin_threshold, unsigned long long part_pairs, uint32_t max_results) {
  static const char *env = getenv("CUKING_DENSE");
  if (part_pairs == 0 || part_pairs > max_results) return false;
  if (env != nullptr && *env != 0) return atoi(env) != 0;
  return kin_threshold < 0.f;
}

// Decides sparse / dense for one evaluation, sizes the emit buffer, uploads the band -> output-slot table.
int plan_output(ck_planes *pl, const KingLaunch &k, uint32_t part, uint32_t parts, float thr, uint32_t max_results,
                bool allow_dense, ResultPlan *plan) {
  ck_ctx *ctx = pl->ctx;
  cudaStream_t s = ctx->stream;
  const uint32_t num_bands = ceil_div(k.num_rows, kDenseBandRows);
  plan->regions.clear();
  plan->part_pairs = 0;
  std::vector<unsigned long long> &base = ctx->dense_host;
  base.assign(std::max<uint32_t>(num_bands, 1), ~0ull);
  // i < j pairs exist only where the column range reaches past the row: a triangular or rows-first rectangular view
  for (uint32_t b = 0; b < num_bands; ++b) {
    if (band_owner(b, parts) != part) continue;
    base[b] = plan->part_pairs;
    plan->part_pairs += dense_band_pairs(b * kDenseBandRows, k.num_rows, k.num_cols, k.triangular != 0);
  }
  plan->dense = allow_dense && want_dense(thr, plan->part_pairs, max_results);
  // max_results records either way (dense output needs part_pairs <= max_results of them): the buffer then keeps its
  // size from part to part and call to call
  int rc = ensure_result_buf(ctx, size_t(max_results));
  if (rc != CK_OK) return rc;
  events_reset(ctx);
  CK_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long) * (plan->dense ? 2 + ck_ctx::kHoleSlots : 1), s));
  if (plan->dense) {
    if (ctx->dense_table_entries < base.size()) {
      if (ctx->dense_table) cudaFree(ctx->dense_table);
      ctx->dense_table = nullptr;
      ctx->dense_table_entries = 0;
      CK_CUDA(dev_alloc(ctx, reinterpret_cast<void **>(&ctx->dense_table), base.size() * 8));
      ctx->dense_table_entries = base.size();
    }
    CK_CUDA(cudaMemcpyAsync(ctx->dense_table, base.data(), base.size() * 8, cudaMemcpyHostToDevice, s));
  }
  return CK_OK;
}

// Launches this part's bands among [band_lo, band_hi) in maximal runs (the whole range when num_parts == 1).  Dense
// output: a run is also an output region - contiguous because consecutive owned bands are consecutive in the output -
// capped in length so that the device -> host copies of finished regions overlap the kernels of the next ones.
int launch_bands(ck_planes *pl, KingLaunch k, int variant, const std::vector<uint64_t> &band_prefix, uint32_t band_lo,
                 uint32_t band_hi, uint32_t part, uint32_t parts, uint32_t max_run, bool organ_pipe, ResultPlan *plan) {
  ck_ctx *ctx = pl->ctx;
  cudaStream_t s = ctx->stream;
  k.results = ctx->result_buf;
  k.counter = ctx->d_counter;
  k.dense_band_base = plan->dense ? ctx->dense_table : nullptr;
  std::vector<std::pair<uint32_t, uint32_t>> runs;
  for (uint32_t b = band_lo; b < band_hi;) {
    if (band_owner(b, parts) != part) { ++b; continue; }
    uint32_t e = b + 1;
    while (e < band_hi && e - b < max_run && band_owner(e, parts) == part) ++e;
    runs.emplace_back(b, e);
    b = e;
  }
  // The later bands of a triangular shard hold fewer pairs.  Launch order does not change the result; it decides how
  // the copy-out of finished regions overlaps the kernels - a two-machine flow shop, compute then copy.  When the copy is
  // the slower machine (several GPUs sharing the host's ingest) the regions should grow (Johnson's rule: the copy engine
  // starts at once and never runs dry), when the kernels are slower they should shrink (only a small region's copy is
  // left exposed at the end).  Which regime holds depends on the platform and on how many GPUs copy at once, so the
  // order is an organ pipe - every other region in growing order, then the rest in shrinking order - which is within one
  // small region of the best order in both regimes.
  if (organ_pipe && runs.size() > 2) {
    std::vector<std::pair<uint32_t, uint32_t>> up, down;  // runs are in shrinking order (band index ascending)
    for (size_t q = runs.size(); q-- > 0;) ((runs.size() - 1 - q) % 2 == 0 ? up : down).push_back(runs[q]);
    runs = up;
    runs.insert(runs.end(), down.rbegin(), down.rend());
  }
  for (const auto &run : runs) {
    const uint32_t b = run.first, e = run.second;
    k.tile_begin = band_prefix[b];
    k.tile_end = band_prefix[e];
    OutRegion region{};
    if (plan->dense) {
      region.offset = ctx->dense_host[b];
      for (uint32_t q = b; q < e; ++q) region.count += dense_band_pairs(q * kDenseBandRows, k.num_rows, k.num_cols, k.triangular != 0);
      region.holes_slot = uint32_t(std::min<size_t>(plan->regions.size(), ck_ctx::kHoleSlots - 1));
      k.holes = ctx->d_counter + 2 + region.holes_slot;
    }
    if (k.tile_end > k.tile_begin) CK_CUDA(dispatch_king(variant, pl, k, s, &ctx->timings.king_launches));
    if (plan->dense && region.count > 0) {
      CK_CUDA(pool_event(ctx, &region.ready));
      CK_CUDA(cudaEventRecord(region.ready, s));
      plan->regions.push_back(region);
    }
  }
  return CK_OK;
}

struct Dest {  // where the records go: a caller buffer (host or device) or a sink fed chunk by chunk
  ck_result *results = nullptr;
  int on_device = 0;
  ck_result_sink sink = nullptr;
  void *user = nullptr;
  size_t chunk_records = 0;
};

struct Piece {  // a run of consecutive output records in device memory, for chunked delivery
  const ck_result *src;
  size_t count;
  cudaEvent_t ready;   // nullptr: already complete on the compute stream
  int holes_slot;      // -1: no holes possible
};

// removes the below-threshold holes (sample_i == 0xffffffff) of a dense output range in place; returns the new length
size_t squeeze_holes(ck_result *r, size_t n) {
  size_t w = 0;
  for (size_t i = 0; i < n; ++i)
    if (r[i].sample_i != 0xffffffffu) {
      if (w != i) r[w] = r[i];
      ++w;
    }
  return w;
}

// Feeds the pieces, in order, through the ctx's page-locked double buffer to the sink: the copy of chunk c + 1 runs
// while the sink consumes chunk c, and host memory stays bounded by two chunks whatever the result size.
int deliver_pieces(ck_ctx *ctx, const std::vector<Piece> &pieces, const Dest &dst, uint64_t *delivered) {
  size_t chunk = std::max<size_t>(dst.chunk_records ? dst.chunk_records : (size_t(4) << 20), 1024);
  size_t largest = 0;
  for (const Piece &p : pieces) largest = std::max(largest, p.count);
  // no larger than what there is to deliver: page-locking 2 x 96 MB for the few thousand records of a sparse run costs more
  // than the run's kernel (bin/cuking, 10,000 samples: 0.1 s of a 0.11-s phase)
  chunk = std::min(chunk, std::max<size_t>(largest, 1024));
  int rc = ensure_out_pinned(ctx, chunk);
  if (rc != CK_OK) return rc;
  struct Job { const ck_result *src; size_t count; cudaEvent_t ready; int holes_slot; };
  std::vector<Job> jobs;
  for (const Piece &p : pieces)
    for (size_t off = 0; off < p.count; off += chunk) jobs.push_back({p.src + off, std::min(chunk, p.count - off), p.ready, p.holes_slot});
  cudaStream_t ds = ctx->d2h_stream;
  cudaEvent_t done[2];
  CK_CUDA(pool_event(ctx, &done[0]));
  CK_CUDA(pool_event(ctx, &done[1]));
  auto issue = [&](size_t j) -> int {
    if (jobs[j].ready) CK_CUDA(cudaStreamWaitEvent(ds, jobs[j].ready, 0));
    CK_CUDA(cudaMemcpyAsync(ctx->out_pinned[j & 1], jobs[j].src, jobs[j].count * sizeof(ck_result), cudaMemcpyDeviceToHost, ds));
    CK_CUDA(cudaEventRecord(done[j & 1], ds));
    return CK_OK;
  };
  *delivered = 0;
  if (!jobs.empty() && (rc = issue(0)) != CK_OK) return rc;
  for (size_t j = 0; j < jobs.size(); ++j) {
    if (j + 1 < jobs.size() && (rc = issue(j + 1)) != CK_OK) return rc;
    CK_CUDA(cudaEventSynchronize(done[j & 1]));
    ck_result *buf = static_cast<ck_result *>(ctx->out_pinned[j & 1]);
    size_t n = jobs[j].count;
    if (jobs[j].holes_slot >= 0) {
      unsigned long long holes = 0;  // the region is complete (its event preceded the copy): its hole count is final
      CK_CUDA(cudaMemcpyAsync(&holes, ctx->d_counter + 2 + jobs[j].holes_slot, sizeof(holes), cudaMemcpyDeviceToHost, ds));
      CK_CUDA(cudaStreamSynchronize(ds));
      if (holes) n = squeeze_holes(buf, n);
    }
    if (n && dst.sink(dst.user, buf, n) != 0) {
      cudaStreamSynchronize(ds);
      return fail(CK_ERR_INVALID_ARGUMENT, "the result sink aborted the delivery");
    }
    *delivered += n;
  }
  return CK_OK;
}

// Common tail of the sparse path: reads the emitted-pair counter, turns an overflow into the reference's error
// (cuking.cu:747-751), sorts on the device and copies the records out (or feeds the sink).  Expects ev[0] / ev[1]
// recorded around the kernel launches on the ctx stream.
int finish_sparse(ck_ctx *ctx, ck_result *d_emit, uint32_t max_results, const Dest &dst, uint64_t *num_results, int sort) {
  cudaStream_t s = ctx->stream;
  static const bool dbg = getenv("CUKING_DEBUG_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  auto tp0 = now();
  unsigned long long count = 0;
  CK_CUDA(cudaMemcpyAsync(&count, ctx->d_counter, sizeof(count), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  ctx->timings.king_ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
  ctx->timings.sort_ms = ctx->timings.d2h_ms = 0.f;
  if (dbg) fprintf(stderr, "[ck] king: host wait %.2f ms, kernel %.2f ms\n", ms_since(tp0), ctx->timings.king_ms);
  *num_results = count;
  if (count > max_results)  // cuking.cu:747-751
    return fail(CK_ERR_RESULT_OVERFLOW, "Could not store all results: try increasing the --max_results parameter.");
  const size_t n = size_t(count);
  if (n == 0) return CK_OK;

  struct Events {  // destroyed on every exit path
    cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
    ~Events() {
      for (cudaEvent_t x : e)
        if (x) cudaEventDestroy(x);
    }
  } ev;
  for (cudaEvent_t &x : ev.e) CK_CUDA(cudaEventCreate(&x));
  cudaEvent_t t0 = ev.e[0], t1 = ev.e[1], t2 = ev.e[2];
  cudaEventRecord(t0, s);
  const ck_result *d_final = d_emit;
  if (sort) {
    int rc = sort_results(ctx, d_emit, n, (!dst.sink && dst.on_device) ? dst.results : nullptr, &d_final);
    if (rc != CK_OK) return rc;
    ctx->timings.king_launches += 3;  // key build, radix sort (one logical launch), gather
  } else if (!dst.sink && dst.on_device && d_emit != dst.results)  // unsorted, device destination, emitted into our buffer
    CK_CUDA(cudaMemcpyAsync(dst.results, d_emit, n * sizeof(ck_result), cudaMemcpyDeviceToDevice, s));
  cudaEventRecord(t1, s);
  if (dst.sink) {
    CK_CUDA(cudaStreamSynchronize(s));
    uint64_t delivered = 0;
    int rc = deliver_pieces(ctx, {Piece{d_final, n, nullptr, -1}}, dst, &delivered);
    if (rc != CK_OK) return rc;
  } else if (!dst.on_device) {
    CK_CUDA(cudaMemcpyAsync(dst.results, d_final, n * sizeof(ck_result), cudaMemcpyDeviceToHost, s));
  }
  cudaEventRecord(t2, s);
  cudaError_t e = cudaStreamSynchronize(s);
  ctx->timings.sort_ms = elapsed_ms(t0, t1);
  ctx->timings.d2h_ms = elapsed_ms(t1, t2);
  if (dbg) fprintf(stderr, "[ck] sort+d2h: device sort %.2f, d2h %.2f\n", ctx->timings.sort_ms, ctx->timings.d2h_ms);
  CK_CUDA(e);
  return CK_OK;
}

// Tail of the dense path: every region's records are already in sorted position in the emit buffer; copy them out as
// their kernels finish.  Expects ev[0] recorded before the first launch.
int finish_dense(ck_ctx *ctx, ResultPlan &plan, const Dest &dst, uint64_t *num_results) {
  cudaStream_t s = ctx->stream, ds = ctx->d2h_stream;
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  const ck_result *d_out = ctx->result_buf;
  if (dst.sink) {
    std::sort(plan.regions.begin(), plan.regions.end(), [](const OutRegion &a, const OutRegion &b) { return a.offset < b.offset; });
    std::vector<Piece> pieces;
    for (const OutRegion &r : plan.regions) pieces.push_back({d_out + r.offset, size_t(r.count), r.ready, int(r.holes_slot)});
    uint64_t delivered = 0;
    int rc = deliver_pieces(ctx, pieces, dst, &delivered);
    cudaError_t e = cudaStreamSynchronize(s);
    ctx->timings.king_ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
    ctx->timings.sort_ms = ctx->timings.d2h_ms = 0.f;
    if (rc != CK_OK) return rc;
    CK_CUDA(e);
    *num_results = delivered;
    return CK_OK;
  }
  static const bool dbg = getenv("CUKING_DEBUG_TIMING") != nullptr;
  const auto tp0 = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  for (const OutRegion &r : plan.regions) {  // caller's host buffer: optimistic copies to the hole-free positions
    CK_CUDA(cudaStreamWaitEvent(ds, r.ready, 0));
    CK_CUDA(cudaMemcpyAsync(dst.results + r.offset, d_out + r.offset, size_t(r.count) * sizeof(ck_result), cudaMemcpyDeviceToHost, ds));
  }
  if (dbg) fprintf(stderr, "[ck] dense: %zu regions, copies queued after %.2f ms\n", plan.regions.size(), ms_since(tp0));
  cudaEvent_t copied;
  CK_CUDA(cudaEventCreate(&copied));
  cudaEventRecord(copied, ds);
  const size_t slots = std::min<size_t>(plan.regions.size(), ck_ctx::kHoleSlots);
  if (slots) CK_CUDA(cudaMemcpyAsync(ctx->h_holes, ctx->d_counter + 2, slots * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  cudaError_t e = cudaStreamSynchronize(s);
  if (dbg) fprintf(stderr, "[ck] dense: kernels done after %.2f ms\n", ms_since(tp0));
  if (e == cudaSuccess) e = cudaStreamSynchronize(ds);
  if (dbg) fprintf(stderr, "[ck] dense: copies done after %.2f ms\n", ms_since(tp0));
  ctx->timings.king_ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
  ctx->timings.sort_ms = 0.f;
  ctx->timings.d2h_ms = std::max(0.f, elapsed_ms(ctx->ev[1], copied));  // what was not hidden behind the kernels
  cudaEventDestroy(copied);
  CK_CUDA(e);
  unsigned long long holes = 0;
  for (size_t q = 0; q < slots; ++q) holes += ctx->h_holes[q];
  size_t n = size_t(plan.part_pairs);
  if (holes) n = squeeze_holes(dst.results, n);  // rare (a negative threshold keeps nearly every pair): one host pass
  if (n != plan.part_pairs - holes) return fail(CK_ERR_CUDA, "dense output: hole count does not match the records found");
  *num_results = n;
  return CK_OK;
}

bool host_bitset_can_pipeline(const ck_planes *pl) {
  static const bool off = getenv("CUKING_NO_PIPELINE") != nullptr;
  return !off && planes_variant(pl) >= 2 && sm_rows(pl->map.sm) >= 4 * kFp4BandRows;
}

// One evaluation: the planes' sub-matrix or a view of it, one part of num_parts, into a buffer or a sink.
int eval_view(ck_planes *pl, const ck_submatrix *view, uint32_t part, uint32_t parts, float thr, uint32_t max_results,
              const Dest &dst, uint64_t *num_results, int sort) {
  *num_results = 0;
  if (parts == 0 || part >= parts) return fail(CK_ERR_INVALID_ARGUMENT, "part_index outside [0, num_parts)");
  if (pl->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "a stream session is open on these planes");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  int rc = ensure_compute(pl);  // may allocate the buffer view_launch() points at
  if (rc != CK_OK) return rc;
  const int variant = planes_variant(pl);
  KingLaunch k{};
  rc = view_launch(pl, view, variant, &k);
  if (rc != CK_OK) return rc;
  k.kin_threshold = thr;
  k.max_results = max_results;
  ctx->timings.king_launches = 0;
  if (variant < 2) {  // LOP3+POPC kernels: a contiguous slice of the 64 x 64 tile grid, appended and sorted
    const bool direct = !dst.sink && dst.on_device && !sort;
    ck_result *d_emit = direct ? dst.results : nullptr;
    if (!direct) {
      rc = ensure_result_buf(ctx, max_results);
      if (rc != CK_OK) return rc;
      d_emit = ctx->result_buf;
    }
    const uint64_t tiles = variant_num_tiles(variant, k);
    k.tile_begin = tiles * part / parts;
    k.tile_end = tiles * (part + 1) / parts;
    k.results = d_emit;
    k.counter = ctx->d_counter;
    events_reset(ctx);
    CK_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), s));
    CK_CUDA(cudaEventRecord(ctx->ev[0], s));
    if (k.tile_end > k.tile_begin) CK_CUDA(dispatch_king(variant, pl, k, s, &ctx->timings.king_launches));
    CK_CUDA(cudaEventRecord(ctx->ev[1], s));
    return finish_sparse(ctx, d_emit, max_results, dst, num_results, sort);
  }
  // tensor-core kernels: bands of 1024 rows dealt to the parts in snake order
  static const bool dbg = getenv("CUKING_DEBUG_TIMING") != nullptr;
  const auto tp0 = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  std::vector<uint64_t> band_prefix;
  CK_CUDA(band_prepare(k, variant_tile_cols(variant), ctx, s, &band_prefix, nullptr));
  const uint32_t num_bands = uint32_t(band_prefix.size()) - 1;
  ResultPlan plan;
  rc = plan_output(pl, k, part, parts, thr, max_results, /*allow_dense=*/sort && (dst.sink || !dst.on_device), &plan);
  if (rc != CK_OK) return rc;
  uint32_t owned = 0;
  for (uint32_t b = 0; b < num_bands; ++b) owned += band_owner(b, parts) == part;
  const uint32_t max_run = plan.dense ? std::max<uint32_t>(1, ceil_div(owned, 24u)) : 0xffffffffu;
  CK_CUDA(cudaEventRecord(ctx->ev[0], s));
  if (dbg) fprintf(stderr, "[ck] eval: planned after %.2f ms (dense %d, %llu pairs)\n", ms_since(tp0), int(plan.dense), plan.part_pairs);
  // dense output into a caller buffer: launch order chosen for the overlap of the copy-out (see launch_bands); a sink
  // wants the regions in output order
  const bool organ_pipe = plan.dense && !dst.sink;
  rc = launch_bands(pl, k, variant, band_prefix, 0, num_bands, part, parts, max_run, organ_pipe, &plan);
  if (rc != CK_OK) {
    cudaStreamSynchronize(s);
    return rc;
  }
  if (dbg) fprintf(stderr, "[ck] eval: launched after %.2f ms\n", ms_since(tp0));
  if (plan.dense) {
    rc = finish_dense(ctx, plan, dst, num_results);
    if (dbg) fprintf(stderr, "[ck] eval: finished after %.2f ms (kernel %.2f ms)\n", ms_since(tp0), ctx->timings.king_ms);
    return rc;
  }
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  return finish_sparse(ctx, ctx->result_buf, max_results, dst, num_results, sort);
}

int stream_begin_impl(ck_planes *pl, float kin_threshold, uint32_t max_results, uint32_t part_index, uint32_t num_parts,
                      bool allow_rect = false) {
  ck_ctx *ctx = pl->ctx;
  if (num_parts == 0 || part_index >= num_parts) return fail(CK_ERR_INVALID_ARGUMENT, "part_index outside [0, num_parts)");
  if (!sm_diagonal(pl->map.sm) && !allow_rect) return fail(CK_ERR_INVALID_ARGUMENT, "streaming delivery needs a diagonal shard");
  const int variant = planes_variant(pl);
  if (variant < 2) return fail(CK_ERR_INVALID_ARGUMENT, "streaming delivery needs a tensor-core kernel variant (2 or 3): their band-ordered tiles");
  if (pl->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "a stream session is already open on these planes");
  if (pl->codes == nullptr) {
    pl->codes_bytes = std::max<size_t>(pl->codes_alloc_words(), 1) * 4;
    CK_CUDA(ctx_alloc(ctx, reinterpret_cast<void **>(&pl->codes), pl->codes_bytes));
  }
  KingStream *st = new (std::nothrow) KingStream();
  if (!st) return fail(CK_ERR_OUT_OF_MEMORY, "host allocation failed");
  int rc = view_launch(pl, nullptr, variant, &st->k);
  if (rc == CK_OK) {
    st->k.kin_threshold = kin_threshold;
    st->k.max_results = max_results;
    st->variant = variant;
    st->part_index = part_index;
    st->num_parts = num_parts;
    st->max_results = max_results;
    st->next_end = sm_rows(pl->map.sm);
    rc = plan_output(pl, st->k, part_index, num_parts, kin_threshold, max_results, /*allow_dense=*/true, &st->plan);
  }
  if (rc == CK_OK) {
    cudaError_t e = band_prepare(st->k, variant_tile_cols(variant), ctx, ctx->stream, &st->band_prefix, nullptr);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[0], ctx->stream);
    if (e != cudaSuccess) rc = fail_cuda(e, "ck_king_stream_begin", __FILE__, __LINE__);
  }
  if (rc != CK_OK) {
    delete st;
    return rc;
  }
  ctx->timings.king_launches = 0;
  pl->mark_stale();
  pl->stream_state = st;
  return CK_OK;
}

// Rows [s0, s1) of the shard, in device memory behind d_rows: transpose, derive the codes, launch this part's bands among
// them.  Everything is queued on the ctx stream; nothing is synchronised.
int stream_rows_device(ck_planes *pl, const uint64_t *d_rows, uint32_t s0, uint32_t s1) {
  KingStream *st = pl->stream_state;
  ck_ctx *ctx = pl->ctx;
  cudaStream_t s = ctx->stream;
  const uint32_t n = sm_rows(pl->map.sm);
  if (s1 != st->next_end || s0 >= s1 || s0 % kFp4BandRows != 0)
    return fail(CK_ERR_INVALID_ARGUMENT, "stream rows must arrive in descending ranges that tile the shard at multiples of "
                                         "ck_king_stream_granularity()");
  const uint32_t block0 = s0 / kTileSamples, num_blocks = ceil_div(s1, kTileSamples) - block0;
  CK_CUDA(launch_import_ref_range(*pl, d_rows, s0, block0, num_blocks, s));
  // variant 5: the first piece delivered (the last rows) stands for the cohort when the screen level is chosen - one small
  // synchronisation per session, behind which the uploads of the next pieces are already running
  const bool sample_stats = st->variant == 5 && s1 == n && ctx->screen_level == 0;
  if (sample_stats) CK_CUDA(cudaMemsetAsync(pl->totals_sums(), 0, 2 * sizeof(unsigned long long), s));
  CK_CUDA(launch_finalize_codes_range(*pl, st->variant >= 3 ? 3 : st->variant, block0, num_blocks, s, sample_stats));
  if (s1 == n) pl->screen1_floor = -1.f;
  if (sample_stats) {
    unsigned long long sums[2] = {0, 0};
    CK_CUDA(cudaMemcpyAsync(sums, pl->totals_sums(), sizeof(sums), cudaMemcpyDeviceToHost, s));
    CK_CUDA(cudaStreamSynchronize(s));
    pl->screen1_floor = screen1_floor_from_sums(sums, double(s1 - s0), double(pl->num_sites));
  }
  ctx->timings.king_launches += 2;
  const uint32_t band_lo = s0 / kFp4BandRows, band_hi = ceil_div(std::min(s1, n), kFp4BandRows);
  st->k.codes = pl->codes;
  st->k.sample_totals = pl->sample_totals();
  int rc = launch_bands(pl, st->k, st->variant, st->band_prefix, band_lo, band_hi, st->part_index, st->num_parts, 0xffffffffu, false, &st->plan);
  if (rc != CK_OK) return rc;
  st->next_end = s0;
  return CK_OK;
}

int stream_end_impl(ck_planes *pl, ck_result *results, uint64_t *num_results) {
  KingStream *st = pl->stream_state;
  ck_ctx *ctx = pl->ctx;
  const bool complete = st->next_end == 0;
  const uint32_t max_results = st->max_results;
  const int variant = st->variant;
  ResultPlan plan = std::move(st->plan);
  struct Staged {  // device copies of host-delivered rows: back to the ctx cache once the stream has drained (every exit
    ck_ctx *ctx;   // path below synchronises it)
    std::vector<std::pair<void *, size_t>> bufs;
    ~Staged() {
      if (!bufs.empty()) cudaStreamSynchronize(ctx->stream);
      for (auto &b : bufs) ctx_release(ctx, b.first, b.second);
    }
  } staged{ctx, std::move(st->staged)};
  delete st;
  pl->stream_state = nullptr;
  if (!complete) {
    cudaStreamSynchronize(ctx->stream);
    return fail(CK_ERR_INVALID_ARGUMENT, "ck_king_stream_end before every row of the shard was delivered");
  }
  pl->compute_stale = true;
  pl->codes_stale = false;
  pl->codes_kind = variant >= 3 ? 3 : variant;
  Dest dst;
  dst.results = results;
  if (plan.dense) return finish_dense(ctx, plan, dst, num_results);
  CK_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  return finish_sparse(ctx, ctx->result_buf, max_results, dst, num_results, 1);
}

// ck_king_host_bitset with a tensor-core kernel: the upload of the reference-layout bit set overlaps the pairwise kernel
// instead of preceding it.  Diagonal shard: a band of the tile enumeration (kFp4BandRows rows) only needs the samples at
// or after its first row (i < j), so the sample range is uploaded LAST CHUNK FIRST on the copy stream and every chunk's
// bands are launched as soon as its rows have been transposed and coded - by then every column they pair with is
// already on the device; the bottom chunks hold few tiles, so only the first small upload is exposed.  Off-diagonal
// shard: every band needs all the column samples, so those go first (one piece, exposed) and the row samples follow in
// chunks behind the kernels of the chunks before them - half of the upload is hidden.
int king_host_bitset_pipelined(ck_planes *pl, const uint64_t *bit_set, float kin_threshold, uint32_t max_results,
                               ck_result *results, uint64_t *num_results, uint32_t part_index, uint32_t num_parts) {
  *num_results = 0;
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
  const ck_submatrix &sm = pl->map.sm;
  const bool rect = !sm_diagonal(sm);
  const uint32_t n = sm_rows(sm);
  const size_t words_per_sample = ref_words_per_sample(pl->num_sites);  // u64
  const size_t bytes = words_per_sample * sm_samples(sm) * 8;
  struct Staging {  // device copy of the host bit set, returned to the ctx cache on scope exit
    ck_ctx *ctx;
    void *p = nullptr;
    size_t bytes = 0;
    std::vector<cudaEvent_t> events;
    ~Staging() {
      for (cudaEvent_t e : events) cudaEventDestroy(e);
      ctx_release(ctx, p, bytes);
    }
  } st{ctx};
  CK_CUDA(ctx_alloc(ctx, &st.p, bytes));
  st.bytes = bytes;
  int rc = stream_begin_impl(pl, kin_threshold, max_results, part_index, num_parts, /*allow_rect=*/true);
  if (rc != CK_OK) return rc;
  auto abandon = [&](int code) {
    cudaStreamSynchronize(cs);
    stream_discard(pl);
    return code;
  };
  const uint32_t num_bands = ceil_div(n, kFp4BandRows);
  const uint32_t chunk_bands = std::max<uint32_t>(1, ceil_div(num_bands, 24u));

  // the copy stream starts after everything already queued on the compute stream (the buffers come from the ctx cache)
  cudaEvent_t fork;
  CK_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  st.events.push_back(fork);
  CK_CUDA(cudaEventRecord(fork, s));
  CK_CUDA(cudaStreamWaitEvent(cs, fork, 0));
  uint64_t *d_bits = static_cast<uint64_t *>(st.p);
  if (rect) {  // the column samples: reference slots [rows, rows + cols), plane blocks from col_slot0 on
    const size_t off = size_t(n) * words_per_sample;
    cudaEvent_t cols_ready;
    CK_CUDA(cudaEventCreateWithFlags(&cols_ready, cudaEventDisableTiming));
    st.events.push_back(cols_ready);
    cudaError_t e = cudaMemcpyAsync(d_bits + off, bit_set + off, size_t(sm_cols(sm)) * words_per_sample * 8, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) e = cudaEventRecord(cols_ready, cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, cols_ready, 0);
    const uint32_t col_block0 = pl->map.col_slot0 / kTileSamples, col_blocks = ceil_div(sm_cols(sm), kTileSamples);
    if (e == cudaSuccess) e = launch_import_ref_range(*pl, d_bits + off, n, col_block0, col_blocks, s);
    if (e == cudaSuccess) e = launch_finalize_codes_range(*pl, pl->stream_state->variant >= 3 ? 3 : pl->stream_state->variant, col_block0, col_blocks, s);
    if (e != cudaSuccess) return abandon(fail_cuda(e, "column upload", __FILE__, __LINE__));
    ctx->timings.king_launches += 2;
  }
  struct Chunk { uint32_t s0, s1; cudaEvent_t ready; };
  std::vector<Chunk> chunks;
  for (uint32_t hi = num_bands; hi > 0;) {
    const uint32_t lo = hi > chunk_bands ? hi - chunk_bands : 0;
    Chunk c{lo * kFp4BandRows, std::min<uint32_t>(hi * kFp4BandRows, n), nullptr};
    CK_CUDA(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
    st.events.push_back(c.ready);
    const size_t off = size_t(c.s0) * words_per_sample;
    CK_CUDA(cudaMemcpyAsync(d_bits + off, bit_set + off, size_t(c.s1 - c.s0) * words_per_sample * 8, cudaMemcpyHostToDevice, cs));
    CK_CUDA(cudaEventRecord(c.ready, cs));
    chunks.push_back(c);
    hi = lo;
  }
  for (const Chunk &c : chunks) {
    cudaError_t e = cudaStreamWaitEvent(s, c.ready, 0);
    rc = e == cudaSuccess ? stream_rows_device(pl, d_bits + size_t(c.s0) * words_per_sample, c.s0, c.s1)
                          : fail_cuda(e, "cudaStreamWaitEvent", __FILE__, __LINE__);
    if (rc != CK_OK) return abandon(rc);
  }
  rc = stream_end_impl(pl, results, num_results);
  ctx->timings.h2d_ms = 0.f;     // overlapped: the whole upload + transpose + kernel span is reported as king_ms
  ctx->timings.import_ms = 0.f;
  return rc;
}

uint32_t clamp_count(uint64_t n) { return n > 0xffffffffull ? 0xffffffffu : uint32_t(n); }

}  // namespace

void stream_discard(ck_planes *pl) {
  KingStream *st = pl->stream_state;
  if (!st) return;
  cudaStreamSynchronize(pl->ctx->stream);
  for (auto &b : st->staged) ctx_release(pl->ctx, b.first, b.second);
  delete st;
  pl->stream_state = nullptr;
}
}  // namespace ck

using namespace ck;

extern "C" {

int ck_planes_king_variant(const ck_planes *pl, int *variant) {
  if (!pl || !variant) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  *variant = planes_variant(pl);
  return CK_OK;
}

int ck_ctx_fp4_selftest(ck_ctx *ctx, int *exact) {
  if (!ctx || !exact) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard guard(ctx->device);
  std::string detail;
  const int rc = fp4_selftest(ctx, exact, &detail);
  if (rc != CK_OK) return rc;
  ctx->fp4_state = *exact ? 1 : -1;
  set_error("kind::mxf4 accumulation self-test: " + detail + (*exact ? "" : " - variant 3 is routed to the int8 kernel (variant 2) on this ctx"));
  return CK_OK;
}

int ck_king_num_tiles(const ck_planes *pl, uint64_t *num_tiles) {
  if (!pl || !num_tiles) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  const int variant = planes_variant(pl);
  KingLaunch k{};
  int rc = view_launch(pl, nullptr, variant, &k);
  if (rc != CK_OK) return rc;
  *num_tiles = variant_num_tiles(variant, k);
  return CK_OK;
}

int ck_king_tiles(ck_planes *pl, uint64_t tile_begin, uint64_t tile_end, float kin_threshold, uint32_t max_results,
                  ck_result *results, int results_on_device, uint32_t *num_results, int sort) {
  if (!pl || !num_results) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (max_results > 0 && !results) return fail(CK_ERR_INVALID_ARGUMENT, "results is NULL");
  *num_results = 0;
  if (pl->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "a stream session is open on these planes");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  int rc = ensure_compute(pl);  // may allocate the buffer view_launch() points at
  if (rc != CK_OK) return rc;
  const int variant = planes_variant(pl);
  KingLaunch k{};
  rc = view_launch(pl, nullptr, variant, &k);
  if (rc != CK_OK) return rc;
  const uint64_t total = variant_num_tiles(variant, k);
  if (tile_begin > tile_end || tile_end > total) return fail(CK_ERR_INVALID_ARGUMENT, "tile range outside the tile grid");

  // Pairs are appended to a device buffer: the caller's when it is device memory and no sort is needed, else ours.
  ck_result *d_emit = nullptr;
  const bool direct = results_on_device && !sort;
  if (direct) {
    d_emit = results;
  } else {
    rc = ensure_result_buf(ctx, max_results);
    if (rc != CK_OK) return rc;
    d_emit = ctx->result_buf;
  }
  k.tile_begin = tile_begin;
  k.tile_end = tile_end;
  k.kin_threshold = kin_threshold;
  k.max_results = max_results;
  k.results = d_emit;
  k.counter = ctx->d_counter;
  ctx->timings.king_launches = 0;
  events_reset(ctx);
  CK_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), s));
  CK_CUDA(cudaEventRecord(ctx->ev[0], s));
  if (tile_end > tile_begin) CK_CUDA(dispatch_king(variant, pl, k, s, &ctx->timings.king_launches));
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  Dest dst;
  dst.results = results;
  dst.on_device = results_on_device;
  uint64_t n = 0;
  rc = finish_sparse(ctx, d_emit, max_results, dst, &n, sort);
  *num_results = clamp_count(n);
  return rc;
}

int ck_king_view(ck_planes *pl, const ck_submatrix *view, uint32_t part_index, uint32_t num_parts, float kin_threshold,
                 uint32_t max_results, ck_result *results, int results_on_device, uint32_t *num_results, int sort) {
  if (!pl || !num_results) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (max_results > 0 && !results) return fail(CK_ERR_INVALID_ARGUMENT, "results is NULL");
  Dest dst;
  dst.results = results;
  dst.on_device = results_on_device;
  uint64_t n = 0;
  const int rc = eval_view(pl, view, part_index, num_parts, kin_threshold, max_results, dst, &n, sort);
  *num_results = clamp_count(n);
  return rc;
}

int ck_king_view_sink(ck_planes *pl, const ck_submatrix *view, uint32_t part_index, uint32_t num_parts, float kin_threshold,
                      uint32_t max_results, size_t chunk_records, ck_result_sink sink, void *user, uint64_t *num_results) {
  if (!pl || !num_results || !sink) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  Dest dst;
  dst.sink = sink;
  dst.user = user;
  dst.chunk_records = chunk_records;
  return eval_view(pl, view, part_index, num_parts, kin_threshold, max_results, dst, num_results, 1);
}

int ck_king(ck_planes *pl, float kin_threshold, uint32_t max_results, ck_result *results, int results_on_device,
            uint32_t *num_results, int sort) {
  return ck_king_view(pl, nullptr, 0, 1, kin_threshold, max_results, results, results_on_device, num_results, sort);
}

int ck_king_counts(ck_planes *pl, const uint32_t *sample_i, const uint32_t *sample_j, size_t num_pairs,
                   ck_counts *counts, float *kin) {
  if (!pl || !sample_i || !sample_j || !counts || !kin) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  const ck_submatrix &sm = pl->map.sm;
  const size_t rows = sm_rows(sm), cols = sm_cols(sm);
  if (rows * cols > (size_t(1) << 28)) return fail(CK_ERR_INVALID_ARGUMENT, "ck_king_counts is a parity hook for shards of at most 2^28 pairs");
  for (size_t q = 0; q < num_pairs; ++q) {
    const bool ok = sample_i[q] >= sm.i_begin && sample_i[q] < sm.i_end && sample_j[q] >= sm.j_begin &&
                    sample_j[q] < sm.j_end && sample_i[q] < sample_j[q];
    if (!ok) return fail(CK_ERR_INVALID_ARGUMENT, "pair " + std::to_string(q) + " is not an i < j pair of this shard");
  }
  int rc = ensure_compute(pl);
  if (rc != CK_OK) return rc;
  const int variant = planes_variant(pl);
  DevBuf d_counts(ctx), d_kin(ctx);
  CK_CUDA(d_counts.alloc(rows * cols * sizeof(ck_counts)));
  CK_CUDA(d_kin.alloc(rows * cols * sizeof(float)));
  CK_CUDA(cudaMemsetAsync(d_counts.p, 0, rows * cols * sizeof(ck_counts), s));
  CK_CUDA(cudaMemsetAsync(d_kin.p, 0, rows * cols * sizeof(float), s));
  KingLaunch k{};
  rc = view_launch(pl, nullptr, variant, &k);
  if (rc != CK_OK) return rc;
  k.tile_begin = 0;
  k.tile_end = variant_num_tiles(variant, k);
  k.kin_threshold = 2.f;  // nothing is emitted: kin <= 0.5
  k.max_results = 0;
  k.results = nullptr;
  k.counter = ctx->d_counter;
  k.dump_counts = d_counts.as<ck_counts>();
  k.dump_kin = d_kin.as<float>();
  CK_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), s));
  CK_CUDA(dispatch_king(variant, pl, k, s, nullptr));
  std::vector<ck_counts> h_counts(rows * cols);
  std::vector<float> h_kin(rows * cols);
  CK_CUDA(cudaMemcpyAsync(h_counts.data(), d_counts.p, rows * cols * sizeof(ck_counts), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaMemcpyAsync(h_kin.data(), d_kin.p, rows * cols * sizeof(float), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  for (size_t q = 0; q < num_pairs; ++q) {
    const size_t idx = size_t(sample_i[q] - sm.i_begin) * cols + (sample_j[q] - sm.j_begin);
    counts[q] = h_counts[idx];
    kin[q] = h_kin[idx];
  }
  return CK_OK;
}

uint32_t ck_king_stream_granularity(void) { return kFp4BandRows; }

int ck_king_stream_begin(ck_planes *pl, float kin_threshold, uint32_t max_results, uint32_t part_index, uint32_t num_parts) {
  if (!pl) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
  DeviceGuard guard(pl->ctx->device);
  return stream_begin_impl(pl, kin_threshold, max_results, part_index, num_parts);
}

int ck_king_stream_rows(ck_planes *pl, const uint64_t *rows, int on_device, uint32_t sample_begin, uint32_t sample_end) {
  if (!pl || !rows) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!pl->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "no stream session is open on these planes");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  if (on_device) return stream_rows_device(pl, rows, sample_begin, sample_end);
  if (sample_end <= sample_begin) return fail(CK_ERR_INVALID_ARGUMENT, "empty row range");
  // Host rows: staged through a device buffer from the ctx cache (no cudaMalloc / cudaFree per call once warm) that the
  // session keeps until ck_king_stream_end; the copy is queued on the ctx stream ahead of the kernels that read it and
  // nothing is synchronised.  Pageable memory makes cudaMemcpyAsync synchronous for the host; page-locked rows
  // (ck_host_alloc) return at once - keep them alive until stream_end.
  const size_t bytes = size_t(sample_end - sample_begin) * ref_words_per_sample(pl->num_sites) * 8;
  void *tmp = nullptr;
  CK_CUDA(ctx_alloc(ctx, &tmp, bytes));
  pl->stream_state->staged.push_back({tmp, bytes});  // returned to the ctx cache by ck_king_stream_end
  CK_CUDA(cudaMemcpyAsync(tmp, rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return stream_rows_device(pl, static_cast<const uint64_t *>(tmp), sample_begin, sample_end);
}

int ck_king_stream_end(ck_planes *pl, ck_result *results, uint32_t *num_results) {
  if (!pl || !num_results) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!pl->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "no stream session is open on these planes");
  if (pl->stream_state->max_results > 0 && !results) return fail(CK_ERR_INVALID_ARGUMENT, "results is NULL");
  *num_results = 0;
  DeviceGuard guard(pl->ctx->device);
  uint64_t n = 0;
  const int rc = stream_end_impl(pl, results, &n);
  *num_results = clamp_count(n);
  return rc;
}

int ck_king_host_bitset(ck_ctx *ctx, uint32_t num_samples, uint32_t split_factor, uint32_t shard_index,
                        uint32_t num_sites, const uint64_t *bit_set, float kin_threshold, uint32_t max_results,
                        ck_result *results, uint32_t *num_results) {
  return ck_king_host_bitset_part(ctx, num_samples, split_factor, shard_index, num_sites, bit_set, kin_threshold,
                                  max_results, results, num_results, 0, 1);
}

int ck_king_host_bitset_part(ck_ctx *ctx, uint32_t num_samples, uint32_t split_factor, uint32_t shard_index,
                             uint32_t num_sites, const uint64_t *bit_set, float kin_threshold, uint32_t max_results,
                             ck_result *results, uint32_t *num_results, uint32_t part_index, uint32_t num_parts) {
  if (!ctx || !bit_set || !num_results) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (max_results > 0 && !results) return fail(CK_ERR_INVALID_ARGUMENT, "results is NULL");
  *num_results = 0;
  if (num_parts == 0 || part_index >= num_parts) return fail(CK_ERR_INVALID_ARGUMENT, "part_index outside [0, num_parts)");
  ck_submatrix sm;
  int rc = ck_submatrix_init(num_samples, split_factor, shard_index, &sm);
  if (rc != CK_OK) return rc;
  ck_planes *pl = nullptr;
  rc = ck_planes_create(ctx, &sm, num_sites, &pl);
  if (rc != CK_OK) return rc;
  if (host_bitset_can_pipeline(pl)) {
    uint64_t n = 0;
    rc = king_host_bitset_pipelined(pl, bit_set, kin_threshold, max_results, results, &n, part_index, num_parts);
    *num_results = clamp_count(n);
  } else {  // small or off-diagonal shards, LOP3+POPC variants: plain upload, then the part
    rc = ck_planes_import_bitset(pl, bit_set, 0);
    if (rc == CK_OK) rc = ck_king_view(pl, nullptr, part_index, num_parts, kin_threshold, max_results, results, 0, num_results, 1);
  }
  ck_planes_destroy(pl);
  return rc;
}

}  // extern "C"
