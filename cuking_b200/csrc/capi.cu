// C ABI of libcuking_b200.so (include/cuking_b200.h), first half: library, contexts, plane storage, pack / import /
// export, synthetic inputs.  The pairwise entry points live in king_api.cu.  Reference seam:
// /root/reference/cuking.cu:505-523 (planning + allocation), :675-703 (pack).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_scan.cuh>
#include <new>
#include <string>
#include <vector>

#include "internal.cuh"
#include "synth.cuh"

namespace ck {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", int(e), cudaGetErrorString(e), what, file, line);
  g_last_error = buf;
  return e == cudaErrorMemoryAllocation ? CK_ERR_OUT_OF_MEMORY : CK_ERR_CUDA;
}

static bool dbg_alloc() {
  static const bool on = getenv("CUKING_DEBUG_ALLOC") != nullptr;
  return on;
}

cudaError_t dev_alloc(ck_ctx *ctx, void **ptr, size_t bytes) {
  cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 1);
  if (e == cudaErrorMemoryAllocation && ctx != nullptr) {  // make room: drop everything cached and retry once
    cudaGetLastError();
    bool dropped = false;
    for (int i = 0; i < ck_ctx::kCacheSlots; ++i)
      if (ctx->cache_ptr[i]) {
        cudaFree(ctx->cache_ptr[i]);
        ctx->cache_ptr[i] = nullptr;
        ctx->cache_bytes[i] = 0;
        dropped = true;
      }
    if (dropped) e = cudaMalloc(ptr, bytes ? bytes : 1);
  }
  return e;
}

cudaError_t pool_event(ck_ctx *ctx, cudaEvent_t *out) {
  if (ctx->events_used == ctx->event_pool.size()) {
    cudaEvent_t e = nullptr;
    cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (err != cudaSuccess) return err;
    ctx->event_pool.push_back(e);
  }
  *out = ctx->event_pool[ctx->events_used++];
  return cudaSuccess;
}

cudaError_t ctx_alloc(ck_ctx *ctx, void **ptr, size_t bytes) {
  bytes = bytes ? bytes : 1;
  for (int i = 0; i < ck_ctx::kCacheSlots; ++i)
    if (ctx->cache_ptr[i] && ctx->cache_bytes[i] == bytes) {
      *ptr = ctx->cache_ptr[i];
      ctx->cache_ptr[i] = nullptr;
      ctx->cache_bytes[i] = 0;
      if (dbg_alloc()) fprintf(stderr, "[ck] ctx_alloc %zu -> cached %p\n", bytes, *ptr);
      return cudaSuccess;
    }
  cudaError_t e = dev_alloc(ctx, ptr, bytes);
  if (dbg_alloc()) fprintf(stderr, "[ck] ctx_alloc %zu -> new %p\n", bytes, *ptr);
  return e;
}

void ctx_release(ck_ctx *ctx, void *ptr, size_t bytes) {
  if (!ptr) return;
  static const bool no_cache = getenv("CUKING_NO_CACHE") != nullptr;
  if (no_cache) {
    cudaFree(ptr);
    return;
  }
  bytes = bytes ? bytes : 1;
  int slot = -1;
  for (int i = 0; i < ck_ctx::kCacheSlots; ++i)
    if (!ctx->cache_ptr[i]) { slot = i; break; }
  if (slot < 0) {  // full: evict the smallest entry if it is smaller than the newcomer
    int smallest = 0;
    for (int i = 1; i < ck_ctx::kCacheSlots; ++i)
      if (ctx->cache_bytes[i] < ctx->cache_bytes[smallest]) smallest = i;
    if (ctx->cache_bytes[smallest] >= bytes) {
      cudaFree(ptr);
      return;
    }
    cudaFree(ctx->cache_ptr[smallest]);
    slot = smallest;
  }
  ctx->cache_ptr[slot] = ptr;
  ctx->cache_bytes[slot] = bytes;
  if (dbg_alloc()) fprintf(stderr, "[ck] ctx_release %p %zu -> slot %d\n", ptr, bytes, slot);
}

namespace {

int king_variant_from_env() {
  const char *v = getenv("CUKING_KING_VARIANT");
  if (v == nullptr || *v == 0) return -1;
  return atoi(v);
}

}  // namespace
}  // namespace ck

using namespace ck;

// default pairwise kernel variant: 0 = 5 POPC per pair-word, 1 = carry-save (2.5 POPC + 5 more LOP3),
// 2 = tcgen05 int8 tensor-core formulation, 3 = tcgen05 mxf4 (E2M1, fp32 accumulation) formulation,
// 5 = the mxf4 kernel behind the one- / three-product screens (king_screen1_kernel.cu, king_screen_kernel.cu)
static int g_default_variant = 5;
static int active_variant(const ck_ctx *ctx) { return ctx->king_variant >= 0 ? ctx->king_variant : g_default_variant; }

// The mxf4 kernel relies on the tensor core adding E2M1 products into its fp32 accumulator without losing low bits - a
// property the PTX ISA does not spell out.  So the first use of variant 3 on a ctx runs fp4_selftest (adversarial
// accumulation patterns, a few hundred microseconds) on that GPU; if any accumulator differs from integer arithmetic
// the ctx takes the int8 kernel (s32 accumulators, exact by specification) from then on and says so on stderr.
static bool fp4_usable(ck_ctx *ctx) {
  if (ctx->fp4_state == 0) {
    int exact = 0;
    std::string detail;
    DeviceGuard guard(ctx->device);
    const int rc = fp4_selftest(ctx, &exact, &detail);
    ctx->fp4_state = (rc == CK_OK && exact) ? 1 : -1;
    if (ctx->fp4_state < 0)
      fprintf(stderr, "[cuking_b200] kind::mxf4 accumulation self-test failed on device %d (%s): using the int8 tensor-core kernel\n",
              ctx->device, detail.c_str());
  }
  return ctx->fp4_state > 0;
}

namespace ck {
float screen1_floor_from_sums(const unsigned long long sums[2], double samples, double sites) {
  if (samples <= 0 || sites <= 0 || sums[0] == 0) return -1.f;
  const double r = std::max(0.0, 1.0 - double(sums[0] + sums[1]) / (samples * sites)), het = double(sums[0]) / samples;
  return float(r + r * r * sites / (2.0 * het));
}
// The variant that actually runs on these planes: the fp32 accumulators of the mxf4 kernel are exact only up to the
// site count the probe verified (kFp4MaxSites) and only if this GPU passed the self-test; otherwise the int8 kernel.
int planes_variant(const ck_planes *pl) {
  const int v = active_variant(pl->ctx);
  if (v < 3) return v;  // 3 = mxf4 kernel, 4 = its CTA-pair form, 5 = screen + mxf4: all rest on the fp32 accumulation of kind::mxf4
  return (pl->num_sites > kFp4MaxSites || !fp4_usable(pl->ctx)) ? 2 : v;
}
}  // namespace ck

extern "C" {

int ck_abi_version(void) { return CK_ABI_VERSION; }
const char *ck_last_error(void) { return g_last_error.c_str(); }

int ck_device_count(int *count) {
  if (!count) return fail(CK_ERR_INVALID_ARGUMENT, "count is NULL");
  CK_CUDA(cudaGetDeviceCount(count));
  return CK_OK;
}

/* ---- shard planning ---- */

int ck_submatrix_init(uint32_t num_samples, uint32_t split_factor, uint32_t shard_index, ck_submatrix *out) {
  if (!out) return fail(CK_ERR_INVALID_ARGUMENT, "out is NULL");
  if (split_factor == 0) return fail(CK_ERR_INVALID_ARGUMENT, "Invalid split factor");
  if (!make_submatrix(num_samples, split_factor, shard_index, out))
    return fail(CK_ERR_INVALID_ARGUMENT, "Invalid shard index");
  return CK_OK;
}
uint32_t ck_num_shards(uint32_t k) { return uint32_t(uint64_t(k) * (uint64_t(k) + 1) / 2); }
uint32_t ck_submatrix_num_rows(const ck_submatrix *sm) { return sm_rows(*sm); }
uint32_t ck_submatrix_num_cols(const ck_submatrix *sm) { return sm_cols(*sm); }
uint32_t ck_submatrix_num_samples(const ck_submatrix *sm) { return sm_samples(*sm); }
uint32_t ck_submatrix_contains(const ck_submatrix *sm, uint32_t s) { return sm_contains(*sm, s) ? 1u : 0u; }
uint32_t ck_submatrix_sample_offset(const ck_submatrix *sm, uint32_t s) { return sm_ref_offset(*sm, s); }
uint32_t ck_words_per_sample(uint32_t num_sites) { return ref_words_per_sample(num_sites); }

int ck_plan_work(uint32_t num_samples, uint32_t split_factor, uint32_t first_shard, uint32_t num_run, uint32_t num_gpus,
                 ck_work_item *items, uint32_t max_items, uint32_t *num_items) {
  if (!num_items || (max_items > 0 && !items)) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (split_factor == 0) return fail(CK_ERR_INVALID_ARGUMENT, "Invalid split factor");
  if (num_gpus == 0) return fail(CK_ERR_INVALID_ARGUMENT, "num_gpus is 0");
  const uint64_t total_shards = uint64_t(split_factor) * (uint64_t(split_factor) + 1) / 2;
  if (uint64_t(first_shard) + num_run > total_shards) return fail(CK_ERR_INVALID_ARGUMENT, "Invalid shard index");
  struct Shard { uint32_t index; uint64_t pairs; uint32_t parts; };
  std::vector<Shard> shards;
  long double total = 0;
  for (uint32_t q = 0; q < num_run; ++q) {
    ck_submatrix sm{};
    make_submatrix(num_samples, split_factor, first_shard + q, &sm);
    const uint64_t r = sm_rows(sm), c = sm_cols(sm);
    const uint64_t pairs = sm_diagonal(sm) ? r * (r ? r - 1 : 0) / 2 : r * c;
    shards.push_back({first_shard + q, pairs, 1});
    total += pairs;
  }
  const long double share = total / num_gpus;
  std::vector<ck_work_item> all;
  for (Shard &sh : shards) {
    // no item above the per-GPU share; a part is made of whole 1024-row bands, so more parts than bands are useless
    ck_submatrix sm{};
    make_submatrix(num_samples, split_factor, sh.index, &sm);
    const uint32_t bands = std::max<uint32_t>(1, ceil_div(sm_rows(sm), 1024u));
    const long double cap = share * 1.02L;  // 2 % slack: diagonal shards are a shade under half an off-diagonal one
    uint64_t parts = cap > 0 ? uint64_t(sh.pairs / cap) : 1;
    if (parts * cap < sh.pairs) ++parts;    // ceil(pairs / cap)
    sh.parts = uint32_t(std::min<uint64_t>(std::max<uint64_t>(parts, 1), std::min<uint32_t>(bands, 4 * num_gpus)));
    for (uint32_t p = 0; p < sh.parts; ++p) all.push_back({sh.index, p, sh.parts, 0, sh.pairs});
  }
  // LPT: longest item first (ties: lower shard, lower part) onto the least loaded GPU (ties: lower GPU)
  std::stable_sort(all.begin(), all.end(), [](const ck_work_item &a, const ck_work_item &b) {
    return a.pairs * b.num_parts > b.pairs * a.num_parts;  // pairs / parts, compared without division
  });
  std::vector<long double> load(num_gpus, 0);
  for (ck_work_item &it : all) {
    uint32_t best = 0;
    for (uint32_t g = 1; g < num_gpus; ++g)
      if (load[g] < load[best]) best = g;
    it.gpu = best;
    load[best] += (long double)it.pairs / it.num_parts;
  }
  std::stable_sort(all.begin(), all.end(), [](const ck_work_item &a, const ck_work_item &b) { return a.gpu < b.gpu; });
  *num_items = uint32_t(all.size());
  for (uint32_t q = 0; q < all.size() && q < max_items; ++q) items[q] = all[q];
  return CK_OK;
}

/* ---- context ---- */

int ck_ctx_create(int device, ck_ctx **out) {
  if (!out) return fail(CK_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  int count = 0;
  CK_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(CK_ERR_INVALID_ARGUMENT, "no such CUDA device");
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  CK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(CK_ERR_CUDA, std::string("libcuking_b200 is built for sm_100a only; device is ") + prop.name);
  ck_ctx *ctx = new (std::nothrow) ck_ctx();
  if (!ctx) return fail(CK_ERR_OUT_OF_MEMORY, "host allocation failed");
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  cudaError_t e;
  constexpr size_t kCounterBytes = (2 + ck_ctx::kHoleSlots) * sizeof(unsigned long long);
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev[0])) != cudaSuccess || (e = cudaEventCreate(&ctx->ev[1])) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_counter, kCounterBytes)) != cudaSuccess ||
      (e = cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_holes), kCounterBytes, cudaHostAllocDefault)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_pack_err, 4 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_screen_flagged, sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMemset(ctx->d_screen_flagged, 0, sizeof(unsigned long long))) != cudaSuccess) {
    ck_ctx_destroy(ctx);
    return fail_cuda(e, "ck_ctx_create", __FILE__, __LINE__);
  }
  ctx->stream = ctx->own_stream;
  ctx->king_variant = king_variant_from_env();
  if (const char *v = getenv("CUKING_SCREEN_LEVEL")) ctx->screen_level = (atoi(v) == 1) ? 1 : (atoi(v) == 3) ? 3 : 0;
  *out = ctx;
  return CK_OK;
}

int ck_ctx_set_stream(ck_ctx *ctx, void *cuda_stream) {
  if (!ctx) return fail(CK_ERR_INVALID_ARGUMENT, "ctx is NULL");
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return CK_OK;
}

int ck_ctx_set_king_variant(ck_ctx *ctx, int variant) {
  if (!ctx) return fail(CK_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (variant < -1 || variant > 5) return fail(CK_ERR_INVALID_ARGUMENT, "unknown pairwise kernel variant");
  ctx->king_variant = variant;
  return CK_OK;
}

int ck_ctx_synchronize(ck_ctx *ctx) {
  if (!ctx) return fail(CK_ERR_INVALID_ARGUMENT, "ctx is NULL");
  DeviceGuard guard(ctx->device);
  CK_CUDA(cudaStreamSynchronize(ctx->stream));
  return CK_OK;
}

int ck_ctx_screen_stats(ck_ctx *ctx, uint64_t *tiles_screened, uint64_t *tiles_flagged, int *level) {
  if (!ctx) return fail(CK_ERR_INVALID_ARGUMENT, "ctx is NULL");
  DeviceGuard guard(ctx->device);
  unsigned long long flagged = 0;
  CK_CUDA(cudaStreamSynchronize(ctx->stream));
  CK_CUDA(cudaMemcpy(&flagged, ctx->d_screen_flagged, sizeof(flagged), cudaMemcpyDeviceToHost));
  if (tiles_screened) *tiles_screened = ctx->screen_tiles;
  if (tiles_flagged) *tiles_flagged = flagged;
  if (level) *level = ctx->screen_level_used;
  return CK_OK;
}

int ck_ctx_get_timings(ck_ctx *ctx, ck_timings *out) {
  if (!ctx || !out) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = ctx->timings;
  return CK_OK;
}

int ck_ctx_destroy(ck_ctx *ctx) {
  if (!ctx) return CK_OK;
  DeviceGuard guard(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  for (int i = 0; i < 2; ++i) {
    if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
    if (ctx->staging[i]) cudaFree(ctx->staging[i]);
    if (ctx->pinned_free[i]) cudaEventDestroy(ctx->pinned_free[i]);
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  }
  if (ctx->d_counter) cudaFree(ctx->d_counter);
  if (ctx->h_holes) cudaFreeHost(ctx->h_holes);
  if (ctx->d_pack_err) cudaFree(ctx->d_pack_err);
  for (ck_ctx::IngestLane *lane : ctx->lanes_all) {
    if (lane->stream) cudaStreamDestroy(lane->stream);
    if (lane->dep) cudaEventDestroy(lane->dep);
    for (cudaEvent_t ev : lane->ev)
      if (ev) cudaEventDestroy(ev);
    if (lane->staging) cudaFree(lane->staging);
    if (lane->d_err) cudaFree(lane->d_err);
    if (lane->h_err) cudaFreeHost(lane->h_err);
    delete lane;
  }
  if (ctx->dense_table) cudaFree(ctx->dense_table);
  for (void *p : ctx->out_pinned)
    if (p) cudaFreeHost(p);
  for (cudaEvent_t ev : ctx->event_pool) cudaEventDestroy(ev);
  if (ctx->syn_row) cudaFree(ctx->syn_row);
  if (ctx->syn_col) cudaFree(ctx->syn_col);
  if (ctx->syn_alt) cudaFree(ctx->syn_alt);
  if (ctx->result_buf) cudaFree(ctx->result_buf);
  if (ctx->tile_table) cudaFree(ctx->tile_table);
  if (ctx->tile_flags) cudaFree(ctx->tile_flags);
  if (ctx->d_screen_flagged) cudaFree(ctx->d_screen_flagged);
  if (ctx->sort_scratch) cudaFree(ctx->sort_scratch);
  for (int i = 0; i < ck_ctx::kCacheSlots; ++i)
    if (ctx->cache_ptr[i]) cudaFree(ctx->cache_ptr[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  delete ctx;
  return CK_OK;
}

/* ---- planes ---- */

int ck_planes_create(ck_ctx *ctx, const ck_submatrix *sm, uint32_t num_sites, ck_planes **out) {
  if (!ctx || !sm || !out) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = nullptr;
  if (sm->i_end < sm->i_begin || sm->j_end < sm->j_begin) return fail(CK_ERR_INVALID_ARGUMENT, "inverted sample range");
  if (!sm_diagonal(*sm) && !(sm->i_end <= sm->j_begin || sm->j_end <= sm->i_begin))
    return fail(CK_ERR_INVALID_ARGUMENT, "row and column ranges must be identical or disjoint");
  if (num_sites == 0) return fail(CK_ERR_INVALID_ARGUMENT, "num_sites is 0");
  DeviceGuard guard(ctx->device);
  ck_planes *pl = new (std::nothrow) ck_planes();
  if (!pl) return fail(CK_ERR_OUT_OF_MEMORY, "host allocation failed");
  pl->ctx = ctx;
  pl->map = make_slot_map(*sm);
  pl->num_sites = num_sites;
  pl->words = padded_words(num_sites);
  if (pl->map.num_blocks > 65535u) {
    delete pl;
    return fail(CK_ERR_INVALID_ARGUMENT, "more than 65535 x 64 samples in one shard; raise --split_factor");
  }
  pl->raw_bytes = std::max<size_t>(pl->raw_words(), 1) * 4;
  cudaError_t e = ctx_alloc(ctx, reinterpret_cast<void **>(&pl->raw), pl->raw_bytes);
  if (e != cudaSuccess) {  // the derived buffers (compute planes / genotype codes) are allocated on first use
    ck_planes_destroy(pl);
    return fail_cuda(e, "cudaMalloc(planes)", __FILE__, __LINE__);
  }
  *out = pl;
  return ck_planes_reset(pl);
}

int ck_planes_reset(ck_planes *pl) {
  if (!pl) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
  DeviceGuard guard(pl->ctx->device);
  if (pl->raw_words()) CK_CUDA(launch_fill_missing(pl->raw, pl->raw_words(), pl->ctx->stream));
  pl->mark_stale();
  CK_CUDA(cudaStreamSynchronize(pl->ctx->stream));
  return CK_OK;
}

int ck_planes_destroy(ck_planes *pl) {
  if (!pl) return CK_OK;
  DeviceGuard guard(pl->ctx->device);
  stream_discard(pl);
  cudaStreamSynchronize(pl->ctx->stream);
  ctx_release(pl->ctx, pl->raw, pl->raw_bytes);
  ctx_release(pl->ctx, pl->compute, pl->compute_bytes);
  ctx_release(pl->ctx, pl->codes, pl->codes_bytes);
  delete pl;
  return CK_OK;
}

int ck_planes_num_sites(const ck_planes *pl, uint32_t *num_sites) {
  if (!pl || !num_sites) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  *num_sites = pl->num_sites;
  return CK_OK;
}

int ck_planes_device_bytes(const ck_planes *pl, uint64_t *bytes) {
  if (!pl || !bytes) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  *bytes = uint64_t(pl->raw_words() + (pl->compute ? pl->compute_words() : 0) + (pl->codes ? pl->codes_words() : 0)) * 4;
  return CK_OK;
}

// Derives what the active pairwise kernel reads from the raw planes: the H/D/A compute planes (variants 0, 1) or the
// nibble-coded genotypes (variants 2 and 3 — different nibble values —, allocated on first use).
}  // extern "C"

namespace ck {
int ensure_compute(ck_planes *pl) {
  ck_ctx *ctx = pl->ctx;
  const int variant = planes_variant(pl);
  const bool want_codes = variant >= 2;
  const int kind = variant >= 3 ? 3 : variant;  // nibble encoding: 2 = int8 selectors, 3 = E2M1 (both mxf4 kernels)
  if (want_codes ? (!pl->codes_stale && pl->codes_kind == kind) : !pl->compute_stale) return CK_OK;
  if (want_codes && pl->codes == nullptr) {
    pl->codes_bytes = std::max<size_t>(pl->codes_alloc_words(), 1) * 4;
    CK_CUDA(ctx_alloc(ctx, reinterpret_cast<void **>(&pl->codes), pl->codes_bytes));
  }
  if (!want_codes && pl->compute == nullptr) {
    pl->compute_bytes = std::max<size_t>(pl->compute_words(), 1) * 4;
    CK_CUDA(ctx_alloc(ctx, reinterpret_cast<void **>(&pl->compute), pl->compute_bytes));
  }
  CK_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (want_codes && kind == 3) CK_CUDA(cudaMemsetAsync(pl->totals_sums(), 0, 2 * sizeof(unsigned long long), ctx->stream));
  if (pl->raw_words()) CK_CUDA(want_codes ? launch_finalize_codes(*pl, kind, ctx->stream) : launch_finalize(*pl, ctx->stream));
  CK_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  CK_CUDA(cudaEventSynchronize(ctx->ev[1]));
  ctx->timings.finalize_ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
  pl->screen1_floor = -1.f;
  if (want_codes && kind == 3 && variant == 5 && pl->raw_words()) {
    // cohort call rate r and mean het count -> the kinship an unrelated pair's one-product bound reaches: about
    // r + r^2 S / (2 het) (king_screen1_kernel.cu: the bound loses the sites where one sample is missing and the other not hom)
    unsigned long long sums[2] = {0, 0};
    CK_CUDA(cudaMemcpy(sums, pl->totals_sums(), sizeof(sums), cudaMemcpyDeviceToHost));
    pl->screen1_floor = screen1_floor_from_sums(sums, double(sm_samples(pl->map.sm)), double(pl->num_sites));
  }
  (want_codes ? pl->codes_stale : pl->compute_stale) = false;
  if (want_codes) pl->codes_kind = kind;
  return CK_OK;
}
}  // namespace ck

extern "C" {

int ck_planes_finalize(ck_planes *pl) {
  if (!pl) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
  DeviceGuard guard(pl->ctx->device);
  return ensure_compute(pl);
}

static int ensure_staging(ck_ctx *ctx, size_t bytes) {
  if (ctx->pinned_bytes >= bytes) return CK_OK;
  for (int i = 0; i < 2; ++i) {
    if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
    if (ctx->staging[i]) cudaFree(ctx->staging[i]);
    ctx->pinned[i] = ctx->staging[i] = nullptr;
  }
  ctx->pinned_bytes = 0;
  for (int i = 0; i < 2; ++i) {
    CK_CUDA(cudaHostAlloc(&ctx->pinned[i], bytes, cudaHostAllocDefault));
    CK_CUDA(dev_alloc(ctx, &ctx->staging[i], bytes));
    if (!ctx->pinned_free[i]) CK_CUDA(cudaEventCreateWithFlags(&ctx->pinned_free[i], cudaEventDisableTiming));
  }
  ctx->pinned_bytes = bytes;
  return CK_OK;
}

int ck_pack_triples(ck_planes *pl, const int64_t *row_idx, const int64_t *col_idx, const int32_t *n_alt_alleles,
                    size_t n, int on_device) {
  if (!pl) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
  if (n == 0) return CK_OK;
  if (!row_idx || !col_idx || !n_alt_alleles) return fail(CK_ERR_INVALID_ARGUMENT, "NULL triple array");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  CK_CUDA(cudaMemsetAsync(ctx->d_pack_err, 0xff, 2 * sizeof(unsigned long long), s));
  pl->mark_stale();
  CK_CUDA(cudaEventRecord(ctx->ev[0], s));
  auto is_pinned = [](const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    return at.type == cudaMemoryTypeHost && at.devicePointer != nullptr;
  };
  if (on_device) {
    CK_CUDA(launch_pack(*pl, row_idx, col_idx, n_alt_alleles, n, 0, ctx->d_pack_err, s));
  } else if (is_pinned(row_idx) && is_pinned(col_idx) && is_pinned(n_alt_alleles)) {
    // Page-locked host arrays are addressable from the device (UVA): the kernel streams them over PCIe in place.
    CK_CUDA(launch_pack(*pl, row_idx, col_idx, n_alt_alleles, n, 0, ctx->d_pack_err, s));
  } else {
    // Host arrays: memcpy into one of two pinned buffers while the GPU consumes the other (H2D + pack kernel).
    constexpr size_t kChunk = size_t(4) << 20;  // triples per chunk (80 MiB of staging per buffer)
    constexpr size_t kBytesPer = 8 + 8 + 4;
    const size_t chunk = std::min(n, kChunk);
    const size_t chunk_pad = (chunk + 1) & ~size_t(1);
    int rc = ensure_staging(ctx, chunk_pad * kBytesPer);
    if (rc != CK_OK) return rc;
    size_t done = 0;
    for (int buf = 0; done < n; buf ^= 1) {
      const size_t m = std::min(chunk, n - done);
      CK_CUDA(cudaEventSynchronize(ctx->pinned_free[buf]));  // a never-recorded event is complete
      char *h = static_cast<char *>(ctx->pinned[buf]);
      char *d = static_cast<char *>(ctx->staging[buf]);
      memcpy(h, row_idx + done, m * 8);
      memcpy(h + chunk_pad * 8, col_idx + done, m * 8);
      memcpy(h + chunk_pad * 16, n_alt_alleles + done, m * 4);
      CK_CUDA(cudaMemcpyAsync(d, h, chunk_pad * kBytesPer, cudaMemcpyHostToDevice, s));
      CK_CUDA(launch_pack(*pl, reinterpret_cast<int64_t *>(d), reinterpret_cast<int64_t *>(d + chunk_pad * 8),
                          reinterpret_cast<int32_t *>(d + chunk_pad * 16), m, done, ctx->d_pack_err, s));
      CK_CUDA(cudaEventRecord(ctx->pinned_free[buf], s));
      done += m;
    }
  }
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  unsigned long long err[2];
  CK_CUDA(cudaMemcpyAsync(err, ctx->d_pack_err, sizeof(err), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  ctx->timings.pack_ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
  if (err[0] != ~0ull) {
    const size_t idx = size_t(err[0]) - 1;
    int32_t value = 0;
    if (on_device)
      cudaMemcpy(&value, n_alt_alleles + idx, sizeof(value), cudaMemcpyDeviceToHost);
    else
      value = n_alt_alleles[idx];
    return fail(CK_ERR_INVALID_GENOTYPE, "Invalid value for n_alt_alleles (" + std::to_string(value) +
                                             ") encountered at triple " + std::to_string(idx));
  }
  if (err[1] != ~0ull)
    return fail(CK_ERR_OUT_OF_RANGE, "row_idx out of range [0, num_sites) at triple " + std::to_string(size_t(err[1]) - 1));
  return CK_OK;
}

int ck_pack_triples_narrow(ck_planes *pl, const uint32_t *row_idx, const uint32_t *col_idx, const uint8_t *n_alt_alleles,
                           size_t n, int on_device) {
  if (!pl) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
  if (n == 0) return CK_OK;
  if (!row_idx || !col_idx || !n_alt_alleles) return fail(CK_ERR_INVALID_ARGUMENT, "NULL triple array");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  auto device_readable = [](const void *p) {  // device memory, or page-locked host memory mapped into the device
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    return at.devicePointer != nullptr && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
  };
  struct Staged {  // pageable host arrays: one device copy of the three arrays, returned to the ctx cache on scope exit
    ck_ctx *ctx;
    void *p = nullptr;
    size_t bytes = 0;
    ~Staged() { ctx_release(ctx, p, bytes); }
  } staged{ctx};
  const uint32_t *d_row = row_idx, *d_col = col_idx;
  const uint8_t *d_alt = n_alt_alleles;
  CK_CUDA(cudaMemsetAsync(ctx->d_pack_err, 0xff, 2 * sizeof(unsigned long long), s));
  pl->mark_stale();
  CK_CUDA(cudaEventRecord(ctx->ev[0], s));
  if (!on_device && !(device_readable(row_idx) && device_readable(col_idx) && device_readable(n_alt_alleles))) {
    const size_t n4 = (n + 3) & ~size_t(3);
    staged.bytes = n4 * 9;
    CK_CUDA(ctx_alloc(ctx, &staged.p, staged.bytes));
    char *d = static_cast<char *>(staged.p);
    CK_CUDA(cudaMemcpyAsync(d, row_idx, n * 4, cudaMemcpyHostToDevice, s));
    CK_CUDA(cudaMemcpyAsync(d + n4 * 4, col_idx, n * 4, cudaMemcpyHostToDevice, s));
    CK_CUDA(cudaMemcpyAsync(d + n4 * 8, n_alt_alleles, n, cudaMemcpyHostToDevice, s));
    d_row = reinterpret_cast<const uint32_t *>(d);
    d_col = reinterpret_cast<const uint32_t *>(d + n4 * 4);
    d_alt = reinterpret_cast<const uint8_t *>(d + n4 * 8);
  }
  CK_CUDA(launch_pack_narrow(*pl, d_row, d_col, d_alt, n, 0, ctx->d_pack_err, s));
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  unsigned long long err[2];
  CK_CUDA(cudaMemcpyAsync(err, ctx->d_pack_err, sizeof(err), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  ctx->timings.pack_ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
  if (err[0] != ~0ull) {
    const size_t idx = size_t(err[0]) - 1;
    uint8_t value = 0;
    if (on_device) cudaMemcpy(&value, n_alt_alleles + idx, 1, cudaMemcpyDeviceToHost);
    else value = n_alt_alleles[idx];
    return fail(CK_ERR_INVALID_GENOTYPE, "Invalid value for n_alt_alleles (" + std::to_string(unsigned(value)) +
                                             ") encountered at triple " + std::to_string(idx));
  }
  if (err[1] != ~0ull)
    return fail(CK_ERR_OUT_OF_RANGE, "row_idx out of range [0, num_sites) at triple " + std::to_string(size_t(err[1]) - 1));
  return CK_OK;
}

int ck_host_alloc(size_t bytes, void **out) {
  if (!out) return fail(CK_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  CK_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return CK_OK;
}

int ck_host_free(void *ptr) {
  if (ptr) CK_CUDA(cudaFreeHost(ptr));
  return CK_OK;
}

int ck_planes_import_bitset(ck_planes *pl, const uint64_t *bit_set, int on_device) {
  if (!pl || !bit_set) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  const size_t bytes = size_t(ref_words_per_sample(pl->num_sites)) * sm_samples(pl->map.sm) * 8;
  struct Staging {  // device copy of a host bit set, returned to the ctx cache on scope exit
    ck_ctx *ctx;
    void *p = nullptr;
    size_t bytes = 0;
    ~Staging() { ctx_release(ctx, p, bytes); }
  } tmp{ctx};
  const uint64_t *d_src = bit_set;
  CK_CUDA(cudaEventRecord(ctx->ev[0], s));
  if (!on_device) {
    CK_CUDA(ctx_alloc(ctx, &tmp.p, bytes));
    tmp.bytes = bytes;
    CK_CUDA(cudaMemcpyAsync(tmp.p, bit_set, bytes, cudaMemcpyHostToDevice, s));
    d_src = static_cast<const uint64_t *>(tmp.p);
  }
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  if (pl->raw_words()) {
    CK_CUDA(launch_fill_missing(pl->raw, pl->raw_words(), s));
    CK_CUDA(launch_import_ref(*pl, d_src, s));
  }
  cudaEvent_t ev2;
  CK_CUDA(cudaEventCreate(&ev2));
  cudaEventRecord(ev2, s);
  cudaError_t e = cudaStreamSynchronize(s);
  ctx->timings.h2d_ms = elapsed_ms(ctx->ev[0], ctx->ev[1]);
  ctx->timings.import_ms = elapsed_ms(ctx->ev[1], ev2);
  cudaEventDestroy(ev2);
  CK_CUDA(e);
  pl->mark_stale();
  return CK_OK;
}

int ck_planes_export_bitset(ck_planes *pl, uint64_t *bit_set, int on_device) {
  if (!pl || !bit_set) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  const size_t bytes = size_t(ref_words_per_sample(pl->num_sites)) * sm_samples(pl->map.sm) * 8;
  DevBuf tmp(ctx);
  uint64_t *d_dst = bit_set;
  if (!on_device) {
    CK_CUDA(tmp.alloc(bytes));
    d_dst = tmp.as<uint64_t>();
  }
  // words of the reference layout that have no plane word behind them do not exist (Wp >= W), so every word is written
  if (pl->raw_words()) CK_CUDA(launch_export_ref(*pl, d_dst, s));
  if (!on_device) CK_CUDA(cudaMemcpyAsync(bit_set, d_dst, bytes, cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  return CK_OK;
}

int ck_planes_synthesize(ck_planes *pl, const ck_synth_params *params) {
  if (!pl || !params) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  if (pl->raw_words()) {
    CK_CUDA(launch_fill_missing(pl->raw, pl->raw_words(), ctx->stream));
    CK_CUDA(launch_synth_planes(*pl, params->seed, missing_threshold(params->missing_rate), ctx->stream));
  }
  pl->mark_stale();
  CK_CUDA(cudaStreamSynchronize(ctx->stream));
  return CK_OK;
}

/* ---- synthetic inputs ---- */

int ck_synth_genotypes_host(const ck_synth_params *params, uint32_t sample_begin, uint32_t sample_end,
                            uint32_t site_begin, uint32_t site_end, int8_t *out) {
  if (!params || !out) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (sample_end < sample_begin || site_end < site_begin) return fail(CK_ERR_INVALID_ARGUMENT, "inverted range");
  const uint32_t thr = missing_threshold(params->missing_rate);
  const size_t num_sites = site_end - site_begin;
  if (sample_begin == sample_end) return CK_OK;
  for (uint32_t group = sample_begin / 8; group <= (sample_end - 1) / 8; ++group) {
    const PedigreeKeys keys = pedigree_keys(params->seed, group);
    for (uint32_t site = site_begin; site < site_end; ++site) {
      int8_t g[8];
      pedigree_genotypes(keys, site, thr, g);
      for (uint32_t m = 0; m < 8; ++m) {
        const uint32_t s = group * 8 + m;
        if (s >= sample_begin && s < sample_end) out[size_t(s - sample_begin) * num_sites + (site - site_begin)] = g[m];
      }
    }
  }
  return CK_OK;
}

int ck_synth_triples_device(ck_ctx *ctx, const ck_synth_params *params, uint32_t sample_begin, uint32_t sample_end,
                            uint32_t site_begin, uint32_t site_end, const int64_t **row_idx, const int64_t **col_idx,
                            const int32_t **n_alt_alleles, size_t *num_triples) {
  if (!ctx || !params || !row_idx || !col_idx || !n_alt_alleles || !num_triples)
    return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (sample_end < sample_begin || site_end <= site_begin) return fail(CK_ERR_INVALID_ARGUMENT, "empty or inverted range");
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  const uint32_t sites = site_end - site_begin;
  const uint32_t thr = missing_threshold(params->missing_rate);
  DevBuf counts(ctx), offsets(ctx), tmp(ctx);
  CK_CUDA(counts.alloc(size_t(sites) * 8));
  CK_CUDA(offsets.alloc(size_t(sites) * 8));
  CK_CUDA(launch_synth_count(params->seed, thr, sample_begin, sample_end, site_begin, site_end,
                             counts.as<unsigned long long>(), s));
  size_t tmp_bytes = 0;
  CK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.as<unsigned long long>(),
                                        offsets.as<unsigned long long>(), int(sites), s));
  CK_CUDA(tmp.alloc(tmp_bytes));
  CK_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts.as<unsigned long long>(),
                                        offsets.as<unsigned long long>(), int(sites), s));
  unsigned long long last_off = 0, last_cnt = 0;
  CK_CUDA(cudaMemcpyAsync(&last_off, offsets.as<unsigned long long>() + (sites - 1), 8, cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaMemcpyAsync(&last_cnt, counts.as<unsigned long long>() + (sites - 1), 8, cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  const size_t total = size_t(last_off + last_cnt);
  if (total > ctx->syn_cap) {
    if (ctx->syn_row) cudaFree(ctx->syn_row);
    if (ctx->syn_col) cudaFree(ctx->syn_col);
    if (ctx->syn_alt) cudaFree(ctx->syn_alt);
    ctx->syn_row = ctx->syn_col = nullptr;
    ctx->syn_alt = nullptr;
    ctx->syn_cap = 0;
    const size_t cap = (total + 1) & ~size_t(1);
    CK_CUDA(dev_alloc(ctx, reinterpret_cast<void **>(&ctx->syn_row), cap * 8));
    CK_CUDA(dev_alloc(ctx, reinterpret_cast<void **>(&ctx->syn_col), cap * 8));
    CK_CUDA(dev_alloc(ctx, reinterpret_cast<void **>(&ctx->syn_alt), cap * 4));
    ctx->syn_cap = cap;
  }
  if (total)
    CK_CUDA(launch_synth_emit(params->seed, thr, sample_begin, sample_end, site_begin, site_end,
                              offsets.as<unsigned long long>(), ctx->syn_row, ctx->syn_col, ctx->syn_alt, s));
  CK_CUDA(cudaStreamSynchronize(s));
  *row_idx = ctx->syn_row;
  *col_idx = ctx->syn_col;
  *n_alt_alleles = ctx->syn_alt;
  *num_triples = total;
  return CK_OK;
}

}  // extern "C"
