// C ABI of libcuking_b200.so (include/cuking_b200.h): contexts, plane storage, pack / import / export, the pairwise
// call with its result compaction, sort and copy-out.  Reference seam: /root/reference/cuking.cu:505-523 (planning +
// allocation), :675-703 (pack), :713-765 (result buffer, launch, overflow check, sort).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
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

namespace {

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
  void *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T>
  T *as() const { return static_cast<T *>(p); }
};

__global__ void make_sort_keys_kernel(const ck_result *res, uint32_t n, unsigned long long *keys, uint32_t *idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = (static_cast<unsigned long long>(res[i].sample_i) << 32) | res[i].sample_j;
    idx[i] = i;
  }
}
__global__ void gather_results_kernel(const ck_result *in, const uint32_t *idx, uint32_t n, ck_result *out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

}  // namespace

static bool dbg_alloc() {
  static const bool on = getenv("CUKING_DEBUG_ALLOC") != nullptr;
  return on;
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
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaErrorMemoryAllocation) {  // make room: drop everything cached and retry once
    cudaGetLastError();
    for (int i = 0; i < ck_ctx::kCacheSlots; ++i)
      if (ctx->cache_ptr[i]) {
        cudaFree(ctx->cache_ptr[i]);
        ctx->cache_ptr[i] = nullptr;
        ctx->cache_bytes[i] = 0;
      }
    e = cudaMalloc(ptr, bytes);
  }
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
// 2 = tcgen05 int8 tensor-core formulation, 3 = tcgen05 mxf4 (E2M1, fp32 accumulation) formulation
static int g_default_variant = 3;
static int active_variant(const ck_ctx *ctx) { return ctx->king_variant >= 0 ? ctx->king_variant : g_default_variant; }
// The variant that actually runs on these planes: the fp32 accumulators of the mxf4 kernel are exact only up to the
// site count the probe verified (kFp4MaxSites); longer genotype vectors take the int8 kernel (s32 accumulators).
static int planes_variant(const ck_planes *pl) {
  const int v = active_variant(pl->ctx);
  return (v == 3 && pl->num_sites > kFp4MaxSites) ? 2 : v;
}

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
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev[0])) != cudaSuccess || (e = cudaEventCreate(&ctx->ev[1])) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_counter, 2 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_pack_err, 2 * sizeof(uint32_t))) != cudaSuccess) {
    ck_ctx_destroy(ctx);
    return fail_cuda(e, "ck_ctx_create", __FILE__, __LINE__);
  }
  ctx->stream = ctx->own_stream;
  ctx->king_variant = king_variant_from_env();
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
  if (variant < -1 || variant > 3) return fail(CK_ERR_INVALID_ARGUMENT, "unknown pairwise kernel variant");
  ctx->king_variant = variant;
  return CK_OK;
}

int ck_ctx_synchronize(ck_ctx *ctx) {
  if (!ctx) return fail(CK_ERR_INVALID_ARGUMENT, "ctx is NULL");
  DeviceGuard guard(ctx->device);
  CK_CUDA(cudaStreamSynchronize(ctx->stream));
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
  if (ctx->d_pack_err) cudaFree(ctx->d_pack_err);
  if (ctx->syn_row) cudaFree(ctx->syn_row);
  if (ctx->syn_col) cudaFree(ctx->syn_col);
  if (ctx->syn_alt) cudaFree(ctx->syn_alt);
  if (ctx->result_buf) cudaFree(ctx->result_buf);
  if (ctx->tile_table) cudaFree(ctx->tile_table);
  if (ctx->sort_scratch) cudaFree(ctx->sort_scratch);
  for (int i = 0; i < ck_ctx::kCacheSlots; ++i)
    if (ctx->cache_ptr[i]) cudaFree(ctx->cache_ptr[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
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
  if (pl && pl->stream_state) {
    cudaStreamSynchronize(pl->ctx->stream);
    delete pl->stream_state;
    pl->stream_state = nullptr;
  }
  if (!pl) return CK_OK;
  DeviceGuard guard(pl->ctx->device);
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
static int ensure_compute(ck_planes *pl) {
  ck_ctx *ctx = pl->ctx;
  const int variant = planes_variant(pl);
  const bool want_codes = variant >= 2;
  if (want_codes ? (!pl->codes_stale && pl->codes_kind == variant) : !pl->compute_stale) return CK_OK;
  if (want_codes && pl->codes == nullptr) {
    pl->codes_bytes = std::max<size_t>(pl->codes_words(), 1) * 4;
    CK_CUDA(ctx_alloc(ctx, reinterpret_cast<void **>(&pl->codes), pl->codes_bytes));
  }
  if (!want_codes && pl->compute == nullptr) {
    pl->compute_bytes = std::max<size_t>(pl->compute_words(), 1) * 4;
    CK_CUDA(ctx_alloc(ctx, reinterpret_cast<void **>(&pl->compute), pl->compute_bytes));
  }
  CK_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (pl->raw_words()) CK_CUDA(want_codes ? launch_finalize_codes(*pl, variant, ctx->stream) : launch_finalize(*pl, ctx->stream));
  CK_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  CK_CUDA(cudaEventSynchronize(ctx->ev[1]));
  ctx->timings.finalize_ms = elapsed(ctx->ev[0], ctx->ev[1]);
  (want_codes ? pl->codes_stale : pl->compute_stale) = false;
  if (want_codes) pl->codes_kind = variant;
  return CK_OK;
}

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
    CK_CUDA(cudaMalloc(&ctx->staging[i], bytes));
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
  CK_CUDA(cudaMemsetAsync(ctx->d_pack_err, 0xff, 2 * sizeof(uint32_t), s));
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
  uint32_t err[2];
  CK_CUDA(cudaMemcpyAsync(err, ctx->d_pack_err, sizeof(err), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  ctx->timings.pack_ms = elapsed(ctx->ev[0], ctx->ev[1]);
  if (err[0] != 0xffffffffu) {
    const size_t idx = size_t(err[0]) - 1;
    int32_t value = 0;
    if (on_device)
      cudaMemcpy(&value, n_alt_alleles + idx, sizeof(value), cudaMemcpyDeviceToHost);
    else
      value = n_alt_alleles[idx];
    return fail(CK_ERR_INVALID_GENOTYPE, "Invalid value for n_alt_alleles (" + std::to_string(value) +
                                             ") encountered at triple " + std::to_string(idx));
  }
  if (err[1] != 0xffffffffu)
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
  ctx->timings.h2d_ms = elapsed(ctx->ev[0], ctx->ev[1]);
  ctx->timings.import_ms = elapsed(ctx->ev[1], ev2);
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
  DevBuf tmp;
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

/* ---- pairwise ---- */

static int ensure_result_buf(ck_ctx *ctx, size_t records) {
  if (ctx->result_cap >= records) return CK_OK;
  if (ctx->result_buf) cudaFree(ctx->result_buf);
  ctx->result_buf = nullptr;
  ctx->result_cap = 0;
  CK_CUDA(cudaMalloc(&ctx->result_buf, std::max<size_t>(records, 1) * sizeof(ck_result)));
  ctx->result_cap = records;
  return CK_OK;
}

static int ensure_sort_scratch(ck_ctx *ctx, size_t bytes) {
  if (ctx->sort_scratch_bytes >= bytes) return CK_OK;
  if (ctx->sort_scratch) cudaFree(ctx->sort_scratch);
  ctx->sort_scratch = nullptr;
  ctx->sort_scratch_bytes = 0;
  const size_t want = bytes + bytes / 4;  // head-room so that slightly larger result sets do not reallocate
  CK_CUDA(cudaMalloc(&ctx->sort_scratch, want));
  ctx->sort_scratch_bytes = want;
  return CK_OK;
}

// Sorts n device records by (sample_i, sample_j) — the pair is unique, so this equals the reference's
// (sample_i, sample_j, kin) order (cuking.cu:761-765).  `out` (device) receives the sorted records; when it is NULL
// they stay in the ctx scratch and *sorted points at them.
static int sort_results(ck_ctx *ctx, const ck_result *in, uint32_t n, ck_result *out, const ck_result **sorted) {
  cudaStream_t s = ctx->stream;
  auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t tmp_bytes = 0;
  {  // size query only (no work is launched); any valid device address serves as the pointer arguments
    auto *k64 = reinterpret_cast<unsigned long long *>(ctx->d_counter);
    auto *v32 = reinterpret_cast<uint32_t *>(ctx->d_counter);
    CK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k64, k64, v32, v32, int(n), 0, 64, s));
  }
  const size_t keys_b = align(size_t(n) * 8), idx_b = align(size_t(n) * 4), tmp_b = align(tmp_bytes),
               rec_b = out ? 0 : align(size_t(n) * sizeof(ck_result));
  int rc = ensure_sort_scratch(ctx, 2 * keys_b + 2 * idx_b + tmp_b + rec_b);
  if (rc != CK_OK) return rc;
  char *base = static_cast<char *>(ctx->sort_scratch);
  auto *keys_a = reinterpret_cast<unsigned long long *>(base);
  auto *keys_o = reinterpret_cast<unsigned long long *>(base + keys_b);
  auto *idx_a = reinterpret_cast<uint32_t *>(base + 2 * keys_b);
  auto *idx_o = reinterpret_cast<uint32_t *>(base + 2 * keys_b + idx_b);
  void *tmp = base + 2 * keys_b + 2 * idx_b;
  ck_result *dst = out ? out : reinterpret_cast<ck_result *>(base + 2 * keys_b + 2 * idx_b + tmp_b);
  const unsigned grid = (n + 255) / 256;
  make_sort_keys_kernel<<<grid, 256, 0, s>>>(in, n, keys_a, idx_a);
  CK_CUDA(cudaGetLastError());
  CK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_o, idx_a, idx_o, int(n), 0, 64, s));
  gather_results_kernel<<<grid, 256, 0, s>>>(in, idx_o, n, dst);
  CK_CUDA(cudaGetLastError());
  if (sorted) *sorted = dst;
  return CK_OK;
}

static uint64_t variant_num_tiles(const ck_planes *pl, const KingLaunch &k) {
  const int variant = planes_variant(pl);
  if (variant == 3) return king_fp4_num_tiles(k);
  if (variant == 2) return king_umma_num_tiles(k);
  return king_num_tiles(k.num_row_blocks, k.num_col_blocks, k.triangular != 0);
}

static cudaError_t dispatch_king(const ck_planes *pl, const KingLaunch &k, cudaStream_t s, uint32_t *launches) {
  const int variant = planes_variant(pl);
  if (variant == 3) return launch_king_fp4(k, pl->map.num_blocks, pl->ctx, s, launches);
  if (variant == 2) return launch_king_umma(k, pl->map.num_blocks, pl->ctx, s, launches);
  return launch_king(k, variant, s, launches);
}

static KingLaunch base_launch(const ck_planes *pl) {
  const ck_submatrix &sm = pl->map.sm;
  KingLaunch k{};
  k.compute = pl->compute;
  k.codes = pl->codes;
  k.words = pl->words;
  k.row_block0 = 0;
  k.num_row_blocks = ceil_div(sm_rows(sm), kTileSamples);
  k.col_block0 = pl->map.col_slot0 / kTileSamples;
  k.num_col_blocks = ceil_div(sm_cols(sm), kTileSamples);
  k.row_global0 = sm.i_begin;
  k.col_global0 = sm.j_begin;
  k.num_rows = sm_rows(sm);
  k.num_cols = sm_cols(sm);
  k.triangular = sm_diagonal(sm) ? 1u : 0u;
  return k;
}

int ck_planes_king_variant(const ck_planes *pl, int *variant) {
  if (!pl || !variant) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  *variant = planes_variant(pl);
  return CK_OK;
}

int ck_king_num_tiles(const ck_planes *pl, uint64_t *num_tiles) {
  if (!pl || !num_tiles) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  const KingLaunch k = base_launch(pl);
  *num_tiles = variant_num_tiles(pl, k);
  return CK_OK;
}

// Common tail of the pairwise entry points: reads the emitted-pair counter, turns an overflow into the reference's
// error (cuking.cu:747-751), sorts on the device and copies the records out.  Expects ev[0] / ev[1] recorded around
// the kernel launches on the ctx stream.
static int finish_results(ck_ctx *ctx, ck_result *d_emit, uint32_t max_results, ck_result *results, int results_on_device,
                          uint32_t *num_results, int sort) {
  cudaStream_t s = ctx->stream;
  static const bool dbg = getenv("CUKING_DEBUG_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  auto tp0 = now();
  unsigned long long count = 0;
  CK_CUDA(cudaMemcpyAsync(&count, ctx->d_counter, sizeof(count), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  ctx->timings.king_ms = elapsed(ctx->ev[0], ctx->ev[1]);
  if (dbg) fprintf(stderr, "[ck] king: host wait %.2f ms, kernel %.2f ms\n", ms_since(tp0), ctx->timings.king_ms);
  auto tp1 = now();
  *num_results = count > 0xffffffffull ? 0xffffffffu : uint32_t(count);
  if (count > max_results)  // cuking.cu:747-751
    return fail(CK_ERR_RESULT_OVERFLOW, "Could not store all results: try increasing the --max_results parameter.");
  const uint32_t n = uint32_t(count);
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
    int rc = sort_results(ctx, d_emit, n, results_on_device ? results : nullptr, &d_final);
    if (rc != CK_OK) return rc;
    ctx->timings.king_launches += 3;  // key build, radix sort (one logical launch), gather
  }
  if (dbg) fprintf(stderr, "[ck] sort: host %.2f ms\n", ms_since(tp1));
  cudaEventRecord(t1, s);
  if (!results_on_device) CK_CUDA(cudaMemcpyAsync(results, d_final, size_t(n) * sizeof(ck_result), cudaMemcpyDeviceToHost, s));
  cudaEventRecord(t2, s);
  cudaError_t e = cudaStreamSynchronize(s);
  ctx->timings.sort_ms = elapsed(t0, t1);
  ctx->timings.d2h_ms = elapsed(t1, t2);
  if (dbg) fprintf(stderr, "[ck] sort+d2h: host %.2f ms (device sort %.2f, d2h %.2f)\n", ms_since(tp1), ctx->timings.sort_ms, ctx->timings.d2h_ms);
  CK_CUDA(e);
  return CK_OK;
}

int ck_king_tiles(ck_planes *pl, uint64_t tile_begin, uint64_t tile_end, float kin_threshold, uint32_t max_results,
                  ck_result *results, int results_on_device, uint32_t *num_results, int sort) {
  if (!pl || !num_results) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (max_results > 0 && !results) return fail(CK_ERR_INVALID_ARGUMENT, "results is NULL");
  *num_results = 0;
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream;
  int rc = ensure_compute(pl);  // may allocate the buffer base_launch() points at
  if (rc != CK_OK) return rc;
  KingLaunch k = base_launch(pl);
  const uint64_t total = variant_num_tiles(pl, k);
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
  k.dump_counts = nullptr;
  k.dump_kin = nullptr;
  ctx->timings.king_launches = 0;
  CK_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), s));
  CK_CUDA(cudaEventRecord(ctx->ev[0], s));
  if (tile_end > tile_begin) CK_CUDA(dispatch_king(pl, k, s, &ctx->timings.king_launches));
  CK_CUDA(cudaEventRecord(ctx->ev[1], s));
  return finish_results(ctx, d_emit, max_results, results, results_on_device, num_results, sort);
}

int ck_king(ck_planes *pl, float kin_threshold, uint32_t max_results, ck_result *results, int results_on_device,
            uint32_t *num_results, int sort) {
  if (!pl) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
  uint64_t tiles = 0;
  int rc = ck_king_num_tiles(pl, &tiles);
  if (rc != CK_OK) return rc;
  return ck_king_tiles(pl, 0, tiles, kin_threshold, max_results, results, results_on_device, num_results, sort);
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
  DevBuf d_counts, d_kin;
  CK_CUDA(d_counts.alloc(rows * cols * sizeof(ck_counts)));
  CK_CUDA(d_kin.alloc(rows * cols * sizeof(float)));
  CK_CUDA(cudaMemsetAsync(d_counts.p, 0, rows * cols * sizeof(ck_counts), s));
  CK_CUDA(cudaMemsetAsync(d_kin.p, 0, rows * cols * sizeof(float), s));
  KingLaunch k = base_launch(pl);
  k.tile_begin = 0;
  k.tile_end = variant_num_tiles(pl, k);
  k.kin_threshold = 2.f;  // nothing is emitted: kin <= 0.5
  k.max_results = 0;
  k.results = nullptr;
  k.counter = ctx->d_counter;
  k.dump_counts = d_counts.as<ck_counts>();
  k.dump_kin = d_kin.as<float>();
  CK_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), s));
  CK_CUDA(dispatch_king(pl, k, s, nullptr));
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

// ---- pipelined host-buffer path ------------------------------------------------------------------------------------
// ck_king_host_bitset on a diagonal shard with the mxf4 kernel: the upload of the reference-layout bit set overlaps
// the pairwise kernel instead of preceding it.  A band of the tile enumeration (kFp4BandRows rows) only needs the
// samples at or after its first row (i < j), so the sample range is uploaded LAST CHUNK FIRST on the copy stream and
// every chunk's bands are launched as soon as its rows have been transposed and coded - by then every column they
// pair with is already on the device.  The bottom chunks hold few tiles, so only the first small upload is exposed.
static bool host_bitset_can_pipeline(const ck_planes *pl) {
  static const bool off = getenv("CUKING_NO_PIPELINE") != nullptr;
  return !off && planes_variant(pl) >= 2 && sm_diagonal(pl->map.sm) && sm_rows(pl->map.sm) >= 4 * kFp4BandRows;
}

// Owner of band b when the shard is split into num_parts parts: bands are dealt in snake order (0 .. P-1, P-1 .. 0, ...):
// the tile count of a band falls linearly with its index, so every pair (g, 2P-1-g) of a group carries the same work.
static uint32_t band_owner(uint32_t band, uint32_t num_parts) {
  const uint32_t g = band % (2 * num_parts);
  return g < num_parts ? g : 2 * num_parts - 1 - g;
}

}  // extern "C"

static int stream_begin_impl(ck_planes *pl, float kin_threshold, uint32_t max_results, uint32_t part_index, uint32_t num_parts) {
  ck_ctx *ctx = pl->ctx;
  if (num_parts == 0 || part_index >= num_parts) return fail(CK_ERR_INVALID_ARGUMENT, "part_index outside [0, num_parts)");
  if (!sm_diagonal(pl->map.sm)) return fail(CK_ERR_INVALID_ARGUMENT, "streaming delivery needs a diagonal shard");
  const int variant = planes_variant(pl);
  if (variant < 2) return fail(CK_ERR_INVALID_ARGUMENT, "streaming delivery needs a tensor-core kernel variant (2 or 3): their band-ordered tiles");
  if (pl->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "a stream session is already open on these planes");
  if (pl->codes == nullptr) {
    pl->codes_bytes = std::max<size_t>(pl->codes_words(), 1) * 4;
    CK_CUDA(ctx_alloc(ctx, reinterpret_cast<void **>(&pl->codes), pl->codes_bytes));
  }
  int rc = ensure_result_buf(ctx, max_results);
  if (rc != CK_OK) return rc;
  KingStream *st = new (std::nothrow) KingStream();
  if (!st) return fail(CK_ERR_OUT_OF_MEMORY, "host allocation failed");
  st->k = base_launch(pl);
  st->k.kin_threshold = kin_threshold;
  st->k.max_results = max_results;
  st->k.results = ctx->result_buf;
  st->k.counter = ctx->d_counter;
  st->variant = variant;
  st->part_index = part_index;
  st->num_parts = num_parts;
  st->max_results = max_results;
  st->next_end = sm_rows(pl->map.sm);
  cudaError_t e = band_prepare(st->k, kBandTileCols, ctx, ctx->stream, &st->band_prefix, nullptr);
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), ctx->stream);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[0], ctx->stream);
  if (e != cudaSuccess) {
    delete st;
    return fail_cuda(e, "ck_king_stream_begin", __FILE__, __LINE__);
  }
  ctx->timings.king_launches = 0;
  pl->mark_stale();
  pl->stream_state = st;
  return CK_OK;
}

// Rows [s0, s1) of the shard, in device memory behind d_rows: transpose, derive the codes, launch this part's bands among
// them.  Everything is queued on the ctx stream; nothing is synchronised.
static int stream_rows_device(ck_planes *pl, const uint64_t *d_rows, uint32_t s0, uint32_t s1) {
  KingStream *st = pl->stream_state;
  ck_ctx *ctx = pl->ctx;
  cudaStream_t s = ctx->stream;
  const uint32_t n = sm_rows(pl->map.sm);
  if (s1 != st->next_end || s0 >= s1 || s0 % kFp4BandRows != 0)
    return fail(CK_ERR_INVALID_ARGUMENT, "stream rows must arrive in descending ranges that tile the shard at multiples of "
                                         "ck_king_stream_granularity()");
  const uint32_t block0 = s0 / kTileSamples, num_blocks = ceil_div(s1, kTileSamples) - block0;
  CK_CUDA(launch_import_ref_range(*pl, d_rows, s0, block0, num_blocks, s));
  CK_CUDA(launch_finalize_codes_range(*pl, st->variant, block0, num_blocks, s));
  ctx->timings.king_launches += 2;
  const uint32_t band_lo = s0 / kFp4BandRows, band_hi = ceil_div(std::min(s1, n), kFp4BandRows);
  KingLaunch k = st->k;
  for (uint32_t b = band_lo; b < band_hi;) {  // maximal runs of this part's bands (the whole range when num_parts == 1)
    if (band_owner(b, st->num_parts) != st->part_index) { ++b; continue; }
    uint32_t e = b + 1;
    while (e < band_hi && band_owner(e, st->num_parts) == st->part_index) ++e;
    k.tile_begin = st->band_prefix[b];
    k.tile_end = st->band_prefix[e];
    if (k.tile_end > k.tile_begin)
      CK_CUDA(st->variant == 3 ? launch_king_fp4(k, pl->map.num_blocks, ctx, s, &ctx->timings.king_launches)
                               : launch_king_umma(k, pl->map.num_blocks, ctx, s, &ctx->timings.king_launches));
    b = e;
  }
  st->next_end = s0;
  return CK_OK;
}

static int stream_end_impl(ck_planes *pl, ck_result *results, uint32_t *num_results) {
  KingStream *st = pl->stream_state;
  ck_ctx *ctx = pl->ctx;
  const bool complete = st->next_end == 0;
  const uint32_t max_results = st->max_results;
  const int variant = st->variant;
  delete st;
  pl->stream_state = nullptr;
  CK_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (!complete) {
    cudaStreamSynchronize(ctx->stream);
    return fail(CK_ERR_INVALID_ARGUMENT, "ck_king_stream_end before every row of the shard was delivered");
  }
  pl->compute_stale = true;
  pl->codes_stale = false;
  pl->codes_kind = variant;
  return finish_results(ctx, ctx->result_buf, max_results, results, 0, num_results, 1);
}

static int king_host_bitset_pipelined(ck_planes *pl, const uint64_t *bit_set, float kin_threshold, uint32_t max_results,
                                      ck_result *results, uint32_t *num_results, uint32_t part_index, uint32_t num_parts) {
  if (!num_results) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (max_results > 0 && !results) return fail(CK_ERR_INVALID_ARGUMENT, "results is NULL");
  *num_results = 0;
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
  const uint32_t n = sm_rows(pl->map.sm);
  const size_t words_per_sample = ref_words_per_sample(pl->num_sites);  // u64
  const size_t bytes = words_per_sample * n * 8;
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
  int rc = stream_begin_impl(pl, kin_threshold, max_results, part_index, num_parts);
  if (rc != CK_OK) return rc;
  const uint32_t num_bands = ceil_div(n, kFp4BandRows);
  const uint32_t chunk_bands = std::max<uint32_t>(1, ceil_div(num_bands, 24u));

  // the copy stream starts after everything already queued on the compute stream (the buffers come from the ctx cache)
  cudaEvent_t fork;
  CK_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  st.events.push_back(fork);
  CK_CUDA(cudaEventRecord(fork, s));
  CK_CUDA(cudaStreamWaitEvent(cs, fork, 0));
  struct Chunk { uint32_t s0, s1; cudaEvent_t ready; };
  std::vector<Chunk> chunks;
  for (uint32_t hi = num_bands; hi > 0;) {
    const uint32_t lo = hi > chunk_bands ? hi - chunk_bands : 0;
    Chunk c{lo * kFp4BandRows, std::min<uint32_t>(hi * kFp4BandRows, n), nullptr};
    CK_CUDA(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
    st.events.push_back(c.ready);
    const size_t off = size_t(c.s0) * words_per_sample;
    CK_CUDA(cudaMemcpyAsync(static_cast<uint64_t *>(st.p) + off, bit_set + off, size_t(c.s1 - c.s0) * words_per_sample * 8,
                            cudaMemcpyHostToDevice, cs));
    CK_CUDA(cudaEventRecord(c.ready, cs));
    chunks.push_back(c);
    hi = lo;
  }
  for (const Chunk &c : chunks) {
    cudaError_t e = cudaStreamWaitEvent(s, c.ready, 0);
    rc = e == cudaSuccess ? stream_rows_device(pl, static_cast<const uint64_t *>(st.p) + size_t(c.s0) * words_per_sample, c.s0, c.s1)
                          : fail_cuda(e, "cudaStreamWaitEvent", __FILE__, __LINE__);
    if (rc != CK_OK) {
      cudaStreamSynchronize(cs);
      cudaStreamSynchronize(s);
      delete pl->stream_state;
      pl->stream_state = nullptr;
      return rc;
    }
  }
  rc = stream_end_impl(pl, results, num_results);
  ctx->timings.h2d_ms = 0.f;     // overlapped: the whole upload + transpose + kernel span is reported as king_ms
  ctx->timings.import_ms = 0.f;
  return rc;
}

extern "C" {

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
  const size_t bytes = size_t(sample_end - sample_begin) * ref_words_per_sample(pl->num_sites) * 8;
  DevBuf tmp;  // host rows: staged synchronously (callers that want overlap deliver device memory)
  CK_CUDA(tmp.alloc(bytes));
  CK_CUDA(cudaMemcpyAsync(tmp.p, rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = stream_rows_device(pl, tmp.as<uint64_t>(), sample_begin, sample_end);
  CK_CUDA(cudaStreamSynchronize(ctx->stream));  // tmp dies with this frame
  return rc;
}

int ck_king_stream_end(ck_planes *pl, ck_result *results, uint32_t *num_results) {
  if (!pl || !num_results) return fail(CK_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!pl->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "no stream session is open on these planes");
  if (pl->stream_state->max_results > 0 && !results) return fail(CK_ERR_INVALID_ARGUMENT, "results is NULL");
  *num_results = 0;
  DeviceGuard guard(pl->ctx->device);
  return stream_end_impl(pl, results, num_results);
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
  if (num_parts == 0 || part_index >= num_parts) return fail(CK_ERR_INVALID_ARGUMENT, "part_index outside [0, num_parts)");
  ck_submatrix sm;
  int rc = ck_submatrix_init(num_samples, split_factor, shard_index, &sm);
  if (rc != CK_OK) return rc;
  ck_planes *pl = nullptr;
  rc = ck_planes_create(ctx, &sm, num_sites, &pl);
  if (rc != CK_OK) return rc;
  if (host_bitset_can_pipeline(pl)) {
    rc = king_host_bitset_pipelined(pl, bit_set, kin_threshold, max_results, results, num_results, part_index, num_parts);
  } else {  // small or off-diagonal shards, other kernel variants: plain upload, contiguous slice of the tile grid
    rc = ck_planes_import_bitset(pl, bit_set, 0);
    uint64_t tiles = 0;
    if (rc == CK_OK) rc = ck_king_num_tiles(pl, &tiles);
    if (rc == CK_OK)
      rc = ck_king_tiles(pl, tiles * part_index / num_parts, tiles * (part_index + 1) / num_parts, kin_threshold, max_results,
                         results, 0, num_results, 1);
  }
  ck_planes_destroy(pl);
  return rc;
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
  DevBuf counts, offsets, tmp;
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
    CK_CUDA(cudaMalloc(&ctx->syn_row, cap * 8));
    CK_CUDA(cudaMalloc(&ctx->syn_col, cap * 8));
    CK_CUDA(cudaMalloc(&ctx->syn_alt, cap * 4));
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
