// ck_planes_and_reduce: AND-all-reduce of the raw genotype planes of several GPUs over NVLink peer memory.
//
// The pack step (cuking.cu:675-703) only ever CLEARS bits of an all-ones bit set, so it distributes over AND: if the
// triples are dealt to G GPUs - each packing its share into a full-size, all-missing plane set - the cohort's planes
// are the bitwise AND of the G partial plane sets.  That turns "every GPU packs every triple" (G x the PCIe traffic and
// G x the atomics) into "every triple is packed once" plus one exchange step, which is this file.
//
// The exchange is ONE kernel per GPU, reduce-scatter and all-gather fused: GPU g owns the g-th slice of the words,
// loads that slice from every peer (P2P loads over NVLink / NVSwitch), ANDs, and stores the result into every peer's
// copy (P2P stores).  Slices are disjoint, so the G kernels run concurrently without ordering among themselves; events
// order them after every GPU's packs and order every GPU's later work after all of them.  Per GPU: (G-1)/G of the plane
// bytes in and out over NVLink, nothing through the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "internal.cuh"

namespace ck {

namespace {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  uint4 *p[kMaxPeers];
};

__global__ void __launch_bounds__(512) and_reduce_kernel(PeerPtrs bufs, const uint4 *mine, int count, int self, size_t begin, size_t end) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = begin + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < end; i += stride) {
    uint4 v = mine[i];
    // all remote loads of a batch are in flight together (NVLink latency is microseconds); constant indices keep the
    // pointer table in the parameter bank
#pragma unroll
    for (int g0 = 0; g0 < kMaxPeers; g0 += 8) {
      if (g0 >= count) break;
      uint4 w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        w[q] = (g0 + q < count && g0 + q != self) ? __ldcs(bufs.p[g0 + q] + i) : make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        v.x &= w[q].x; v.y &= w[q].y; v.z &= w[q].z; v.w &= w[q].w;
      }
    }
#pragma unroll
    for (int g = 0; g < kMaxPeers; ++g)
      if (g < count) bufs.p[g][i] = v;
  }
}

}  // namespace
}  // namespace ck

using namespace ck;

extern "C" int ck_planes_and_reduce(ck_planes *const *planes, uint32_t count) {
  if (!planes || count == 0) return fail(CK_ERR_INVALID_ARGUMENT, "no planes");
  if (count > uint32_t(kMaxPeers)) return fail(CK_ERR_INVALID_ARGUMENT, "at most 16 GPUs per reduction");
  for (uint32_t g = 0; g < count; ++g) {
    if (!planes[g]) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
    if (planes[g]->raw_bytes != planes[0]->raw_bytes || planes[g]->num_sites != planes[0]->num_sites ||
        planes[g]->map.num_blocks != planes[0]->map.num_blocks)
      return fail(CK_ERR_INVALID_ARGUMENT, "planes of different shapes cannot be reduced");
    if (planes[g]->stream_state) return fail(CK_ERR_INVALID_ARGUMENT, "a stream session is open on these planes");
    for (uint32_t h = 0; h < g; ++h)
      if (planes[h]->ctx->device == planes[g]->ctx->device)
        return fail(CK_ERR_INVALID_ARGUMENT, "ck_planes_and_reduce needs one plane set per GPU");
  }
  if (count == 1) return CK_OK;
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  struct Restore {
    int dev;
    ~Restore() {
      if (dev >= 0) cudaSetDevice(dev);
    }
  } restore{prev_device};

  // peer access in both directions between every pair
  for (uint32_t g = 0; g < count; ++g) {
    CK_CUDA(cudaSetDevice(planes[g]->ctx->device));
    for (uint32_t h = 0; h < count; ++h) {
      if (h == g) continue;
      int can = 0;
      CK_CUDA(cudaDeviceCanAccessPeer(&can, planes[g]->ctx->device, planes[h]->ctx->device));
      if (!can)
        return fail(CK_ERR_CUDA, "GPU " + std::to_string(planes[g]->ctx->device) + " cannot access the memory of GPU " +
                                     std::to_string(planes[h]->ctx->device) + " (no peer access): planes cannot be reduced over NVLink");
      const cudaError_t e = cudaDeviceEnablePeerAccess(planes[h]->ctx->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
      else CK_CUDA(e);
    }
  }
  PeerPtrs bufs{};
  for (uint32_t g = 0; g < count; ++g) bufs.p[g] = reinterpret_cast<uint4 *>(planes[g]->raw);
  const size_t n4 = planes[0]->raw_words() / 4;  // plane rows are 64 words: a multiple of 4
  std::vector<cudaEvent_t> packed(count, nullptr), reduced(count, nullptr);
  struct Events {
    std::vector<cudaEvent_t> *a, *b;
    ~Events() {
      for (auto *v : {a, b})
        for (cudaEvent_t e : *v)
          if (e) cudaEventDestroy(e);
    }
  } events{&packed, &reduced};
  for (uint32_t g = 0; g < count; ++g) {  // everything already queued on a GPU (its packs) precedes the exchange
    CK_CUDA(cudaSetDevice(planes[g]->ctx->device));
    CK_CUDA(cudaEventCreateWithFlags(&packed[g], cudaEventDisableTiming));
    CK_CUDA(cudaEventCreateWithFlags(&reduced[g], cudaEventDisableTiming));
    CK_CUDA(cudaEventRecord(packed[g], planes[g]->ctx->stream));
  }
  for (uint32_t g = 0; g < count; ++g) {
    ck_ctx *ctx = planes[g]->ctx;
    CK_CUDA(cudaSetDevice(ctx->device));
    for (uint32_t h = 0; h < count; ++h)
      if (h != g) CK_CUDA(cudaStreamWaitEvent(ctx->stream, packed[h], 0));
    const size_t begin = n4 * g / count, end = n4 * (g + 1) / count;
    if (end > begin) {
      const unsigned grid = unsigned(std::min<size_t>((end - begin + 511) / 512, size_t(ctx->num_sms) * 4));
      and_reduce_kernel<<<grid, 512, 0, ctx->stream>>>(bufs, bufs.p[g], int(count), int(g), begin, end);
      CK_CUDA(cudaGetLastError());
    }
    CK_CUDA(cudaEventRecord(reduced[g], ctx->stream));
  }
  for (uint32_t g = 0; g < count; ++g) {  // nobody touches its planes before every slice has landed in them
    ck_ctx *ctx = planes[g]->ctx;
    CK_CUDA(cudaSetDevice(ctx->device));
    for (uint32_t h = 0; h < count; ++h)
      if (h != g) CK_CUDA(cudaStreamWaitEvent(ctx->stream, reduced[h], 0));
    planes[g]->mark_stale();
  }
  for (uint32_t g = 0; g < count; ++g) {
    CK_CUDA(cudaSetDevice(planes[g]->ctx->device));
    CK_CUDA(cudaStreamSynchronize(planes[g]->ctx->stream));
  }
  return CK_OK;
}
