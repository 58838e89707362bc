// Plane construction kernels: pack (sparse triples -> raw planes), finalize (raw -> compute planes), exchange
// with the reference bit-set layout, and the on-device synthetic cohort.
//
// Reference behaviour restated here: /root/reference/cuking.cu:519-523 (all-ones = missing), :675-703 (per-triple
// bit clears, AND-accumulation, Contains() filter, int64->int32 truncation), :204-212 + :507-513 (bit-set layout).
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.cuh"
#include "synth.cuh"

namespace ck {

namespace {

// ---- fill -----------------------------------------------------------------------------------------------------
__global__ void fill_missing_kernel(uint4 *raw4, size_t n4) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride)
    raw4[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);  // cuking.cu:520-523
}

// ---- pack -----------------------------------------------------------------------------------------------------
// One triple per lane per load.  For Hail-ordered input (site-major, sample-minor) the 32 lanes of a warp hit 32
// consecutive words of one (block, word, plane) row, i.e. each RED.AND warp instruction touches one 128-byte line.
__device__ __forceinline__ void pack_one(uint32_t *raw, const SlotMap &map, uint32_t words, uint32_t num_sites,
                                         int64_t row64, int64_t col64, int32_t n_alt, size_t index, unsigned long long *err) {
  const uint32_t col = uint32_t(int32_t(col64));                     // cuking.cu:676
  if (!sm_contains(map.sm, col)) return;                             // cuking.cu:677-679
  const uint32_t site = uint32_t(int32_t(row64));                    // cuking.cu:680
  const unsigned long long tag = (unsigned long long)index + 1ull;  // 64-bit: the 'no error' sentinel ~0 is never a real tag
  if (uint32_t(n_alt) > 2u) {                                        // cuking.cu:698-701
    atomicMin(&err[0], tag);
    return;
  }
  if (site >= num_sites) {  // the reference writes out of bounds here; we refuse
    atomicMin(&err[1], tag);
    return;
  }
  const uint32_t slot = map.slot(col);
  uint32_t *het = raw + plane_index(slot / kTileSamples, site >> 5, 0, slot % kTileSamples, words, kRawPlanes);
  const uint32_t mask = ~(1u << (site & 31u));
  if (n_alt != 1) atomicAnd(het, mask);                   // 0 and 2 clear the het bit      (cuking.cu:689, :696)
  if (n_alt != 2) atomicAnd(het + kTileSamples, mask);    // 0 and 1 clear the hom-alt bit  (cuking.cu:690, :693)
}

// RowT / ColT / AltT: int64 / int64 / int32 as decoded from the reference's Parquet columns (cuking.cu:603-672), or the narrow
// form uint32 / uint32 / uint8 (the same values after the int32 truncation of cuking.cu:676,:680; 9 instead of 20 bytes per
// triple over PCIe when the kernel reads page-locked host memory in place)
template <typename RowT, typename ColT, typename AltT>
__global__ void __launch_bounds__(256) pack_kernel(uint32_t *raw, SlotMap map, uint32_t words, uint32_t num_sites,
                                                   const RowT *__restrict__ row, const ColT *__restrict__ col,
                                                   const AltT *__restrict__ alt, size_t n, size_t index_base,
                                                   unsigned long long *err) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  // One triple per lane per load group, kUnroll independent groups (20 B each) in flight per thread before the first
  // atomic: the kernel is latency-bound (78 % long-scoreboard stalls with one group per iteration,
  // profiles/r01_pack_kernel_ncu.txt).  Lane-contiguous triples keep both sides coalesced: a warp's loads are whole
  // 256 / 256 / 128-byte runs, and for Hail-ordered input its RED.AND hits 32 consecutive words = one 128-byte line
  // (two triples per lane - 16-byte loads - made every RED touch two lines at half sector efficiency).
  constexpr int kUnroll = 8;
  size_t i = tid;
  for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
    RowT r[kUnroll];
    ColT c[kUnroll];
    AltT a[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      r[u] = __ldcs(row + i + u * stride);  // streaming: every triple is read exactly once
      c[u] = __ldcs(col + i + u * stride);
      a[u] = __ldcs(alt + i + u * stride);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      pack_one(raw, map, words, num_sites, int64_t(r[u]), int64_t(c[u]), int32_t(a[u]), index_base + i + u * stride, err);
  }
  for (; i < n; i += stride)
    pack_one(raw, map, words, num_sites, int64_t(__ldcs(row + i)), int64_t(__ldcs(col + i)), int32_t(__ldcs(alt + i)), index_base + i, err);
}

// ---- decode + pack: Parquet page payloads -> raw planes (ck_pack_encoded, page_decode.cu) ---------------------------
// One thread decodes kDecodeRows consecutive rows of the window: per column one binary search over the run table for its
// first value, then a walk.  RLE runs cost nothing, bit-packed dictionary indices one funnel shift over two aligned words
// of the payload buffer, PLAIN values one load.  The decoded triples go through pack_one like every other triple.
constexpr int kDecodeRows = 8;

__device__ __forceinline__ uint32_t decode_rows(const EncodedColumnDev &c, uint32_t row0, uint32_t n, int64_t (&out)[kDecodeRows]) {
  const uint32_t v0 = c.skip + row0;
  uint32_t lo = 0, hi = c.num_runs;  // runs[lo].first_value <= v0 < runs[hi].first_value (runs[num_runs] is the sentinel)
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(&c.runs[mid].first_value) <= v0) lo = mid; else hi = mid;
  }
  uint4 run = __ldg(reinterpret_cast<const uint4 *>(c.runs + lo));  // first_value, kind, bit_width, payload
  uint32_t next = __ldg(&c.runs[lo + 1].first_value);
  uint32_t bad = 0xffffffffu;  // first row of this thread whose dictionary index is out of range
#pragma unroll
  for (int k = 0; k < kDecodeRows; ++k) {
    out[k] = 0;
    if (uint32_t(k) < n) {
      const uint32_t v = v0 + k;
      while (v >= next) {
        ++lo;
        run = __ldg(reinterpret_cast<const uint4 *>(c.runs + lo));
        next = __ldg(&c.runs[lo + 1].first_value);
      }
      const uint32_t rel = v - run.x;
      if (run.y == CK_RUN_PLAIN) {
        const uint8_t *p = reinterpret_cast<const uint8_t *>(c.words) + run.w;
        out[k] = c.width == 8 ? __ldg(reinterpret_cast<const long long *>(p) + rel) : (long long)__ldg(reinterpret_cast<const int *>(p) + rel);
      } else {
        uint32_t idx = run.w;
        if (run.y == CK_RUN_BITPACKED) {
          const unsigned long long bit = (unsigned long long)run.w * 8ull + (unsigned long long)rel * run.z;
          const size_t w = size_t(bit >> 5);
          idx = __funnelshift_r(__ldg(c.words + w), __ldg(c.words + w + 1), uint32_t(bit) & 31u);  // the buffer is padded by 8 bytes
          if (run.z < 32u) idx &= (1u << run.z) - 1u;
        }
        if (idx >= c.dict_len) {
          if (bad == 0xffffffffu) bad = uint32_t(k);
        } else {
          out[k] = c.width == 8 ? __ldg(reinterpret_cast<const long long *>(c.dict) + idx) : (long long)__ldg(reinterpret_cast<const int *>(c.dict) + idx);
        }
      }
    }
  }
  return bad;
}

__global__ void __launch_bounds__(256) decode_pack_kernel(uint32_t *raw, SlotMap map, uint32_t words, uint32_t num_sites,
                                                          EncodedColumnDev c_row, EncodedColumnDev c_col, EncodedColumnDev c_alt,
                                                          uint32_t num_rows, unsigned long long *err) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t groups = (size_t(num_rows) + kDecodeRows - 1) / kDecodeRows;
  for (size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; t < groups; t += stride) {
    const uint32_t row0 = uint32_t(t * kDecodeRows), n = min(uint32_t(kDecodeRows), num_rows - row0);
    int64_t r[kDecodeRows], c[kDecodeRows], a[kDecodeRows];
    const uint32_t bad = min(min(decode_rows(c_row, row0, n, r), decode_rows(c_col, row0, n, c)), decode_rows(c_alt, row0, n, a));
    if (bad != 0xffffffffu) atomicMin(&err[2], (unsigned long long)(row0 + bad) + 1ull);
#pragma unroll
    for (int k = 0; k < kDecodeRows; ++k)
      if (uint32_t(k) < n && uint32_t(k) < bad)  // nothing is packed from a corrupt row on (the call fails anyway)
        pack_one(raw, map, words, num_sites, r[k], c[k], int32_t(a[k]), size_t(row0) + k, err);
  }
}

// ---- finalize: raw (het, alt) -> compute (H, D, A) --------------------------------------------------------------
// One thread per 4 lanes (16 bytes) of one (block, word) row; pure streaming, 8 B read + 12 B written per sample-word.
__global__ void __launch_bounds__(256) finalize_kernel(const uint4 *__restrict__ raw4, uint4 *__restrict__ out4,
                                                       size_t num_rows /* blocks * words */) {
  constexpr uint32_t kVec = kTileSamples / 4;  // 16 uint4 per plane row
  const size_t total = num_rows * kVec;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / kVec, v = i % kVec;
    const uint4 het = raw4[(r * kRawPlanes + 0) * kVec + v];
    const uint4 alt = raw4[(r * kRawPlanes + 1) * kVec + v];
    uint4 H, D, A;
    H.x = het.x & ~alt.x; H.y = het.y & ~alt.y; H.z = het.z & ~alt.z; H.w = het.w & ~alt.w;
    D.x = ~(het.x & alt.x); D.y = ~(het.y & alt.y); D.z = ~(het.z & alt.z); D.w = ~(het.w & alt.w);  // cuking.cu:229
    A.x = alt.x & ~het.x; A.y = alt.y & ~het.y; A.z = alt.z & ~het.z; A.w = alt.w & ~het.w;
    out4[(r * kComputePlanes + kPlaneH) * kVec + v] = H;
    out4[(r * kComputePlanes + kPlaneD) * kVec + v] = D;
    out4[(r * kComputePlanes + kPlaneA) * kVec + v] = A;
  }
}

// ---- finalize for the tensor-core kernel: raw (het, alt) -> nibble-coded genotypes ---------------------------------
// spread8: bit i of the low byte -> bit 4i
__device__ __forceinline__ uint32_t spread8(uint32_t b) {
  b &= 0xffu;
  b = (b | (b << 12)) & 0x000f000fu;
  b = (b | (b << 6)) & 0x03030303u;
  b = (b | (b << 3)) & 0x11111111u;
  return b;
}
// kFp4 = false: bit 0 het, bit 1 hom-alt, bit 2 hom-ref (PRMT selectors of the int8 kernel);
// kFp4 = true:  the nibble IS the E2M1 operand superposition of the mxf4 kernel: 1 het (0.5), 2 hom-alt (+1), 0xA hom-ref (-1)
template <bool kFp4>
__global__ void __launch_bounds__(256) finalize_codes_kernel(const uint32_t *__restrict__ raw, uint4 *__restrict__ codes,
                                                             size_t num_rows /* blocks * words */) {
  const size_t total = num_rows * kTileSamples;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / kTileSamples, lane = i % kTileSamples;
    const uint32_t het = raw[(r * kRawPlanes + 0) * kTileSamples + lane];
    const uint32_t alt = raw[(r * kRawPlanes + 1) * kTileSamples + lane];
    const uint32_t H = het & ~alt, A = alt & ~het, R = ~het & ~alt;  // (1,1) = missing -> code 0
    uint32_t z[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      z[t] = kFp4 ? spread8(H >> (8 * t)) | (spread8((A | R) >> (8 * t)) << 1) | (spread8(R >> (8 * t)) << 3)
                  : spread8(H >> (8 * t)) | (spread8(A >> (8 * t)) << 1) | (spread8(R >> (8 * t)) << 2);
    codes[i] = make_uint4(z[0], z[1], z[2], z[3]);
  }
}

// (het, hom) site counts of every plane slot over all its sites, from the E2M1 nibble codes (het = 0x1, hom = 0x2 | sign):
// what the screen kernels bound kinship with (king_screen_kernel.cu, king_screen1_kernel.cu).  One CTA per 64-sample block,
// four word-strided partial sums per lane; the block's sums also go into the two cohort-wide counters.
__global__ void __launch_bounds__(256) sample_totals_kernel(const uint4 *__restrict__ codes, uint2 *__restrict__ totals,
                                                            unsigned long long *__restrict__ sums, uint32_t words) {
  __shared__ uint32_t partial[2][4][kTileSamples];
  const uint32_t lane = threadIdx.x % kTileSamples, part = threadIdx.x / kTileSamples;
  const uint4 *src = codes + size_t(blockIdx.x) * words * kTileSamples + lane;
  uint32_t het = 0, hom = 0;
  for (uint32_t w = part; w < words; w += 4) {
    const uint4 z = __ldg(src + size_t(w) * kTileSamples);
    het += __popc(z.x & 0x11111111u) + __popc(z.y & 0x11111111u) + __popc(z.z & 0x11111111u) + __popc(z.w & 0x11111111u);
    hom += __popc(z.x & 0x22222222u) + __popc(z.y & 0x22222222u) + __popc(z.z & 0x22222222u) + __popc(z.w & 0x22222222u);
  }
  partial[0][part][lane] = het;
  partial[1][part][lane] = hom;
  __syncthreads();
  if (part == 0) {
    het = partial[0][0][lane] + partial[0][1][lane] + partial[0][2][lane] + partial[0][3][lane];
    hom = partial[1][0][lane] + partial[1][1][lane] + partial[1][2][lane] + partial[1][3][lane];
    totals[size_t(blockIdx.x) * kTileSamples + lane] = make_uint2(het, hom);
    if (sums != nullptr) {
      atomicAdd(&sums[0], (unsigned long long)het);
      atomicAdd(&sums[1], (unsigned long long)hom);
    }
  }
}

// ---- reference layout <-> raw planes ----------------------------------------------------------------------------
// The reference bit set is sample-major (cuking.cu:204-212): slot o, plane p, 64-bit word q at
// bit_set[o*W + p*W/2 + q]; as little-endian uint32 the 32-site word k sits at index 2*(o*W + p*W/2) + k.
// A CTA transposes a 64-sample x 32-word tile of one plane through shared memory so that both sides are coalesced.
template <bool kImport>
__global__ void __launch_bounds__(256) ref_transpose_kernel(uint32_t *raw, uint32_t *ref32, SlotMap map, uint32_t words,
                                                            uint32_t ref_words_u64, uint32_t num_ref_slots, uint32_t block0,
                                                            uint32_t ref_slot0 /* slot of the first row behind ref32 */) {
  __shared__ uint32_t tile[kTileSamples][33];
  const uint32_t block = block0 + blockIdx.y, plane = blockIdx.z;
  const uint32_t k0 = blockIdx.x * 32;
  const uint32_t ref_k = ref_words_u64;  // uint32 words per plane in the reference layout: 2 * (W/2)
  const uint32_t rows = sm_rows(map.sm);
  // reference slot of plane-slot s (inverse of SlotMap::slot_of_ref); 0xffffffff = padding lane
  auto ref_slot_of = [&](uint32_t s) -> uint32_t {
    if (s < rows) return s;
    if (map.col_slot0 != 0 && s >= map.col_slot0 && s - map.col_slot0 + rows < num_ref_slots) return s - map.col_slot0 + rows;
    return 0xffffffffu;
  };
  if (kImport) {
    for (uint32_t e = threadIdx.x; e < kTileSamples * 32; e += blockDim.x) {
      const uint32_t lane = e / 32, kk = e % 32;
      const uint32_t o = ref_slot_of(block * kTileSamples + lane);
      uint32_t v = 0xffffffffu;  // padding lanes / padding words are missing
      if (o != 0xffffffffu && k0 + kk < ref_k)
        v = ref32[(size_t(o - ref_slot0) * 2 + plane) * ref_k + k0 + kk];
      tile[lane][kk] = v;
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < kTileSamples * 32; e += blockDim.x) {
      const uint32_t kk = e / kTileSamples, lane = e % kTileSamples;
      if (k0 + kk < words) raw[plane_index(block, k0 + kk, plane, lane, words, kRawPlanes)] = tile[lane][kk];
    }
  } else {
    for (uint32_t e = threadIdx.x; e < kTileSamples * 32; e += blockDim.x) {
      const uint32_t kk = e / kTileSamples, lane = e % kTileSamples;
      tile[lane][kk] = (k0 + kk < words) ? raw[plane_index(block, k0 + kk, plane, lane, words, kRawPlanes)] : 0xffffffffu;
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < kTileSamples * 32; e += blockDim.x) {
      const uint32_t lane = e / 32, kk = e % 32;
      const uint32_t o = ref_slot_of(block * kTileSamples + lane);
      if (o != 0xffffffffu && k0 + kk < ref_k) ref32[(size_t(o - ref_slot0) * 2 + plane) * ref_k + k0 + kk] = tile[lane][kk];
    }
  }
}

// ---- synthetic cohort -------------------------------------------------------------------------------------------
// One thread per (pedigree block of 8 samples, 32-site word): evaluates 32 x 8 genotypes and writes the raw planes
// in the reference encoding, exactly what packing the same cohort's triples would leave behind.
__global__ void __launch_bounds__(128) synth_planes_kernel(uint32_t *raw, SlotMap map, uint32_t words, uint32_t num_sites,
                                                           uint64_t seed, uint32_t miss_thr, uint32_t first_group,
                                                           uint32_t num_groups, uint32_t range_begin, uint32_t range_end) {
  const uint32_t num_words = ceil_div(num_sites, 32u);
  const size_t total = size_t(num_groups) * num_words;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    // consecutive threads -> consecutive groups of the same word: writes of a warp cover 256 consecutive lanes
    const uint32_t word = uint32_t(t / num_groups), group = first_group + uint32_t(t % num_groups);
    const PedigreeKeys keys = pedigree_keys(seed, group);
    uint32_t het[8], alt[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) het[m] = alt[m] = 0xffffffffu;
    const uint32_t site0 = word * 32;
    const uint32_t nbits = min(32u, num_sites - site0);
    for (uint32_t b = 0; b < nbits; ++b) {
      int8_t g[8];
      pedigree_genotypes(keys, site0 + b, miss_thr, g);
      const uint32_t clear = ~(1u << b);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        if (g[m] == 0 || g[m] == 2) het[m] &= clear;
        if (g[m] == 0 || g[m] == 1) alt[m] &= clear;
      }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const uint32_t sample = group * 8 + m;
      if (sample < range_begin || sample >= range_end) continue;
      const uint32_t slot = map.slot(sample);
      const size_t idx = plane_index(slot / kTileSamples, word, 0, slot % kTileSamples, words, kRawPlanes);
      raw[idx] = het[m];
      raw[idx + kTileSamples] = alt[m];
    }
  }
}

// Triples in Hail order: one warp per site walks the samples 32 at a time.
__global__ void __launch_bounds__(256) synth_count_kernel(uint64_t seed, uint32_t miss_thr, uint32_t sample_begin,
                                                          uint32_t sample_end, uint32_t site_begin, uint32_t site_end,
                                                          unsigned long long *site_counts) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
  const uint32_t num_warps = gridDim.x * blockDim.x / 32;
  for (uint32_t site = site_begin + warp; site < site_end; site += num_warps) {
    unsigned long long count = 0;
    for (uint32_t s0 = sample_begin; s0 < sample_end; s0 += 32) {
      const uint32_t s = s0 + lane;
      bool present = false;
      if (s < sample_end) present = !(uint32_t(hash_at(stream_key(seed, kTagMiss, s), site) >> 32) < miss_thr);
      count += __popc(__ballot_sync(0xffffffffu, present));
    }
    if (lane == 0) site_counts[site - site_begin] = count;
  }
}

__global__ void __launch_bounds__(256) synth_emit_kernel(uint64_t seed, uint32_t miss_thr, uint32_t sample_begin,
                                                         uint32_t sample_end, uint32_t site_begin, uint32_t site_end,
                                                         const unsigned long long *site_offsets, int64_t *row,
                                                         int64_t *col, int32_t *alt) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
  const uint32_t num_warps = gridDim.x * blockDim.x / 32;
  for (uint32_t site = site_begin + warp; site < site_end; site += num_warps) {
    unsigned long long base = site_offsets[site - site_begin];
    for (uint32_t s0 = sample_begin; s0 < sample_end; s0 += 32) {
      const uint32_t s = s0 + lane;
      int8_t mine = -1;
      if (s < sample_end) {
        int8_t g[8];
        pedigree_genotypes(pedigree_keys(seed, s / 8), site, miss_thr, g);
        mine = g[s % 8];
      }
      const uint32_t ballot = __ballot_sync(0xffffffffu, mine >= 0);
      if (mine >= 0) {
        const unsigned long long o = base + __popc(ballot & ((1u << lane) - 1u));
        row[o] = site;
        col[o] = s;
        alt[o] = mine;
      }
      base += __popc(ballot);
    }
  }
}

inline unsigned grid_for(size_t work_items, unsigned threads, unsigned max_blocks = 148 * 32) {
  size_t b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  return unsigned(b < max_blocks ? b : max_blocks);
}

}  // namespace

cudaError_t launch_fill_missing(uint32_t *raw, size_t num_words, cudaStream_t s) {
  const size_t n4 = num_words / 4;  // plane rows are 64 words: always a multiple of 4
  fill_missing_kernel<<<grid_for(n4, 256), 256, 0, s>>>(reinterpret_cast<uint4 *>(raw), n4);
  return cudaGetLastError();
}

cudaError_t launch_pack(const ck_planes &pl, const int64_t *row, const int64_t *col, const int32_t *alt, size_t n,
                        size_t index_base, unsigned long long *d_err, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  pack_kernel<<<grid_for(n / 8 + 1, 256, 148 * 8), 256, 0, s>>>(pl.raw, pl.map, pl.words, pl.num_sites, row, col, alt, n,
                                                              index_base, d_err);
  return cudaGetLastError();
}

cudaError_t launch_pack_narrow(const ck_planes &pl, const uint32_t *row, const uint32_t *col, const uint8_t *alt, size_t n,
                               size_t index_base, unsigned long long *d_err, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  pack_kernel<<<grid_for(n / 8 + 1, 256, 148 * 8), 256, 0, s>>>(pl.raw, pl.map, pl.words, pl.num_sites, row, col, alt, n,
                                                              index_base, d_err);
  return cudaGetLastError();
}

cudaError_t launch_decode_pack(const ck_planes &pl, const EncodedColumnDev (&cols)[3], uint32_t num_rows, unsigned long long *d_err,
                               cudaStream_t s) {
  if (num_rows == 0) return cudaSuccess;
  decode_pack_kernel<<<grid_for((size_t(num_rows) + kDecodeRows - 1) / kDecodeRows, 256), 256, 0, s>>>(
      pl.raw, pl.map, pl.words, pl.num_sites, cols[0], cols[1], cols[2], num_rows, d_err);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const ck_planes &pl, cudaStream_t s) {
  const size_t rows = size_t(pl.map.num_blocks) * pl.words;
  finalize_kernel<<<grid_for(rows * (kTileSamples / 4), 256), 256, 0, s>>>(
      reinterpret_cast<const uint4 *>(pl.raw), reinterpret_cast<uint4 *>(pl.compute), rows);
  return cudaGetLastError();
}

cudaError_t launch_finalize_codes(const ck_planes &pl, int kind, cudaStream_t s) {
  return launch_finalize_codes_range(pl, kind, 0, pl.map.num_blocks, s, /*add_to_sums=*/true);
}
cudaError_t launch_finalize_codes_range(const ck_planes &pl, int kind, uint32_t block0, uint32_t num_blocks, cudaStream_t s, bool add_to_sums) {
  if (num_blocks == 0) return cudaSuccess;
  const size_t rows = size_t(num_blocks) * pl.words, row0 = size_t(block0) * pl.words;
  const uint32_t *raw = pl.raw + row0 * kRawPlanes * kTileSamples;
  uint4 *codes = reinterpret_cast<uint4 *>(pl.codes) + row0 * kTileSamples;
  if (kind == 3) {
    finalize_codes_kernel<true><<<grid_for(rows * kTileSamples, 256), 256, 0, s>>>(raw, codes, rows);
    // the sums over the slots are only wanted where the caller has zeroed them and knows which samples they cover
    sample_totals_kernel<<<num_blocks, 256, 0, s>>>(codes, pl.sample_totals() + size_t(block0) * kTileSamples,
                                                    add_to_sums ? pl.totals_sums() : nullptr, pl.words);
  }
  else
    finalize_codes_kernel<false><<<grid_for(rows * kTileSamples, 256), 256, 0, s>>>(raw, codes, rows);
  return cudaGetLastError();
}

cudaError_t launch_import_ref_range(const ck_planes &pl, const uint64_t *d_rows, uint32_t ref_slot0, uint32_t block0,
                                    uint32_t num_blocks, cudaStream_t s) {
  if (num_blocks == 0) return cudaSuccess;
  const uint32_t ref_k = ref_words_per_sample(pl.num_sites);  // u64 words per sample == u32 words per plane
  dim3 grid(ceil_div(pl.words, 32u), num_blocks, kRawPlanes);
  ref_transpose_kernel<true><<<grid, 256, 0, s>>>(pl.raw, const_cast<uint32_t *>(reinterpret_cast<const uint32_t *>(d_rows)),
                                                 pl.map, pl.words, ref_k, sm_samples(pl.map.sm), block0, ref_slot0);
  return cudaGetLastError();
}
cudaError_t launch_import_ref(const ck_planes &pl, const uint64_t *d_bit_set, cudaStream_t s) {
  return launch_import_ref_range(pl, d_bit_set, 0, 0, pl.map.num_blocks, s);
}

cudaError_t launch_export_ref(const ck_planes &pl, uint64_t *d_bit_set, cudaStream_t s) {
  const uint32_t ref_k = ref_words_per_sample(pl.num_sites);
  dim3 grid(ceil_div(pl.words, 32u), pl.map.num_blocks, kRawPlanes);
  ref_transpose_kernel<false><<<grid, 256, 0, s>>>(pl.raw, reinterpret_cast<uint32_t *>(d_bit_set), pl.map, pl.words,
                                                  ref_k, sm_samples(pl.map.sm), 0, 0);
  return cudaGetLastError();
}

cudaError_t launch_synth_planes(const ck_planes &pl, uint64_t seed, uint32_t miss_thr, cudaStream_t s) {
  const ck_submatrix &sm = pl.map.sm;
  const uint32_t ranges[2][2] = {{sm.i_begin, sm.i_end}, {sm.j_begin, sm.j_end}};
  const int num_ranges = sm_diagonal(sm) ? 1 : 2;
  for (int r = 0; r < num_ranges; ++r) {
    const uint32_t b = ranges[r][0], e = ranges[r][1];
    if (b >= e) continue;
    const uint32_t first_group = b / 8, num_groups = (e - 1) / 8 - first_group + 1;
    const size_t total = size_t(num_groups) * ceil_div(pl.num_sites, 32u);
    synth_planes_kernel<<<grid_for(total, 128, 148 * 64), 128, 0, s>>>(pl.raw, pl.map, pl.words, pl.num_sites, seed,
                                                                      miss_thr, first_group, num_groups, b, e);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
  }
  return cudaSuccess;
}

cudaError_t launch_synth_count(uint64_t seed, uint32_t miss_thr, uint32_t sample_begin, uint32_t sample_end,
                               uint32_t site_begin, uint32_t site_end, unsigned long long *d_site_counts,
                               cudaStream_t s) {
  const uint32_t sites = site_end - site_begin;
  synth_count_kernel<<<grid_for(size_t(sites) * 32, 256), 256, 0, s>>>(seed, miss_thr, sample_begin, sample_end,
                                                                        site_begin, site_end, d_site_counts);
  return cudaGetLastError();
}

cudaError_t launch_synth_emit(uint64_t seed, uint32_t miss_thr, uint32_t sample_begin, uint32_t sample_end,
                              uint32_t site_begin, uint32_t site_end, const unsigned long long *d_site_offsets,
                              int64_t *row, int64_t *col, int32_t *alt, cudaStream_t s) {
  const uint32_t sites = site_end - site_begin;
  synth_emit_kernel<<<grid_for(size_t(sites) * 32, 256), 256, 0, s>>>(seed, miss_thr, sample_begin, sample_end,
                                                                       site_begin, site_end, d_site_offsets, row, col, alt);
  return cudaGetLastError();
}

}  // namespace ck
