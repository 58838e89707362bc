// Shard planning: which rows/columns of the relatedness matrix one shard covers, and where a sample of that shard
// lives in plane storage.  Same contract as struct Submatrix, /root/reference/cuking.cu:129-179 — row-major walk
// over the upper triangle of the split_factor x split_factor block grid, block edge ceil(n / k) — written as free
// functions over the plain-C ck_submatrix so they can cross the C ABI and run on host and device.
#pragma once
#include <cstdint>

#include "../../include/cuking_b200.h"
#include "layout.cuh"

namespace ck {

// Upper-triangular linear shard index -> block coordinates (same row-major walk as cuking.cu:136-144).
__host__ __device__ inline void shard_to_block(uint32_t k, uint32_t shard, uint32_t *block_i, uint32_t *block_j) {
  // Row b of the triangle starts at offset b*k - b*(b-1)/2 and holds k - b shards.
  uint32_t b = 0, start = 0;
  while (b + 1 < k && start + (k - b) <= shard) {
    start += k - b;
    ++b;
  }
  *block_i = b;
  *block_j = b + (shard - start);
}

__host__ __device__ inline bool make_submatrix(uint32_t n, uint32_t k, uint32_t shard, ck_submatrix *out) {
  if (k == 0) return false;                                          // cuking.cu:454-457
  const uint64_t num_shards = uint64_t(k) * (uint64_t(k) + 1) / 2;   // cuking.cu:459 (without its u32 overflow)
  if (shard >= num_shards) return false;                             // cuking.cu:459-462
  uint32_t bi, bj;
  shard_to_block(k, shard, &bi, &bj);
  const uint32_t size = ceil_div(n, k);                              // cuking.cu:147
  const uint64_t ib = uint64_t(bi) * size, jb = uint64_t(bj) * size;
  // Clamp begin as well as end so that an empty trailing block is [n, n) instead of underflowing (cuking.cu:148-151
  // yields i_end_ < i_begin for e.g. n = 5, k = 4).
  out->i_begin = uint32_t(ib < n ? ib : n);
  out->i_end = uint32_t(ib + size < n ? ib + size : n);
  out->j_begin = uint32_t(jb < n ? jb : n);
  out->j_end = uint32_t(jb + size < n ? jb + size : n);
  return true;
}

__host__ __device__ inline uint32_t sm_rows(const ck_submatrix &s) { return s.i_end - s.i_begin; }
__host__ __device__ inline uint32_t sm_cols(const ck_submatrix &s) { return s.j_end - s.j_begin; }
__host__ __device__ inline bool sm_diagonal(const ck_submatrix &s) { return s.i_begin == s.j_begin; }
__host__ __device__ inline uint32_t sm_samples(const ck_submatrix &s) {
  return sm_diagonal(s) ? sm_rows(s) : sm_rows(s) + sm_cols(s);
}
__host__ __device__ inline bool sm_contains(const ck_submatrix &s, uint32_t x) {
  return (s.i_begin <= x && x < s.i_end) || (s.j_begin <= x && x < s.j_end);
}
// Slot in the REFERENCE bit set (rows first, then columns): cuking.cu:171-175.
__host__ __device__ inline uint32_t sm_ref_offset(const ck_submatrix &s, uint32_t x) {
  return (x < s.i_end) ? (x - s.i_begin) : (s.i_end - s.i_begin + x - s.j_begin);
}

// Slot in PLANE storage: rows at [0, rows); columns of an off-diagonal shard start on the next 64-sample block
// boundary so that every pairwise tile is one whole row block x one whole column block.
struct SlotMap {
  ck_submatrix sm;
  uint32_t col_slot0;   // first slot of the column range (0 for a diagonal shard)
  uint32_t num_blocks;  // 64-sample blocks allocated
  __host__ __device__ uint32_t slot(uint32_t x) const {
    return (x >= sm.i_begin && x < sm.i_end) ? (x - sm.i_begin) : (col_slot0 + (x - sm.j_begin));
  }
  // reference slot (cuking.cu:171-175) -> plane slot
  __host__ __device__ uint32_t slot_of_ref(uint32_t ref_slot) const {
    const uint32_t rows = sm_rows(sm);
    return ref_slot < rows ? ref_slot : col_slot0 + (ref_slot - rows);
  }
};

__host__ __device__ inline SlotMap make_slot_map(const ck_submatrix &s) {
  SlotMap m;
  m.sm = s;
  const uint32_t row_blocks = ceil_div(sm_rows(s), kTileSamples);
  if (sm_diagonal(s)) {
    m.col_slot0 = 0;
    m.num_blocks = row_blocks;
  } else {
    m.col_slot0 = row_blocks * kTileSamples;
    m.num_blocks = row_blocks + ceil_div(sm_cols(s), kTileSamples);
  }
  return m;
}

}  // namespace ck
