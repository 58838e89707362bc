// Harness around the reference's own ComputeKingKernel (sliced from /root/reference/cuking.cu:100-314 into
// oracle/_ref/cuking_ref_extract.cuh by oracle/build_ref.sh).  TEST / BASELINE INFRASTRUCTURE ONLY.
//
// It launches the unmodified reference kernel with the reference's launch geometry (cuking.cu:734-741) on a bit
// set in the reference layout, either (mode 0) from plain device memory or (mode 1) as shipped: cudaMallocManaged
// memory first touched on the host (cuking.cu:113, :523, :719-722), so the timed launch includes page migration.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <type_traits>

#include "cuking_ref_extract.cuh"

namespace {
constexpr uint32_t kWarpSize = 32;  // cuking.cu:498
}

#define REF_CK(x)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "ref_harness: %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

// host_bit_sets: reference layout for this shard (words_per_sample x Submatrix::NumSamples() u64).
// results: caller-allocated host buffer of max_results 24-byte records.  Returns 0 on success.
extern "C" int ref_king(uint32_t num_samples, uint32_t split_factor, uint32_t shard_index, uint32_t words_per_sample,
                        const uint64_t *host_bit_sets, float kin_threshold, uint32_t max_results, void *results_out,
                        uint32_t *result_count, uint32_t *result_overflow, float *kernel_ms, int mode) {
  static_assert(sizeof(KingResult) == 24, "KingResult layout");
  const Submatrix submatrix(num_samples, split_factor, shard_index);
  const size_t bit_set_size = size_t(words_per_sample) * submatrix.NumSamples();
  uint64_t *bit_set = nullptr;
  KingResult *results = nullptr;
  uint32_t *index_and_flag = nullptr;
  if (mode == 1) {
    REF_CK(cudaMallocManaged(&bit_set, bit_set_size * sizeof(uint64_t)));
    REF_CK(cudaMallocManaged(&results, sizeof(KingResult) * max_results));
    REF_CK(cudaMallocManaged(&index_and_flag, 2 * sizeof(uint32_t)));
    memcpy(bit_set, host_bit_sets, bit_set_size * sizeof(uint64_t));  // host first touch
    memset(results, 0, sizeof(KingResult) * max_results);             // cuking.cu:719
    index_and_flag[0] = index_and_flag[1] = 0;                        // cuking.cu:722
  } else {
    REF_CK(cudaMalloc(&bit_set, bit_set_size * sizeof(uint64_t)));
    REF_CK(cudaMalloc(&results, sizeof(KingResult) * max_results));
    REF_CK(cudaMalloc(&index_and_flag, 2 * sizeof(uint32_t)));
    REF_CK(cudaMemcpy(bit_set, host_bit_sets, bit_set_size * sizeof(uint64_t), cudaMemcpyHostToDevice));
    REF_CK(cudaMemset(results, 0, sizeof(KingResult) * max_results));
    REF_CK(cudaMemset(index_and_flag, 0, 2 * sizeof(uint32_t)));
  }
  const uint32_t num_rows = submatrix.NumRows();
  const uint32_t num_cols = submatrix.NumCols();
  const dim3 num_blocks(num_rows, CeilIntDiv(num_cols, kMaxBlocksYZ), std::min(num_cols, kMaxBlocksYZ));  // :734-735
  constexpr uint32_t kNumBlockThreads = 4 * kWarpSize;                                                    // :737
  cudaEvent_t e0, e1;
  REF_CK(cudaEventCreate(&e0));
  REF_CK(cudaEventCreate(&e1));
  REF_CK(cudaEventRecord(e0));
  ComputeKingKernel<<<num_blocks, kNumBlockThreads>>>(submatrix, words_per_sample, bit_set, kin_threshold,
                                                      max_results, results, &index_and_flag[0], &index_and_flag[1]);
  REF_CK(cudaEventRecord(e1));
  REF_CK(cudaGetLastError());
  REF_CK(cudaDeviceSynchronize());  // cuking.cu:744
  REF_CK(cudaEventElapsedTime(kernel_ms, e0, e1));
  uint32_t host_flags[2];
  REF_CK(cudaMemcpy(host_flags, index_and_flag, sizeof(host_flags), cudaMemcpyDefault));
  *result_count = host_flags[0];
  *result_overflow = host_flags[1];
  const uint32_t n = std::min(host_flags[0], max_results);
  REF_CK(cudaMemcpy(results_out, results, sizeof(KingResult) * n, cudaMemcpyDefault));
  // cuking.cu:761-765
  KingResult *out = static_cast<KingResult *>(results_out);
  std::sort(out, out + n, [](const KingResult &lhs, const KingResult &rhs) {
    return std::tie(lhs.sample_i, lhs.sample_j, lhs.kin) < std::tie(rhs.sample_i, rhs.sample_j, rhs.kin);
  });
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(bit_set);
  cudaFree(results);
  cudaFree(index_and_flag);
  return 0;
}

// Rectangle timing helper for bench.py: rows [0, num_rows) x cols of an OFF-DIAGONAL shard built from device
// planes already resident (dev_bit_sets in the reference layout, rows stored first then cols).  Times only the
// kernel; results are discarded (max_results sized by the caller).
extern "C" int ref_king_device_resident(uint32_t num_samples, uint32_t split_factor, uint32_t shard_index,
                                        uint32_t words_per_sample, const uint64_t *dev_bit_sets, float kin_threshold,
                                        uint32_t max_results, void *dev_results, uint32_t *dev_index_and_flag,
                                        float *kernel_ms) {
  const Submatrix submatrix(num_samples, split_factor, shard_index);
  const uint32_t num_rows = submatrix.NumRows();
  const uint32_t num_cols = submatrix.NumCols();
  const dim3 num_blocks(num_rows, CeilIntDiv(num_cols, kMaxBlocksYZ), std::min(num_cols, kMaxBlocksYZ));
  constexpr uint32_t kNumBlockThreads = 4 * kWarpSize;
  cudaEvent_t e0, e1;
  REF_CK(cudaEventCreate(&e0));
  REF_CK(cudaEventCreate(&e1));
  REF_CK(cudaMemset(dev_index_and_flag, 0, 2 * sizeof(uint32_t)));
  REF_CK(cudaEventRecord(e0));
  ComputeKingKernel<<<num_blocks, kNumBlockThreads>>>(submatrix, words_per_sample, dev_bit_sets, kin_threshold,
                                                      max_results, static_cast<KingResult *>(dev_results),
                                                      &dev_index_and_flag[0], &dev_index_and_flag[1]);
  REF_CK(cudaEventRecord(e1));
  REF_CK(cudaGetLastError());
  REF_CK(cudaDeviceSynchronize());
  REF_CK(cudaEventElapsedTime(kernel_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}
