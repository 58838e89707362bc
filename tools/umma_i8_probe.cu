// Probe: one tcgen05.mma.kind::i8 GEMM tile (M=128, N, K) with operands in the no-swizzle K-major canonical shared
// memory layout, accumulator in TMEM, read back with tcgen05.ld — checked against a CPU integer matmul.
// Purpose: validate the instruction/smem descriptors and TMEM addressing used by the int8 tensor-core formulation of
// the pairwise KING kernel before building the full pipeline.  Also times a long K loop to estimate the sustained
// int8 MMA rate of one SM and of the whole chip.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_i8_probe tools/umma_i8_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3fff);
  d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= uint64_t(1) << 46;  // descriptor version (Blackwell)
  return d;                // layout_type 0 = no swizzle
}

__host__ __device__ constexpr uint32_t make_idesc_i8(uint32_t M, uint32_t N, bool a_signed, bool b_signed) {
  return (2u << 4) | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// canonical K-major no-swizzle layout: 8-row x 16-byte core matrices; K-adjacent core matrices LBO apart, 8-row groups SBO apart
__device__ __forceinline__ uint32_t canon_off(uint32_t row, uint32_t kbyte, uint32_t lbo, uint32_t sbo) {
  return (row >> 3) * sbo + (kbyte >> 4) * lbo + (row & 7) * 16 + (kbyte & 15);
}

template <int N, int NACC>
__global__ void __launch_bounds__(128) probe_kernel(const int8_t *A, const int8_t *B, int32_t *D, int K, int reps, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  constexpr int M = 128;
  const uint32_t kbytes = K;                       // int8: 1 byte per k
  const uint32_t LBO = 128, SBO = (kbytes / 16) * 128;  // an 8-row group holds kbytes/16 core matrices
  uint8_t *sA = smem, *sB = smem + M * kbytes;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t e = tid; e < M * kbytes; e += blockDim.x) sA[canon_off(e / kbytes, e % kbytes, LBO, SBO)] = uint8_t(A[e]);
  for (uint32_t e = tid; e < N * kbytes; e += blockDim.x) sB[canon_off(e / kbytes, e % kbytes, LBO, SBO)] = uint8_t(B[e]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  unsigned long long t0 = clock64();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_i8(M, N, true, true);
    for (int rep = 0; rep < reps; ++rep) {
      for (uint32_t ks = 0; ks < kbytes / 32; ++ks) {
        const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2 * LBO, LBO, SBO);
        const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2 * LBO, LBO, SBO);
        // NACC independent accumulators round-robin (the KING formulation has five); accumulator 0 gets every NACC-th MMA
        const uint32_t acc = (rep * (kbytes / 32) + ks) % NACC;
        umma_i8(tmem_base + acc * N, da, db, idesc, (rep > 0 || ks >= NACC) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  unsigned long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;

  // epilogue: warp w reads TMEM lanes 32w..32w+31 (rows), 8 columns at a time
  for (int n0 = 0; n0 < N; n0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + n0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0)
      for (int q = 0; q < 8; ++q) D[tid * N + n0 + q] = int32_t(v[q]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}


// ---- TS mode: A operand in TMEM (written with tcgen05.st, thread = lane = row; K byte k -> column k/4, byte k%4) ----
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

template <int N, int NACC>
__global__ void __launch_bounds__(128) probe_ts_kernel(const int8_t *A, const int8_t *B, int32_t *D, int K, int reps, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  constexpr int M = 128;
  const uint32_t kbytes = K;
  const uint32_t LBO = 128, SBO = (kbytes / 16) * 128;
  uint8_t *sB = smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t e = tid; e < N * kbytes; e += blockDim.x) sB[canon_off(e / kbytes, e % kbytes, LBO, SBO)] = uint8_t(B[e]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  // A: row tid, K bytes -> columns [a_col0, a_col0 + K/4)
  const uint32_t a_col0 = NACC * N;  // after the accumulators
  for (uint32_t c = 0; c < kbytes / 4; c += 8) {
    uint32_t v[8];
    for (int q = 0; q < 8; ++q) {
      uint32_t w = 0;
      for (int b = 0; b < 4; ++b) w |= uint32_t(uint8_t(A[tid * kbytes + (c + q) * 4 + b])) << (8 * b);
      v[q] = w;
    }
    const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + a_col0 + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  unsigned long long t0 = clock64();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_i8(M, N, true, true);
    for (int rep = 0; rep < reps; ++rep) {
      for (uint32_t ks = 0; ks < kbytes / 32; ++ks) {
        const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2 * LBO, LBO, SBO);
        const uint32_t acc = (rep * (kbytes / 32) + ks) % NACC;
        umma_i8_ts(tmem_base + acc * N, tmem_base + a_col0 + ks * 8, db, idesc, (rep > 0 || ks >= NACC) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  unsigned long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  for (int n0 = 0; n0 < N; n0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + n0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0)
      for (int q = 0; q < 8; ++q) D[tid * N + n0 + q] = int32_t(v[q]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

template <int N, int NACC>
static int run_ts(int K, int reps_time) {
  constexpr int M = 128;
  std::vector<int8_t> hA(M * K), hB(N * K);
  uint32_t s = 777u + N;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return int8_t(int((s >> 24) % 3) - 1); };
  for (auto &x : hA) x = rnd();
  for (auto &x : hB) x = rnd();
  int8_t *dA, *dB; int32_t *dD; unsigned long long *dC;
  int dev_sms = 148;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dC, 8 * 1024));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const size_t smem = size_t(N) * K + 1024;
  CK(cudaFuncSetAttribute(probe_ts_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  probe_ts_kernel<N, NACC><<<1, 128, smem>>>(dA, dB, dD, K, 1, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD(M * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  if (NACC == 1)
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        int32_t ref = 0;
        for (int k = 0; k < K; ++k) ref += int32_t(hA[m * K + k]) * int32_t(hB[n * K + k]);
        if (ref != hD[m * N + n] && bad++ < 5) printf("  TS mismatch N=%d (m=%d,n=%d): got %d want %d\n", N, m, n, hD[m * N + n], ref);
      }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe_ts_kernel<N, NACC><<<dev_sms, 128, smem>>>(dA, dB, dD, K, reps_time, dC);
  CK(cudaEventRecord(e0));
  probe_ts_kernel<N, NACC><<<dev_sms, 128, smem>>>(dA, dB, dD, K, reps_time, dC);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  const double macs = double(M) * N * K * reps_time;
  printf("{\"probe\": \"umma_i8_ts\", \"accumulators\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"mismatches\": %d, \"macs_per_clk_per_sm\": %.1f, \"clk_per_mma\": %.1f, \"chip_int8_tops\": %.1f}\n",
         NACC, M, N, K, bad, macs / double(cyc), double(cyc) / (double(reps_time) * K / 32), 2.0 * macs * dev_sms / (ms * 1e-3) / 1e12);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad;
}

// ---- two issuing threads (warps 0 and 1), each with its own accumulator: is the ~100 clk/MMA floor per issuer or per SM? ----
template <int N>
__global__ void __launch_bounds__(128) probe_dual_kernel(const int8_t *A, const int8_t *B, int K, int reps, unsigned long long *cycles, int issuers, int commit_each) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t dummy[4];
  __shared__ uint32_t tmem_base_smem;
  constexpr int M = 128;
  const uint32_t kbytes = K;
  const uint32_t LBO = 128, SBO = (kbytes / 16) * 128;
  uint8_t *sA = smem, *sB = smem + M * kbytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, issuers);
    for (int q = 0; q < 4; ++q) mbar_init(&dummy[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t e = tid; e < M * kbytes; e += blockDim.x) sA[canon_off(e / kbytes, e % kbytes, LBO, SBO)] = uint8_t(A[e]);
  for (uint32_t e = tid; e < N * kbytes; e += blockDim.x) sB[canon_off(e / kbytes, e % kbytes, LBO, SBO)] = uint8_t(B[e]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  unsigned long long t0 = clock64();
  if ((tid & 31) == 0 && warp < issuers) {
    const uint32_t idesc = make_idesc_i8(M, N, true, true);
    for (int rep = 0; rep < reps; ++rep)
      for (uint32_t ks = 0; ks < kbytes / 32; ++ks) {
        const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2 * LBO, LBO, SBO);
        const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2 * LBO, LBO, SBO);
        umma_i8(tmem_base + warp * N, da, db, idesc, (rep > 0 || ks > 0) ? 1u : 0u);
        if (commit_each) umma_commit(&dummy[warp]);
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  unsigned long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

template <int N>
static void run_dual(int K, int reps, int issuers, int commit_each = 0) {
  constexpr int M = 128;
  int8_t *dA, *dB; unsigned long long *dC;
  CK(cudaMalloc(&dA, M * K)); CK(cudaMalloc(&dB, N * K)); CK(cudaMalloc(&dC, 8 * 1024));
  CK(cudaMemset(dA, 1, M * K)); CK(cudaMemset(dB, 1, N * K));
  const size_t smem = size_t(M + N) * K;
  CK(cudaFuncSetAttribute(probe_dual_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  probe_dual_kernel<N><<<148, 128, smem>>>(dA, dB, K, reps, dC, issuers, commit_each);
  probe_dual_kernel<N><<<148, 128, smem>>>(dA, dB, K, reps, dC, issuers, commit_each);
  CK(cudaDeviceSynchronize());
  unsigned long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  printf("{\"probe\": \"umma_i8_dual\", \"commit_each\": %d, \"issuers\": %d, \"N\": %d, \"clk_per_mma_aggregate\": %.1f, \"macs_per_clk_per_sm\": %.1f}\n", commit_each, issuers, N,
         double(cyc) / (double(reps) * (K / 32) * issuers), double(M) * N * K * reps * issuers / double(cyc));
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

template <int N, int NACC>
static int run(int K, int reps_time) {
  constexpr int M = 128;
  std::vector<int8_t> hA(M * K), hB(N * K);
  uint32_t s = 12345u + N;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return int8_t(int((s >> 24) % 3) - 1); };
  for (auto &x : hA) x = rnd();
  for (auto &x : hB) x = rnd();
  int8_t *dA, *dB; int32_t *dD; unsigned long long *dC;
  int dev_sms = 148;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dC, 8 * 1024));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const size_t smem = size_t(M + N) * K;
  CK(cudaFuncSetAttribute(probe_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  probe_kernel<N, NACC><<<1, 128, smem>>>(dA, dB, dD, K, 1, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD(M * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  if (NACC == 1)
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += int32_t(hA[m * K + k]) * int32_t(hB[n * K + k]);
      if (ref != hD[m * N + n] && bad++ < 5) printf("  mismatch N=%d (m=%d,n=%d): got %d want %d\n", N, m, n, hD[m * N + n], ref);
    }
  // timing: whole chip, one CTA per SM, reps_time x K/32 MMAs each
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe_kernel<N, NACC><<<dev_sms, 128, smem>>>(dA, dB, dD, K, reps_time, dC);
  CK(cudaEventRecord(e0));
  probe_kernel<N, NACC><<<dev_sms, 128, smem>>>(dA, dB, dD, K, reps_time, dC);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  const double macs = double(M) * N * K * reps_time;
  printf("{\"probe\": \"umma_i8\", \"accumulators\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"mismatches\": %d, \"macs_per_clk_per_sm\": %.1f, \"chip_int8_tops\": %.1f, \"ms\": %.3f}\n",
         NACC, M, N, K, bad, macs / double(cyc), 2.0 * macs * dev_sms / (ms * 1e-3) / 1e12, ms);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad;
}

int main() {
  int bad = 0;
  bad += run<64, 1>(128, 2000);
  bad += run<96, 1>(128, 2000);
  bad += run<128, 1>(128, 2000);
  bad += run<192, 1>(128, 2000);
  bad += run<256, 1>(128, 2000);
  // independent accumulators (correctness is checked by the single-accumulator runs above)
  run<64, 2>(128, 2000);
  run<64, 5>(128, 2000);
  run<96, 2>(128, 2000);
  run<96, 5>(128, 2000);
  run<128, 4>(128, 2000);
  run<32, 5>(128, 2000);
  run_dual<96>(128, 2000, 1);
  run_dual<96>(128, 2000, 2);
  run_dual<96>(128, 2000, 4);
  run_dual<32>(128, 2000, 4);
  run_dual<96>(128, 2000, 1, 1);
  run_dual<96>(128, 2000, 2, 1);
  run_dual<96>(128, 2000, 3, 1);
  run_dual<160>(128, 2000, 3, 1);
  run_dual<160>(128, 2000, 3, 0);
  // A operand from TMEM
  bad += run_ts<64, 1>(128, 2000);
  bad += run_ts<80, 1>(128, 2000);
  bad += run_ts<96, 1>(128, 2000);
  bad += run_ts<160, 1>(128, 2000);
  bad += run_ts<192, 1>(128, 2000);
  run_ts<80, 5>(128, 2000);
  run_ts<96, 4>(128, 2000);
  run_ts<160, 2>(128, 2000);
  run_ts<32, 5>(128, 2000);
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad ? 1 : 0;
}
