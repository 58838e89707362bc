// Probe: tcgen05.mma.cta_group::2.kind::mxf4.block_scale with the A operand in TMEM (TS mode) - the building block of
// a CTA-pair version of the pairwise kernel (M = 256 across two SMs sharing the B operand).  Questions:
//   1. does the pair instruction work with A from each CTA's own TMEM and unit block scales, and how is B split between
//      the two CTAs' shared memories (hypothesis: rows [0, N/2) from the leader, [N/2, N) from its peer, both at the
//      shared-memory offset the leader's descriptor names);
//   2. commit multicast to both CTAs' mbarriers;
//   3. sustained rate of the KING issue pattern in pair mode.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_pair_probe tools/umma_pair_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((smem_addr >> 4) & 0x3fff) | (uint64_t((lbo_bytes >> 4) & 0x3fff) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3fff) << 32) | (uint64_t(1) << 46);
}
__host__ __device__ constexpr uint32_t make_idesc_mxf4(uint32_t M, uint32_t N) {
  return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma2_mxf4_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
      : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
// arrives on the barrier at this shared-memory offset in every CTA of the mask once the MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_multicast(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}

constexpr uint32_t kSfCol = 480, kACol = 416;

// A: [256][kbytes] packed E2M1 (row m of the pair tile; CTA r holds rows 128 r .. 128 r + 127), B: [N][kbytes],
// D: [256][N] fp32.  b_split = 0: rows [0, N/2) in CTA 0 and [N/2, N) in CTA 1; 1: every CTA holds all N rows.
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_probe_kernel(const uint8_t *A, const uint8_t *B, float *D, int kbytes, int reps,
                                                                                   int b_split, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t LBO = 128, SBO = (kbytes / 16) * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int rows_here = b_split == 0 ? N / 2 : N, row0 = b_split == 0 ? int(rank) * (N / 2) : 0;
  for (uint32_t e = tid; e < uint32_t(rows_here * kbytes); e += blockDim.x) {
    const uint32_t row = e / kbytes, kbyte = e % kbytes;
    smem[(row >> 3) * SBO + (kbyte >> 4) * LBO + (row & 7) * 16 + (kbyte & 15)] = B[size_t(row0 + row) * kbytes + kbyte];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t lane_base = tmem_base + (uint32_t(warp * 32) << 16);
  for (uint32_t c = 0; c < 32; ++c) tmem_st1(lane_base + kSfCol + c, 0x7f7f7f7fu);
  for (uint32_t c = 0; c < uint32_t(kbytes) / 4; ++c) {
    uint32_t w = 0;
    for (int b = 0; b < 4; ++b) w |= uint32_t(A[size_t(rank * 128 + tid) * kbytes + c * 4 + b]) << (8 * b);
    tmem_st1(lane_base + kACol + c, w);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();  // both CTAs' operands are in place
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  unsigned long long t0 = clock64();
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_mxf4(256, N);
    for (int rep = 0; rep < reps; ++rep)
      for (uint32_t ks = 0; ks < uint32_t(kbytes) / 32; ++ks) {
        const uint64_t db = make_smem_desc(smem_u32(smem) + ks * 2 * LBO, LBO, SBO);
        umma2_mxf4_ts(tmem_base, tmem_base + kACol + ks * 8, db, idesc, tmem_base + kSfCol, tmem_base + kSfCol + 16, (rep > 0 || ks > 0) ? 1u : 0u);
      }
    umma2_commit_multicast(&bar, 3);
  }
  mbar_wait(&bar, 0);
  unsigned long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  for (int n0 = 0; n0 < N; n0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(lane_base + n0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x < 2)
      for (int q = 0; q < 8; ++q) D[size_t(rank * 128 + tid) * N + n0 + q] = __uint_as_float(v[q]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

static int e2m1_value4(uint32_t nib) { return nib == 0x2 ? 4 : nib == 0xA ? -4 : nib == 0x1 ? 2 : 0; }

template <int N>
static int run(int kbytes, int reps, int b_split, const char *label) {
  constexpr int M = 256;
  std::vector<uint8_t> hA(size_t(M) * kbytes), hB(size_t(N) * kbytes);
  uint32_t s = 777u + N + b_split;
  auto nib = [&]() -> uint8_t {
    s = s * 1664525u + 1013904223u;
    const uint32_t q = (s >> 20) % 4;
    return q == 0 ? 0x0 : q == 1 ? 0x1 : q == 2 ? 0x2 : 0xA;
  };
  for (auto &x : hA) { uint8_t lo = nib(), hi = nib(); x = uint8_t(lo | (hi << 4)); }
  for (auto &x : hB) { uint8_t lo = nib(), hi = nib(); x = uint8_t(lo | (hi << 4)); }
  uint8_t *dA, *dB; float *dD; unsigned long long *dC;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dC, 8 * 1024));
  CK(cudaMemset(dD, 0xff, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const size_t smem = size_t(N) * kbytes;
  CK(cudaFuncSetAttribute(pair_probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  pair_probe_kernel<N><<<2, 128, smem>>>(dA, dB, dD, kbytes, reps, b_split, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"probe\": \"umma_pair\", \"test\": \"%s\", \"N\": %d, \"b_split\": %d, \"error\": \"%s\"}\n", label, N, b_split, cudaGetErrorString(e));
    return 1 << 20;
  }
  std::vector<float> hD(M * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  unsigned long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  int bad = 0, bad_half[2][2] = {{0, 0}, {0, 0}};
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      long long ref16 = 0;
      for (int k = 0; k < 2 * kbytes; ++k)
        ref16 += e2m1_value4((hA[size_t(m) * kbytes + (k >> 1)] >> (4 * (k & 1))) & 0xf) * e2m1_value4((hB[size_t(n) * kbytes + (k >> 1)] >> (4 * (k & 1))) & 0xf);
      ref16 *= reps;
      if (double(ref16) != 16.0 * double(hD[m * N + n])) {
        ++bad_half[m / 128][n / (N / 2)];
        if (bad++ < 3) printf("  %s mismatch (m=%d,n=%d): got %.2f want %.4f\n", label, m, n, hD[m * N + n], double(ref16) / 16.0);
      }
    }
  printf("{\"probe\": \"umma_pair\", \"test\": \"%s\", \"N\": %d, \"K\": %d, \"reps\": %d, \"b_split\": %d, \"mismatches\": %d, \"by_quadrant_mhalf_nhalf\": [[%d, %d], [%d, %d]], \"clk_per_mma\": %.1f}\n",
         label, N, 2 * kbytes, reps, b_split, bad, bad_half[0][0], bad_half[0][1], bad_half[1][0], bad_half[1][1], double(cyc) / (double(reps) * kbytes / 32));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad;
}

// sustained rate, KING issue pattern in pair mode: issuers 0..2 of the leader CTA (one lane of warps 0-2) own one
// accumulator each (N = n0, n1, n1); A from TMEM, B from shared memory (each CTA holds half the rows)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_rate_kernel(int n0, int n1, int issuers, int steps, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t kbytes = 128, LBO = 128, SBO = (kbytes / 16) * 128;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, issuers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t e = tid; e < 128 * kbytes / 4; e += blockDim.x) reinterpret_cast<uint32_t *>(smem)[e] = 0x22222222u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t lane_base = tmem_base + (uint32_t(warp * 32) << 16);
  for (uint32_t c = 0; c < 32; ++c) tmem_st1(lane_base + kSfCol + c, 0x7f7f7f7fu);
  for (uint32_t c = 0; c < 64; ++c) tmem_st1(lane_base + kACol + c, 0x22222222u);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  unsigned long long t0 = clock64();
  if (rank == 0 && (tid & 31) == 0 && warp < issuers) {
    const uint32_t n = warp == 0 ? n0 : n1;
    const uint32_t idesc = make_idesc_mxf4(256, n);
    const uint32_t d = tmem_base + (warp == 0 ? 0 : n0 + (warp - 1) * n1);
    for (int st = 0; st < steps; ++st) {
      const uint32_t ks = st & 3;
      const uint64_t db = make_smem_desc(smem_u32(smem) + ks * 2 * LBO, LBO, SBO);
      umma2_mxf4_ts(d, tmem_base + kACol + ks * 8 + warp * 8, db, idesc, tmem_base + kSfCol, tmem_base + kSfCol + 16, st > 0);
    }
    umma2_commit_multicast(&bar, 3);
  }
  mbar_wait(&bar, 0);
  unsigned long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

static void run_rate(int n0, int n1, int issuers) {
  unsigned long long *dC;
  CK(cudaMalloc(&dC, 8 * 1024));
  const size_t smem = 128 * 128;
  const int steps = 20000;
  pair_rate_kernel<<<148, 128, smem>>>(n0, n1, issuers, steps, dC);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  pair_rate_kernel<<<148, 128, smem>>>(n0, n1, issuers, steps, dC);
  CK(cudaEventRecord(e1));
  cudaError_t e = cudaEventSynchronize(e1);
  if (e != cudaSuccess) { printf("{\"probe\": \"umma_pair_rate\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return; }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  const double macs_per_step_per_sm = 128.0 * 64 * (n0 + double(issuers - 1) * n1);  // each SM of the pair does its 128 rows
  printf("{\"probe\": \"umma_pair_rate\", \"issuers\": %d, \"N0\": %d, \"N1\": %d, \"clk_per_step\": %.1f, \"macs_per_clk_per_sm\": %.1f, \"chip_tops\": %.1f, \"ms\": %.3f}\n",
         issuers, n0, n1, double(cyc) / steps, macs_per_step_per_sm * steps / double(cyc), 2.0 * macs_per_step_per_sm * steps * 148 / (ms * 1e-3) / 1e12, ms);
  cudaFree(dC);
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  if (mode == 0) {  // B split between the CTAs (the expected semantics)
    int bad = run<80>(64, 1, 0, "pair_ts_split");
    bad += run<160>(64, 1, 0, "pair_ts_split");
    bad += run<64>(128, 4, 0, "pair_ts_split");
    bad += run<128>(128, 4, 0, "pair_ts_split");
    printf(bad ? "PAIR PROBE: B-split hypothesis FAILED\n" : "PAIR PROBE OK (B rows [0, N/2) from the leader, [N/2, N) from the peer)\n");
  } else if (mode == 1) {  // every CTA holds all rows (alternative hypothesis)
    int bad = run<80>(64, 1, 1, "pair_ts_full");
    printf(bad ? "full-B hypothesis FAILED\n" : "full-B hypothesis OK\n");
  } else {
    run_rate(80, 160, 3);
    run_rate(64, 128, 3);
    run_rate(64, 128, 1);
    run_rate(128, 128, 3);
    run_rate(256, 256, 1);
    run_rate(208, 208, 2);
  }
  return 0;
}
