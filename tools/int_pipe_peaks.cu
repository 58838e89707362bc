// Integer-pipe peak microbenchmark for sm_100a: measures sustained issue rates (lane-ops / clk / SM) of the
// instructions the pairwise KING kernel is built from (POPC, LOP3, IADD3, IMAD) and a few mixes.
// the POPC roofline denominator must be measured, not assumed.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_pipe_peaks tools/int_pipe_peaks.cu
// Output: one JSON object per line on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

enum Op { OP_POPC = 0, OP_LOP3, OP_IADD3, OP_IMAD, OP_POPC_LOP3, OP_POPC_2LOP3, OP_POPC_IADD, OP_KING5, OP_KINGCSA, OP_SHF, OP_NUM };
static const char* kOpName[] = {"popc", "lop3", "iadd3", "imad", "popc+lop3", "popc+2lop3", "popc+iadd", "king5_plain", "king5_csa", "shf"};
// lane-ops counted per inner iteration per chain (for the headline "ops/clk/SM" figure)
static const int kOpsPerIter[] = {1, 1, 1, 1, 2, 3, 2, 0, 0, 1};

constexpr int kChains = 8;

template <int OP>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t* out, int iters, unsigned long long* cyc) {
  uint32_t a[kChains];
  uint32_t b = threadIdx.x * 2654435761u + blockIdx.x, c = b ^ 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < kChains; ++i) a[i] = b + i * 0x01010101u;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        if (OP == OP_POPC) {
          asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
        } else if (OP == OP_LOP3) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
        } else if (OP == OP_IADD3) {
          asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        } else if (OP == OP_IMAD) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        } else if (OP == OP_SHF) {
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
        } else if (OP == OP_POPC_LOP3) {
          uint32_t t;
          asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(a[i]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(t), "r"(c));
        } else if (OP == OP_POPC_2LOP3) {
          uint32_t t;
          asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(a[i]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(t), "r"(c));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(a[i]) : "r"(b), "r"(c));
        } else if (OP == OP_POPC_IADD) {
          uint32_t t;
          asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(a[i]));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(t));
        }
      }
    }
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Realistic inner loops on register operands: 2x2 pair block, 5 counts per pair-word.
// plain: 5 POPC + 6 LOP3 + adds ; csa: per 2 words: 5 POPC + 12+10 LOP3 + adds (one-level carry-save).
template <bool CSA>
__global__ void __launch_bounds__(256) king_kernel(uint32_t* out, int iters, unsigned long long* cyc) {
  uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  uint32_t acc[4][5], ones[4][5];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 5; ++c) { acc[p][c] = 0; ones[p][c] = 0; }
  uint32_t H[4], D[4], A[4], Y[4];   // [0,1] = rows, [2,3] = cols
#pragma unroll
  for (int s = 0; s < 4; ++s) { H[s] = seed * (s + 3); D[s] = ~(seed >> (s + 1)); A[s] = seed ^ (0x5bd1e995u * (s + 1)); Y[s] = D[s] & ~H[s]; }
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // fake "next word" operands: cheap register perturbation (1 op per plane-word, amortised like an LDS)
      uint32_t H2[4], D2[4], A2[4], Y2[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) { H2[s] = H[s] + 0x9e3779b9u; D2[s] = D[s] ^ H2[s]; A2[s] = A[s] + D2[s]; Y2[s] = Y[s] ^ A2[s]; }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int cidx = 0; cidx < 2; ++cidx) {
          const int p = r * 2 + cidx, i = r, j = 2 + cidx;
          if (!CSA) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const uint32_t Hi = w ? H2[i] : H[i], Hj = w ? H2[j] : H[j], Di = w ? D2[i] : D[i], Dj = w ? D2[j] : D[j];
              const uint32_t Ai = w ? A2[i] : A[i], Aj = w ? A2[j] : A[j], Yi = w ? Y2[i] : Y[i], Yj = w ? Y2[j] : Y[j];
              acc[p][0] += __popc(Hi & Hj);
              acc[p][1] += __popc(Hi & Dj);
              acc[p][2] += __popc(Di & Hj);
              acc[p][3] += __popc(Di & Dj);
              acc[p][4] += __popc((Ai ^ Aj) & Yi & Yj);
            }
          } else {
            uint32_t x0, x1, o;
#define CSA_STEP(c, e0, e1) x0 = (e0); x1 = (e1); o = ones[p][c]; ones[p][c] = o ^ x0 ^ x1; acc[p][c] += __popc((o & x0) | (o & x1) | (x0 & x1));
            CSA_STEP(0, H[i] & H[j], H2[i] & H2[j])
            CSA_STEP(1, H[i] & D[j], H2[i] & D2[j])
            CSA_STEP(2, D[i] & H[j], D2[i] & H2[j])
            CSA_STEP(3, D[i] & D[j], D2[i] & D2[j])
            CSA_STEP(4, (A[i] ^ A[j]) & Y[i] & Y[j], (A2[i] ^ A2[j]) & Y2[i] & Y2[j])
#undef CSA_STEP
          }
        }
#pragma unroll
      for (int s = 0; s < 4; ++s) { H[s] = H2[s] ^ 0x85ebca6bu; D[s] = D2[s] + 0xc2b2ae35u; A[s] = A2[s] ^ D[s]; Y[s] = Y2[s] + H[s]; }
    }
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 5; ++c) s ^= acc[p][c] * 2 + __popc(ones[p][c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, double lane_ops_per_thread, int threads, int blocks, int sms, unsigned long long* d_cyc, double pairwords_per_thread) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  std::vector<unsigned long long> cyc(blocks);
  CK(cudaMemcpy(cyc.data(), d_cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::sort(cyc.begin(), cyc.end());
  const double med_cyc = (double)cyc[blocks / 2];
  const double total_ops = lane_ops_per_thread * threads * (double)blocks;
  const int blocks_per_sm = blocks / sms;
  // per-SM rate from in-kernel cycle counter: all resident CTAs of an SM run concurrently for ~med_cyc
  const double ops_per_clk_sm = lane_ops_per_thread * threads * blocks_per_sm / med_cyc;
  const double sm_mhz = med_cyc / (best * 1e-3) / 1e6;
  printf("{\"bench\": \"%s\", \"ms\": %.4f, \"lane_ops_per_s\": %.4e, \"lane_ops_per_clk_per_sm\": %.2f, \"sm_mhz_est\": %.0f, \"pair_sites_per_s\": %.4e}\n",
         name, best, total_ops / (best * 1e-3), ops_per_clk_sm, sm_mhz, pairwords_per_thread * 32.0 * threads * blocks / (best * 1e-3));
  fflush(stdout);
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  int ctas_per_sm = argc > 1 ? atoi(argv[1]) : 4;      // 4 x 256 threads = 32 warps / SM
  const int threads = 256, blocks = sms * ctas_per_sm;
  int iters = argc > 2 ? atoi(argv[2]) : 4096;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"ctas_per_sm\": %d, \"iters\": %d}\n", prop.name, sms, prop.clockRate, ctas_per_sm, iters);
  uint32_t* d_out; unsigned long long* d_cyc;
  CK(cudaMalloc(&d_out, (size_t)threads * blocks * 4)); CK(cudaMalloc(&d_cyc, blocks * 8));
#define RUN(OP) run(kOpName[OP], [&] { pipe_kernel<OP><<<blocks, threads>>>(d_out, iters, d_cyc); }, (double)iters * 4 * kChains * kOpsPerIter[OP], threads, blocks, sms, d_cyc, 0.0)
  RUN(OP_POPC); RUN(OP_LOP3); RUN(OP_IADD3); RUN(OP_IMAD); RUN(OP_SHF); RUN(OP_POPC_LOP3); RUN(OP_POPC_2LOP3); RUN(OP_POPC_IADD);
  // king loops: per outer iter: 4 unrolled steps x 4 pairs x 2 words = 32 pair-words; POPC lane-ops: plain 5/pair-word, csa 2.5
  run("king5_plain", [&] { king_kernel<false><<<blocks, threads>>>(d_out, iters / 4, d_cyc); }, (double)(iters / 4) * 32 * 5, threads, blocks, sms, d_cyc, (double)(iters / 4) * 32);
  run("king5_csa", [&] { king_kernel<true><<<blocks, threads>>>(d_out, iters / 4, d_cyc); }, (double)(iters / 4) * 32 * 2.5, threads, blocks, sms, d_cyc, (double)(iters / 4) * 32);
  CK(cudaGetLastError());
  return 0;
}
