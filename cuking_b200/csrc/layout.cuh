// Device data layout of the genotype bit planes (see DESIGN.md "Data layout in HBM").
//
// The reference keeps one sample-major uint64 bit set (cuking.cu:507-523): per sample a het plane then a hom-alt
// plane, so every sample pair streams 2 x 25 KB with no reuse (cuking.cu:218-240).  Here samples are grouped in
// blocks of kTileSamples = 64 and the 32-bit site words of a block are interleaved across its samples:
//
//     plane_word(block b, word k, plane p, lane s)  =  base[ ((b * Wp + k) * P + p) * 64 + s ]      (uint32)
//
// so that (1) the K-chunk [k, k+kChunkWords) of a block is ONE contiguous 64*P*4*kChunkWords-byte run that a
// single bulk-async copy (TMA, cp.async.bulk) stages into shared memory, (2) inside shared memory the four
// consecutive samples a thread owns are one conflict-free 16-byte LDS.128, and (3) the pack kernel's atomics for
// Hail-ordered triples (site-major, sample-minor) land on consecutive 32-bit words.
//
// Two buffers use this layout:
//   raw planes     P = 2: (het, hom-alt) in the reference encoding — (1,1) = missing — filled by AND-accumulation
//                  exactly like cuking.cu:675-703;
//   compute planes P = 3: H = het & ~alt (true het), D = ~(het & alt) (defined), A = alt & ~het (true hom-alt),
//                  derived once per pack by finalize_planes_kernel so the pair loop needs no per-sample decode.
//   genotype codes (tcgen05 kernel only): 4 bits per genotype — bit 0 het, bit 1 hom-alt, bit 2 hom-ref, 0 = missing —
//                  eight sites per uint32 (site 8t+u of a 32-site word in nibble u of code word t), laid out
//                  codes[((b * Wp + k) * 64 + s) * 4 + t] so that one thread fetches a sample's 32 sites with one
//                  16-byte load and a warp's loads are contiguous.  A nibble is directly a PRMT byte selector, which
//                  turns 4 genotypes into 4 int8 operand bytes in one instruction (king_umma_kernel.cu).
#pragma once
#include <cstddef>
#include <cstdint>

namespace ck {

constexpr uint32_t kTileSamples = 64;  // samples per block == pairwise tile edge
constexpr uint32_t kChunkWords = 16;   // 32-bit site words per pipeline stage (512 sites)
constexpr uint32_t kRawPlanes = 2;
constexpr uint32_t kComputePlanes = 3;
constexpr uint32_t kPlaneH = 0, kPlaneD = 1, kPlaneA = 2;

__host__ __device__ inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
__host__ __device__ inline uint32_t padded_sites(uint32_t num_sites) { return ceil_div(num_sites, 32u) * 32u; }  // cuking.cu:498-500
__host__ __device__ inline uint32_t ref_words_per_sample(uint32_t num_sites) {                                   // cuking.cu:513
  return 2u * ceil_div(padded_sites(num_sites), 64u);
}
// 32-bit words per plane after padding to whole pipeline chunks; the padding words are "missing".
__host__ __device__ inline uint32_t padded_words(uint32_t num_sites) {
  return ceil_div(ceil_div(num_sites, 32u), kChunkWords) * kChunkWords;
}
__host__ __device__ inline size_t plane_index(uint32_t block, uint32_t word, uint32_t plane, uint32_t lane,
                                              uint32_t padded_words_, uint32_t num_planes) {
  return ((size_t(block) * padded_words_ + word) * num_planes + plane) * kTileSamples + lane;
}

}  // namespace ck
