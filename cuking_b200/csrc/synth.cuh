// Counter-based synthetic cohort: genotype g(seed, sample, site) is a pure function, evaluated by
// the same code on host and device so any cell can be regenerated anywhere.
//
//   site allele frequency   p_s = 0.05 + 0.45 * u(seed, AF, site)
//   founders                two haplotypes, each carries the alt allele with probability p_s (HWE)
//   pedigrees               blocks of 8 consecutive samples: members 0,1,4,6 are founders; 2 and 3 are children of
//                           (0,1); 5 is the child of (2,4); 7 is the child of (5,6); a child inherits one
//                           hash-selected haplotype from each parent, independently per site (unlinked sites)
//   missingness             genotype absent with probability m, independently
//
// All comparisons are done on 32-bit integers (no floating point), so host and device agree bit for bit.
#pragma once
#include <cstdint>

namespace ck {

enum : uint32_t { kTagAf = 1, kTagHap = 2, kTagSel = 3, kTagMiss = 4 };

__host__ __device__ inline uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
// key of (seed, tag, sample); hash_at() then folds in the site
__host__ __device__ inline uint64_t stream_key(uint64_t seed, uint32_t tag, uint32_t sample) {
  return mix64(seed ^ mix64((uint64_t(tag) << 32) | sample));
}
__host__ __device__ inline uint64_t hash_at(uint64_t key, uint32_t site) {
  return mix64(key + 0x9e3779b97f4a7c15ull * (uint64_t(site) + 1));
}

constexpr uint32_t kAfBase = 214748365u;    // round(0.05 * 2^32)
constexpr uint32_t kAfRange = 1932735283u;  // round(0.45 * 2^32)

__host__ __device__ inline uint32_t missing_threshold(double m) {
  if (!(m > 0.0)) return 0u;
  if (m >= 1.0) return 0xffffffffu;
  return uint32_t(m * 4294967296.0);
}

// Genotypes (0,1,2 or -1 = missing) of the 8 members of pedigree block `block` at `site`.
struct PedigreeKeys {
  uint64_t hap[8];   // founders: haplotype stream; children: selector stream
  uint64_t miss[8];
  uint64_t af;
};

__host__ __device__ inline PedigreeKeys pedigree_keys(uint64_t seed, uint32_t block) {
  PedigreeKeys k;
  k.af = stream_key(seed, kTagAf, 0);
  for (uint32_t m = 0; m < 8; ++m) {
    const uint32_t sample = block * 8 + m;
    const bool founder = (m == 0 || m == 1 || m == 4 || m == 6);
    k.hap[m] = stream_key(seed, founder ? kTagHap : kTagSel, sample);
    k.miss[m] = stream_key(seed, kTagMiss, sample);
  }
  return k;
}

__host__ __device__ inline void pedigree_genotypes(const PedigreeKeys &k, uint32_t site, uint32_t miss_thr,
                                                   int8_t g[8]) {
  const uint32_t p = kAfBase + uint32_t(((hash_at(k.af, site) >> 32) * uint64_t(kAfRange)) >> 32);
  uint8_t h[8][2];
  const int founders[4] = {0, 1, 4, 6};
  for (int f = 0; f < 4; ++f) {
    const int m = founders[f];
    const uint64_t x = hash_at(k.hap[m], site);
    h[m][0] = uint32_t(x) < p;
    h[m][1] = uint32_t(x >> 32) < p;
  }
  const int child[4] = {2, 3, 5, 7}, pa[4] = {0, 0, 2, 5}, pb[4] = {1, 1, 4, 6};
  for (int c = 0; c < 4; ++c) {
    const int m = child[c];
    const uint64_t x = hash_at(k.hap[m], site);
    h[m][0] = h[pa[c]][x & 1];
    h[m][1] = h[pb[c]][(x >> 1) & 1];
  }
  for (int m = 0; m < 8; ++m) {
    const bool missing = uint32_t(hash_at(k.miss[m], site) >> 32) < miss_thr;
    g[m] = missing ? int8_t(-1) : int8_t(h[m][0] + h[m][1]);
  }
}

}  // namespace ck
