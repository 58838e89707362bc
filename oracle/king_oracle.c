/*
 * king_oracle.c — CPU restatement of cuKING's pairwise-KING hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA product in cuking_b200/.  Nothing in the product path may
 * import, link, call or execute it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it (as the checker / the timed CPU baseline, never as the thing shipped).
 *
 * PARITY PINNING: the reference repository (populationgenomics/cuKING) contains NO tests, golden vectors or
 * fixtures — "parity unpinned" by the reference's own tests.  The pins used instead are
 *   (1) the hand-derived known-answer vectors KAT1-5 of SURVEY.md §4 (tests/golden/kat_vectors.json),
 *   (2) three-way agreement on a B200 between this file, the reference's own ComputeKingKernel compiled from
 *       /root/reference/cuking.cu:100-314 into oracle/_ref/ (see oracle/build_ref.sh), and the product kernel.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- shard planning: cuking.cu:129-179 (struct Submatrix) ------------------------------------------------- */

typedef struct {
  uint32_t i_begin, i_end; /* sample row range    (cuking.cu:177) */
  uint32_t j_begin, j_end; /* sample column range (cuking.cu:178) */
} ko_submatrix;

static inline uint32_t ceil_div_u32(uint32_t a, uint32_t b) { return (a + b - 1) / b; } /* cuking.cu:123-126 */

/* cuking.cu:130-152.  Row-major walk over the upper triangle (diagonal included) of the k x k block grid. */
void ko_submatrix_init(ko_submatrix *sm, uint32_t num_samples, uint32_t split_factor, uint32_t shard_index) {
  uint32_t tri_sum = 0, block_i = 0, block_j = 0;
  for (uint32_t i = 0; i < split_factor; ++i) {
    tri_sum += split_factor - i;
    if (shard_index < tri_sum) {
      block_i = i;
      block_j = split_factor - tri_sum + shard_index;
      break;
    }
  }
  const uint32_t size = ceil_div_u32(num_samples, split_factor);
  sm->i_begin = block_i * size;
  sm->i_end = (sm->i_begin + size < num_samples) ? sm->i_begin + size : num_samples;
  sm->j_begin = block_j * size;
  sm->j_end = (sm->j_begin + size < num_samples) ? sm->j_begin + size : num_samples;
}

uint32_t ko_num_rows(const ko_submatrix *sm) { return sm->i_end - sm->i_begin; }  /* cuking.cu:154 */
uint32_t ko_num_cols(const ko_submatrix *sm) { return sm->j_end - sm->j_begin; }  /* cuking.cu:156 */
uint32_t ko_num_samples(const ko_submatrix *sm) {                                 /* cuking.cu:159-162 */
  return (sm->i_begin == sm->j_begin) ? ko_num_rows(sm) : ko_num_rows(sm) + ko_num_cols(sm);
}
uint32_t ko_contains(const ko_submatrix *sm, uint32_t index) {                    /* cuking.cu:165-168 */
  return (sm->i_begin <= index && index < sm->i_end) || (sm->j_begin <= index && index < sm->j_end);
}
uint32_t ko_sample_offset(const ko_submatrix *sm, uint32_t index) {               /* cuking.cu:171-175 */
  return (index < sm->i_end) ? (index - sm->i_begin) : (sm->i_end - sm->i_begin + index - sm->j_begin);
}

/* ---- bit-set geometry: cuking.cu:496-513 ------------------------------------------------------------------ */

uint32_t ko_padded_sites(uint32_t num_sites) { return ceil_div_u32(num_sites, 32u) * 32u; }       /* :498-500 */
uint32_t ko_words_per_sample(uint32_t padded_sites) { return 2u * ceil_div_u32(padded_sites, 64u); } /* :513 */

/* ---- pack: cuking.cu:519-523 (all-ones = missing) and :675-703 (clear bits per triple) -------------------- */

void ko_bitset_init(uint64_t *bit_set, size_t num_words) { memset(bit_set, 0xFF, num_words * sizeof(uint64_t)); }

static inline void clear_bit(uint64_t *bit_set, uint64_t index) { /* cuking.cu:317-323 (AtomicClearBit) */
  uint64_t *ptr = bit_set + (index >> 6);
  __atomic_and_fetch(ptr, ~((uint64_t)1 << (index & 63)), __ATOMIC_RELAXED);
}

/* Returns -1 on success, else the index of the first triple whose n_alt_alleles is not 0/1/2
 * (cuking.cu:698-701 turns that into FailedPrecondition).  int64 -> int32 truncation as cuking.cu:676,680. */
int64_t ko_pack(uint64_t *bit_set, uint32_t words_per_sample, const ko_submatrix *sm, const int64_t *row_idx,
                const int64_t *col_idx, const int32_t *n_alt_alleles, size_t num_triples) {
  for (size_t row = 0; row < num_triples; ++row) {
    const int32_t col = (int32_t)col_idx[row];
    if (!ko_contains(sm, (uint32_t)col)) continue;                                    /* :677-679 */
    const int32_t site = (int32_t)row_idx[row];
    uint64_t *is_het = bit_set + (uint64_t)ko_sample_offset(sm, (uint32_t)col) * words_per_sample; /* :683-685 */
    uint64_t *is_hom_var = is_het + words_per_sample / 2;                             /* :686 */
    switch (n_alt_alleles[row]) {
      case 0: clear_bit(is_het, (uint64_t)site); clear_bit(is_hom_var, (uint64_t)site); break; /* :688-691 */
      case 1: clear_bit(is_hom_var, (uint64_t)site); break;                           /* :692-694 */
      case 2: clear_bit(is_het, (uint64_t)site); break;                               /* :695-697 */
      default: return (int64_t)row;                                                   /* :698-701 */
    }
  }
  return -1;
}

/* ---- the pairwise kernel: cuking.cu:191-314 --------------------------------------------------------------- */

typedef struct {            /* cuking.cu:182-186 */
  uint32_t sample_i, sample_j;
  float kin;
  uint32_t ibs0, ibs1, ibs2;
} ko_result;

typedef struct {
  uint32_t het_i, het_j, both_het, opposing_hom, concordant_hom, shared_sites;
} ko_counts;

/* The hot loop, cuking.cu:216-240, for one pair given storage slots (SampleOffset values). */
static inline ko_counts pair_counts(const uint64_t *bit_sets, uint32_t words_per_sample, uint32_t slot_i,
                                    uint32_t slot_j) {
  const uint32_t num_entries = words_per_sample / 2;                                  /* :204 */
  const uint64_t *het_i_e = bit_sets + (uint64_t)slot_i * words_per_sample;           /* :205-209 */
  const uint64_t *alt_i_e = het_i_e + num_entries;                                    /* :210 */
  const uint64_t *het_j_e = bit_sets + (uint64_t)slot_j * words_per_sample;           /* :211 */
  const uint64_t *alt_j_e = het_j_e + num_entries;                                    /* :212 */
  ko_counts c = {0, 0, 0, 0, 0, 0};
  for (uint32_t k = 0; k < num_entries; ++k) {
    const uint64_t het_i = het_i_e[k], alt_i = alt_i_e[k];
    const uint64_t ref_i = (~het_i) & (~alt_i);                                       /* :221 */
    const uint64_t het_j = het_j_e[k], alt_j = alt_j_e[k];
    const uint64_t ref_j = (~het_j) & (~alt_j);                                       /* :225 */
    const uint64_t defined = ~(het_i & alt_i) & ~(het_j & alt_j);                     /* :229 */
    c.het_i += (uint32_t)__builtin_popcountll(het_i & defined);                       /* :232 */
    c.het_j += (uint32_t)__builtin_popcountll(het_j & defined);                       /* :233 */
    c.both_het += (uint32_t)__builtin_popcountll(het_i & het_j & defined);            /* :234 */
    c.opposing_hom += (uint32_t)__builtin_popcountll(((ref_i & alt_j) | (alt_i & ref_j)) & defined);  /* :235-236 */
    c.concordant_hom += (uint32_t)__builtin_popcountll(((ref_i & ref_j) | (alt_i & alt_j)) & defined); /* :237-238 */
    c.shared_sites += (uint32_t)__builtin_popcountll(defined);                        /* :239 */
  }
  return c;
}

/* cuking.cu:286-294: "between-family" estimator, fp32, in the reference's expression order.
 * volatile temporaries pin the evaluation order / rounding against the host compiler (no contraction, no
 * double-precision intermediates).  All integer->float conversions are exact for counts < 2^24. */
static inline float kinship_fp32(const ko_counts *c) {
  const uint32_t min_hets = c->het_i < c->het_j ? c->het_i : c->het_j;                /* :286-287 */
  volatile float a = 2.f * (float)c->both_het;
  volatile float b = 4.f * (float)c->opposing_hom;
  volatile float num = a - b;
  num = num - (float)c->het_i;
  num = num - (float)c->het_j;
  volatile float den = 4.f * (float)min_hets;
  volatile float q = num / den;                                                       /* IEEE RN division */
  return 0.5f + q;
}

void ko_pair_counts(const uint64_t *bit_sets, uint32_t words_per_sample, uint32_t slot_i, uint32_t slot_j,
                    ko_counts *out, float *kin) {
  *out = pair_counts(bit_sets, words_per_sample, slot_i, slot_j);
  *kin = kinship_fp32(out);
}

/* Whole kernel launch, cuking.cu:191-314 + launch geometry :734-741: every (i, j) with i in rows, j in cols,
 * i < j (global indices).  Result order is unspecified in the reference (atomic slot reservation, :299);
 * ko_sort restores the order the reference writes (:761-765).  result_index keeps counting past max_results
 * like the reference's atomicAdd (:299); overflow is raised instead of writing (:308-311). */
void ko_king(const ko_submatrix *sm, uint32_t words_per_sample, const uint64_t *bit_sets, float kin_threshold,
             uint32_t max_results, ko_result *results, uint32_t *result_index, uint32_t *result_overflow) {
  const int64_t i_begin = sm->i_begin, i_end = sm->i_end;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t ii = i_begin; ii < i_end; ++ii) {
    const uint32_t i = (uint32_t)ii;
    for (uint32_t j = sm->j_begin; j < sm->j_end; ++j) {
      if (i >= j) continue;                                                           /* :199 (Contains(j) holds) */
      ko_counts c = pair_counts(bit_sets, words_per_sample, ko_sample_offset(sm, i), ko_sample_offset(sm, j));
      const float kin = kinship_fp32(&c);
      if (kin > kin_threshold) {                                                      /* :297, strict; NaN fails */
        uint32_t reserved;
#pragma omp atomic capture
        { reserved = *result_index; *result_index += 1u; }                            /* :299 */
        if (reserved < max_results) {                                                 /* :300-307 */
          ko_result *r = &results[reserved];
          r->sample_i = i;
          r->sample_j = j;
          r->kin = kin;
          r->ibs0 = c.opposing_hom;
          r->ibs2 = c.concordant_hom + c.both_het;
          r->ibs1 = c.shared_sites - r->ibs0 - r->ibs2;
        } else {
#pragma omp atomic write
          *result_overflow = 1u;                                                      /* :308-311 */
        }
      }
    }
  }
}

/* cuking.cu:761-765: sort by (sample_i, sample_j, kin). */
static int cmp_result(const void *pa, const void *pb) {
  const ko_result *a = (const ko_result *)pa, *b = (const ko_result *)pb;
  if (a->sample_i != b->sample_i) return a->sample_i < b->sample_i ? -1 : 1;
  if (a->sample_j != b->sample_j) return a->sample_j < b->sample_j ? -1 : 1;
  if (a->kin < b->kin) return -1;
  if (a->kin > b->kin) return 1;
  return 0;
}
void ko_sort(ko_result *results, size_t n) { qsort(results, n, sizeof(ko_result), cmp_result); }

/* Timed-baseline helper for bench.py: evaluate the rows x cols rectangle of storage slots with the reference
 * loop and return a checksum so the work cannot be elided.  Same arithmetic as ko_king (counts + kinship +
 * threshold test), no result buffer.  Threads = omp_get_max_threads(). */
uint64_t ko_bench_rect(const uint64_t *bit_sets, uint32_t words_per_sample, uint32_t row_slot0, uint32_t num_rows,
                       uint32_t col_slot0, uint32_t num_cols, float kin_threshold) {
  uint64_t checksum = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : checksum)
  for (int64_t r = 0; r < (int64_t)num_rows; ++r) {
    for (uint32_t c = 0; c < num_cols; ++c) {
      ko_counts k = pair_counts(bit_sets, words_per_sample, row_slot0 + (uint32_t)r, col_slot0 + c);
      const float kin = kinship_fp32(&k);
      checksum += (uint64_t)k.shared_sites + k.opposing_hom + (kin > kin_threshold ? 1u : 0u);
    }
  }
  return checksum;
}

int ko_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ko_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- the synthetic cohort of SURVEY.md section 8d, for the CPU baseline's inputs -------------------------------------
 * Restatement in plain C of the bench generator (cuking_b200/csrc/synth.cuh - the repo's own definition, not reference
 * code) so that bench.py's --impl reference arm can build its sample of the workload without loading the product
 * library: splitmix64-keyed streams, per-site allele frequency 0.05 + 0.45 u, HWE founders, pedigrees in blocks of 8
 * samples (members 0, 1, 4, 6 founders; 2, 3 children of (0, 1); 5 child of (2, 4); 7 child of (5, 6)), independent
 * missingness.  tests/test_oracle.py pins it against ck_synth_genotypes_host cell by cell. */
static uint64_t syn_mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
static uint64_t syn_key(uint64_t seed, uint32_t tag, uint32_t sample) {
  return syn_mix64(seed ^ syn_mix64(((uint64_t)tag << 32) | sample));
}
static uint64_t syn_at(uint64_t key, uint32_t site) { return syn_mix64(key + 0x9e3779b97f4a7c15ull * ((uint64_t)site + 1)); }

static void syn_block(uint64_t seed, uint32_t block, uint32_t site, uint32_t miss_thr, int8_t g[8]) {
  enum { TAG_AF = 1, TAG_HAP = 2, TAG_SEL = 3, TAG_MISS = 4 };
  static const int founder[8] = {1, 1, 0, 0, 1, 0, 1, 0};
  static const int pa[8] = {0, 0, 0, 0, 0, 2, 0, 5}, pb[8] = {0, 0, 1, 1, 0, 4, 0, 6};
  static const int order[8] = {0, 1, 4, 6, 2, 3, 5, 7}; /* parents before children */
  const uint32_t p = 214748365u + (uint32_t)(((syn_at(syn_key(seed, TAG_AF, 0), site) >> 32) * 1932735283ull) >> 32);
  uint8_t h[8][2];
  for (int q = 0; q < 8; ++q) {
    const int m = order[q];
    const uint64_t x = syn_at(syn_key(seed, founder[m] ? TAG_HAP : TAG_SEL, block * 8 + (uint32_t)m), site);
    if (founder[m]) {
      h[m][0] = (uint32_t)x < p;
      h[m][1] = (uint32_t)(x >> 32) < p;
    } else {
      h[m][0] = h[pa[m]][x & 1];
      h[m][1] = h[pb[m]][(x >> 1) & 1];
    }
  }
  for (int m = 0; m < 8; ++m) {
    const int missing = (uint32_t)(syn_at(syn_key(seed, TAG_MISS, block * 8 + (uint32_t)m), site) >> 32) < miss_thr;
    g[m] = missing ? (int8_t)-1 : (int8_t)(h[m][0] + h[m][1]);
  }
}

static uint32_t syn_missing_threshold(double m) {
  if (!(m > 0.0)) return 0u;
  if (m >= 1.0) return 0xffffffffu;
  return (uint32_t)(m * 4294967296.0);
}

/* dense genotypes out[(s - sample_begin) * sites + (r - site_begin)] in {0, 1, 2} or -1 */
void ko_synth_genotypes(uint64_t seed, double missing_rate, uint32_t sample_begin, uint32_t sample_end, uint32_t site_begin,
                        uint32_t site_end, int8_t *out) {
  if (sample_end <= sample_begin || site_end <= site_begin) return;
  const uint32_t thr = syn_missing_threshold(missing_rate);
  const size_t sites = site_end - site_begin;
#pragma omp parallel for schedule(static)
  for (int64_t block = sample_begin / 8; block <= (int64_t)((sample_end - 1) / 8); ++block)
    for (uint32_t site = site_begin; site < site_end; ++site) {
      int8_t g[8];
      syn_block(seed, (uint32_t)block, site, thr, g);
      for (uint32_t m = 0; m < 8; ++m) {
        const uint32_t s = (uint32_t)block * 8 + m;
        if (s >= sample_begin && s < sample_end) out[(size_t)(s - sample_begin) * sites + (site - site_begin)] = g[m];
      }
    }
}

/* the same cohort straight into a reference-layout bit set (cuking.cu:507-523) for samples [sample_begin, sample_end) at
 * slots 0 .. n-1: bit_sets must hold n * words_per_sample words; padding sites stay missing */
void ko_synth_bitset(uint64_t seed, double missing_rate, uint32_t sample_begin, uint32_t sample_end, uint32_t num_sites,
                     uint64_t *bit_sets) {
  if (sample_end <= sample_begin) return;
  const uint32_t thr = syn_missing_threshold(missing_rate);
  const uint32_t wps = ko_words_per_sample(ko_padded_sites(num_sites)), half = wps / 2;
  memset(bit_sets, 0xff, (size_t)(sample_end - sample_begin) * wps * sizeof(uint64_t)); /* cuking.cu:519-523 */
#pragma omp parallel for schedule(static)
  for (int64_t block = sample_begin / 8; block <= (int64_t)((sample_end - 1) / 8); ++block)
    for (uint32_t site = 0; site < num_sites; ++site) {
      int8_t g[8];
      syn_block(seed, (uint32_t)block, site, thr, g);
      for (uint32_t m = 0; m < 8; ++m) {
        const uint32_t s = (uint32_t)block * 8 + m;
        if (s < sample_begin || s >= sample_end || g[m] < 0) continue;
        uint64_t *het = bit_sets + (size_t)(s - sample_begin) * wps, *alt = het + half;
        const uint64_t bit = 1ull << (site & 63);
        if (g[m] != 1) het[site >> 6] &= ~bit; /* cuking.cu:689, :696 */
        if (g[m] != 2) alt[site >> 6] &= ~bit; /* cuking.cu:690, :693 */
      }
    }
}
