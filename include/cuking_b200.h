/*
 * cuking_b200.h — C ABI of the B200-native pairwise-KING hot path (libcuking_b200.so).
 *
 * The reference (populationgenomics/cuKING) is one executable with no plugin/FFI layer; the seam this ABI
 * replaces is the host -> device boundary inside its Run() (reference paths relative to /root/reference):
 *
 *   shard planning        struct Submatrix                       cuking.cu:129-179
 *   bit-set allocation    NewCudaArray<uint64_t> + memset 0xFF   cuking.cu:513-523
 *   transpose / pack      AtomicClearBit loop                    cuking.cu:317-323, :675-703
 *   the kernel launch     ComputeKingKernel<<<grid,128>>>(...)   cuking.cu:191-195, :734-744
 *   result record         struct KingResult                      cuking.cu:182-186
 *   overflow check        result_index_and_flag[1]               cuking.cu:747-751
 *   result sort           std::sort by (i, j, kin)               cuking.cu:761-765
 *
 * Conventions: every call returns an int status (CK_OK == 0) and never throws or exits; the message for the
 * last failure on the calling thread is available from ck_last_error().  The caller owns host buffers; the
 * library owns device buffers behind opaque handles.  One ck_ctx per GPU; calls on one ctx are not thread-safe,
 * different ctxs are independent.  Calls are stream-ordered on the ctx's stream and synchronous at return unless
 * documented otherwise.  There is no CPU fallback: without a CUDA device every device call fails with CK_ERR_CUDA.
 */
#ifndef CUKING_B200_H_
#define CUKING_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CK_ABI_VERSION 3

enum ck_status {
  CK_OK = 0,
  CK_ERR_INVALID_ARGUMENT = 1, /* absl::InvalidArgumentError at cuking.cu:437-462 */
  CK_ERR_CUDA = 2,             /* any CUDA runtime failure (the reference ignores them, cuking.cu:744) */
  CK_ERR_RESULT_OVERFLOW = 3,  /* absl::ResourceExhaustedError at cuking.cu:747-751 */
  CK_ERR_INVALID_GENOTYPE = 4, /* absl::FailedPreconditionError at cuking.cu:698-701 */
  CK_ERR_OUT_OF_RANGE = 5,     /* row_idx >= num_sites: undefined behaviour in the reference, an error here */
  CK_ERR_OUT_OF_MEMORY = 6     /* the reference exit(1)s at cuking.cu:114-118 */
};

/* Bounds of the relatedness sub-matrix one shard computes.  Mirrors struct Submatrix, cuking.cu:129-179. */
typedef struct ck_submatrix {
  uint32_t i_begin, i_end; /* sample row range    */
  uint32_t j_begin, j_end; /* sample column range */
} ck_submatrix;

/* One retained pair.  Bit-compatible with struct KingResult, cuking.cu:182-186 (24 bytes). */
typedef struct ck_result {
  uint32_t sample_i, sample_j; /* global sample indices, sample_i < sample_j */
  float kin;                   /* KING between-family kinship, fp32, expression order of cuking.cu:289-294 */
  uint32_t ibs0, ibs1, ibs2;   /* cuking.cu:305-307 */
} ck_result;

/* Raw per-pair counters of cuking.cu:214-240, for parity tests (ck_king_counts). */
typedef struct ck_counts {
  uint32_t het_i, het_j, both_het, opposing_hom, concordant_hom, shared_sites;
} ck_counts;

typedef struct ck_ctx ck_ctx;       /* one GPU + one stream */
typedef struct ck_planes ck_planes; /* device-resident genotype bit planes for a set of sample slots */

/* Parameters of the synthetic genotype generator: HWE founders, planted pedigrees in blocks of
 * 8 consecutive samples, per-genotype missingness.  Pure function of (seed, sample, site). */
typedef struct ck_synth_params {
  uint64_t seed;
  double missing_rate; /* m: P(genotype missing) */
} ck_synth_params;

/* Timing of the most recent calls on a ctx, measured with CUDA events on the ctx's stream. */
typedef struct ck_timings {
  float pack_ms;     /* last ck_pack_triples: pack kernel only */
  float finalize_ms; /* last plane finalisation (raw -> compute planes) */
  float import_ms;   /* last ck_planes_import_bitset transpose kernel(s) */
  float king_ms;     /* last pairwise kernel */
  float sort_ms;     /* last device result sort */
  float h2d_ms, d2h_ms;
  uint32_t king_launches; /* kernels launched by the last ck_king* call (pairwise + sort helpers) */
} ck_timings;

/* ---- library ------------------------------------------------------------------------------------------- */

int ck_abi_version(void);
const char *ck_last_error(void);
int ck_device_count(int *count);

/* ---- shard planning: cuking.cu:129-179, validation :454-462 ---------------------------------------------- */

/* Fails with CK_ERR_INVALID_ARGUMENT for split_factor == 0 ("Invalid split factor", cuking.cu:454-457) and for
 * shard_index >= k(k+1)/2 ("Invalid shard index", :459-462).  Unlike the reference, a block that starts past
 * num_samples yields an empty range instead of an underflowed one (SURVEY.md §8a4 latent bug). */
int ck_submatrix_init(uint32_t num_samples, uint32_t split_factor, uint32_t shard_index, ck_submatrix *out);
uint32_t ck_num_shards(uint32_t split_factor);                         /* k(k+1)/2, cloud_batch_submit.py:73 */
uint32_t ck_submatrix_num_rows(const ck_submatrix *sm);                /* cuking.cu:154 */
uint32_t ck_submatrix_num_cols(const ck_submatrix *sm);                /* cuking.cu:156 */
uint32_t ck_submatrix_num_samples(const ck_submatrix *sm);             /* cuking.cu:159-162 */
uint32_t ck_submatrix_contains(const ck_submatrix *sm, uint32_t s);    /* cuking.cu:165-168 */
uint32_t ck_submatrix_sample_offset(const ck_submatrix *sm, uint32_t s); /* cuking.cu:171-175 */
/* u64 words per sample in the REFERENCE layout: 2 * ceil(pad32(num_sites) / 64), cuking.cu:498-500, :513. */
uint32_t ck_words_per_sample(uint32_t num_sites);

/* Orchestration of several shards on the GPUs of one box - the local counterpart of one Cloud Batch task per shard,
 * cloud_batch_submit.py:45,73.  Shards [first_shard, first_shard + num_run) of the split are cut into work items and
 * assigned to GPUs: a shard whose pair count exceeds the per-GPU share is split into parts (ck_king_view's part_index /
 * num_parts) so that no item is larger than the share, then the items go to the GPUs longest-first onto the least
 * loaded GPU (LPT; diagonal shards cost about half an off-diagonal one).  `items` receives at most max_items entries,
 * grouped by GPU in execution order; *num_items the number needed. */
typedef struct ck_work_item {
  uint32_t shard_index, part_index, num_parts, gpu;
  uint64_t pairs; /* i < j pairs of the whole shard (an item's cost is pairs / num_parts) */
} ck_work_item;
int ck_plan_work(uint32_t num_samples, uint32_t split_factor, uint32_t first_shard, uint32_t num_run, uint32_t num_gpus,
                 ck_work_item *items, uint32_t max_items, uint32_t *num_items);

/* ---- context --------------------------------------------------------------------------------------------- */

int ck_ctx_create(int device, ck_ctx **out);
/* Run all later work of this ctx on an existing cudaStream_t (e.g. torch's current stream).  NULL restores the
 * ctx-owned stream. */
int ck_ctx_set_stream(ck_ctx *ctx, void *cuda_stream);
/* Pairwise kernel variant: 0 = LOP3 + 5 POPC per pair and 32 sites, 1 = carry-save (2.5 POPC + 5 more LOP3),
 * 2 = tcgen05 int8 tensor-core formulation (five exact s32 GEMMs of indicator vectors), 3 = the same five GEMMs on
 * the FP4 tensor path (kind::mxf4 E2M1 operands, unit block scales, fp32 accumulation - exact for counts <= 2^23; planes
 * with more than 2^23 sites are routed to variant 2), 4 = variant 3 on CTA pairs (tcgen05 cta_group::2, 256-row tiles
 * sharing the B operand; experimental, slower today: profiles/r02_pair_kernel.md), 5 = variant 3 behind a screen: a
 * pair can only pass the threshold if a bound on its squared genotype distance - one or three exact products instead of
 * five, chosen from the cohort's call rate - lies under a multiple of the samples' het counts, and only tiles holding
 * such a pair go on to the five-product kernel (dense output, the count dump and thresholds too low for a
 * screen to reject anything take variant 3 directly), -1 = library default (= 5; also settable with the
 * CUKING_KING_VARIANT environment variable).  Results are bit-identical across variants. */
int ck_ctx_set_king_variant(ck_ctx *ctx, int variant);
/* Variant 3 relies on the tensor core adding E2M1 products into its fp32 accumulator without losing low bits, which the
 * PTX ISA does not spell out.  The first use of variant 3 on a ctx therefore runs an on-device self-test (about a
 * millisecond: accumulators pre-loaded just below 2^21 .. 2^24, addends of 1, 1/2 and 1/4, alternating signs, single
 * non-zero operands, compared with integer arithmetic); if any accumulator differs, the ctx runs variant 2 (int8, s32
 * accumulators, exact by specification) wherever variant 3 was asked for, and says so on stderr.  This call runs the
 * self-test now: *exact = 1 / 0, and ck_last_error() holds a one-line report. */
int ck_ctx_fp4_selftest(ck_ctx *ctx, int *exact);
/* Variant 5 bookkeeping since the ctx was created: tiles its screens were launched over, tiles they flagged (the ones the
 * five-product kernel then ran on), and the screen of the last evaluation (1 = one product, 3 = three products, 0 = none:
 * dense output, count dump, or a threshold too low for a screen).  Synchronises the ctx's stream. */
int ck_ctx_screen_stats(ck_ctx *ctx, uint64_t *tiles_screened, uint64_t *tiles_flagged, int *level);
int ck_ctx_synchronize(ck_ctx *ctx);
int ck_ctx_get_timings(ck_ctx *ctx, ck_timings *out);
/* Measures, on this GPU and now, the sustained issue rate of POPC.32 and LOP3 (lane-ops per second, whole chip).
 * The pairwise kernel is bound by these pipes, not by HBM or the tensor cores; bench.py quotes its roofline against
 * the POPC figure.  Takes a few milliseconds. */
int ck_measure_int_peaks(ck_ctx *ctx, double *popc_lane_ops_per_s, double *lop3_lane_ops_per_s);
/* Measures, on this GPU and now, the dense tcgen05 kind::mxf4 rate (E2M1 operands, ops = 2 x MACs per second, whole
 * chip): every SM streams M = 128, N = 208, K = 64 instructions from resident operands for a few milliseconds.  This
 * is what bounds pairwise kernel variant 3 (10 fp4 ops per pair-site); bench.py quotes its roofline against it because
 * MEASURED_PEAKS.json holds no fp4 figure. */
int ck_measure_fp4_peak(ck_ctx *ctx, double *ops_per_s);
/* The same rate sustained: the kernel is launched back to back for `seconds` (0 < seconds <= 30) and the second half is
 * timed, i.e. at the clock the board settles at under its power limit with a saturated tensor pipe - the denominator
 * for a pairwise pass that itself runs for a second or longer. */
int ck_measure_fp4_peak_sustained(ck_ctx *ctx, double seconds, double *ops_per_s);
int ck_ctx_destroy(ck_ctx *ctx);

/* ---- planes: replaces the managed bit_set of cuking.cu:513-523 --------------------------------------------- */

/* Allocates planes for the samples of one shard; every genotype starts missing (cuking.cu:520-523).  Sample
 * s of the sub-matrix lives in slot ck_submatrix_sample_offset(sm, s), exactly as in the reference. */
int ck_planes_create(ck_ctx *ctx, const ck_submatrix *sm, uint32_t num_sites, ck_planes **out);
int ck_planes_reset(ck_planes *planes); /* back to all-missing */
int ck_planes_destroy(ck_planes *planes);
/* Derives the compute planes (true-het / defined / true-hom-alt) from the packed raw planes.  Implicit in the first
 * ck_king* call after a pack / import / synthesize; exposed so that it can be timed or moved off the critical path. */
int ck_planes_finalize(ck_planes *planes);
int ck_planes_num_sites(const ck_planes *planes, uint32_t *num_sites);
int ck_planes_device_bytes(const ck_planes *planes, uint64_t *bytes);

/* Transpose / pack, cuking.cu:675-703.  For each triple whose col_idx is in the sub-matrix (cuking.cu:677):
 * n_alt_alleles 0 clears both bits, 1 clears the hom-alt bit, 2 clears the het bit (AND-accumulation, so
 * duplicate or conflicting triples combine exactly as in the reference and order never matters).  row_idx and
 * col_idx are truncated to 32 bits like cuking.cu:676,:680.  Any other n_alt_alleles value fails the call with
 * CK_ERR_INVALID_GENOTYPE (cuking.cu:698-701); row_idx >= num_sites fails with CK_ERR_OUT_OF_RANGE.
 * The three arrays are host pointers (on_device == 0) or device pointers.  Pageable host arrays are staged through
 * pinned memory chunk by chunk; page-locked host arrays (ck_host_alloc) are read by the kernel in place. */
int ck_pack_triples(ck_planes *planes, const int64_t *row_idx, const int64_t *col_idx, const int32_t *n_alt_alleles,
                    size_t num_triples, int on_device);

/* The same for triples already narrowed by the caller: row_idx and col_idx as the 32-bit values the reference truncates
 * them to (cuking.cu:676,:680), n_alt_alleles as one byte (any value other than 0, 1, 2 - use 255 for an int32 that does
 * not fit a byte - fails the call like ck_pack_triples).  9 instead of 20 bytes per triple: when the kernel reads
 * page-locked host memory in place, ingest is bound by the PCIe link, so a host that decodes Parquet should narrow its
 * columns once and use this entry point (host/parquet_io.cc does). */
int ck_pack_triples_narrow(ck_planes *planes, const uint32_t *row_idx, const uint32_t *col_idx, const uint8_t *n_alt_alleles,
                           size_t num_triples, int on_device);

/* ---- Parquet pages decoded on the device (replaces the ReadBatch loops of cuking.cu:603-672) ----
 * The host only decompresses the pages of the three columns (parquet::PageReader, any codec) and walks the run headers of
 * their RLE / bit-packed hybrid streams; bit unpacking, dictionary lookup, the int32 truncations and the pack itself run in
 * one kernel.  About 2 bytes per triple cross PCIe instead of 9 (narrowed) or 20 (as decoded), and the host skips
 * libparquet's value decoding altogether.
 *
 * A column of one window of rows is described by a table of runs over a byte buffer of page payloads:
 *   CK_RUN_RLE        `count` copies of the dictionary index `payload`
 *   CK_RUN_BITPACKED  dictionary indices of `bit_width` bits each, LSB first, starting at byte `payload` of the buffer
 *                     (Parquet Encodings.md, "RLE/Bit-Packing Hybrid", the encoding of RLE_DICTIONARY / PLAIN_DICTIONARY
 *                     data pages behind their bit-width byte)
 *   CK_RUN_PLAIN      little-endian values of the column's physical width starting at byte `payload` (PLAIN data pages:
 *                     the writer's fallback when a dictionary grows too large); `payload` is a multiple of that width
 * Run r covers the values [first_value, runs[r + 1].first_value); first_value is strictly increasing and the table ends
 * with a sentinel entry (kind ignored) whose first_value is the column's value count. */
enum ck_run_kind { CK_RUN_RLE = 0, CK_RUN_BITPACKED = 1, CK_RUN_PLAIN = 2 };
typedef struct ck_run {
  uint32_t first_value; /* index of the run's first value in this column's window table */
  uint32_t kind;        /* enum ck_run_kind */
  uint32_t bit_width;   /* CK_RUN_BITPACKED: 0 .. 32 */
  uint32_t payload;     /* see above */
} ck_run;

typedef struct ck_encoded_column {
  const uint8_t *bytes; /* page payloads of the window (host memory; page-locked memory is copied faster) */
  uint64_t num_bytes;
  const ck_run *runs;   /* num_runs entries + the sentinel (host memory) */
  uint32_t num_runs;
  const void *dict;     /* PLAIN dictionary page: dict_len values of value_width bytes (NULL if no run needs it) */
  uint32_t dict_len;
  uint32_t value_width; /* 8 = INT64 (row_idx, col_idx), 4 = INT32 (n_alt_alleles) */
  uint32_t skip;        /* leading values of the table that precede the window's first row (pages rarely end together
                           in the three columns: a page that straddles two windows is simply described in both) */
} ck_encoded_column;

/* Walks one RLE / bit-packed hybrid stream of `num_values` values and appends its runs to runs[*num_runs ...] (capacity
 * max_runs entries, sentinel not included and not written).  first_value = table index of the stream's first value,
 * payload_base = byte offset of `data` inside the column's payload buffer.  Empty runs are dropped; the values of a
 * last bit-packed group beyond num_values (padding) are not counted.  Pure host code.  CK_ERR_INVALID_ARGUMENT: the stream is malformed or ends early; CK_ERR_OUT_OF_RANGE: the
 * table is full (num_bytes + 1 free entries always suffice). */
int ck_rle_scan(const uint8_t *data, size_t num_bytes, uint32_t bit_width, uint32_t num_values, uint32_t first_value,
                uint32_t payload_base, ck_run *runs, uint32_t max_runs, uint32_t *num_runs);

/* Decodes rows [0, num_rows) of a window - row r is value skip + r of each column; cols[0] = row_idx, cols[1] = col_idx,
 * cols[2] = n_alt_alleles - and packs them exactly like ck_pack_triples (same filter, truncations and errors; the
 * "triple" index of an error message is the row inside the window).  A table that is inconsistent with its buffers
 * (run outside the payload bytes, dictionary index >= dict_len, unsorted runs) fails with CK_ERR_INVALID_ARGUMENT
 * before or instead of touching the planes' neighbours.  Synchronous like ck_pack_triples.  Unlike the other calls of a
 * ctx this one may be issued by several host threads at once on the same planes (the decode threads of a host): every call
 * runs on its own stream, staging buffer and error slots of the ctx, so the windows overlap on the GPU - the pack is a pure
 * AND-accumulation - instead of queueing behind a host lock.  Pieces that lie close together in host memory (a window laid
 * out in one arena) are uploaded with a single copy. */
int ck_pack_encoded(ck_planes *planes, const ck_encoded_column cols[3], uint32_t num_rows);

/* Page-locked host memory for triple buffers.  ck_pack_triples recognises it (and any other cudaHostAlloc /
 * cudaHostRegister memory) and lets the pack kernel stream the triples straight over PCIe, skipping the staging copy;
 * decode threads should read their Parquet columns directly into buffers from here. */
int ck_host_alloc(size_t bytes, void **out);
int ck_host_free(void *ptr);

/* AND-all-reduce of the raw planes of `count` plane sets of identical shape, one per GPU, over NVLink peer memory.
 * The pack only clears bits of an all-missing bit set (cuking.cu:689-696), so it distributes over AND: deal the triples
 * to the GPUs (each triple to exactly one), let each pack its share into its own full-size planes, then call this:
 * afterwards every GPU holds the planes of ALL triples.  One fused reduce-scatter + all-gather kernel per GPU (P2P
 * loads and stores), nothing through the host.  Called by one host thread while no other call runs on the ctxs
 * involved; fails with CK_ERR_CUDA when the GPUs have no peer access to each other. */
int ck_planes_and_reduce(ck_planes *const *planes, uint32_t count);

/* Exchange with the reference bit-set layout (cuking.cu:204-212, :507-523): sample-major uint64 words, slot o at
 * [o*W, (o+1)*W), het plane first, hom-alt plane second, W = ck_words_per_sample(num_sites); site r is bit r&63 of
 * word r>>6; (het, hom-alt) = (1,1) means missing, including the padding sites.  The buffer covers all
 * ck_submatrix_num_samples slots.  Import REPLACES the planes' contents. */
int ck_planes_import_bitset(ck_planes *planes, const uint64_t *bit_set, int on_device);
int ck_planes_export_bitset(ck_planes *planes, uint64_t *bit_set, int on_device);

/* Fills the planes with the synthetic cohort of SURVEY.md §8d (global sample index = sub-matrix sample). */
int ck_planes_synthesize(ck_planes *planes, const ck_synth_params *params);

/* ---- the pairwise kernel: replaces the launch at cuking.cu:734-744 ---------------------------------------- */

/* For every pair (i, j), i in the sub-matrix rows, j in its columns, i < j: the six counters of cuking.cu:232-239,
 * kin (cuking.cu:289-294), and a record for each pair with kin > kin_threshold (strict, cuking.cu:297).
 * results: caller buffer of max_results records, host (results_on_device == 0) or device memory.
 * *num_results receives the number of pairs above the threshold.  If it exceeds max_results the call returns
 * CK_ERR_RESULT_OVERFLOW (cuking.cu:747-751) and the buffer contents are unspecified.
 * sort != 0 orders the records by (sample_i, sample_j, kin) like cuking.cu:761-765; otherwise order is unspecified. */
int ck_king(ck_planes *planes, float kin_threshold, uint32_t max_results, ck_result *results, int results_on_device,
            uint32_t *num_results, int sort);

/* A shard as a VIEW into planes that hold a larger sample range, and one PART of it (cuking.cu:129-152 allocates and
 * packs one bit set per shard process; here a cohort is packed once and every shard of a --split_factor run reads it):
 *   planes      created over a diagonal sub-matrix, i.e. one contiguous sample range (the whole cohort, typically);
 *   view        any sub-matrix inside that range whose rows and columns are identical (diagonal shard) or disjoint with
 *               the rows first (off-diagonal shard) - exactly what ck_submatrix_init yields; NULL = the planes' own
 *               sub-matrix (then off-diagonal planes work too);
 *   part_index  of num_parts: the bands of 1024 rows of the view are dealt to the parts in snake order, so the parts
 *               carry equal work; the parts are disjoint and their union is the whole view (one part per GPU of a box;
 *               max_results bounds each part).
 * Everything else as ck_king.  Views need kernel variant 2 or 3.
 * Dense output: when kin_threshold < 0 and max_results >= the number of i < j pairs of the part (BASELINE configs[4],
 * --kin_threshold -1), sorted host results are produced WITHOUT the append counter or a sort: every pair is written to
 * the position it has in the sorted output, the rare below-threshold pairs are squeezed out afterwards, and finished
 * rows are copied to the host (page-locked `results` recommended: ck_host_alloc) while later rows are still computed. */
int ck_king_view(ck_planes *planes, const ck_submatrix *view, uint32_t part_index, uint32_t num_parts, float kin_threshold,
                 uint32_t max_results, ck_result *results, int results_on_device, uint32_t *num_results, int sort);
/* The same with the sorted records delivered chunk by chunk (at most chunk_records per call, 0 = 4 Mi) to `sink`, in
 * order, through a page-locked double buffer the library owns: host memory stays bounded whatever the result size, and
 * the copy of the next chunk overlaps the sink's work on this one (cuking.cu:761-862 sorts and writes from one
 * max_results-sized managed array).  A non-zero return value of the sink aborts the call. */
typedef int (*ck_result_sink)(void *user, const ck_result *records, size_t count);
int ck_king_view_sink(ck_planes *planes, const ck_submatrix *view, uint32_t part_index, uint32_t num_parts, float kin_threshold,
                      uint32_t max_results, size_t chunk_records, ck_result_sink sink, void *user, uint64_t *num_results);

/* The pairwise kernel variant that ck_king* will run on these planes (the ctx's variant, except that variants 3, 4 and 5
 * fall back to 2 beyond 2^23 sites or when the ctx's GPU failed the kind::mxf4 self-test). */
int ck_planes_king_variant(const ck_planes *pl, int *variant);
/* Same, restricted to the linear range [tile_begin, tile_end) of the sub-matrix's tile grid (row-major over the tiles
 * that can contain an i < j pair; the tile shape belongs to the active kernel variant, so tile counts are only
 * comparable under one variant).  This is how one shard is split across the GPUs of a box:
 * each GPU holds the planes and takes a contiguous slice of ck_king_num_tiles().  Results of all slices together
 * equal ck_king's. */
int ck_king_num_tiles(const ck_planes *planes, uint64_t *num_tiles);
int ck_king_tiles(ck_planes *planes, uint64_t tile_begin, uint64_t tile_end, float kin_threshold,
                  uint32_t max_results, ck_result *results, int results_on_device, uint32_t *num_results, int sort);

/* Parity hook: raw counters and kin for an explicit list of (sample_i, sample_j) pairs (global indices, both in
 * the sub-matrix), bypassing threshold and compaction.  counts/kin are host arrays of num_pairs entries. */
int ck_king_counts(ck_planes *planes, const uint32_t *sample_i, const uint32_t *sample_j, size_t num_pairs,
                   ck_counts *counts, float *kin);

/* The reference seam in one call, host buffers in and out (what a maintainer would call from Run() in place of
 * cuking.cu:713-765): bit_set in the reference layout for Submatrix(num_samples, split_factor, shard_index),
 * copied host->device, evaluated, sorted, and copied back. */
int ck_king_host_bitset(ck_ctx *ctx, uint32_t num_samples, uint32_t split_factor, uint32_t shard_index,
                        uint32_t num_sites, const uint64_t *bit_set, float kin_threshold, uint32_t max_results,
                        ck_result *results, uint32_t *num_results);
/* The same seam for several GPUs of one box: every GPU calls it on its own ctx with the same host bit set and its own
 * part_index in [0, num_parts); the library picks the partition of the shard's pair matrix (bands of rows dealt in
 * snake order so that the parts carry equal work and every part can overlap its upload with its kernel) and returns
 * that part's retained pairs, sorted.  The parts are disjoint and their union is ck_king_host_bitset's result;
 * max_results bounds each part (the caller applies the reference's overflow rule to the total if it wants it global). */
int ck_king_host_bitset_part(ck_ctx *ctx, uint32_t num_samples, uint32_t split_factor, uint32_t shard_index,
                             uint32_t num_sites, const uint64_t *bit_set, float kin_threshold, uint32_t max_results,
                             ck_result *results, uint32_t *num_results, uint32_t part_index, uint32_t num_parts);

/* Streaming form of the seam, for callers that deliver the bit set in pieces (several readers + an NCCL all-gather over
 * NVLink, a file reader, ...).  Diagonal shards and the tensor-core kernel variants (2, 3) only.
 *   ck_king_stream_begin(planes, thr, max_results, part_index, num_parts)
 *   ck_king_stream_rows(planes, rows, on_device, sample_begin, sample_end)   repeatedly, for DESCENDING ranges of shard-
 *       local sample indices that tile [0, rows of the shard): boundaries are multiples of ck_king_stream_granularity()
 *       (the shard's end excepted); `rows` points at the reference-layout row (cuking.cu:507-513) of sample_begin.  The
 *       call transposes the rows, derives the genotype codes and launches this part's bands among them - with device
 *       memory everything is queued on the ctx stream and the call returns at once (keep `rows` alive until end);
 *   ck_king_stream_end(planes, results, &num_results)   waits, applies the overflow rule, sorts, copies out.
 * A band only pairs its rows with samples at or after them, which is why descending delivery lets the kernel start on
 * the first piece.  ck_king_host_bitset[_part] is this API fed by chunked cudaMemcpyAsync from one host buffer. */
uint32_t ck_king_stream_granularity(void);
int ck_king_stream_begin(ck_planes *planes, float kin_threshold, uint32_t max_results, uint32_t part_index, uint32_t num_parts);
int ck_king_stream_rows(ck_planes *planes, const uint64_t *rows, int on_device, uint32_t sample_begin, uint32_t sample_end);
int ck_king_stream_end(ck_planes *planes, ck_result *results, uint32_t *num_results);

/* ---- synthetic inputs (bench / tests) --------------------------------------------------------------------- */

/* Dense genotypes of the synthetic cohort on the HOST: out[(s - sample_begin) * num_sites_out + (r - site_begin)]
 * in {0,1,2} or -1 for missing.  Same function the device generator evaluates. */
int ck_synth_genotypes_host(const ck_synth_params *params, uint32_t sample_begin, uint32_t sample_end,
                            uint32_t site_begin, uint32_t site_end, int8_t *out);

/* Sparse triples of the synthetic cohort, generated ON THE DEVICE in Hail's order (site-major, sample-minor,
 * missing entries absent; mt_to_cuking_inputs.py:28-30) for samples [sample_begin, sample_end) x sites
 * [site_begin, site_end).  The three device arrays are library-owned and valid until the next call or
 * ck_ctx_destroy. */
int ck_synth_triples_device(ck_ctx *ctx, const ck_synth_params *params, uint32_t sample_begin, uint32_t sample_end,
                            uint32_t site_begin, uint32_t site_end, const int64_t **row_idx, const int64_t **col_idx,
                            const int32_t **n_alt_alleles, size_t *num_triples);

#ifdef __cplusplus
}
#endif
#endif /* CUKING_B200_H_ */
