// Parquet input decode and output encode for the `cuking` binary — the reference's I/O surface.
//   input : /root/reference/cuking.cu:529-545 (listing), :574-672 (low-level column decode)
//   output: /root/reference/cuking.cu:770-862 (schema, Snappy, one row group), file name :868-870
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/cuking_b200.h"

namespace cuking {

// One chunk of decoded triples.  libparquet decodes the three columns at their physical width (INT64, INT64, INT32,
// cuking.cu:603-672) into pageable scratch; the chunk is then narrowed to what the pack step really uses - the 32-bit values
// the reference truncates row_idx / col_idx to (cuking.cu:676,:680) and one byte of n_alt_alleles (values that do not fit a
// byte become 255, which the pack rejects like any value other than 0, 1, 2) - in page-locked host memory (ck_host_alloc)
// that the GPU pack kernel reads in place: 9 instead of 20 bytes per triple over PCIe, which is what bounds ingest.
// Buffers grow geometrically and are reused from file to file by the owning reader thread.
class Triples {
 public:
  // narrow = true: decode into pageable scratch, then narrow into page-locked memory (9 bytes per triple over PCIe);
  // narrow = false: decode straight into page-locked arrays of the physical column widths (20 bytes per triple, no extra
  // pass over the chunk on the host)
  explicit Triples(bool narrow = true) : narrow_(narrow) {}
  Triples(const Triples &) = delete;
  Triples &operator=(const Triples &) = delete;
  ~Triples();
  // Makes room for n rows (contents are not preserved).  Returns "" or an error message.
  std::string Reserve(size_t n);
  // Narrows rows [0, size) of the decode scratch into the page-locked arrays (no-op in wide mode).
  void Narrow();
  bool narrow() const { return narrow_; }
  int64_t *row_idx = nullptr, *col_idx = nullptr;  // decoded columns (pageable scratch in narrow mode, page-locked in wide mode)
  int32_t *n_alt_alleles = nullptr;
  uint32_t *row32 = nullptr, *col32 = nullptr;     // page-locked, what ck_pack_triples_narrow reads (narrow mode)
  uint8_t *alt8 = nullptr;
  size_t size = 0;

 private:
  bool narrow_;
  void *block_ = nullptr;
  std::vector<int64_t> wide64_;
  std::vector<int32_t> wide32_;
  size_t capacity_ = 0;
};

// Non-recursive listing of <dir>/*.parquet, sorted; everything else is skipped (cuking.cu:530-540).
std::string ListParquetFiles(const std::string &dir, std::vector<std::string> *files);

// Decodes one file: exactly 3 columns INT64, INT64, INT32 in that order (cuking.cu:585-590, :608, :630, :652), any
// number of row groups, any codec Arrow was built with.  OPTIONAL columns are accepted as long as they hold no nulls.
// The file is streamed in chunks of at most `chunk_rows` rows through `buf` (page-locked, reused): after each chunk
// `consume(first_row_of_chunk)` is called with buf->size rows valid; a non-empty return value aborts the read.
// (The reference decodes the whole file into three vectors first, cuking.cu:596-672; the rows are the same.)
// Returns "" or an error message; *rows_out = rows delivered.
std::string ReadTriples(const std::string &path, size_t chunk_rows, Triples *buf,
                        const std::function<std::string(size_t)> &consume, size_t *rows_out);

// ---- pages for the device decoder (ck_pack_encoded) -----------------------------------------------------------------
// One window of rows of the three columns as page payloads + run tables + dictionaries.  The host only runs the page codec
// (parquet::PageReader) and walks the run headers (ck_rle_scan); bit unpacking, dictionary lookup and the narrowing happen
// in the pack kernel.  Pages are gathered in pageable buffers; a finished window is laid out as ONE block
// [payload | runs + sentinel | dictionary] x 3 in page-locked memory - a slice the caller lends (so that the window can be
// queued for the GPU while the reader goes on), else the window's own arena - which ck_pack_encoded uploads with one copy.
class EncodedWindow {
 public:
  EncodedWindow() = default;
  EncodedWindow(const EncodedWindow &) = delete;
  EncodedWindow &operator=(const EncodedWindow &) = delete;
  ~EncodedWindow();
  ck_encoded_column cols[3] = {};  // valid inside `consume`
  uint32_t num_rows = 0;           // rows of the window

  struct Page {  // one data page inside a column's buffers
    uint32_t first_value, num_values, byte_begin, byte_end, run_begin, run_end;
  };
  struct Column {
    uint8_t *bytes = nullptr;
    size_t bytes_cap = 0, bytes_size = 0;
    ck_run *runs = nullptr;
    size_t runs_cap = 0, runs_size = 0;
    uint8_t *dict = nullptr;
    size_t dict_cap = 0;
    uint32_t dict_len = 0, value_width = 0;
    uint64_t first_row = 0;    // row (inside the row group) of table value 0
    uint32_t num_values = 0;   // values described by the table
    std::vector<Page> pages;
    std::string GrowBytes(size_t want);
    std::string GrowRuns(size_t want);
    std::string GrowDict(size_t want);
    void Reset(uint32_t width);
    void DropBefore(uint64_t row);  // forgets the pages that end at or before `row` (a straddling page stays)
  };
  Column col[3];
  // Bytes the block of the window that ends at row `end_row` takes (the pages gathered beyond it are left out).
  size_t StagedBytes(uint64_t end_row) const;
  // Lays that block out at `dst` (NULL: in the window's own page-locked arena, grown as needed) and points cols[] at it.
  std::string Stage(uint64_t first_row, uint64_t end_row, uint8_t *dst);
  // where this reader's time went, in seconds (CUKING_INGEST_STATS=1 prints the sums over the threads)
  double s_pages = 0, s_scan = 0, s_stage = 0, s_wait = 0;

 private:
  uint8_t *arena_ = nullptr;
  size_t arena_cap_ = 0;
};

// Streams one file through `win` in windows of at most `window_rows` rows: after each window `consume(first_row, slice)` is
// called with win->cols / win->num_rows describing it.  `acquire` (may be empty) lends a page-locked slice of `slice_bytes`
// bytes per window - it may block, and NULL aborts the read with "Aborted"; the slice is the consumer's from `consume` on,
// so the window can be queued.  A window that would not fit a slice is halved (down to a few thousand rows); if it still
// does not fit, or without `acquire`, it is staged in the window's own arena and `consume` gets slice == NULL: that memory
// is reused by the next window, so the consumer must be done with it when it returns.  Same checks and error messages as
// ReadTriples.  *unsupported is set (and "" returned, nothing more delivered) when the file uses something the device
// decoder does not take - an encoding other than PLAIN / dictionary, legacy BIT_PACKED levels, a page of more than 2^31
// bytes: the caller then reads the file with ReadTriples instead (packing a triple twice is harmless, the pack is an AND).
std::string ReadEncoded(const std::string &path, size_t window_rows, size_t slice_bytes, EncodedWindow *win,
                        const std::function<uint8_t *()> &acquire,
                        const std::function<std::string(size_t first_row, uint8_t *slice)> &consume, size_t *rows_out,
                        bool *unsupported);

// Writer of <dir>/part-<%05d shard>.snappy.parquet with the reference schema (all REQUIRED): i, j BYTE_ARRAY/String,
// kin FLOAT, ibs0, ibs1, ibs2 INT32; Snappy (cuking.cu:770-798, :868-870).  Records are appended in sorted order chunk by
// chunk, so a shard's output never has to sit in host memory as a whole (the reference sorts and writes from one
// max_results-sized array, one value per WriteBatch call, :761-862).  One row group by default like the reference
// (:804-805) - its column pages are buffered compressed until Close; `row_group_rows` > 0 closes a row group every that
// many rows and bounds the buffering for outputs of billions of rows.  The file appears under its final name only on
// a successful Close (written as .tmp, then renamed), so readers never see a partial part file.
class ResultWriter {
 public:
  ResultWriter();
  ~ResultWriter();  // abandons (deletes) an unfinished file
  ResultWriter(const ResultWriter &) = delete;
  ResultWriter &operator=(const ResultWriter &) = delete;
  std::string Open(const std::string &dir, uint32_t shard_index, const std::vector<std::string> *sample_ids, uint64_t row_group_rows);
  std::string Append(const ck_result *records, size_t n);
  std::string Close(std::string *path_out, size_t *bytes_written);
  uint64_t rows() const { return rows_; }

 private:
  struct Impl;
  Impl *impl_ = nullptr;
  uint64_t rows_ = 0;
};

// One-shot form: Open + Append + Close.
std::string WriteResults(const std::string &dir, uint32_t shard_index, const std::vector<std::string> &sample_ids,
                         const ck_result *results, size_t num_results, std::string *path_out, size_t *bytes_written);

}  // namespace cuking
