// Parquet input decode and output encode for the `cuking` binary — the reference's I/O surface.
//   input : /root/reference/cuking.cu:529-545 (listing), :574-672 (low-level column decode)
//   output: /root/reference/cuking.cu:770-862 (schema, Snappy, one row group), file name :868-870
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cuking_b200.h"

namespace cuking {

struct Triples {
  std::vector<int64_t> row_idx, col_idx;
  std::vector<int32_t> n_alt_alleles;
};

// Non-recursive listing of <dir>/*.parquet, sorted; everything else is skipped (cuking.cu:530-540).
std::string ListParquetFiles(const std::string &dir, std::vector<std::string> *files);

// Decodes one file: exactly 3 columns INT64, INT64, INT32 in that order (cuking.cu:585-590, :608, :630, :652), any
// number of row groups, any codec Arrow was built with.  OPTIONAL columns are accepted as long as they hold no nulls.
// Returns "" or an error message.
std::string ReadTriples(const std::string &path, Triples *out);

// Writes <dir>/part-<%05d shard>.snappy.parquet with the reference schema (all REQUIRED): i, j BYTE_ARRAY/String,
// kin FLOAT, ibs0, ibs1, ibs2 INT32; Snappy; one row group.  Returns "" or an error; *bytes_written = file size.
std::string WriteResults(const std::string &dir, uint32_t shard_index, const std::vector<std::string> &sample_ids,
                         const ck_result *results, size_t num_results, std::string *path_out, size_t *bytes_written);

}  // namespace cuking
