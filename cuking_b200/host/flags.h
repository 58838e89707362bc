// Command-line flags of the `cuking` binary.  Same names, defaults and validation as the ABSL_FLAGs of
// /root/reference/cuking.cu:27-52 and :437-462.  Abseil is not available in this environment, so parsing is done by
// hand with Abseil's syntax: --flag=value, --flag value, -flag, and "--" ends flag parsing.  The hyphenated spellings
// used by cloud_batch_submit.py:28-32 (--kin-threshold, --split-factor, ...) are accepted as aliases.
#pragma once
#include <cstdint>
#include <string>

namespace cuking {

struct Flags {
  std::string input_uri;                    // cuking.cu:27-29
  std::string output_uri;                   // cuking.cu:30-32
  std::string requester_pays_project;       // cuking.cu:33-35 (accepted, unused without GCS)
  size_t num_reader_threads = 36;           // cuking.cu:36-38
  uint32_t max_results = uint32_t(10) << 20;  // cuking.cu:39-41
  float kin_threshold = 0.0884f;            // cuking.cu:42-45
  uint32_t split_factor = 1;                // cuking.cu:46-48
  uint32_t shard_index = 0;                 // cuking.cu:49-52
  // extensions (default to reference behaviour)
  uint32_t num_gpus = 1;    // GPUs of this box to use: shards are scheduled across them, a lone shard is split into parts
  bool all_shards = false;  // decode the input once and compute every shard of --split_factor (one part file each)
  bool write_success_file = false;  // --all_shards: write <output>/_SUCCESS at the end (cloud_batch_submit.py:103-127)
  uint64_t row_group_rows = 0;      // rows per Parquet row group; 0 = one row group (cuking.cu:804-805)
  int device = 0;           // first CUDA device to use
  bool help = false;
};

// Returns an empty string on success, else the error message (main prints "ERROR: ..." and exits 1 like absl).
std::string ParseFlags(int argc, char **argv, Flags *flags);
std::string Usage();

}  // namespace cuking
