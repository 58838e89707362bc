// `cuking` — host side of the B200-native pairwise KING path: same command line, input layout, output schema and
// phase log as the reference's Run(), /root/reference/cuking.cu:435-882, with the device work behind the C ABI of
// include/cuking_b200.h.
//
//   validate flags (:437-462) -> metadata.json (:475-500) -> shard planning (:505) -> plane allocation (:513-523)
//   -> list *.parquet (:529-545) -> parallel decode (:550-672) -> pack (:675-703, on the GPU here)
//   -> pairwise kernel + overflow check (:713-751) -> sort (:761-765, on the GPU here) -> Parquet write (:770-875)
//
// Differences, all deliberate: local directories (or file://) replace gs:// (no google-cloud-cpp here; gs:// is
// rejected with the reference's own "Unsupported URI" error class); every CUDA call is checked; extensions
// --num_gpus / --all_shards default to the reference behaviour of one shard on one GPU.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cuking_b200.h"
#include "flags.h"
#include "json_min.h"
#include "parquet_io.h"

namespace {

using cuking::Flags;

struct Status {
  std::string code;  // absl::StatusCode spelling, e.g. INVALID_ARGUMENT; empty = OK
  std::string message;
  bool ok() const { return code.empty(); }
};
Status Ok() { return {}; }
Status InvalidArgument(const std::string &m) { return {"INVALID_ARGUMENT", m}; }
Status FailedPrecondition(const std::string &m) { return {"FAILED_PRECONDITION", m}; }
Status ResourceExhausted(const std::string &m) { return {"RESOURCE_EXHAUSTED", m}; }
Status Unknown(const std::string &m) { return {"UNKNOWN", m}; }
Status Internal(const std::string &m) { return {"INTERNAL", m}; }

Status FromCk(int rc) {
  const std::string msg = ck_last_error();
  switch (rc) {
    case CK_OK: return Ok();
    case CK_ERR_INVALID_ARGUMENT: return InvalidArgument(msg);
    case CK_ERR_RESULT_OVERFLOW: return ResourceExhausted(msg);  // cuking.cu:747-751
    case CK_ERR_INVALID_GENOTYPE:
    case CK_ERR_OUT_OF_RANGE: return FailedPrecondition(msg);     // cuking.cu:698-701
    case CK_ERR_OUT_OF_MEMORY: return ResourceExhausted(msg);
    default: return Internal(msg);
  }
}

class StopWatch {  // cuking.cu:326-337
 public:
  std::string ElapsedAndReset() {
    const auto now = std::chrono::steady_clock::now();
    const double s = std::chrono::duration<double>(now - last_).count();
    last_ = now;
    char buf[32];
    if (s < 1.0) snprintf(buf, sizeof(buf), "%.3gms", s * 1e3);
    else snprintf(buf, sizeof(buf), "%.4gs", s);
    return buf;
  }

 private:
  std::chrono::steady_clock::time_point last_ = std::chrono::steady_clock::now();
};

// The reference insists on gs:// (SplitGcsUri, cuking.cu:339-353).  Here a URI is a local directory.
Status ResolveLocalUri(const std::string &uri, std::string *path) {
  if (uri.rfind("gs://", 0) == 0)
    return InvalidArgument("Unsupported URI: " + uri + " (this build has no GCS transport; pass a local directory or file:// URI)");
  *path = uri.rfind("file://", 0) == 0 ? uri.substr(7) : uri;
  while (path->size() > 1 && path->back() == '/') path->pop_back();
  if (path->empty()) return InvalidArgument("Incomplete blob URI " + uri);
  return Ok();
}

struct Metadata {
  uint32_t num_sites = 0;
  std::vector<std::string> sample_ids;
};

Status ReadMetadata(const std::string &dir, Metadata *md) {  // cuking.cu:475-500
  const std::string path = dir + "/metadata.json";
  std::ifstream in(path, std::ios::binary);
  if (!in) return FailedPrecondition("Failed to read metadata: cannot open " + path);
  std::stringstream ss;
  ss << in.rdbuf();
  cuking::JsonValue doc;
  std::string err;
  if (!cuking::ParseJson(ss.str(), &doc, &err)) return FailedPrecondition("Failed to parse metadata JSON: " + err);
  const cuking::JsonValue *samples = doc.Find("samples");
  const cuking::JsonValue *num_sites = doc.Find("num_sites");
  if (doc.kind != cuking::JsonValue::kObject || !samples || samples->kind != cuking::JsonValue::kArray || !num_sites ||
      num_sites->kind != cuking::JsonValue::kNumber || !num_sites->number_is_integer || num_sites->integer <= 0 ||
      num_sites->integer > 0xffffffffll)
    return FailedPrecondition("Failed to parse metadata JSON: expected {\"num_sites\": <int>, \"samples\": [...]}");
  md->num_sites = uint32_t(num_sites->integer);
  md->sample_ids.reserve(samples->array.size());
  for (const auto &s : samples->array) {
    if (s.kind != cuking::JsonValue::kString) return FailedPrecondition("Failed to parse metadata JSON: sample IDs must be strings");
    md->sample_ids.push_back(s.string);
  }
  return Ok();
}

// One GPU of the box.  Calls on one ck_ctx are not thread-safe, so the decode threads serialise on `mu`.
struct Gpu {
  ck_ctx *ctx = nullptr;
  std::mutex mu;
};

// Planes of one shard on one GPU.
struct ShardOnGpu {
  Gpu *gpu = nullptr;
  ck_planes *planes = nullptr;
};

struct ShardJob {
  uint32_t shard_index = 0;
  ck_submatrix sm{};
  std::vector<ShardOnGpu *> replicas;  // one per GPU that works on this shard
};

Status Run(const Flags &flags) {
  // ---- flag validation, cuking.cu:437-462 ----
  if (flags.input_uri.empty()) return InvalidArgument("No input URI specified");
  std::string input_dir, output_dir;
  if (Status s = ResolveLocalUri(flags.input_uri, &input_dir); !s.ok()) return s;
  if (flags.output_uri.empty()) return InvalidArgument("No output URI specified");
  if (Status s = ResolveLocalUri(flags.output_uri, &output_dir); !s.ok()) return s;
  if (flags.num_reader_threads == 0) return InvalidArgument("Invalid number of reader threads");
  if (flags.split_factor == 0) return InvalidArgument("Invalid split factor");
  const uint64_t num_shards = uint64_t(flags.split_factor) * (uint64_t(flags.split_factor) + 1) / 2;
  if (flags.shard_index >= num_shards) return InvalidArgument("Invalid shard index");

  StopWatch stop_watch;
  std::cout << "Reading metadata...";
  std::cout.flush();
  Metadata md;
  if (Status s = ReadMetadata(input_dir, &md); !s.ok()) return s;
  const uint32_t num_samples = uint32_t(md.sample_ids.size());
  std::cout << " (" << stop_watch.ElapsedAndReset() << ")" << std::endl;

  // ---- devices ----
  std::cout << "Initializing CUDA...";
  std::cout.flush();
  int device_count = 0;
  if (int rc = ck_device_count(&device_count); rc != CK_OK) return FromCk(rc);
  if (device_count <= 0) return Internal("No CUDA device found (this program has no CPU fallback)");
  if (flags.device + int(flags.num_gpus) > device_count)
    return InvalidArgument("--device/--num_gpus exceed the " + std::to_string(device_count) + " visible CUDA devices");
  std::vector<Gpu> gpus(flags.num_gpus);
  struct CtxCloser {
    std::vector<Gpu> *v;
    ~CtxCloser() {
      for (Gpu &g : *v) ck_ctx_destroy(g.ctx);
    }
  } ctx_closer{&gpus};
  for (uint32_t g = 0; g < flags.num_gpus; ++g)
    if (int rc = ck_ctx_create(flags.device + int(g), &gpus[g].ctx); rc != CK_OK) return FromCk(rc);
  std::cout << " " << flags.num_gpus << " GPU(s) (" << stop_watch.ElapsedAndReset() << ")" << std::endl;

  // ---- shard planning (cuking.cu:505) and plane allocation (:513-523) ----
  // One shard: all GPUs hold its planes and split its tile grid.  --all_shards: shards are dealt round-robin to the
  // GPUs, each shard living on one GPU; the input is decoded once for all of them (the reference decodes the whole
  // input once per shard process, cuking.cu:677).
  std::vector<ShardJob> jobs;
  std::vector<std::unique_ptr<ShardOnGpu>> storage;
  struct PlanesCloser {
    std::vector<std::unique_ptr<ShardOnGpu>> *v;
    ~PlanesCloser() {
      for (auto &p : *v) ck_planes_destroy(p->planes);
    }
  } planes_closer{&storage};
  uint64_t plane_bytes = 0;
  std::cout << "Allocating memory for bit set...";
  std::cout.flush();
  const uint32_t first_shard = flags.all_shards ? 0 : flags.shard_index;
  const uint32_t last_shard = flags.all_shards ? uint32_t(num_shards) : flags.shard_index + 1;
  for (uint32_t shard = first_shard; shard < last_shard; ++shard) {
    ShardJob job;
    job.shard_index = shard;
    if (int rc = ck_submatrix_init(num_samples, flags.split_factor, shard, &job.sm); rc != CK_OK) return FromCk(rc);
    const uint32_t g_begin = flags.all_shards ? (shard - first_shard) % flags.num_gpus : 0;
    const uint32_t g_end = flags.all_shards ? g_begin + 1 : flags.num_gpus;
    for (uint32_t g = g_begin; g < g_end; ++g) {
      auto rep = std::make_unique<ShardOnGpu>();
      rep->gpu = &gpus[g];
      if (int rc = ck_planes_create(gpus[g].ctx, &job.sm, md.num_sites, &rep->planes); rc != CK_OK) return FromCk(rc);
      uint64_t b = 0;
      ck_planes_device_bytes(rep->planes, &b);
      plane_bytes += b;
      job.replicas.push_back(rep.get());
      storage.push_back(std::move(rep));
    }
    jobs.push_back(std::move(job));
  }
  std::cout << " " << ((plane_bytes + (1 << 20) - 1) >> 20) << " MiB on " << flags.num_gpus << " GPU(s) ("
            << stop_watch.ElapsedAndReset() << ")" << std::endl;

  // ---- list and decode input files, pack on the GPU ----
  std::cout << "Listing input files...";
  std::cout.flush();
  std::vector<std::string> files;
  if (std::string e = cuking::ListParquetFiles(input_dir, &files); !e.empty()) return FailedPrecondition(e);
  std::cout << " (" << stop_watch.ElapsedAndReset() << ")" << std::endl;
  if (files.empty()) return FailedPrecondition("No input files found");  // cuking.cu:542-544
  std::cout << "Found " << files.size() << " input files." << std::endl;

  std::cout << "Processing Parquet tables...";
  std::cout.flush();
  {
    std::atomic<size_t> next(0), processed(0), total_triples(0);
    std::mutex err_mu;
    Status first_error;  // first error wins, like ParallelFor (cuking.cu:415-433)
    constexpr size_t kChunkRows = size_t(1) << 20;  // 20 MiB of page-locked memory per reader thread
    auto worker = [&]() {
      cuking::Triples t;
      for (;;) {
        const size_t f = next.fetch_add(1);
        if (f >= files.size()) return;
        {
          std::lock_guard<std::mutex> l(err_mu);
          if (!first_error.ok()) return;
        }
        Status st;
        // Stream the file through this thread's page-locked chunk buffer: decode a chunk, let every shard that needs it
        // pack it on its GPU (the kernel reads the pinned chunk in place), decode the next chunk.
        auto consume = [&](size_t first_row) -> std::string {
          for (ShardJob &job : jobs) {
            for (ShardOnGpu *rep : job.replicas) {
              std::lock_guard<std::mutex> l(rep->gpu->mu);
              const int rc = ck_pack_triples(rep->planes, t.row_idx, t.col_idx, t.n_alt_alleles, t.size, /*on_device=*/0);
              if (rc != CK_OK) {
                st = FromCk(rc);
                st.message += " (chunk starting at row " + std::to_string(first_row) + ") in " + files[f];
                return st.message;
              }
            }
          }
          return "";
        };
        size_t rows = 0;
        if (std::string e = cuking::ReadTriples(files[f], kChunkRows, &t, consume, &rows); !e.empty() && st.ok())
          st = (e.rfind("Error reading", 0) == 0) ? Unknown(e) : FailedPrecondition(e);
        total_triples += rows;
        if (!st.ok()) {
          std::lock_guard<std::mutex> l(err_mu);
          if (first_error.ok()) first_error = st;
          return;
        }
        if ((++processed & ((size_t(1) << 10) - 1)) == 0) {  // progress indicator, cuking.cu:705-708
          std::cout << ".";
          std::cout.flush();
        }
      }
    };
    const size_t num_threads = std::min(flags.num_reader_threads, files.size());
    std::vector<std::thread> threads;
    for (size_t i = 0; i < num_threads; ++i) threads.emplace_back(worker);
    for (auto &th : threads) th.join();
    if (!first_error.ok()) return first_error;
    std::cout << " " << total_triples.load() << " entries (" << stop_watch.ElapsedAndReset() << ")" << std::endl;
  }

  // ---- pairwise kernel per shard ----
  const uint32_t max_results = flags.max_results;
  std::vector<ck_result> results(max_results);
  for (ShardJob &job : jobs) {
    const uint32_t rows = ck_submatrix_num_rows(&job.sm), cols = ck_submatrix_num_cols(&job.sm);
    std::cout << "Running KING CUDA kernel for " << rows << " x " << cols << " matrix";
    if (flags.all_shards) std::cout << " (shard " << job.shard_index << ")";
    std::cout << "...";
    std::cout.flush();
    uint32_t num_results = 0;
    const size_t reps = job.replicas.size();
    if (reps == 1) {
      const int rc = ck_king(job.replicas[0]->planes, flags.kin_threshold, max_results, results.data(), 0, &num_results, 1);
      if (rc != CK_OK) return FromCk(rc);
    } else {
      // split the tile grid across GPUs; each GPU returns its retained pairs, merged and sorted on the host
      uint64_t tiles = 0;
      ck_king_num_tiles(job.replicas[0]->planes, &tiles);
      std::vector<std::vector<ck_result>> part(reps);
      std::vector<uint32_t> counts(reps, 0);
      std::vector<int> rcs(reps, CK_OK);
      std::vector<std::string> errs(reps);
      std::vector<std::thread> threads;
      for (size_t g = 0; g < reps; ++g)
        threads.emplace_back([&, g]() {
          part[g].resize(max_results);
          rcs[g] = ck_king_tiles(job.replicas[g]->planes, tiles * g / reps, tiles * (g + 1) / reps, flags.kin_threshold,
                                 max_results, part[g].data(), 0, &counts[g], 1);
          if (rcs[g] != CK_OK) errs[g] = ck_last_error();
        });
      for (auto &th : threads) th.join();
      uint64_t total = 0;
      for (size_t g = 0; g < reps; ++g) {
        if (rcs[g] != CK_OK && rcs[g] != CK_ERR_RESULT_OVERFLOW) return Internal(errs[g]);
        total += counts[g];
      }
      if (total > max_results)  // cuking.cu:747-751
        return ResourceExhausted("Could not store all results: try increasing the --max_results parameter.");
      // tile slices are contiguous in (row block, column block) order, but rows of one block interleave across
      // slices only at slice boundaries: a k-way merge by (i, j) restores the global order
      std::vector<size_t> pos(reps, 0);
      for (uint64_t o = 0; o < total; ++o) {
        size_t best = reps;
        for (size_t g = 0; g < reps; ++g) {
          if (pos[g] >= counts[g]) continue;
          if (best == reps) { best = g; continue; }
          const ck_result &a = part[g][pos[g]], &b = part[best][pos[best]];
          if (a.sample_i < b.sample_i || (a.sample_i == b.sample_i && a.sample_j < b.sample_j)) best = g;
        }
        results[o] = part[best][pos[best]++];
      }
      num_results = uint32_t(total);
    }
    ck_timings tm{};
    ck_ctx_get_timings(job.replicas[0]->gpu->ctx, &tm);
    std::cout << " (" << stop_watch.ElapsedAndReset() << "; kernel " << tm.king_ms << " ms on GPU " << flags.device << ")"
              << std::endl;

    std::cout << "Processing " << num_results << " results...";
    std::cout.flush();
    std::string path;
    size_t bytes = 0;
    if (std::string e = cuking::WriteResults(output_dir, job.shard_index, md.sample_ids, results.data(), num_results,
                                             &path, &bytes);
        !e.empty())
      return Unknown(e);
    std::cout << " (" << stop_watch.ElapsedAndReset() << ")" << std::endl;
    std::cout << "Wrote " << ((bytes + (1 << 20) - 1) >> 20) << " MiB to " << path << "." << std::endl;
  }
  return Ok();
}

}  // namespace

int main(int argc, char **argv) {
  Flags flags;
  if (const std::string err = cuking::ParseFlags(argc, argv, &flags); !err.empty()) {
    std::cerr << "ERROR: " << err << std::endl;  // absl::ParseCommandLine's format
    return 1;
  }
  if (flags.help) {
    std::cout << cuking::Usage();
    return 1;  // absl exits 1 after --help
  }
  if (const Status status = Run(flags); !status.ok()) {
    std::cerr << std::endl << "Error: " << status.code << ": " << status.message << std::endl;  // cuking.cu:889-892
    return 1;
  }
  return 0;
}
