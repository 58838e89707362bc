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
// --num_gpus / --all_shards / --write_success_file default to the reference behaviour of one shard on one GPU.
//
// Several GPUs (--num_gpus N) - the box-local form of cloud_batch_submit.py's one VM per shard:
//   * the input is decoded ONCE; every decoded chunk goes to ONE GPU (round-robin), which packs it into its own
//     full-size plane set; the N partial plane sets are then AND-reduced over NVLink (ck_planes_and_reduce), so every
//     triple is packed once and every GPU ends up with all planes;
//   * with --all_shards the planes hold the whole cohort and every shard is a view of them; the k(k+1)/2 shards
//     (cloud_batch_submit.py:73) are cut into work items and scheduled longest-first onto the least loaded GPU
//     (ck_plan_work), one worker thread per GPU, all GPUs busy at once; one part file per shard, then _SUCCESS;
//   * a lone shard is split into N parts (bands of rows dealt in snake order), merged on the host.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <memory>
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
  ck_planes *planes = nullptr;
  std::mutex mu;
};

// Results of one shard: written straight to the part file when the shard is one work item, else collected per part and
// merged once the last part is in.
struct ShardOutput {
  uint32_t shard_index = 0;
  ck_submatrix sm{};
  uint32_t num_parts = 1;
  std::vector<std::vector<ck_result>> parts;
  std::atomic<uint32_t> parts_done{0};
};

struct SinkState {
  cuking::ResultWriter *writer = nullptr;  // single-part shard: append to the file
  std::vector<ck_result> *collect = nullptr;  // part of a multi-part shard
  std::string error;
};

int SinkTrampoline(void *user, const ck_result *records, size_t count) {
  SinkState *st = static_cast<SinkState *>(user);
  if (st->writer) {
    st->error = st->writer->Append(records, count);
    return st->error.empty() ? 0 : 1;
  }
  st->collect->insert(st->collect->end(), records, records + count);
  return 0;
}

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
  if (num_shards > 0xffffffffull) return InvalidArgument("Invalid split factor");

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
  const uint32_t num_gpus = flags.num_gpus;
  std::vector<Gpu> gpus(num_gpus);
  struct Closer {  // planes before their ctx
    std::vector<Gpu> *v;
    ~Closer() {
      for (Gpu &g : *v) ck_planes_destroy(g.planes);
      for (Gpu &g : *v) ck_ctx_destroy(g.ctx);
    }
  } closer{&gpus};
  {  // one thread per GPU: creating a CUDA context takes about half a second, and eight in a row would be the longest phase
    std::vector<Status> created(num_gpus);
    std::vector<std::thread> init;
    for (uint32_t g = 0; g < num_gpus; ++g)
      init.emplace_back([&, g] {
        if (int rc = ck_ctx_create(flags.device + int(g), &gpus[g].ctx); rc != CK_OK) created[g] = FromCk(rc);  // ck_last_error is per thread
      });
    for (auto &th : init) th.join();
    for (const Status &st : created)
      if (!st.ok()) return st;
  }
  std::cout << " " << num_gpus << " GPU(s) (" << stop_watch.ElapsedAndReset() << ")" << std::endl;

  // ---- shard planning (cuking.cu:505) and plane allocation (:513-523) ----
  // One shard: the planes hold the shard's samples (reference behaviour).  --all_shards: they hold the whole cohort and
  // every shard is a view of them (the reference decodes and packs the whole input once per shard process, :677).
  const uint32_t first_shard = flags.all_shards ? 0 : flags.shard_index;
  const uint32_t shards_to_run = flags.all_shards ? uint32_t(num_shards) : 1;
  ck_submatrix plane_sm{};
  if (int rc = flags.all_shards ? ck_submatrix_init(num_samples, 1, 0, &plane_sm)
                                : ck_submatrix_init(num_samples, flags.split_factor, flags.shard_index, &plane_sm);
      rc != CK_OK)
    return FromCk(rc);
  std::cout << "Allocating memory for bit set...";
  std::cout.flush();
  uint64_t plane_bytes = 0;
  for (Gpu &g : gpus) {
    if (int rc = ck_planes_create(g.ctx, &plane_sm, md.num_sites, &g.planes); rc != CK_OK) return FromCk(rc);
    uint64_t b = 0;
    ck_planes_device_bytes(g.planes, &b);
    plane_bytes += b;
  }
  // Several GPUs: probe the NVLink exchange on the still all-missing planes (AND of all-ones changes nothing).  Without
  // peer access every GPU packs every chunk itself instead.
  bool exchange = num_gpus > 1;
  if (exchange) {
    std::vector<ck_planes *> all;
    for (Gpu &g : gpus) all.push_back(g.planes);
    if (ck_planes_and_reduce(all.data(), num_gpus) != CK_OK) {
      exchange = false;
      std::cout << " [no peer access between the GPUs: every GPU packs every chunk]";
    }
  }
  std::cout << " " << ((plane_bytes + (1 << 20) - 1) >> 20) << " MiB on " << num_gpus << " GPU(s) ("
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
    std::atomic<size_t> next(0), processed(0), total_triples(0), next_chunk(0);
    std::mutex err_mu;
    Status first_error;  // first error wins, like ParallelFor (cuking.cu:415-433)
    std::atomic<bool> failed(false);
    std::mutex q_mu;  // guards the window queue and the slice pool below
    std::condition_variable q_jobs, q_slices;
    auto fail_with = [&](const Status &st) {
      {
        std::lock_guard<std::mutex> l(err_mu);
        if (first_error.ok()) first_error = st;
      }
      {
        std::lock_guard<std::mutex> l(q_mu);  // readers test `failed` under q_mu before they sleep on q_slices
        failed = true;
      }
      q_slices.notify_all();
    };
    constexpr size_t kChunkRows = size_t(1) << 20;  // 20 MiB of page-locked memory per reader thread
    // CUKING_WIDE_TRIPLES=1: hand the pack kernel the columns at their physical widths (20 bytes per triple over PCIe, no
    // narrowing pass on the host) instead of the narrowed 9 bytes
    const bool narrow_triples = getenv("CUKING_WIDE_TRIPLES") == nullptr;
    // Default: the pages go to the GPU still encoded.  The reader threads only run the page codec and walk the run headers;
    // a finished window is laid out in a page-locked slice and queued, and two submitter threads per GPU hand the windows to
    // ck_pack_encoded (bit unpacking, dictionary lookup, narrowing and the pack in one kernel).  The readers never wait for
    // the GPU unless every slice is in flight.  CUKING_HOST_DECODE=1 keeps libparquet's value decoding on the host (the
    // path every unsupported file takes by itself).
    const bool device_decode = getenv("CUKING_HOST_DECODE") == nullptr;
    size_t window_rows = size_t(1) << 20;
    if (const char *v = getenv("CUKING_DECODE_WINDOW_ROWS")) window_rows = std::max<size_t>(1, strtoull(v, nullptr, 10));
    std::atomic<size_t> host_decoded_files(0);
    const size_t num_threads = std::min(flags.num_reader_threads, files.size());

    // ---- page-locked slices and the window queue ----
    constexpr size_t kSliceBytes = size_t(4) << 20;
    const bool queued = device_decode && (exchange || num_gpus == 1);  // every window goes to exactly one GPU
    const size_t num_submitters = queued ? 2 * num_gpus : 0;
    const size_t num_slices = queued ? std::min<size_t>(48, std::max<size_t>(16, 6 * num_submitters)) : 0;
    struct Job {
      ck_encoded_column cols[3];
      uint32_t num_rows = 0;
      uint8_t *slice = nullptr;
      size_t file = 0, first_row = 0;
    };
    std::deque<Job> jobs;
    std::vector<uint8_t *> free_slices, all_slices;
    bool q_closed = false;
    size_t slices_made = 0;
    // page-locking memory is slow and holds up the other CUDA calls of the process: a helper makes the slices one by one
    // while the readers are already decompressing their first pages
    std::thread slice_maker;
    if (queued)
      slice_maker = std::thread([&] {
        for (size_t i = 0; i < num_slices && !failed; ++i) {
          void *p = nullptr;
          const bool ok = ck_host_alloc(kSliceBytes, &p) == CK_OK;
          std::lock_guard<std::mutex> l(q_mu);
          ++slices_made;
          if (ok) {
            all_slices.push_back(static_cast<uint8_t *>(p));
            free_slices.push_back(static_cast<uint8_t *>(p));
          }
          q_slices.notify_all();
        }
      });
    auto acquire_slice = [&]() -> uint8_t * {
      std::unique_lock<std::mutex> l(q_mu);
      q_slices.wait(l, [&] { return failed || !free_slices.empty() || (slices_made == num_slices && all_slices.empty()); });
      if (!failed && free_slices.empty()) {  // not one slice could be page-locked
        l.unlock();
        fail_with(ResourceExhausted("Cannot allocate pinned host memory for the decode windows"));
        return nullptr;
      }
      if (failed) return nullptr;
      uint8_t *p = free_slices.back();
      free_slices.pop_back();
      return p;
    };
    std::mutex stats_mu;
    double s_pages = 0, s_scan = 0, s_stage = 0, s_wait = 0, s_pack = 0, s_pack_gpu = 0;  // CUKING_INGEST_STATS=1
    size_t windows = 0;
    auto pack_window = [&](Gpu &g, const ck_encoded_column *cols, uint32_t num_rows, size_t file, size_t first_row, double *wall, double *gpu) {
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = ck_pack_encoded(g.planes, cols, num_rows);  // thread-safe: every call runs on its own lane of the ctx
      *wall += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      ck_timings tm{};
      ck_ctx_get_timings(g.ctx, &tm);
      *gpu += tm.pack_ms * 1e-3;
      if (rc == CK_OK) return true;
      Status st = FromCk(rc);
      st.message += " (chunk starting at row " + std::to_string(first_row) + ") in " + files[file];
      fail_with(st);
      return false;
    };
    auto submitter = [&](size_t index) {
      Gpu &g = gpus[index % num_gpus];
      double wall = 0, gpu = 0;
      size_t count = 0;
      for (;;) {
        Job job;
        {
          std::unique_lock<std::mutex> l(q_mu);
          q_jobs.wait(l, [&] { return q_closed || !jobs.empty(); });
          if (jobs.empty()) break;
          job = jobs.front();
          jobs.pop_front();
        }
        if (!failed) {
          pack_window(g, job.cols, job.num_rows, job.file, job.first_row, &wall, &gpu);
          ++count;
        }
        {
          std::lock_guard<std::mutex> l(q_mu);
          free_slices.push_back(job.slice);
        }
        q_slices.notify_one();
      }
      std::lock_guard<std::mutex> l(stats_mu);
      s_pack += wall, s_pack_gpu += gpu, windows += count;
    };
    std::vector<std::thread> submitters;
    for (size_t i = 0; i < num_submitters; ++i) submitters.emplace_back(submitter, i);

    auto worker = [&]() {
      cuking::Triples t(narrow_triples);
      cuking::EncodedWindow win;
      double my_pack = 0, my_pack_gpu = 0;
      size_t my_windows = 0;
      struct Report {
        std::function<void()> fn;
        ~Report() { fn(); }
      } report{[&] {
        std::lock_guard<std::mutex> l(stats_mu);
        s_pages += win.s_pages, s_scan += win.s_scan, s_stage += win.s_stage, s_wait += win.s_wait, s_pack += my_pack, s_pack_gpu += my_pack_gpu;
        windows += my_windows;
      }};
      for (;;) {
        const size_t f = next.fetch_add(1);
        if (f >= files.size() || failed) return;
        Status st;
        // Host decode: stream the file through this thread's chunk buffers - decode a chunk, narrow it to 9 bytes per triple
        // in page-locked memory, let ONE GPU pack it (the kernel reads the pinned chunk in place over PCIe), decode the next.
        auto pack_on = [&](Gpu &g, size_t first_row) -> bool {
          std::lock_guard<std::mutex> l(g.mu);
          const int rc = t.narrow() ? ck_pack_triples_narrow(g.planes, t.row32, t.col32, t.alt8, t.size, /*on_device=*/0)
                                    : ck_pack_triples(g.planes, t.row_idx, t.col_idx, t.n_alt_alleles, t.size, /*on_device=*/0);
          if (rc == CK_OK) return true;
          st = FromCk(rc);
          if (rc == CK_ERR_INVALID_GENOTYPE) {  // report the value as decoded (a byte cannot hold every int32), cuking.cu:698-701
            const std::string msg = ck_last_error();
            const size_t at = msg.rfind("triple ");
            const size_t idx = at == std::string::npos ? t.size : size_t(strtoull(msg.c_str() + at + 7, nullptr, 10));
            if (idx < t.size)
              st.message = "Invalid value for n_alt_alleles (" + std::to_string(t.n_alt_alleles[idx]) + ") encountered at triple " + std::to_string(idx);
          }
          st.message += " (chunk starting at row " + std::to_string(first_row) + ") in " + files[f];
          return false;
        };
        auto consume = [&](size_t first_row) -> std::string {
          if (exchange) {
            if (!pack_on(gpus[next_chunk.fetch_add(1) % num_gpus], first_row)) return st.message;
          } else {
            for (Gpu &g : gpus)
              if (!pack_on(g, first_row)) return st.message;
          }
          return "";
        };
        // Device decode: queue the window (its slice goes with it), or - no queue, or a window too large for a slice - pack it
        // from the window's own arena before going on.
        auto consume_encoded = [&](size_t first_row, uint8_t *slice) -> std::string {
          if (slice) {
            Job job;
            for (int c = 0; c < 3; ++c) job.cols[c] = win.cols[c];
            job.num_rows = win.num_rows, job.slice = slice, job.file = f, job.first_row = first_row;
            {
              std::lock_guard<std::mutex> l(q_mu);
              jobs.push_back(job);
            }
            q_jobs.notify_one();
            return failed ? "Aborted" : "";
          }
          ++my_windows;
          if (exchange || num_gpus == 1)
            return pack_window(gpus[next_chunk.fetch_add(1) % num_gpus], win.cols, win.num_rows, f, first_row, &my_pack, &my_pack_gpu) ? "" : "Aborted";
          for (Gpu &g : gpus)
            if (!pack_window(g, win.cols, win.num_rows, f, first_row, &my_pack, &my_pack_gpu)) return "Aborted";
          return "";
        };
        size_t rows = 0;
        bool host_decode = !device_decode;
        if (device_decode) {
          bool unsupported = false;
          const std::string e = cuking::ReadEncoded(files[f], window_rows, kSliceBytes, &win, queued ? std::function<uint8_t *()>(acquire_slice) : nullptr,
                                                    consume_encoded, &rows, &unsupported);
          if (e == "Aborted") return;  // another thread (or the GPU side of this one) has already recorded the error
          if (!e.empty()) st = (e.rfind("Error reading", 0) == 0) ? Unknown(e) : FailedPrecondition(e);
          host_decode = unsupported && st.ok();  // the windows already packed are packed again: harmless, the pack is an AND
        }
        if (host_decode) {
          ++host_decoded_files;
          rows = 0;
          if (std::string e = cuking::ReadTriples(files[f], kChunkRows, &t, consume, &rows); !e.empty() && st.ok())
            st = (e.rfind("Error reading", 0) == 0) ? Unknown(e) : FailedPrecondition(e);
        }
        total_triples += rows;
        if (!st.ok()) {
          fail_with(st);
          return;
        }
        if ((++processed & ((size_t(1) << 10) - 1)) == 0) {  // progress indicator, cuking.cu:705-708
          std::cout << ".";
          std::cout.flush();
        }
      }
    };
    std::vector<std::thread> threads;
    for (size_t i = 0; i < num_threads; ++i) threads.emplace_back(worker);
    for (auto &th : threads) th.join();
    {
      std::lock_guard<std::mutex> l(q_mu);
      q_closed = true;
    }
    q_jobs.notify_all();
    for (auto &th : submitters) th.join();
    if (slice_maker.joinable()) slice_maker.join();
    for (uint8_t *p : all_slices) ck_host_free(p);
    if (!first_error.ok()) return first_error;
    std::cout << " " << total_triples.load() << " entries";
    if (device_decode && host_decoded_files.load() > 0) std::cout << " [" << host_decoded_files.load() << " file(s) decoded on the host]";
    std::cout << " (" << stop_watch.ElapsedAndReset() << ")" << std::endl;
    if (getenv("CUKING_INGEST_STATS"))
      std::cout << "Ingest thread-seconds over " << num_threads << " readers + " << num_submitters << " submitters: page read + codec " << s_pages
                << ", run scan + copy " << s_scan << ", staging " << s_stage << ", waiting for a free slice " << s_wait << ", ck_pack_encoded "
                << s_pack << " (GPU " << s_pack_gpu << ") in " << windows << " windows" << std::endl;
  }
  if (exchange) {
    std::cout << "Exchanging bit sets between " << num_gpus << " GPUs...";
    std::cout.flush();
    std::vector<ck_planes *> all;
    for (Gpu &g : gpus) all.push_back(g.planes);
    if (int rc = ck_planes_and_reduce(all.data(), num_gpus); rc != CK_OK) return FromCk(rc);
    std::cout << " (" << stop_watch.ElapsedAndReset() << ")" << std::endl;
  }

  // ---- work items: (shard, part) scheduled onto the GPUs ----
  uint32_t num_items = 0;
  if (int rc = ck_plan_work(num_samples, flags.split_factor, first_shard, shards_to_run, num_gpus, nullptr, 0, &num_items); rc != CK_OK)
    return FromCk(rc);
  std::vector<ck_work_item> items(std::max<uint32_t>(num_items, 1));
  if (int rc = ck_plan_work(num_samples, flags.split_factor, first_shard, shards_to_run, num_gpus, items.data(), num_items, &num_items);
      rc != CK_OK)
    return FromCk(rc);
  items.resize(num_items);
  std::vector<std::unique_ptr<ShardOutput>> outputs(shards_to_run);
  for (uint32_t q = 0; q < shards_to_run; ++q) {
    outputs[q] = std::make_unique<ShardOutput>();
    outputs[q]->shard_index = first_shard + q;
    if (int rc = ck_submatrix_init(num_samples, flags.split_factor, first_shard + q, &outputs[q]->sm); rc != CK_OK) return FromCk(rc);
  }
  for (const ck_work_item &it : items) {
    ShardOutput &o = *outputs[it.shard_index - first_shard];
    o.num_parts = it.num_parts;
    o.parts.resize(it.num_parts);
  }

  // ---- pairwise kernel: one worker thread per GPU, every GPU runs its items in order ----
  std::mutex log_mu, err_mu;
  Status first_error;
  auto fail_with = [&](const Status &st) {
    std::lock_guard<std::mutex> l(err_mu);
    if (first_error.ok()) first_error = st;
  };
  auto failed = [&]() {
    std::lock_guard<std::mutex> l(err_mu);
    return !first_error.ok();
  };
  // writes one shard's part file from its merged parts (k-way merge by (i, j): bands of different parts interleave)
  auto write_merged = [&](ShardOutput &o) -> Status {
    uint64_t total = 0;
    for (const auto &p : o.parts) total += p.size();
    if (total > flags.max_results)  // cuking.cu:747-751 applies to the shard as a whole
      return ResourceExhausted("Could not store all results: try increasing the --max_results parameter.");
    cuking::ResultWriter writer;
    if (std::string e = writer.Open(output_dir, o.shard_index, &md.sample_ids, flags.row_group_rows); !e.empty()) return Unknown(e);
    std::vector<size_t> pos(o.parts.size(), 0);
    std::vector<ck_result> batch;
    batch.reserve(1 << 16);
    for (uint64_t done = 0; done < total; ++done) {
      size_t best = o.parts.size();
      for (size_t g = 0; g < o.parts.size(); ++g) {
        if (pos[g] >= o.parts[g].size()) continue;
        if (best == o.parts.size()) { best = g; continue; }
        const ck_result &a = o.parts[g][pos[g]], &b = o.parts[best][pos[best]];
        if (a.sample_i < b.sample_i || (a.sample_i == b.sample_i && a.sample_j < b.sample_j)) best = g;
      }
      batch.push_back(o.parts[best][pos[best]++]);
      if (batch.size() == batch.capacity() || done + 1 == total) {
        if (std::string e = writer.Append(batch.data(), batch.size()); !e.empty()) return Unknown(e);
        batch.clear();
      }
    }
    std::string path;
    size_t bytes = 0;
    if (std::string e = writer.Close(&path, &bytes); !e.empty()) return Unknown(e);
    std::lock_guard<std::mutex> l(log_mu);
    std::cout << "Processing " << total << " results... shard " << o.shard_index << ": wrote " << ((bytes + (1 << 20) - 1) >> 20)
              << " MiB to " << path << "." << std::endl;
    return Ok();
  };
  auto gpu_worker = [&](uint32_t g) {
    for (const ck_work_item &it : items) {
      if (it.gpu != g || failed()) continue;
      ShardOutput &o = *outputs[it.shard_index - first_shard];
      const auto t0 = std::chrono::steady_clock::now();
      SinkState sink;
      cuking::ResultWriter writer;
      if (it.num_parts == 1) {
        if (std::string e = writer.Open(output_dir, o.shard_index, &md.sample_ids, flags.row_group_rows); !e.empty()) return fail_with(Unknown(e));
        sink.writer = &writer;
      } else {
        sink.collect = &o.parts[it.part_index];
      }
      uint64_t count = 0;
      const int rc = ck_king_view_sink(gpus[g].planes, flags.all_shards ? &o.sm : nullptr, it.part_index, it.num_parts,
                                       flags.kin_threshold, flags.max_results, 0, SinkTrampoline, &sink, &count);
      if (rc != CK_OK) return fail_with(sink.error.empty() ? FromCk(rc) : Unknown(sink.error));
      ck_timings tm{};
      ck_ctx_get_timings(gpus[g].ctx, &tm);
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      std::string path;
      size_t bytes = 0;
      if (it.num_parts == 1)
        if (std::string e = writer.Close(&path, &bytes); !e.empty()) return fail_with(Unknown(e));
      {
        std::lock_guard<std::mutex> l(log_mu);
        std::cout << "Running KING CUDA kernel for " << ck_submatrix_num_rows(&o.sm) << " x " << ck_submatrix_num_cols(&o.sm) << " matrix";
        if (flags.all_shards || it.num_parts > 1)
          std::cout << " (shard " << o.shard_index << ", part " << it.part_index + 1 << "/" << it.num_parts << ")";
        char buf[96];
        snprintf(buf, sizeof(buf), "... (%.4gs; kernel %.4g ms on GPU %d)", secs, double(tm.king_ms), flags.device + int(g));
        std::cout << buf << std::endl;
        if (it.num_parts == 1)
          std::cout << "Processing " << count << " results... wrote " << ((bytes + (1 << 20) - 1) >> 20) << " MiB to " << path << "." << std::endl;
      }
      if (it.num_parts > 1 && o.parts_done.fetch_add(1) + 1 == it.num_parts)  // the last part in merges and writes the shard
        if (Status st = write_merged(o); !st.ok()) return fail_with(st);
    }
  };
  {
    std::vector<std::thread> threads;
    for (uint32_t g = 0; g < num_gpus; ++g) threads.emplace_back(gpu_worker, g);
    for (auto &th : threads) th.join();
  }
  if (!first_error.ok()) return first_error;
  std::cout << "Computed " << shards_to_run << " shard(s) in " << items.size() << " work item(s) on " << num_gpus << " GPU(s) ("
            << stop_watch.ElapsedAndReset() << ")" << std::endl;

  if (flags.write_success_file) {  // cloud_batch_submit.py:103-127: an empty _SUCCESS once every shard is written
    const std::string path = output_dir + "/_SUCCESS";
    std::ofstream out(path, std::ios::binary | std::ios::trunc);
    if (!out) return Unknown("Cannot write " + path);
    std::cout << "Wrote " << path << "." << std::endl;
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
