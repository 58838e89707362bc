// Parquet pages decoded on the device: the C-ABI side of ck_rle_scan / ck_pack_encoded (include/cuking_b200.h).
//
// Replaces the value decoding of the reference's ReadBatch loops (/root/reference/cuking.cu:603-672) for the encodings its
// producer writes (mt_to_cuking_inputs.py:28-31: Spark / parquet-mr, dictionary-encoded data pages with a PLAIN fallback).
// The host keeps what is inherently serial or library-bound - file I/O, the page codec (parquet::PageReader) and the walk
// over the run headers of the RLE / bit-packed hybrid streams (Parquet format, Encodings.md) - and hands the device
// (payload bytes, run table, dictionary) per column; decode_pack_kernel (pack_kernels.cu) does the rest.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "internal.cuh"

using namespace ck;

namespace {

// unsigned LEB128 (the run header of the hybrid encoding); false when the bytes run out or the value overflows 32 bits
bool read_varint(const uint8_t *data, size_t n, size_t *pos, uint32_t *out) {
  uint64_t v = 0;
  for (unsigned shift = 0; shift < 35; shift += 7) {
    if (*pos >= n) return false;
    const uint8_t b = data[(*pos)++];
    v |= uint64_t(b & 0x7fu) << shift;
    if (!(b & 0x80u)) {
      if (v > 0xffffffffull) return false;
      *out = uint32_t(v);
      return true;
    }
  }
  return false;
}

// One value of a column on the host (error messages only): the same walk as decode_rows in pack_kernels.cu.
bool host_value(const ck_encoded_column &c, uint32_t v, int64_t *out) {
  const ck_run *end = c.runs + c.num_runs;
  const ck_run *r = std::upper_bound(c.runs, end, v, [](uint32_t x, const ck_run &run) { return x < run.first_value; });
  if (r == c.runs) return false;
  --r;
  const uint32_t rel = v - r->first_value;
  auto wide = [&](const void *base, size_t idx) {
    if (c.value_width == 8) {
      int64_t x;
      memcpy(&x, static_cast<const uint8_t *>(base) + idx * 8, 8);
      return x;
    }
    int32_t x;
    memcpy(&x, static_cast<const uint8_t *>(base) + idx * 4, 4);
    return int64_t(x);
  };
  if (r->kind == CK_RUN_PLAIN) {
    *out = wide(c.bytes + r->payload, rel);
    return true;
  }
  uint32_t idx = r->payload;
  if (r->kind == CK_RUN_BITPACKED) {
    const uint64_t bit = uint64_t(r->payload) * 8 + uint64_t(rel) * r->bit_width;
    uint64_t window = 0;
    const size_t byte = size_t(bit >> 3);
    memcpy(&window, c.bytes + byte, std::min<size_t>(8, size_t(c.num_bytes) - byte));
    idx = uint32_t((window >> (bit & 7)) & (r->bit_width >= 32 ? 0xffffffffull : ((1ull << r->bit_width) - 1ull)));
  }
  if (idx >= c.dict_len) return false;
  *out = wide(c.dict, idx);
  return true;
}

// Consistency of a column description with its own buffers: the kernel trusts every table entry.
int check_column(const ck_encoded_column &c, uint32_t num_rows, int which) {
  const std::string col = "column " + std::to_string(which) + ": ";
  if (c.value_width != 4 && c.value_width != 8) return fail(CK_ERR_INVALID_ARGUMENT, col + "value_width must be 4 or 8");
  if (!c.runs || c.num_runs == 0) return fail(CK_ERR_INVALID_ARGUMENT, col + "empty run table");
  if (c.num_bytes > 0 && !c.bytes) return fail(CK_ERR_INVALID_ARGUMENT, col + "NULL payload buffer");
  if (c.num_bytes > 0xffffffffull) return fail(CK_ERR_INVALID_ARGUMENT, col + "payload buffer larger than 4 GiB");
  if (c.dict_len > 0 && !c.dict) return fail(CK_ERR_INVALID_ARGUMENT, col + "NULL dictionary");
  const uint32_t num_values = c.runs[c.num_runs].first_value;  // sentinel
  if (c.runs[0].first_value != 0) return fail(CK_ERR_INVALID_ARGUMENT, col + "the run table must start at value 0");
  if (uint64_t(c.skip) + num_rows > num_values)
    return fail(CK_ERR_INVALID_ARGUMENT, col + "skip + num_rows exceeds the " + std::to_string(num_values) + " values of the table");
  for (uint32_t r = 0; r < c.num_runs; ++r) {
    const ck_run &run = c.runs[r];
    const uint32_t next = c.runs[r + 1].first_value;
    if (next <= run.first_value) return fail(CK_ERR_INVALID_ARGUMENT, col + "run " + std::to_string(r) + " is empty or out of order");
    const uint64_t count = next - run.first_value;
    if (run.kind == CK_RUN_BITPACKED) {
      if (run.bit_width > 32) return fail(CK_ERR_INVALID_ARGUMENT, col + "bit width " + std::to_string(run.bit_width) + " in run " + std::to_string(r));
      if (uint64_t(run.payload) * 8 + count * run.bit_width > c.num_bytes * 8)
        return fail(CK_ERR_INVALID_ARGUMENT, col + "bit-packed run " + std::to_string(r) + " reaches beyond the payload bytes");
      if (c.dict_len == 0) return fail(CK_ERR_INVALID_ARGUMENT, col + "dictionary-encoded run without a dictionary");
    } else if (run.kind == CK_RUN_PLAIN) {
      if (run.payload % c.value_width != 0) return fail(CK_ERR_INVALID_ARGUMENT, col + "PLAIN run " + std::to_string(r) + " is not aligned to the value width");
      if (uint64_t(run.payload) + count * c.value_width > c.num_bytes)
        return fail(CK_ERR_INVALID_ARGUMENT, col + "PLAIN run " + std::to_string(r) + " reaches beyond the payload bytes");
    } else if (run.kind == CK_RUN_RLE) {
      if (run.payload >= c.dict_len)
        return fail(CK_ERR_INVALID_ARGUMENT, col + "dictionary index " + std::to_string(run.payload) + " of RLE run " + std::to_string(r) +
                                                 " is outside the dictionary of " + std::to_string(c.dict_len) + " values");
    } else {
      return fail(CK_ERR_INVALID_ARGUMENT, col + "unknown run kind " + std::to_string(run.kind));
    }
  }
  return CK_OK;
}

size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// A free ingest lane of the ctx for the duration of one call (created on first need).
struct LaneLease {
  ck_ctx *ctx;
  ck_ctx::IngestLane *lane = nullptr;
  explicit LaneLease(ck_ctx *c) : ctx(c) {}
  LaneLease(const LaneLease &) = delete;
  LaneLease &operator=(const LaneLease &) = delete;
  int acquire() {
    {
      std::lock_guard<std::mutex> l(ctx->lane_mu);
      if (!ctx->lanes_free.empty()) {
        lane = ctx->lanes_free.back();
        ctx->lanes_free.pop_back();
        return CK_OK;
      }
    }
    auto *fresh = new (std::nothrow) ck_ctx::IngestLane();
    if (!fresh) return fail(CK_ERR_OUT_OF_MEMORY, "host allocation failed");
    {
      std::lock_guard<std::mutex> l(ctx->lane_mu);
      ctx->lanes_all.push_back(fresh);  // owned by the ctx from here on, whatever happens below
    }
    CK_CUDA(cudaStreamCreateWithFlags(&fresh->stream, cudaStreamNonBlocking));
    CK_CUDA(cudaEventCreateWithFlags(&fresh->dep, cudaEventDisableTiming));
    CK_CUDA(cudaEventCreate(&fresh->ev[0]));
    CK_CUDA(cudaEventCreate(&fresh->ev[1]));
    CK_CUDA(cudaMalloc(&fresh->d_err, 4 * sizeof(unsigned long long)));
    CK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&fresh->h_err), 4 * sizeof(unsigned long long), cudaHostAllocDefault));
    lane = fresh;
    return CK_OK;
  }
  ~LaneLease() {
    if (!lane) return;
    cudaStreamSynchronize(lane->stream);  // error exits: nothing of this call may still be in flight when the lane is reused
    std::lock_guard<std::mutex> l(ctx->lane_mu);
    ctx->lanes_free.push_back(lane);
  }
};

}  // namespace

extern "C" {

int ck_rle_scan(const uint8_t *data, size_t num_bytes, uint32_t bit_width, uint32_t num_values, uint32_t first_value,
                uint32_t payload_base, ck_run *runs, uint32_t max_runs, uint32_t *num_runs) {
  if (!num_runs || (!runs && max_runs > 0)) return fail(CK_ERR_INVALID_ARGUMENT, "NULL run table");
  if (!data && num_bytes > 0) return fail(CK_ERR_INVALID_ARGUMENT, "NULL stream");
  if (bit_width > 32) return fail(CK_ERR_INVALID_ARGUMENT, "bit width " + std::to_string(bit_width) + " of a hybrid stream (at most 32)");
  const uint32_t value_bytes = (bit_width + 7) / 8;  // an RLE run stores its value in this many bytes
  size_t pos = 0;
  uint32_t done = 0, n = *num_runs;
  while (done < num_values) {
    uint32_t header = 0;
    if (!read_varint(data, num_bytes, &pos, &header))
      return fail(CK_ERR_INVALID_ARGUMENT, "hybrid stream ends after " + std::to_string(done) + " of " + std::to_string(num_values) + " values");
    ck_run run{};
    run.first_value = first_value + done;
    uint64_t count;
    if (header & 1u) {  // bit-packed run: (header >> 1) groups of 8 values
      count = uint64_t(header >> 1) * 8;
      const uint64_t bytes = uint64_t(header >> 1) * bit_width;
      const uint64_t take = std::min<uint64_t>(count, num_values - done);  // the last group of a page may be padding
      // writers may truncate the padding bytes of the last group: what must be present are the bits of the values taken
      if (uint64_t(pos) * 8 + take * bit_width > uint64_t(num_bytes) * 8)
        return fail(CK_ERR_INVALID_ARGUMENT, "bit-packed run reaches beyond the stream");
      run.kind = CK_RUN_BITPACKED;
      run.bit_width = bit_width;
      if (uint64_t(payload_base) + pos > 0xffffffffull) return fail(CK_ERR_INVALID_ARGUMENT, "payload offset overflows 32 bits");
      run.payload = uint32_t(payload_base + pos);
      pos = size_t(std::min<uint64_t>(uint64_t(pos) + bytes, num_bytes));
      count = take;
    } else {  // RLE run: (header >> 1) copies of one value
      count = header >> 1;
      if (pos + value_bytes > num_bytes) return fail(CK_ERR_INVALID_ARGUMENT, "RLE run reaches beyond the stream");
      uint32_t value = 0;
      memcpy(&value, data + pos, value_bytes);  // little-endian hosts only (x86-64 / aarch64)
      pos += value_bytes;
      run.kind = CK_RUN_RLE;
      run.payload = value;
      count = std::min<uint64_t>(count, num_values - done);
    }
    if (count == 0) continue;  // empty runs are legal and carry nothing
    if (n >= max_runs) {
      *num_runs = n;
      return fail(CK_ERR_OUT_OF_RANGE, "run table full (" + std::to_string(max_runs) + " entries)");
    }
    runs[n++] = run;
    done += uint32_t(count);
  }
  *num_runs = n;
  return CK_OK;
}

int ck_pack_encoded(ck_planes *pl, const ck_encoded_column cols[3], uint32_t num_rows) {
  if (!pl) return fail(CK_ERR_INVALID_ARGUMENT, "planes is NULL");
  if (!cols) return fail(CK_ERR_INVALID_ARGUMENT, "cols is NULL");
  if (num_rows == 0) return CK_OK;
  for (int c = 0; c < 3; ++c)
    if (int rc = check_column(cols[c], num_rows, c); rc != CK_OK) return rc;
  ck_ctx *ctx = pl->ctx;
  DeviceGuard guard(ctx->device);
  LaneLease lease(ctx);
  if (int rc = lease.acquire(); rc != CK_OK) return rc;
  ck_ctx::IngestLane *lane = lease.lane;
  cudaStream_t s = lane->stream;
  // the lane's stream starts behind whatever the ctx's stream has been given so far (plane allocation, reset, an earlier pack)
  CK_CUDA(cudaEventRecord(lane->dep, ctx->stream));
  CK_CUDA(cudaStreamWaitEvent(s, lane->dep, 0));

  // Device copy of the window.  When the nine pieces (payload, runs + sentinel, dictionary per column) sit close together
  // in host memory - host/parquet_io.cc stages every window in one page-locked arena - the whole span goes up with ONE
  // copy and the device pointers keep the host offsets; otherwise piece by piece into
  // [payload + 16 zero bytes | runs | dictionary] x 3.  (The funnel shift of decode_rows reads one word past a payload:
  // whatever it finds there is masked off.)
  struct Piece { const void *host; size_t bytes, align; };
  Piece pieces[3][3];
  uintptr_t lo = ~uintptr_t(0), hi = 0;
  size_t sum = 0;
  bool aligned = true;
  for (int c = 0; c < 3; ++c) {
    pieces[c][0] = Piece{cols[c].bytes, size_t(cols[c].num_bytes), 8};
    pieces[c][1] = Piece{cols[c].runs, (size_t(cols[c].num_runs) + 1) * sizeof(ck_run), 16};
    pieces[c][2] = Piece{cols[c].dict, size_t(cols[c].dict_len) * cols[c].value_width, 8};
    for (const Piece &p : pieces[c]) {
      if (p.bytes == 0) continue;
      const uintptr_t a = reinterpret_cast<uintptr_t>(p.host);
      lo = std::min(lo, a);
      hi = std::max(hi, a + p.bytes);
      sum += p.bytes;
      aligned = aligned && a % p.align == 0;
    }
  }
  // the span starts at the first byte of the lowest piece (never before it: that memory is not the caller's); the device copy
  // keeps the span's offset inside a 16-byte unit, so every piece keeps the alignment it has on the host
  const size_t skew = size_t(lo & 15);
  const bool one_copy = aligned && hi - lo <= sum + 4096;
  size_t off[3][3], total = 0;
  if (one_copy) {
    total = align16(skew + (hi - lo)) + 16;
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 3; ++k) off[c][k] = pieces[c][k].bytes ? skew + (reinterpret_cast<uintptr_t>(pieces[c][k].host) - lo) : 0;
  } else {
    for (int c = 0; c < 3; ++c) {
      off[c][0] = total;
      total += align16(pieces[c][0].bytes + 16);
      off[c][1] = total;
      total += align16(pieces[c][1].bytes);
      off[c][2] = total;
      total += align16(pieces[c][2].bytes);
    }
  }
  if (total > lane->staging_bytes) {
    if (lane->staging) {
      CK_CUDA(cudaFree(lane->staging));  // the lane is idle: every call ends with a synchronisation of its stream
      lane->staging = nullptr;
      lane->staging_bytes = 0;
    }
    const size_t want = std::max(total + total / 2, size_t(8) << 20);
    {
      std::lock_guard<std::mutex> l(ctx->lane_mu);  // dev_alloc may drop the ctx's buffer cache
      CK_CUDA(dev_alloc(ctx, &lane->staging, want));
    }
    lane->staging_bytes = want;
  }
  char *d = static_cast<char *>(lane->staging);
  CK_CUDA(cudaMemsetAsync(lane->d_err, 0xff, 4 * sizeof(unsigned long long), s));
  CK_CUDA(cudaEventRecord(lane->ev[0], s));
  if (one_copy) {
    CK_CUDA(cudaMemcpyAsync(d + skew, reinterpret_cast<const void *>(lo), hi - lo, cudaMemcpyHostToDevice, s));
  } else {
    for (int c = 0; c < 3; ++c) {
      const size_t nb = pieces[c][0].bytes;
      CK_CUDA(cudaMemsetAsync(d + off[c][0] + (nb & ~size_t(15)), 0, align16(nb + 16) - (nb & ~size_t(15)), s));
      for (int k = 0; k < 3; ++k)
        if (pieces[c][k].bytes) CK_CUDA(cudaMemcpyAsync(d + off[c][k], pieces[c][k].host, pieces[c][k].bytes, cudaMemcpyHostToDevice, s));
    }
  }
  EncodedColumnDev dev[3];
  for (int c = 0; c < 3; ++c) {
    dev[c].words = reinterpret_cast<const uint32_t *>(d + off[c][0]);
    dev[c].runs = reinterpret_cast<const ck_run *>(d + off[c][1]);
    dev[c].dict = d + off[c][2];
    dev[c].num_runs = cols[c].num_runs;
    dev[c].dict_len = cols[c].dict_len;
    dev[c].width = cols[c].value_width;
    dev[c].skip = cols[c].skip;
  }
  CK_CUDA(launch_decode_pack(*pl, dev, num_rows, lane->d_err, s));
  CK_CUDA(cudaEventRecord(lane->ev[1], s));
  CK_CUDA(cudaMemcpyAsync(lane->h_err, lane->d_err, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK_CUDA(cudaStreamSynchronize(s));
  const unsigned long long *err = lane->h_err;
  {
    std::lock_guard<std::mutex> l(ctx->lane_mu);
    pl->mark_stale();
    ctx->timings.pack_ms = elapsed_ms(lane->ev[0], lane->ev[1]);
  }
  if (err[2] != ~0ull)
    return fail(CK_ERR_INVALID_ARGUMENT, "dictionary index outside the dictionary at row " + std::to_string(size_t(err[2]) - 1) + " of the window (corrupt page)");
  if (err[0] != ~0ull) {
    const size_t idx = size_t(err[0]) - 1;
    int64_t value = 0;
    host_value(cols[2], cols[2].skip + uint32_t(idx), &value);
    return fail(CK_ERR_INVALID_GENOTYPE, "Invalid value for n_alt_alleles (" + std::to_string(int32_t(value)) + ") encountered at triple " + std::to_string(idx));
  }
  if (err[1] != ~0ull)
    return fail(CK_ERR_OUT_OF_RANGE, "row_idx out of range [0, num_sites) at triple " + std::to_string(size_t(err[1]) - 1));
  return CK_OK;
}

}  // extern "C"
