// Test tool (no GPU): checks the host half of the device page decoder.  For every file given, the rows that
// cuking::ReadEncoded describes window by window (page payloads + run tables + dictionaries, interpreted here by a plain
// loop that shares no code with the kernel) must equal the rows cuking::ReadTriples decodes through libparquet.
//   encoded_check <window_rows>[:<slice_bytes>] <file>...      prints "OK <rows> rows, <windows> windows, ..." per file
// With <slice_bytes> the windows are staged in a lent slice of that size (windows that do not fit are halved by the reader).
// ck_host_alloc / ck_host_free are replaced by malloc / free below so that the tool runs without a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../cuking_b200/host/parquet_io.h"

extern "C" int ck_host_alloc(size_t bytes, void **out) {
  *out = malloc(bytes ? bytes : 1);
  return *out ? CK_OK : CK_ERR_OUT_OF_MEMORY;
}
extern "C" int ck_host_free(void *p) {
  free(p);
  return CK_OK;
}

static bool Value(const ck_encoded_column &c, uint32_t v, int64_t *out, uint32_t *cursor) {
  uint32_t r = (*cursor < c.num_runs && c.runs[*cursor].first_value <= v) ? *cursor : 0;
  while (r + 1 < c.num_runs && c.runs[r + 1].first_value <= v) ++r;  // a linear walk on purpose (values are asked for in order)
  *cursor = r;
  const ck_run &run = c.runs[r];
  const uint32_t rel = v - run.first_value;
  auto wide = [&](const uint8_t *base, size_t i) {
    if (c.value_width == 8) { int64_t x; memcpy(&x, base + 8 * i, 8); return x; }
    int32_t x; memcpy(&x, base + 4 * i, 4); return int64_t(x);
  };
  if (run.kind == CK_RUN_PLAIN) { *out = wide(c.bytes + run.payload, rel); return true; }
  uint64_t idx = run.payload;
  if (run.kind == CK_RUN_BITPACKED) {
    idx = 0;
    const uint64_t bit0 = uint64_t(run.payload) * 8 + uint64_t(rel) * run.bit_width;
    for (uint32_t b = 0; b < run.bit_width; ++b) {
      const uint64_t bit = bit0 + b;
      if (bit / 8 >= c.num_bytes) return false;
      idx |= uint64_t((c.bytes[bit / 8] >> (bit % 8)) & 1u) << b;
    }
  }
  if (idx >= c.dict_len) return false;
  *out = wide(static_cast<const uint8_t *>(c.dict), idx);
  return true;
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  char *colon = nullptr;
  const size_t window_rows = strtoull(argv[1], &colon, 10);
  const size_t slice_bytes = (colon && *colon == ':') ? strtoull(colon + 1, nullptr, 10) : 0;
  std::vector<uint8_t> slice_mem(slice_bytes + 64);
  uint8_t *slice_base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(slice_mem.data()) + 63) & ~uintptr_t(63));
  size_t in_slices = 0;
  for (int a = 2; a < argc; ++a) {
    const std::string path = argv[a];
    std::vector<int64_t> row, col, alt;
    cuking::Triples t(false);
    size_t rows = 0;
    std::string e = cuking::ReadTriples(path, size_t(1) << 16, &t, [&](size_t) {
      row.insert(row.end(), t.row_idx, t.row_idx + t.size);
      col.insert(col.end(), t.col_idx, t.col_idx + t.size);
      for (size_t i = 0; i < t.size; ++i) alt.push_back(t.n_alt_alleles[i]);
      return std::string();
    }, &rows);
    if (!e.empty()) { printf("HOST_ERROR %s\n", e.c_str()); }
    cuking::EncodedWindow win;
    size_t seen = 0, windows = 0, runs = 0, rows2 = 0, col_runs[3] = {0, 0, 0}, bytes = 0;
    bool unsupported = false;
    std::string bad;
    std::function<uint8_t *()> acquire;
    if (slice_bytes) acquire = [&]() { return slice_base; };
    std::string e2 = cuking::ReadEncoded(path, window_rows, slice_bytes, &win, acquire, [&](size_t first, uint8_t *slice) {
      if (slice) {
        ++in_slices;
        if (win.cols[0].bytes != slice) bad = "window not staged in the lent slice";
        const uint8_t *last = static_cast<const uint8_t *>(win.cols[2].dict ? win.cols[2].dict : static_cast<const void *>(win.cols[2].runs + win.cols[2].num_runs + 1));
        if (last + size_t(win.cols[2].dict_len) * 4 > slice + slice_bytes) bad = "window overflows the slice";
      }
      if (first != seen) bad = "window starts at " + std::to_string(first) + ", expected " + std::to_string(seen);
      ++windows;
      for (int c = 0; c < 3; ++c) {
        const ck_encoded_column &ec = win.cols[c];
        runs += ec.num_runs;
        col_runs[c] += ec.num_runs;
        bytes += ec.num_bytes + (size_t(ec.num_runs) + 1) * sizeof(ck_run) + size_t(ec.dict_len) * ec.value_width;
        const std::vector<int64_t> &want = c == 0 ? row : c == 1 ? col : alt;
        if (ec.runs[0].first_value != 0 || ec.runs[ec.num_runs].first_value < ec.skip + win.num_rows) bad = "table bounds";
        uint32_t cursor = 0;
        for (uint32_t r = 0; r < win.num_rows && bad.empty(); ++r) {
          int64_t v = 0;
          if (!Value(ec, ec.skip + r, &v, &cursor)) bad = "undecodable value";
          else if (seen + r >= want.size() || v != want[seen + r])
            bad = "column " + std::to_string(c) + " row " + std::to_string(seen + r) + ": " + std::to_string(v);
        }
      }
      seen += win.num_rows;
      return bad;
    }, &rows2, &unsupported);
    if (unsupported) { printf("UNSUPPORTED %s\n", path.c_str()); continue; }
    if (!e2.empty()) { printf("%s %s\n", e.empty() ? "ERROR" : (e == e2 ? "SAME_ERROR" : "OTHER_ERROR"), e2.c_str()); continue; }
    if (!e.empty()) { printf("ERROR host path failed where the encoded path did not\n"); continue; }
    if (rows2 != rows || seen != rows) { printf("ERROR rows %zu vs %zu\n", rows2, rows); continue; }
    printf("OK %zu rows, %zu windows (%zu in slices), %zu runs (%zu + %zu + %zu), %.2f bytes per row to the device\n", rows, windows, in_slices,
           runs, col_runs[0], col_runs[1], col_runs[2], double(bytes) / double(rows ? rows : 1));
    in_slices = 0;
  }
  return 0;
}
