#include "parquet_io.h"

#include <arrow/io/file.h>
#include <parquet/api/reader.h>
#include <parquet/api/writer.h>
#include <parquet/column_page.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>

namespace cuking {
namespace fs = std::filesystem;

Triples::~Triples() { ck_host_free(block_); }

std::string Triples::Reserve(size_t n) {
  size = n;
  if (n <= capacity_) return "";
  ck_host_free(block_);
  block_ = nullptr;
  capacity_ = 0;
  const size_t cap = std::max<size_t>(n, size_t(1) << 12);
  const size_t cap_pad = (cap + 15) & ~size_t(15);  // column starts stay 16-byte aligned
  const size_t per_row = narrow_ ? 4 + 4 + 1 : 8 + 8 + 4;
  if (ck_host_alloc(cap_pad * per_row, &block_) != CK_OK) return std::string("Cannot allocate pinned host memory: ") + ck_last_error();
  capacity_ = cap_pad;
  char *base = static_cast<char *>(block_);
  if (narrow_) {
    row32 = reinterpret_cast<uint32_t *>(base);
    col32 = reinterpret_cast<uint32_t *>(base + cap_pad * 4);
    alt8 = reinterpret_cast<uint8_t *>(base + cap_pad * 8);
    wide64_.resize(2 * cap_pad);
    wide32_.resize(cap_pad);
    row_idx = wide64_.data();
    col_idx = wide64_.data() + cap_pad;
    n_alt_alleles = wide32_.data();
  } else {
    row_idx = reinterpret_cast<int64_t *>(base);
    col_idx = reinterpret_cast<int64_t *>(base + cap_pad * 8);
    n_alt_alleles = reinterpret_cast<int32_t *>(base + cap_pad * 16);
  }
  return "";
}

void Triples::Narrow() {
  if (!narrow_) return;
  for (size_t i = 0; i < size; ++i) row32[i] = uint32_t(row_idx[i]);  // the low 32 bits, like the casts at cuking.cu:676,:680
  for (size_t i = 0; i < size; ++i) col32[i] = uint32_t(col_idx[i]);
  for (size_t i = 0; i < size; ++i) alt8[i] = uint32_t(n_alt_alleles[i]) <= 255u ? uint8_t(n_alt_alleles[i]) : uint8_t(255);
}

std::string ListParquetFiles(const std::string &dir, std::vector<std::string> *files) {
  std::error_code ec;
  if (!fs::is_directory(dir, ec)) return "Input is not a directory: " + dir;
  for (const auto &entry : fs::directory_iterator(dir, ec)) {
    if (!entry.is_regular_file()) continue;  // skips _temporary/ etc. (non-recursive, cuking.cu:530-534)
    const std::string name = entry.path().filename().string();
    if (name.size() < 8 || name.compare(name.size() - 8, 8, ".parquet") != 0) continue;  // cuking.cu:537-539
    files->push_back(entry.path().string());
  }
  if (ec) return "Failed to list " + dir + ": " + ec.message();
  std::sort(files->begin(), files->end());
  return "";
}

namespace {

// Reads exactly `want` values (fewer only at the end of the column chunk) from one typed column reader.
template <typename ReaderT, typename T>
std::string ReadValues(ReaderT *reader, bool optional, const std::string &path, T *dst, size_t want, size_t *got,
                       std::vector<int16_t> *def_scratch) {
  *got = 0;
  while (*got < want && reader->HasNext()) {
    const int64_t room = int64_t(want - *got);
    int64_t values_read = 0, levels_read = 0;
    if (optional) {
      // Spark writes these columns OPTIONAL; the reference passes null def-levels because there are no nulls
      // (cuking.cu:617-623).  Read the levels and insist on that.
      def_scratch->resize(size_t(room));
      levels_read = reader->ReadBatch(room, def_scratch->data(), nullptr, dst + *got, &values_read);
      if (values_read != levels_read) return "Null values in " + path;
    } else {
      levels_read = reader->ReadBatch(room, nullptr, nullptr, dst + *got, &values_read);
    }
    if (values_read == 0 && levels_read == 0) break;
    *got += size_t(values_read);
  }
  return "";
}

std::string CheckColumn(parquet::ColumnReader *column, parquet::Type::type want, const std::string &path) {
  if (column->type() != want)  // cuking.cu:608-612, :630-634, :652-656
    return "Expected " + parquet::TypeToString(want) + " type, found " + parquet::TypeToString(column->type()) + " in " + path;
  if (column->descr()->max_repetition_level() > 0) return "Repeated column in " + path;
  return "";
}

}  // namespace

std::string ReadTriples(const std::string &path, size_t chunk_rows, Triples *buf,
                        const std::function<std::string(size_t)> &consume, size_t *rows_out) {
  size_t delivered = 0;
  try {
    std::unique_ptr<parquet::ParquetFileReader> reader = parquet::ParquetFileReader::OpenFile(path, /*memory_map=*/false);
    const auto md = reader->metadata();
    constexpr int kNumColumns = 3;
    if (md->num_columns() != kNumColumns)  // cuking.cu:585-590
      return "Expected 3 columns, found " + std::to_string(md->num_columns()) + " in " + path;
    if (std::string e = buf->Reserve(chunk_rows); !e.empty()) return e;
    std::vector<int16_t> def_scratch;
    for (int rg = 0; rg < md->num_row_groups(); ++rg) {
      auto group = reader->RowGroup(rg);
      auto c0 = group->Column(0), c1 = group->Column(1), c2 = group->Column(2);
      std::string err;
      if (!(err = CheckColumn(c0.get(), parquet::Type::INT64, path)).empty()) return err;
      if (!(err = CheckColumn(c1.get(), parquet::Type::INT64, path)).empty()) return err;
      if (!(err = CheckColumn(c2.get(), parquet::Type::INT32, path)).empty()) return err;
      auto *r0 = static_cast<parquet::Int64Reader *>(c0.get());
      auto *r1 = static_cast<parquet::Int64Reader *>(c1.get());
      auto *r2 = static_cast<parquet::Int32Reader *>(c2.get());
      const bool o0 = c0->descr()->max_definition_level() > 0, o1 = c1->descr()->max_definition_level() > 0,
                 o2 = c2->descr()->max_definition_level() > 0;
      size_t remaining = size_t(md->RowGroup(rg)->num_rows());
      while (remaining > 0) {
        const size_t want = std::min(chunk_rows, remaining);
        size_t g0 = 0, g1 = 0, g2 = 0;
        if (!(err = ReadValues(r0, o0, path, buf->row_idx, want, &g0, &def_scratch)).empty()) return err;
        if (!(err = ReadValues(r1, o1, path, buf->col_idx, want, &g1, &def_scratch)).empty()) return err;
        if (!(err = ReadValues(r2, o2, path, buf->n_alt_alleles, want, &g2, &def_scratch)).empty()) return err;
        if (g0 != want || g1 != want || g2 != want) return "Column lengths differ from the row count in " + path;
        buf->size = want;
        buf->Narrow();
        if (!(err = consume(delivered)).empty()) return err;
        delivered += want;
        remaining -= want;
      }
    }
    if (delivered != size_t(md->num_rows())) return "Column lengths differ from the row count in " + path;
  } catch (const std::exception &e) {  // parquet::ParquetException, cuking.cu:580-583
    return "Error reading " + path + ": " + e.what();
  }
  if (rows_out) *rows_out = delivered;
  return "";
}

// ---- pages for the device decoder ---------------------------------------------------------------------------------

namespace {
// pageable, geometric growth (the page-locked memory is the window's one arena, below)
std::string GrowHeap(uint8_t **ptr, size_t *cap, size_t used, size_t want) {
  if (want <= *cap) return "";
  const size_t ncap = std::max(want + want / 2, size_t(1) << 16);
  void *fresh = aligned_alloc(64, (ncap + 63) & ~size_t(63));
  if (!fresh) return "Out of host memory";
  if (used) memcpy(fresh, *ptr, used);
  free(*ptr);
  *ptr = static_cast<uint8_t *>(fresh);
  *cap = ncap;
  return "";
}
size_t Align64(size_t x) { return (x + 63) & ~size_t(63); }
}  // namespace

EncodedWindow::~EncodedWindow() {
  for (Column &c : col) {
    free(c.bytes);
    free(c.runs);
    free(c.dict);
  }
  ck_host_free(arena_);
}

std::string EncodedWindow::Column::GrowBytes(size_t want) { return GrowHeap(&bytes, &bytes_cap, bytes_size, want); }
std::string EncodedWindow::Column::GrowDict(size_t want) { return GrowHeap(&dict, &dict_cap, 0, want); }
std::string EncodedWindow::Column::GrowRuns(size_t want) {
  uint8_t *p = reinterpret_cast<uint8_t *>(runs);
  size_t cap = runs_cap * sizeof(ck_run);
  std::string e = GrowHeap(&p, &cap, runs_size * sizeof(ck_run), want * sizeof(ck_run));
  runs = reinterpret_cast<ck_run *>(p);
  runs_cap = cap / sizeof(ck_run);
  return e;
}

namespace {
struct Cut {  // the part of a column's buffers that describes the rows before `end_row`
  size_t bytes, runs;
  uint32_t values;
};
Cut CutAt(const EncodedWindow::Column &c, uint64_t end_row) {
  for (const EncodedWindow::Page &p : c.pages)
    if (c.first_row + p.first_value >= end_row) return Cut{p.byte_begin, p.run_begin, p.first_value};
  return Cut{c.bytes_size, c.runs_size, c.num_values};
}
}  // namespace

size_t EncodedWindow::StagedBytes(uint64_t end_row) const {
  size_t total = 0;
  for (const Column &c : col) {
    const Cut cut = CutAt(c, end_row);
    total += Align64(cut.bytes + 16) + Align64((cut.runs + 1) * sizeof(ck_run)) + Align64(size_t(c.dict_len) * c.value_width);
  }
  return total;
}

std::string EncodedWindow::Stage(uint64_t first_row, uint64_t end_row, uint8_t *dst) {
  if (!dst) {  // the window's own arena: page-locking is slow (and stalls other CUDA calls), so it only ever grows
    const size_t total = StagedBytes(end_row);
    if (total > arena_cap_) {
      ck_host_free(arena_);
      arena_ = nullptr;
      arena_cap_ = 0;
      const size_t want = std::max(total + total / 2, size_t(4) << 20);
      void *fresh = nullptr;
      if (ck_host_alloc(want, &fresh) != CK_OK) return std::string("Cannot allocate pinned host memory: ") + ck_last_error();
      arena_ = static_cast<uint8_t *>(fresh);
      arena_cap_ = want;
    }
    dst = arena_;
  }
  uint8_t *p = dst;
  for (int i = 0; i < 3; ++i) {
    const Column &c = col[i];
    const Cut cut = CutAt(c, end_row);
    ck_encoded_column &out = cols[i];
    out.bytes = p;
    out.num_bytes = cut.bytes;
    if (cut.bytes) memcpy(p, c.bytes, cut.bytes);
    memset(p + cut.bytes, 0, Align64(cut.bytes + 16) - cut.bytes);
    p += Align64(cut.bytes + 16);
    ck_run *runs = reinterpret_cast<ck_run *>(p);
    if (cut.runs) memcpy(runs, c.runs, cut.runs * sizeof(ck_run));
    runs[cut.runs] = ck_run{cut.values, 0, 0, 0};  // sentinel
    out.runs = runs;
    out.num_runs = uint32_t(cut.runs);
    p += Align64((cut.runs + 1) * sizeof(ck_run));
    out.dict = c.dict_len ? p : nullptr;
    out.dict_len = c.dict_len;
    if (c.dict_len) memcpy(p, c.dict, size_t(c.dict_len) * c.value_width);
    p += Align64(size_t(c.dict_len) * c.value_width);
    out.value_width = c.value_width;
    out.skip = uint32_t(first_row - c.first_row);
  }
  num_rows = uint32_t(end_row - first_row);
  return "";
}

void EncodedWindow::Column::Reset(uint32_t width) {
  bytes_size = runs_size = 0;
  dict_len = 0;
  value_width = width;
  first_row = 0;
  num_values = 0;
  pages.clear();
}

void EncodedWindow::Column::DropBefore(uint64_t row) {
  size_t keep = 0;
  while (keep < pages.size() && first_row + pages[keep].first_value + pages[keep].num_values <= row) ++keep;
  if (keep == 0) return;
  if (keep == pages.size()) {
    first_row += num_values;
    bytes_size = runs_size = 0;
    num_values = 0;
    pages.clear();
    return;
  }
  // a page straddles the window's end: it becomes the head of the next window's table
  const Page head = pages[keep];
  const uint32_t dv = head.first_value, db = head.byte_begin, dr = head.run_begin;
  memmove(bytes, bytes + db, bytes_size - db);
  bytes_size -= db;
  memmove(runs, runs + dr, (runs_size - dr) * sizeof(ck_run));
  runs_size -= dr;
  for (size_t r = 0; r < runs_size; ++r) {
    runs[r].first_value -= dv;
    if (runs[r].kind != CK_RUN_RLE) runs[r].payload -= db;
  }
  pages.erase(pages.begin(), pages.begin() + keep);
  for (Page &p : pages) {
    p.first_value -= dv;
    p.byte_begin -= db;
    p.byte_end -= db;
    p.run_begin -= dr;
    p.run_end -= dr;
  }
  first_row += dv;
  num_values -= dv;
}

namespace {

// All definition levels of an OPTIONAL column's page must be 1 (no nulls, cuking.cu:617-623): walks the hybrid stream.
// Returns 1 = no nulls, 0 = nulls, -1 = malformed.
int LevelsAllOne(const uint8_t *data, size_t n, uint32_t num_values, std::vector<ck_run> *scratch) {
  scratch->resize(n + 2);
  uint32_t runs = 0;
  if (ck_rle_scan(data, n, 1, num_values, 0, 0, scratch->data(), uint32_t(scratch->size()), &runs) != CK_OK) return -1;
  for (uint32_t r = 0; r < runs; ++r) {
    const ck_run &run = (*scratch)[r];
    const uint32_t count = (r + 1 < runs ? (*scratch)[r + 1].first_value : num_values) - run.first_value;
    if (run.kind == CK_RUN_RLE) {
      if (run.payload != 1) return 0;
    } else {
      const uint8_t *p = data + run.payload;
      uint32_t full = count / 8, rest = count % 8;
      for (uint32_t i = 0; i < full; ++i)
        if (p[i] != 0xff) return 0;
      if (rest && (p[full] & ((1u << rest) - 1u)) != ((1u << rest) - 1u)) return 0;
    }
  }
  return 1;
}

bool DictionaryEncoding(parquet::Encoding::type e) {
  return e == parquet::Encoding::RLE_DICTIONARY || e == parquet::Encoding::PLAIN_DICTIONARY;
}

// Appends one data page to a column of the window.  Returns "" / error; *unsupported as in ReadEncoded.
std::string AppendPage(EncodedWindow::Column *c, const parquet::Page &page, bool optional, const std::string &path,
                       std::vector<ck_run> *scratch, bool *unsupported) {
  const uint8_t *data = nullptr;
  size_t size = 0;
  uint32_t num_values = 0;
  parquet::Encoding::type enc;
  if (page.type() == parquet::PageType::DATA_PAGE) {
    const auto &p = static_cast<const parquet::DataPageV1 &>(page);
    data = p.data();
    size = size_t(p.size());
    num_values = uint32_t(p.num_values());
    enc = p.encoding();
    if (optional) {  // [u32 length][hybrid definition levels]
      if (p.definition_level_encoding() != parquet::Encoding::RLE) {
        *unsupported = true;
        return "";
      }
      if (size < 4) return "Error reading " + path + ": truncated data page";
      uint32_t len;
      memcpy(&len, data, 4);
      if (size_t(len) + 4 > size) return "Error reading " + path + ": truncated definition levels";
      const int ok = LevelsAllOne(data + 4, len, num_values, scratch);
      if (ok < 0) return "Error reading " + path + ": malformed definition levels";
      if (ok == 0) return "Null values in " + path;
      data += 4 + len;
      size -= 4 + len;
    }
  } else {
    const auto &p = static_cast<const parquet::DataPageV2 &>(page);
    data = p.data();
    size = size_t(p.size());
    num_values = uint32_t(p.num_values());
    enc = p.encoding();
    const size_t rl = size_t(p.repetition_levels_byte_length()), dl = size_t(p.definition_levels_byte_length());
    if (rl + dl > size) return "Error reading " + path + ": truncated data page";
    if (optional) {
      if (p.num_nulls() != 0) return "Null values in " + path;
      const int ok = LevelsAllOne(data + rl, dl, num_values, scratch);
      if (ok < 0) return "Error reading " + path + ": malformed definition levels";
      if (ok == 0) return "Null values in " + path;
    }
    data += rl + dl;
    size -= rl + dl;
  }
  if (num_values == 0) return "";
  if (size >= (size_t(1) << 31) || uint64_t(c->num_values) + num_values > 0x7fffffffull) {
    *unsupported = true;
    return "";
  }
  EncodedWindow::Page rec{};
  rec.first_value = c->num_values;
  rec.num_values = num_values;
  rec.byte_begin = uint32_t((c->bytes_size + 7) & ~size_t(7));  // pages start 8-byte aligned (PLAIN values are read as words)
  rec.run_begin = uint32_t(c->runs_size);
  if (std::string e = c->GrowBytes(rec.byte_begin + size + 16); !e.empty()) return e;
  memset(c->bytes + c->bytes_size, 0, rec.byte_begin - c->bytes_size);
  if (DictionaryEncoding(enc)) {
    if (size < 1) return "Error reading " + path + ": truncated dictionary-encoded page";
    const uint32_t bit_width = data[0];
    memcpy(c->bytes + rec.byte_begin, data + 1, size - 1);
    c->bytes_size = rec.byte_begin + size - 1;
    if (std::string e = c->GrowRuns(c->runs_size + size + 2); !e.empty()) return e;
    uint32_t n = uint32_t(c->runs_size);
    if (ck_rle_scan(c->bytes + rec.byte_begin, size - 1, bit_width, num_values, rec.first_value, rec.byte_begin, c->runs,
                    uint32_t(c->runs_cap - 1), &n) != CK_OK)
      return "Error reading " + path + ": " + ck_last_error();
    c->runs_size = n;
  } else if (enc == parquet::Encoding::PLAIN) {
    if (size < size_t(num_values) * c->value_width) return "Error reading " + path + ": truncated PLAIN page";
    memcpy(c->bytes + rec.byte_begin, data, size_t(num_values) * c->value_width);
    c->bytes_size = rec.byte_begin + size_t(num_values) * c->value_width;
    if (std::string e = c->GrowRuns(c->runs_size + 2); !e.empty()) return e;
    c->runs[c->runs_size++] = ck_run{rec.first_value, CK_RUN_PLAIN, 0, rec.byte_begin};
  } else {
    *unsupported = true;
    return "";
  }
  rec.byte_end = uint32_t(c->bytes_size);
  rec.run_end = uint32_t(c->runs_size);
  c->num_values += num_values;
  c->pages.push_back(rec);
  return "";
}

}  // namespace

std::string ReadEncoded(const std::string &path, size_t window_rows, size_t slice_bytes, EncodedWindow *win,
                        const std::function<uint8_t *()> &acquire,
                        const std::function<std::string(size_t first_row, uint8_t *slice)> &consume, size_t *rows_out,
                        bool *unsupported) {
  size_t delivered = 0;
  *unsupported = false;
  try {
    std::unique_ptr<parquet::ParquetFileReader> reader = parquet::ParquetFileReader::OpenFile(path, /*memory_map=*/false);
    const auto md = reader->metadata();
    constexpr int kNumColumns = 3;
    if (md->num_columns() != kNumColumns)  // cuking.cu:585-590
      return "Expected 3 columns, found " + std::to_string(md->num_columns()) + " in " + path;
    const parquet::Type::type want_type[3] = {parquet::Type::INT64, parquet::Type::INT64, parquet::Type::INT32};
    bool optional[3];
    for (int c = 0; c < kNumColumns; ++c) {
      const parquet::ColumnDescriptor *d = md->schema()->Column(c);
      if (d->physical_type() != want_type[c])  // cuking.cu:608-612, :630-634, :652-656
        return "Expected " + parquet::TypeToString(want_type[c]) + " type, found " + parquet::TypeToString(d->physical_type()) + " in " + path;
      if (d->max_repetition_level() > 0) return "Repeated column in " + path;
      if (d->max_definition_level() > 1) {
        *unsupported = true;
        return "";
      }
      optional[c] = d->max_definition_level() > 0;
    }
    std::vector<ck_run> scratch;
    for (int rg = 0; rg < md->num_row_groups(); ++rg) {
      auto group = reader->RowGroup(rg);
      std::unique_ptr<parquet::PageReader> pages[3];
      for (int c = 0; c < kNumColumns; ++c) {
        pages[c] = group->GetColumnPageReader(c);
        win->col[c].Reset(c == 2 ? 4 : 8);
      }
      const uint64_t rg_rows = uint64_t(md->RowGroup(rg)->num_rows());
      uint64_t done = 0;
      while (done < rg_rows) {
        constexpr uint64_t kMinWindowRows = 4096;
        uint64_t rows = std::min<uint64_t>(window_rows, rg_rows - done), end = 0;
        for (;;) {
          end = done + rows;
          for (int c = 0; c < kNumColumns; ++c) {
            EncodedWindow::Column &col = win->col[c];
            while (col.first_row + col.num_values < end) {
              const auto t0 = std::chrono::steady_clock::now();
              std::shared_ptr<parquet::Page> page = pages[c]->NextPage();  // file read + page codec
              const auto t1 = std::chrono::steady_clock::now();
              win->s_pages += std::chrono::duration<double>(t1 - t0).count();
              struct ScanTimer {
                EncodedWindow *w;
                std::chrono::steady_clock::time_point from;
                ~ScanTimer() { w->s_scan += std::chrono::duration<double>(std::chrono::steady_clock::now() - from).count(); }
              } scan_timer{win, t1};
              if (!page) return "Column lengths differ from the row count in " + path;
              if (page->type() == parquet::PageType::DICTIONARY_PAGE) {
                const auto &dp = static_cast<const parquet::DictionaryPage &>(*page);
                if (dp.encoding() != parquet::Encoding::PLAIN && dp.encoding() != parquet::Encoding::PLAIN_DICTIONARY) {
                  *unsupported = true;
                  return "";
                }
                const size_t bytes = size_t(dp.num_values()) * col.value_width;
                if (size_t(dp.size()) < bytes) return "Error reading " + path + ": truncated dictionary page";
                if (std::string e = col.GrowDict(bytes); !e.empty()) return e;
                memcpy(col.dict, dp.data(), bytes);
                col.dict_len = uint32_t(dp.num_values());
              } else if (page->type() == parquet::PageType::DATA_PAGE || page->type() == parquet::PageType::DATA_PAGE_V2) {
                if (std::string e = AppendPage(&col, *page, optional[c], path, &scratch, unsupported); !e.empty() || *unsupported) return e;
              }  // index pages carry no values
            }
          }
          if (!acquire || win->StagedBytes(end) <= slice_bytes || rows <= kMinWindowRows) break;
          rows = std::max<uint64_t>(kMinWindowRows, rows / 2);  // the pages gathered beyond the new end wait for the next window
        }
        uint8_t *slice = nullptr;
        if (acquire && win->StagedBytes(end) <= slice_bytes) {
          const auto t0 = std::chrono::steady_clock::now();
          slice = acquire();
          win->s_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          if (!slice) return "Aborted";
        }
        {
          const auto t0 = std::chrono::steady_clock::now();
          if (std::string e = win->Stage(done, end, slice); !e.empty()) return e;
          win->s_stage += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (std::string e = consume(delivered, slice); !e.empty()) return e;
        delivered += size_t(end - done);
        done = end;
        for (int c = 0; c < kNumColumns; ++c) win->col[c].DropBefore(done);
      }
    }
    if (delivered != size_t(md->num_rows())) return "Column lengths differ from the row count in " + path;
  } catch (const std::exception &e) {  // parquet::ParquetException, cuking.cu:580-583
    return "Error reading " + path + ": " + e.what();
  }
  if (rows_out) *rows_out = delivered;
  return "";
}

struct ResultWriter::Impl {
  std::string path, tmp;
  const std::vector<std::string> *sample_ids = nullptr;
  std::shared_ptr<arrow::io::FileOutputStream> sink;
  std::shared_ptr<parquet::ParquetFileWriter> writer;
  parquet::RowGroupWriter *rg = nullptr;
  uint64_t row_group_rows = 0, rows_in_group = 0;
  std::vector<parquet::ByteArray> strings;
  std::vector<float> floats;
  std::vector<int32_t> ints;
  bool closed = false;
};

ResultWriter::ResultWriter() = default;

ResultWriter::~ResultWriter() {
  if (impl_ && !impl_->closed) {
    try {
      if (impl_->writer) impl_->writer->Close();
      if (impl_->sink) (void)impl_->sink->Close();
    } catch (...) {
    }
    std::error_code ec;
    fs::remove(impl_->tmp, ec);
  }
  delete impl_;
}

std::string ResultWriter::Open(const std::string &dir, uint32_t shard_index, const std::vector<std::string> *sample_ids,
                               uint64_t row_group_rows) {
  try {
    delete impl_;
    impl_ = new Impl();
    rows_ = 0;
    std::error_code ec;
    fs::create_directories(dir, ec);
    if (ec) return "Cannot create output directory " + dir + ": " + ec.message();
    char name[64];
    snprintf(name, sizeof(name), "part-%05u.snappy.parquet", shard_index);  // cuking.cu:868-870
    impl_->path = (fs::path(dir) / name).string();
    impl_->tmp = impl_->path + ".tmp";
    impl_->sample_ids = sample_ids;
    impl_->row_group_rows = row_group_rows;

    using parquet::schema::PrimitiveNode;
    parquet::schema::NodeVector fields;  // cuking.cu:770-788
    fields.push_back(PrimitiveNode::Make("i", parquet::Repetition::REQUIRED, parquet::LogicalType::String(), parquet::Type::BYTE_ARRAY));
    fields.push_back(PrimitiveNode::Make("j", parquet::Repetition::REQUIRED, parquet::LogicalType::String(), parquet::Type::BYTE_ARRAY));
    fields.push_back(PrimitiveNode::Make("kin", parquet::Repetition::REQUIRED, parquet::LogicalType::None(), parquet::Type::FLOAT));
    fields.push_back(PrimitiveNode::Make("ibs0", parquet::Repetition::REQUIRED, parquet::LogicalType::None(), parquet::Type::INT32));
    fields.push_back(PrimitiveNode::Make("ibs1", parquet::Repetition::REQUIRED, parquet::LogicalType::None(), parquet::Type::INT32));
    fields.push_back(PrimitiveNode::Make("ibs2", parquet::Repetition::REQUIRED, parquet::LogicalType::None(), parquet::Type::INT32));
    auto schema = std::static_pointer_cast<parquet::schema::GroupNode>(
        parquet::schema::GroupNode::Make("schema", parquet::Repetition::REQUIRED, fields));  // cuking.cu:789-791

    auto sink_result = arrow::io::FileOutputStream::Open(impl_->tmp);
    if (!sink_result.ok()) return "Cannot open " + impl_->tmp + ": " + sink_result.status().ToString();
    impl_->sink = *sink_result;
    parquet::WriterProperties::Builder props;
    props.compression(parquet::Compression::SNAPPY);  // Hail's libhadoop has no ZSTD, cuking.cu:797-798
    props.max_row_group_length(int64_t(1) << 62);
    impl_->writer = parquet::ParquetFileWriter::Open(impl_->sink, schema, props.build());
    // a buffered row group takes values for all six columns chunk after chunk (pages are compressed as they fill)
    impl_->rg = impl_->writer->AppendBufferedRowGroup();
  } catch (const std::exception &e) {
    return std::string("Error writing results: ") + e.what();
  }
  return "";
}

std::string ResultWriter::Append(const ck_result *results, size_t n) {
  if (!impl_ || !impl_->rg) return "Error writing results: writer is not open";
  Impl &w = *impl_;
  try {
    constexpr size_t kBatch = 1 << 16;  // the reference writes one value per WriteBatch call (:810-859); batch instead
    for (size_t base = 0; base < n;) {
      size_t m = std::min(kBatch, n - base);
      if (w.row_group_rows > 0) {
        if (w.rows_in_group == w.row_group_rows) {
          w.rg->Close();
          w.rg = w.writer->AppendBufferedRowGroup();
          w.rows_in_group = 0;
        }
        m = size_t(std::min<uint64_t>(m, w.row_group_rows - w.rows_in_group));
      }
      const ck_result *r = results + base;
      w.strings.resize(m);
      for (int which = 0; which < 2; ++which) {  // i, j: sample_ids[sample_i], cuking.cu:811,:821
        for (size_t q = 0; q < m; ++q) {
          const uint32_t s = which == 0 ? r[q].sample_i : r[q].sample_j;
          if (s >= w.sample_ids->size()) return "Result refers to sample " + std::to_string(s) + " beyond metadata.json";
          const std::string &id = (*w.sample_ids)[s];
          w.strings[q] = parquet::ByteArray(uint32_t(id.size()), reinterpret_cast<const uint8_t *>(id.data()));
        }
        static_cast<parquet::ByteArrayWriter *>(w.rg->column(which))->WriteBatch(int64_t(m), nullptr, nullptr, w.strings.data());
      }
      w.floats.resize(m);
      for (size_t q = 0; q < m; ++q) w.floats[q] = r[q].kin;
      static_cast<parquet::FloatWriter *>(w.rg->column(2))->WriteBatch(int64_t(m), nullptr, nullptr, w.floats.data());
      w.ints.resize(m);
      for (int which = 0; which < 3; ++which) {  // ibs0, ibs1, ibs2
        for (size_t q = 0; q < m; ++q) w.ints[q] = int32_t(which == 0 ? r[q].ibs0 : which == 1 ? r[q].ibs1 : r[q].ibs2);
        static_cast<parquet::Int32Writer *>(w.rg->column(3 + which))->WriteBatch(int64_t(m), nullptr, nullptr, w.ints.data());
      }
      w.rows_in_group += m;
      rows_ += m;
      base += m;
    }
  } catch (const std::exception &e) {
    return std::string("Error writing results: ") + e.what();
  }
  return "";
}

std::string ResultWriter::Close(std::string *path_out, size_t *bytes_written) {
  if (!impl_ || !impl_->writer) return "Error writing results: writer is not open";
  Impl &w = *impl_;
  try {
    w.rg->Close();
    w.rg = nullptr;
    w.writer->Close();
    auto st = w.sink->Close();
    if (!st.ok()) return "Cannot close " + w.tmp + ": " + st.ToString();
    std::error_code ec;
    fs::rename(w.tmp, w.path, ec);  // readers never see a partial part file
    if (ec) return "Cannot rename " + w.tmp + ": " + ec.message();
    w.closed = true;
    if (path_out) *path_out = w.path;
    if (bytes_written) *bytes_written = size_t(fs::file_size(w.path, ec));
  } catch (const std::exception &e) {
    return std::string("Error writing results: ") + e.what();
  }
  return "";
}

std::string WriteResults(const std::string &dir, uint32_t shard_index, const std::vector<std::string> &sample_ids,
                         const ck_result *results, size_t n, std::string *path_out, size_t *bytes_written) {
  ResultWriter w;
  if (std::string e = w.Open(dir, shard_index, &sample_ids, 0); !e.empty()) return e;
  if (std::string e = w.Append(results, n); !e.empty()) return e;
  return w.Close(path_out, bytes_written);
}

}  // namespace cuking
