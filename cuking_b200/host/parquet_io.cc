#include "parquet_io.h"

#include <arrow/io/file.h>
#include <parquet/api/reader.h>
#include <parquet/api/writer.h>

#include <algorithm>
#include <cstdio>
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
