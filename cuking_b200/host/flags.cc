#include "flags.h"

#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace cuking {
namespace {

std::string Normalize(std::string name) {
  for (char &c : name)
    if (c == '-') c = '_';
  return name;
}

bool ParseU64(const std::string &v, uint64_t max, uint64_t *out) {
  if (v.empty() || v[0] == '-' || v[0] == '+') return false;
  errno = 0;
  char *end = nullptr;
  const unsigned long long x = strtoull(v.c_str(), &end, 0);
  if (errno != 0 || end == v.c_str() || *end != 0 || x > max) return false;
  *out = x;
  return true;
}

bool ParseBool(const std::string &v, bool *out) {
  if (v == "true" || v == "1" || v == "t" || v == "yes" || v == "y") return *out = true, true;
  if (v == "false" || v == "0" || v == "f" || v == "no" || v == "n") return *out = false, true;
  return false;
}

}  // namespace

std::string Usage() {
  return "cuking: pairwise KING relatedness on B200\n"
         "  --input_uri=DIR            directory with *.parquet (row_idx, col_idx, n_alt_alleles) + metadata.json\n"
         "  --output_uri=DIR           directory for part-<shard>.snappy.parquet\n"
         "  --requester_pays_project=  accepted for compatibility (GCS is not available in this build)\n"
         "  --num_reader_threads=36    threads decoding Parquet files\n"
         "  --max_results=10485760     capacity of the result buffer\n"
         "  --kin_threshold=0.0884     keep pairs with kin strictly above this\n"
         "  --split_factor=1           k: split the relatedness matrix into k(k+1)/2 shards\n"
         "  --shard_index=0            which shard to compute\n"
         "  --num_gpus=1               GPUs of this box: triples are dealt to them and packed once, the planes exchanged\n"
         "                             over NVLink, shards scheduled across them (a lone shard is split into parts)\n"
         "  --all_shards=false         compute every shard of --split_factor (decoding and packing the input once)\n"
         "  --write_success_file=false with --all_shards: write <output>/_SUCCESS when every shard is written\n"
         "  --row_group_rows=0         rows per Parquet row group of the output (0 = a single row group)\n"
         "  --device=0                 first CUDA device\n";
}

std::string ParseFlags(int argc, char **argv, Flags *f) {
  for (int a = 1; a < argc; ++a) {
    std::string arg = argv[a];
    if (arg == "--") break;
    if (arg.size() < 2 || arg[0] != '-') return "unexpected positional argument '" + arg + "'";
    size_t dash = arg[1] == '-' ? 2 : 1;
    std::string name = arg.substr(dash), value;
    bool has_value = false;
    const size_t eq = name.find('=');
    if (eq != std::string::npos) {
      value = name.substr(eq + 1);
      name = name.substr(0, eq);
      has_value = true;
    }
    name = Normalize(name);
    if (name == "help" || name == "helpfull" || name == "h") {
      f->help = true;
      continue;
    }
    {  // boolean flags, Abseil syntax: --flag, --noflag, --flag=true|false
      bool *target = nullptr;
      bool negate = false;
      std::string base = name;
      if (base.rfind("no", 0) == 0 && (base.substr(2) == "all_shards" || base.substr(2) == "write_success_file")) {
        base = base.substr(2);
        negate = true;
      }
      if (base == "all_shards") target = &f->all_shards;
      if (base == "write_success_file") target = &f->write_success_file;
      if (target) {
        if (negate) {
          *target = false;
        } else if (!has_value) {
          *target = true;
        } else if (!ParseBool(value, target)) {
          return "Illegal value '" + value + "' specified for flag '" + base + "'";
        }
        continue;
      }
    }
    if (!has_value) {
      if (a + 1 >= argc) return "Missing the value for the flag '" + name + "'";
      value = argv[++a];
    }
    uint64_t u = 0;
    auto bad = [&]() { return "Illegal value '" + value + "' specified for flag '" + name + "'"; };
    if (name == "input_uri") {
      f->input_uri = value;
    } else if (name == "output_uri") {
      f->output_uri = value;
    } else if (name == "requester_pays_project") {
      f->requester_pays_project = value;
    } else if (name == "num_reader_threads") {
      if (!ParseU64(value, std::numeric_limits<uint32_t>::max(), &u)) return bad();
      f->num_reader_threads = size_t(u);
    } else if (name == "max_results") {
      if (!ParseU64(value, std::numeric_limits<uint32_t>::max(), &u)) return bad();
      f->max_results = uint32_t(u);
    } else if (name == "kin_threshold") {
      errno = 0;
      char *end = nullptr;
      const float x = strtof(value.c_str(), &end);
      if (value.empty() || end == value.c_str() || *end != 0 || errno == ERANGE) return bad();
      f->kin_threshold = x;
    } else if (name == "split_factor") {
      if (!ParseU64(value, std::numeric_limits<uint32_t>::max(), &u)) return bad();
      f->split_factor = uint32_t(u);
    } else if (name == "shard_index") {
      if (!ParseU64(value, std::numeric_limits<uint32_t>::max(), &u)) return bad();
      f->shard_index = uint32_t(u);
    } else if (name == "num_gpus") {
      if (!ParseU64(value, 64, &u) || u == 0) return bad();
      f->num_gpus = uint32_t(u);
    } else if (name == "row_group_rows") {
      if (!ParseU64(value, std::numeric_limits<uint64_t>::max(), &u)) return bad();
      f->row_group_rows = u;
    } else if (name == "device") {
      if (!ParseU64(value, 1024, &u)) return bad();
      f->device = int(u);
    } else {
      return "Unknown command line flag '" + name + "'";
    }
  }
  return "";
}

}  // namespace cuking
