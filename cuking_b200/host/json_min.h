// Minimal JSON reader for <input>/metadata.json = {"num_sites": <int>, "samples": [<str>, ...]}
// (/root/reference/mt_to_cuking_inputs.py:40-47, read at cuking.cu:475-500).  nlohmann-json is not available here.
// Parses any RFC 8259 document into a small DOM; only what metadata.json needs is exposed.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace cuking {

struct JsonValue {
  enum Kind { kNull, kBool, kNumber, kString, kArray, kObject } kind = kNull;
  bool boolean = false;
  double number = 0;
  bool number_is_integer = false;
  int64_t integer = 0;
  std::string string;
  std::vector<JsonValue> array;
  std::map<std::string, JsonValue> object;
  const JsonValue *Find(const std::string &key) const {
    auto it = object.find(key);
    return it == object.end() ? nullptr : &it->second;
  }
};

// Returns false and fills *error on malformed input.
bool ParseJson(const std::string &text, JsonValue *out, std::string *error);

}  // namespace cuking
