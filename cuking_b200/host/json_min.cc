#include "json_min.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace cuking {
namespace {

struct Parser {
  const char *p, *end;
  std::string err;
  int depth = 0;

  bool Fail(const std::string &m) {
    if (err.empty()) err = m;
    return false;
  }
  void Ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  static void AppendUtf8(uint32_t cp, std::string *s) {
    if (cp < 0x80) {
      s->push_back(char(cp));
    } else if (cp < 0x800) {
      s->push_back(char(0xC0 | (cp >> 6)));
      s->push_back(char(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
      s->push_back(char(0xE0 | (cp >> 12)));
      s->push_back(char(0x80 | ((cp >> 6) & 0x3F)));
      s->push_back(char(0x80 | (cp & 0x3F)));
    } else {
      s->push_back(char(0xF0 | (cp >> 18)));
      s->push_back(char(0x80 | ((cp >> 12) & 0x3F)));
      s->push_back(char(0x80 | ((cp >> 6) & 0x3F)));
      s->push_back(char(0x80 | (cp & 0x3F)));
    }
  }
  bool Hex4(uint32_t *out) {
    if (end - p < 4) return Fail("truncated \\u escape");
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
      else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
      else return Fail("bad \\u escape");
    }
    *out = v;
    return true;
  }
  bool String(std::string *out) {
    if (p >= end || *p != '"') return Fail("expected string");
    ++p;
    while (p < end) {
      const unsigned char c = static_cast<unsigned char>(*p++);
      if (c == '"') return true;
      if (c < 0x20) return Fail("control character in string");
      if (c != '\\') {
        out->push_back(char(c));
        continue;
      }
      if (p >= end) break;
      const char e = *p++;
      switch (e) {
        case '"': out->push_back('"'); break;
        case '\\': out->push_back('\\'); break;
        case '/': out->push_back('/'); break;
        case 'b': out->push_back('\b'); break;
        case 'f': out->push_back('\f'); break;
        case 'n': out->push_back('\n'); break;
        case 'r': out->push_back('\r'); break;
        case 't': out->push_back('\t'); break;
        case 'u': {
          uint32_t cp;
          if (!Hex4(&cp)) return false;
          if (cp >= 0xD800 && cp <= 0xDBFF && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
            p += 2;
            uint32_t lo;
            if (!Hex4(&lo)) return false;
            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else return Fail("unpaired surrogate");
          }
          AppendUtf8(cp, out);
          break;
        }
        default: return Fail("bad escape");
      }
    }
    return Fail("unterminated string");
  }
  bool Number(JsonValue *v) {
    const char *s = p;
    if (p < end && *p == '-') ++p;
    if (p >= end || !(*p >= '0' && *p <= '9')) return Fail("bad number");
    while (p < end && *p >= '0' && *p <= '9') ++p;
    bool integer = true;
    if (p < end && *p == '.') {
      integer = false;
      ++p;
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
      integer = false;
      ++p;
      if (p < end && (*p == '+' || *p == '-')) ++p;
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    const std::string tok(s, p);
    v->kind = JsonValue::kNumber;
    v->number = strtod(tok.c_str(), nullptr);
    v->number_is_integer = integer && tok.size() <= 18;
    if (v->number_is_integer) v->integer = strtoll(tok.c_str(), nullptr, 10);
    return true;
  }
  bool Value(JsonValue *v) {
    if (++depth > 256) return Fail("nesting too deep");
    Ws();
    if (p >= end) return Fail("unexpected end of input");
    bool ok = true;
    if (*p == '{') {
      ++p;
      v->kind = JsonValue::kObject;
      Ws();
      if (p < end && *p == '}') {
        ++p;
      } else {
        while (ok) {
          Ws();
          std::string key;
          if (!String(&key)) { ok = false; break; }
          Ws();
          if (p >= end || *p != ':') { ok = Fail("expected ':'"); break; }
          ++p;
          JsonValue child;
          if (!Value(&child)) { ok = false; break; }
          v->object[key] = std::move(child);
          Ws();
          if (p < end && *p == ',') { ++p; continue; }
          if (p < end && *p == '}') { ++p; break; }
          ok = Fail("expected ',' or '}'");
        }
      }
    } else if (*p == '[') {
      ++p;
      v->kind = JsonValue::kArray;
      Ws();
      if (p < end && *p == ']') {
        ++p;
      } else {
        while (ok) {
          JsonValue child;
          if (!Value(&child)) { ok = false; break; }
          v->array.push_back(std::move(child));
          Ws();
          if (p < end && *p == ',') { ++p; continue; }
          if (p < end && *p == ']') { ++p; break; }
          ok = Fail("expected ',' or ']'");
        }
      }
    } else if (*p == '"') {
      v->kind = JsonValue::kString;
      ok = String(&v->string);
    } else if (end - p >= 4 && !strncmp(p, "true", 4)) {
      v->kind = JsonValue::kBool; v->boolean = true; p += 4;
    } else if (end - p >= 5 && !strncmp(p, "false", 5)) {
      v->kind = JsonValue::kBool; v->boolean = false; p += 5;
    } else if (end - p >= 4 && !strncmp(p, "null", 4)) {
      v->kind = JsonValue::kNull; p += 4;
    } else {
      ok = Number(v);
    }
    --depth;
    return ok;
  }
};

}  // namespace

bool ParseJson(const std::string &text, JsonValue *out, std::string *error) {
  Parser ps{text.data(), text.data() + text.size(), {}, 0};
  if (!ps.Value(out)) {
    if (error) *error = ps.err;
    return false;
  }
  ps.Ws();
  if (ps.p != ps.end) {
    if (error) *error = "trailing characters after JSON document";
    return false;
  }
  return true;
}

}  // namespace cuking
