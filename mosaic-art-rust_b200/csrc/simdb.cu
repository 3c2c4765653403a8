// simdb.cu — the similarity-database FILE of the reference (host code only, no CUDA calls).
//
// reader/writer for the JSON file `SimilarityDatabase::save_to_file` /
// `load_from_file` produce and consume (reference src/similarity.rs:31-42 struct, 112-123 I/O):
//
//   {
//     "path_to_index": { "<path>": 0, ... },        HashMap<PathBuf, usize>  (key order arbitrary)
//     "index_to_path": { "0": "<path>", ... },      HashMap<usize, PathBuf>  (integer keys are JSON strings)
//     "lab_colors":   [ { "l": .., "a": .., "b": .. }, ... ],                Vec<SerializableLab>
//     "similarities": [ .. ]                        Vec<f32>, packed upper triangle (similarity.rs:100)
//   }
//
// written by serde_json::to_string_pretty (2-space indent, ": " separators, f32 through ryu's shortest
// round-trip printer; serde_json 1.0.140 / ryu — third-party, restated: [3P]).  The writer here emits the maps in
// index order (the reference's order is the HashMap's, i.e. arbitrary); the reader accepts any order, any
// whitespace, and — like serde — ignores nothing and requires all four fields.
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

struct mosaic_simdb {
  std::vector<std::string> index_to_path;          // by index ("" = no entry for that index)
  std::map<std::string, uint64_t> path_to_index;   // similarity.rs:33-34
  std::vector<float> lab;                           // 3 per row (similarity.rs:38)
  std::vector<float> sim;                           // similarity.rs:41
};

namespace mosaic {
namespace {

// ---- f32 -> text, as ryu's `Buffer::format_finite(f32)` prints it [3P] ----------------------------------
// shortest decimal digit string that parses back to the same f32, then ryu's "pretty" layout:
//   digits d1..dL, decimal exponent k (value = digits * 10^k), kk = L + k
//   0 <= k && kk <= 13   -> d1..dL 0..0 ".0"          (1234e7  -> 12340000000.0)
//   0 <  kk && kk <= 13  -> d1..d_kk "." rest         (1234e-2 -> 12.34)
//   -6 < kk && kk <= 0   -> "0." 0..0 d1..dL          (1234e-7 -> 0.0001234)
//   otherwise            -> d1 ["." rest] "e" (kk-1)  (1234e30 -> 1.234e33, 1e-7 -> 1e-7)
std::string format_f32(float f) {
  if (!std::isfinite(f)) return "null";  // serde_json serialises non-finite floats as null
  std::string out;
  if (std::signbit(f)) out.push_back('-');
  const float a = std::fabs(f);
  if (a == 0.0f) return out + "0.0";
  // shortest round-trip digits in scientific form, "d[.ddd]e[+-]XX" (std::to_chars without a precision IS the shortest
  // form — libstdc++ implements it with Ryu — and, unlike printf / strtof, does not read the process's locale)
  char buf[64];
  const auto r = std::to_chars(buf, buf + sizeof buf, a, std::chars_format::scientific);
  const char* e = std::find(buf, r.ptr, 'e');
  std::string digits((const char*)buf, (size_t)(e - buf));
  digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
  int exp10 = 0;  // scientific exponent: value = d1.d2d3... * 10^exp10
  if (e < r.ptr) {
    const char* q = e + 1;
    if (q < r.ptr && *q == '+') ++q;
    std::from_chars(q, r.ptr, exp10);
  }
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int L = (int)digits.size();
  const int kk = exp10 + 1;   // position of the decimal point relative to the first digit
  const int k = kk - L;
  if (0 <= k && kk <= 13) {
    out += digits;
    out.append((size_t)k, '0');
    out += ".0";
  } else if (0 < kk && kk <= 13) {
    out.append(digits, 0, (size_t)kk);
    out.push_back('.');
    out.append(digits, (size_t)kk, std::string::npos);
  } else if (-6 < kk && kk <= 0) {
    out += "0.";
    out.append((size_t)(-kk), '0');
    out += digits;
  } else {
    out.push_back(digits[0]);
    if (L > 1) {
      out.push_back('.');
      out.append(digits, 1, std::string::npos);
    }
    out.push_back('e');
    out += std::to_string(kk - 1);
  }
  return out;
}

void write_json_string(std::string& o, const std::string& s) {  // serde_json's escaping
  o.push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\b': o += "\\b"; break;
      case '\f': o += "\\f"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) {
          char b[8];
          std::snprintf(b, sizeof b, "\\u%04x", c);
          o += b;
        } else {
          o.push_back((char)c);
        }
    }
  }
  o.push_back('"');
}

// ---- the reader: what `serde_json::from_str::<SimilarityDatabase>(&fs::read_to_string(path)?)` accepts ---------
// (similarity.rs:119-123; serde_json 1.0.140 and serde's derive, [3P] restated).  `load_or_new` (similarity.rs:126-139)
// turns ANY error into an empty database that is rebuilt from scratch, so what loads and what does not is behaviour:
//   * the bytes are UTF-8 (`read_to_string`), the document is one value between JSON whitespace (space \n \r \t);
//   * a struct is an object — unknown fields are skipped (no `deny_unknown_fields` on the reference's structs), a
//     field given twice or not at all is an error — or an array of its fields in declaration order;
//   * numbers follow the JSON grammar (no `+`, no leading zeros, no `inf` / `nan` / hex); an f32 takes any number —
//     integer literals that fit 64 bits go integer -> f32, everything else f64 -> f32 — and a literal whose f64 is
//     infinite is "number out of range"; a `usize` takes a plain non-negative integer literal that fits 64 bits;
//   * the keys of `index_to_path` are such integer literals between the quotes, with nothing else;
//   * strings: no raw control characters, the eight short escapes, \uXXXX with properly paired surrogates;
//   * nesting deeper than 128 is "recursion limit exceeded" (only skipped values can nest here).
// Nothing is allocated from a number in the file: the tables grow with the bytes that are there.
// No libc parsing (`strtod` reads the process's locale): std::from_chars.
struct Parser {
  const char* p;
  const char* end;
  std::string err;
  const char* begin_;
  Parser(const char* b, const char* e) : p(b), end(e), begin_(b) {}
  bool fail(const std::string& m) {
    if (err.empty()) err = m + " at byte " + std::to_string((size_t)(p - begin_));
    return false;
  }
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
  }
  bool expect(char c) {
    ws();
    if (p < end && *p == c) {
      ++p;
      return true;
    }
    return fail(std::string("expected '") + c + "'");
  }
  bool peek(char c) {
    ws();
    return p < end && *p == c;
  }
  // `,` then another element (true) or the closing bracket (false); anything else is an error (ok = false)
  bool more(char close, bool& ok) {
    ws();
    if (p < end && *p == ',') {
      ++p;
      return true;
    }
    if (p < end && *p == close) {
      ++p;
      return false;
    }
    ok = fail(std::string("expected ',' or '") + close + "'");
    return false;
  }
  static void put_utf8(std::string& o, unsigned cp) {
    if (cp < 0x80) o.push_back((char)cp);
    else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 63))); }
    else if (cp < 0x10000) { o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 63))); o.push_back((char)(0x80 | (cp & 63))); }
    else { o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 63))); o.push_back((char)(0x80 | ((cp >> 6) & 63))); o.push_back((char)(0x80 | (cp & 63))); }
  }
  bool hex4(unsigned& v) {
    if (end - p < 4) return fail("truncated \\u escape");
    v = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
      else return fail("bad \\u escape");
    }
    return true;
  }
  bool string(std::string& out) {
    out.clear();
    if (!expect('"')) return false;
    while (p < end) {
      const unsigned char c = (unsigned char)*p++;
      if (c == '"') return true;
      if (c < 0x20) return fail("control character in a string");
      if (c != '\\') {
        out.push_back((char)c);
        continue;
      }
      if (p >= end) break;
      switch (*p++) {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
          unsigned cp = 0;
          if (!hex4(cp)) return false;
          if (cp >= 0xDC00 && cp < 0xE000) return fail("lone trailing surrogate in a \\u escape");
          if (cp >= 0xD800 && cp < 0xDC00) {  // a leading surrogate needs its partner
            if (end - p < 2 || p[0] != '\\' || p[1] != 'u') return fail("lone leading surrogate in a \\u escape");
            p += 2;
            unsigned lo = 0;
            if (!hex4(lo)) return false;
            if (lo < 0xDC00 || lo >= 0xE000) return fail("lone leading surrogate in a \\u escape");
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          put_utf8(out, cp);
          break;
        }
        default: return fail("bad escape");
      }
    }
    return fail("unterminated string");
  }
  // the JSON number grammar: -? (0 | [1-9][0-9]*) (. [0-9]+)? ([eE] [+-]? [0-9]+)?
  struct Num {
    const char* b;
    const char* e;
    bool negative, integer;  // integer: neither fraction nor exponent
  };
  bool scan_number(Num& n) {
    ws();
    n.b = p;
    n.negative = p < end && *p == '-';
    if (n.negative) ++p;
    if (p >= end || *p < '0' || *p > '9') return fail("expected a number");
    if (*p == '0') {
      ++p;
      if (p < end && *p >= '0' && *p <= '9') return fail("invalid number (leading zero)");
    } else {
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    n.integer = true;
    if (p < end && *p == '.') {
      n.integer = false;
      ++p;
      if (p >= end || *p < '0' || *p > '9') return fail("invalid number (no digit after the point)");
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
      n.integer = false;
      ++p;
      if (p < end && (*p == '+' || *p == '-')) ++p;
      if (p >= end || *p < '0' || *p > '9') return fail("invalid number (no digit in the exponent)");
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    n.e = p;
    return true;
  }
  static bool digits_to_u64(const char* b, const char* e, uint64_t& v) {
    v = 0;
    for (; b < e; ++b) {
      const uint64_t d = (uint64_t)(*b - '0');
      if (v > (UINT64_MAX - d) / 10) return false;
      v = v * 10 + d;
    }
    return true;
  }
  // a scanned literal as f64 (serde_json's parse_decimal / parse_exponent / f64_from_parts): an infinite result is an
  // error, one that underflows is a signed zero
  bool to_f64(const Num& n, double& v) {
    const auto r = std::from_chars(n.b, n.e, v, std::chars_format::general);
    if (r.ec == std::errc() && r.ptr == n.e) {
      if (std::isinf(v)) return fail("number out of range");
      return true;
    }
    if (r.ec != std::errc::result_out_of_range) return fail("invalid number");
    // out of f64's range: towards zero or towards infinity?  The position of the first non-zero digit plus the written
    // exponent decides (|exponent of the value| > 300 here, so the sign of that sum is the answer).
    const char* s = n.b + (n.negative ? 1 : 0);
    long long point = 0, first = 0;  // digits in front of the point; index of the first non-zero digit (1-based), 0 = none
    long long seen = 0;
    bool in_frac = false;
    const char* q = s;
    for (; q < n.e && *q != 'e' && *q != 'E'; ++q) {
      if (*q == '.') {
        in_frac = true;
        continue;
      }
      ++seen;
      if (!in_frac) ++point;
      if (!first && *q != '0') first = seen;
    }
    long long ex = 0;
    if (q < n.e) {
      ++q;
      bool eneg = false;
      if (*q == '+' || *q == '-') eneg = *q++ == '-';
      for (; q < n.e; ++q) ex = std::min<long long>(ex * 10 + (*q - '0'), 1000000000000LL);
      if (eneg) ex = -ex;
    }
    if (!first || ex + (point - first) < 0) {
      v = n.negative ? -0.0 : 0.0;
      return true;
    }
    return fail("number out of range");
  }
  // f32 field: visit_u64 / visit_i64 / visit_f64 of serde's f32 visitor, each an `as f32`
  bool f32_value(float& out) {
    Num n;
    if (!scan_number(n)) return false;
    if (n.integer) {
      uint64_t u = 0;
      const bool fits = digits_to_u64(n.b + (n.negative ? 1 : 0), n.e, u);
      if (fits && !n.negative) {
        out = (float)u;
        return true;
      }
      if (fits && n.negative && u != 0 && u <= (uint64_t)INT64_MAX + 1) {
        out = u == (uint64_t)INT64_MAX + 1 ? (float)INT64_MIN : (float)(-(int64_t)u);
        return true;
      }
    }
    double d = 0.0;
    if (!to_f64(n, d)) return false;
    out = (float)d;  // Rust's `as f32`: round to nearest, beyond the range +-inf
    return true;
  }
  // usize field: a plain non-negative integer literal
  bool usize_value(uint64_t& v) {
    Num n;
    const char* at = p;
    if (!scan_number(n)) return false;
    if (n.negative || !n.integer || !digits_to_u64(n.b, n.e, v)) {
      p = at;
      return fail("expected an unsigned integer (usize)");
    }
    return true;
  }
  // usize as a map key: the literal between the quotes, nothing else (serde_json's numeric map keys)
  bool usize_key(uint64_t& v) {
    if (!expect('"')) return false;
    if (p >= end || *p < '0' || *p > '9') return fail("expected a numeric key");
    const char* at = p;
    Num n;
    if (!scan_number(n)) return false;
    if (!n.integer || !digits_to_u64(n.b, n.e, v)) {
      p = at;
      return fail("numeric key is not a usize");
    }
    if (p >= end || *p != '"') return fail("expected '\"' after a numeric key");
    ++p;
    return true;
  }
  bool literal(const char* word) {
    const size_t len = std::strlen(word);
    if ((size_t)(end - p) < len || std::memcmp(p, word, len) != 0) return fail("expected a value");
    p += len;
    return true;
  }
  // any JSON value, validated and dropped (serde's IgnoredAny for a field the struct does not have)
  bool skip_value(int depth) {
    if (depth > 128) return fail("recursion limit exceeded");
    ws();
    if (p >= end) return fail("expected a value");
    bool ok = true;
    switch (*p) {
      case '"': {
        std::string s;
        return string(s);
      }
      case '{':
        ++p;
        if (peek('}')) return expect('}');
        do {
          std::string k;
          if (!string(k) || !expect(':') || !skip_value(depth + 1)) return false;
        } while (more('}', ok));
        return ok;
      case '[':
        ++p;
        if (peek(']')) return expect(']');
        do {
          if (!skip_value(depth + 1)) return false;
        } while (more(']', ok));
        return ok;
      case 't': return literal("true");
      case 'f': return literal("false");
      case 'n': return literal("null");
      default: {
        Num n;
        return scan_number(n);
      }
    }
  }
};

// `read_to_string`: the file has to be UTF-8 (shortest forms, no surrogates, nothing above U+10FFFF)
bool valid_utf8(const std::string& s, size_t& bad_at) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(s.data());
  const size_t n = s.size();
  for (size_t i = 0; i < n;) {
    const unsigned c = b[i];
    size_t len = 0;
    unsigned lo = 0x80, hi = 0xBF;
    if (c < 0x80) { ++i; continue; }
    else if (c >= 0xC2 && c <= 0xDF) len = 1;
    else if (c == 0xE0) len = 2, lo = 0xA0;
    else if (c >= 0xE1 && c <= 0xEC) len = 2;
    else if (c == 0xED) len = 2, hi = 0x9F;
    else if (c == 0xEE || c == 0xEF) len = 2;
    else if (c == 0xF0) len = 3, lo = 0x90;
    else if (c >= 0xF1 && c <= 0xF3) len = 3;
    else if (c == 0xF4) len = 3, hi = 0x8F;
    else { bad_at = i; return false; }
    if (i + len >= n) { bad_at = i; return false; }  // truncated sequence
    if (b[i + 1] < lo || b[i + 1] > hi) { bad_at = i; return false; }
    for (size_t k = 2; k <= len; ++k)
      if ((b[i + k] & 0xC0) != 0x80) { bad_at = i; return false; }
    i += len + 1;
  }
  return true;
}

// SerializableLab (similarity.rs:7-12): {"l": f32, "a": f32, "b": f32}, or [l, a, b]
bool parse_lab(Parser& ps, std::vector<float>& lab_out) {
  float lab[3] = {0.f, 0.f, 0.f};
  bool ok = true;
  if (ps.peek('[')) {
    ps.expect('[');
    int got = 0;
    if (!ps.peek(']'))
      do {
        if (got == 3) return ps.fail("trailing elements in a lab colour");
        if (!ps.f32_value(lab[got++])) return false;
      } while (ps.more(']', ok));
    else
      ps.expect(']');
    if (!ok) return false;
    if (got != 3) return ps.fail("lab colour with fewer than 3 elements");
  } else {
    bool got[3] = {false, false, false};
    if (!ps.expect('{')) return false;
    if (!ps.peek('}'))
      do {
        std::string f;
        if (!ps.string(f) || !ps.expect(':')) return false;
        const int w = f == "l" ? 0 : f == "a" ? 1 : f == "b" ? 2 : -1;
        if (w < 0) {
          if (!ps.skip_value(1)) return false;
          continue;
        }
        if (got[w]) return ps.fail("duplicate field `" + f + "` in a lab colour");
        if (!ps.f32_value(lab[w])) return false;
        got[w] = true;
      } while (ps.more('}', ok));
    else
      ps.expect('}');
    if (!ok) return false;
    if (!(got[0] && got[1] && got[2])) return ps.fail("lab colour without l, a and b");
  }
  lab_out.insert(lab_out.end(), lab, lab + 3);
  return true;
}

// one field of SimilarityDatabase (similarity.rs:31-42), by its position in the declaration
bool parse_field(Parser& ps, int which, mosaic_simdb& db, std::map<uint64_t, std::string>& i2p) {
  bool ok = true;
  switch (which) {
    case 0:  // path_to_index: HashMap<PathBuf, usize> — a key given twice keeps its last value
      if (!ps.expect('{')) return false;
      if (ps.peek('}')) return ps.expect('}');
      do {
        std::string path;
        uint64_t idx = 0;
        if (!ps.string(path) || !ps.expect(':') || !ps.usize_value(idx)) return false;
        db.path_to_index[path] = idx;
      } while (ps.more('}', ok));
      return ok;
    case 1:  // index_to_path: HashMap<usize, PathBuf>
      if (!ps.expect('{')) return false;
      if (ps.peek('}')) return ps.expect('}');
      do {
        uint64_t idx = 0;
        std::string path;
        if (!ps.usize_key(idx) || !ps.expect(':') || !ps.string(path)) return false;
        i2p[idx] = path;
      } while (ps.more('}', ok));
      return ok;
    case 2:  // lab_colors: Vec<SerializableLab>
      if (!ps.expect('[')) return false;
      if (ps.peek(']')) return ps.expect(']');
      do {
        if (!parse_lab(ps, db.lab)) return false;
      } while (ps.more(']', ok));
      return ok;
    default:  // similarities: Vec<f32>
      if (!ps.expect('[')) return false;
      if (ps.peek(']')) return ps.expect(']');
      do {
        float v = 0.f;
        if (!ps.f32_value(v)) return false;
        db.sim.push_back(v);
      } while (ps.more(']', ok));
      return ok;
  }
}

bool parse_db(Parser& ps, mosaic_simdb& db) {
  static const char* const names[4] = {"path_to_index", "index_to_path", "lab_colors", "similarities"};
  bool have[4] = {false, false, false, false};
  std::map<uint64_t, std::string> i2p;
  bool ok = true;
  if (ps.peek('[')) {  // the struct as a sequence of its fields
    ps.expect('[');
    int got = 0;
    if (!ps.peek(']'))
      do {
        if (got == 4) return ps.fail("trailing elements after the four fields");
        if (!parse_field(ps, got, db, i2p)) return false;
        have[got++] = true;
      } while (ps.more(']', ok));
    else
      ps.expect(']');
    if (!ok) return false;
  } else {
    if (!ps.expect('{')) return false;
    if (!ps.peek('}'))
      do {
        std::string key;
        if (!ps.string(key) || !ps.expect(':')) return false;
        int which = -1;
        for (int i = 0; i < 4; ++i)
          if (key == names[i]) which = i;
        if (which < 0) {
          if (!ps.skip_value(1)) return false;
          continue;
        }
        if (have[which]) return ps.fail("duplicate field `" + key + "`");
        if (!parse_field(ps, which, db, i2p)) return false;
        have[which] = true;
      } while (ps.more('}', ok));
    else
      ps.expect('}');
    if (!ok) return false;
  }
  for (int i = 0; i < 4; ++i)
    if (!have[i]) return ps.fail(std::string("missing field `") + names[i] + "`");
  const size_t rows = db.lab.size() / 3;
  db.index_to_path.assign(rows, std::string());
  for (auto& kv : i2p)
    if (kv.first < rows) db.index_to_path[kv.first] = kv.second;
  return true;
}

}  // namespace
}  // namespace mosaic

using namespace mosaic;

extern "C" {

int mosaic_simdb_save_json(const char* file, const char* const* paths, int n, const float* lab, const float* sim) {
  const uint64_t nsim = n > 1 ? (uint64_t)n * (uint64_t)(n - 1) / 2 : 0;
  return mosaic_simdb_save_json_n(file, paths, n, lab, sim, nsim);
}

int mosaic_simdb_save_json_n(const char* file, const char* const* paths, int n, const float* lab, const float* sim, uint64_t nsim) {
  if (!file || n < 0 || (n > 0 && (!paths || !lab))) return fail(nullptr, MOSAIC_ERR_INVALID, "simdb_save: bad arguments");
  if (nsim && !sim) return fail(nullptr, MOSAIC_ERR_INVALID, "simdb_save: similarity matrix is NULL");
  for (int i = 0; i < n; ++i) {  // serde's `impl Serialize for Path`: "path contains invalid UTF-8 characters" — and nothing is written
    size_t bad_at = 0;
    if (!paths[i]) return fail(nullptr, MOSAIC_ERR_INVALID, "simdb_save: path " + std::to_string(i) + " is NULL");
    if (!valid_utf8(paths[i], bad_at))
      return fail(nullptr, MOSAIC_ERR_INVALID, "simdb_save: path " + std::to_string(i) + " contains invalid UTF-8 characters");
  }
  FILE* f = std::fopen(file, "wb");
  if (!f) return fail(nullptr, MOSAIC_ERR_INVALID, std::string("simdb_save: cannot open ") + file + ": " + std::strerror(errno));
  std::string o;
  o.reserve(1 << 20);
  auto flush = [&](bool force) {
    if (force || o.size() > (1u << 20)) {
      std::fwrite(o.data(), 1, o.size(), f);
      o.clear();
    }
  };
  o += "{\n  \"path_to_index\": {";
  for (int i = 0; i < n; ++i) {
    o += i ? ",\n    " : "\n    ";
    write_json_string(o, paths[i]);
    o += ": " + std::to_string(i);
    flush(false);
  }
  o += n ? "\n  },\n  \"index_to_path\": {" : "},\n  \"index_to_path\": {";
  for (int i = 0; i < n; ++i) {
    o += i ? ",\n    " : "\n    ";
    o += "\"" + std::to_string(i) + "\": ";
    write_json_string(o, paths[i]);
    flush(false);
  }
  o += n ? "\n  },\n  \"lab_colors\": [" : "},\n  \"lab_colors\": [";
  for (int i = 0; i < n; ++i) {
    o += i ? ",\n    {\n" : "\n    {\n";
    o += "      \"l\": " + format_f32(lab[3 * i]) + ",\n";
    o += "      \"a\": " + format_f32(lab[3 * i + 1]) + ",\n";
    o += "      \"b\": " + format_f32(lab[3 * i + 2]) + "\n    }";
    flush(false);
  }
  o += n ? "\n  ],\n  \"similarities\": [" : "],\n  \"similarities\": [";
  for (uint64_t i = 0; i < nsim; ++i) {
    o += i ? ",\n    " : "\n    ";
    o += format_f32(sim[i]);
    flush(false);
  }
  o += nsim ? "\n  ]\n}" : "]\n}";
  flush(true);
  const bool ok = std::ferror(f) == 0;
  if (std::fclose(f) != 0 || !ok) return fail(nullptr, MOSAIC_ERR_INVALID, std::string("simdb_save: write error on ") + file);
  return MOSAIC_OK;
}

int mosaic_simdb_load_json(const char* file, mosaic_simdb** out) {
  if (!file || !out) return fail(nullptr, MOSAIC_ERR_INVALID, "simdb_load: bad arguments");
  *out = nullptr;
  FILE* f = std::fopen(file, "rb");
  if (!f) return fail(nullptr, MOSAIC_ERR_INVALID, std::string("simdb_load: cannot open ") + file + ": " + std::strerror(errno));
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
  std::fclose(f);
  size_t bad_at = 0;
  if (!valid_utf8(text, bad_at))  // std::fs::read_to_string (similarity.rs:120)
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string("simdb_load: ") + file + ": not UTF-8 at byte " + std::to_string(bad_at));
  mosaic_simdb* db = new mosaic_simdb();
  Parser ps(text.data(), text.data() + text.size());
  bool ok = parse_db(ps, *db);
  if (ok) {
    ps.ws();
    if (ps.p != ps.end) ok = ps.fail("trailing characters");
  }
  if (!ok) {
    delete db;
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string("simdb_load: ") + file + ": " + ps.err);
  }
  *out = db;
  return MOSAIC_OK;
}

void mosaic_simdb_free(mosaic_simdb* db) { delete db; }

int mosaic_simdb_info(const mosaic_simdb* db, int* n_rows, uint64_t* n_similarities, int* n_paths) {
  if (!db) return MOSAIC_ERR_INVALID;
  if (n_rows) *n_rows = (int)(db->lab.size() / 3);
  if (n_similarities) *n_similarities = (uint64_t)db->sim.size();
  if (n_paths) *n_paths = (int)db->path_to_index.size();
  return MOSAIC_OK;
}

const char* mosaic_simdb_path(const mosaic_simdb* db, int index) {
  if (!db || index < 0 || (size_t)index >= db->index_to_path.size()) return nullptr;
  return db->index_to_path[(size_t)index].c_str();
}

int64_t mosaic_simdb_index_of(const mosaic_simdb* db, const char* path) {
  if (!db || !path) return -1;
  auto it = db->path_to_index.find(path);
  return it == db->path_to_index.end() ? -1 : (int64_t)it->second;
}

int mosaic_simdb_copy(const mosaic_simdb* db, float* lab_out, float* sim_out) {
  if (!db) return MOSAIC_ERR_INVALID;
  if (lab_out && !db->lab.empty()) std::memcpy(lab_out, db->lab.data(), db->lab.size() * sizeof(float));
  if (sim_out && !db->sim.empty()) std::memcpy(sim_out, db->sim.data(), db->sim.size() * sizeof(float));
  return MOSAIC_OK;
}

// SimilarityDatabase::get_similarity (similarity.rs:85-103): Some(0.0) for the same index, the packed entry for
// (min, max), None (-> rc MOSAIC_ERR_STATE) for an unknown path or a position beyond the stored matrix.
int mosaic_simdb_get_similarity(const mosaic_simdb* db, const char* path1, const char* path2, float* out) {
  if (!db || !path1 || !path2 || !out) return MOSAIC_ERR_INVALID;
  const int64_t a = mosaic_simdb_index_of(db, path1), b = mosaic_simdb_index_of(db, path2);
  if (a < 0 || b < 0) return MOSAIC_ERR_STATE;
  if (a == b) {
    *out = 0.0f;
    return MOSAIC_OK;
  }
  const uint64_t i = (uint64_t)std::min(a, b), j = (uint64_t)std::max(a, b), n = db->lab.size() / 3;
  const uint64_t pos = i * n - i * (i + 1) / 2 + j - i - 1;
  if (pos >= db->sim.size()) return MOSAIC_ERR_STATE;
  *out = db->sim[pos];
  return MOSAIC_OK;
}

// Lab means + packed matrix of a context -> the reference's file (what main.rs:122-131 does after building).
int mosaic_simdb_save_ctx(mosaic_ctx* c, const char* file, const char* const* paths) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!c->have_mat_lab) return fail(c, MOSAIC_ERR_STATE, "simdb_save_ctx: material Lab means have not been computed");
  const size_t n = (size_t)c->n, nsim = n > 1 ? n * (n - 1) / 2 : 0;
  std::vector<float> lab(3 * n), sim(nsim);
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  if (c->sim_lab_override) {
    // the matrix is computed from the DATABASE's colours (mosaic_set_similarity_lab): those are the lab_colors the file
    // must hold next to it (similarity.rs:31-42 — a database stores the colours its similarities were computed from)
    std::vector<float4> packed(n);
    MOSAIC_CUDA(c, cudaMemcpyAsync(packed.data(), c->d_sim_lab4_own, sizeof(float4) * n, cudaMemcpyDeviceToHost, c->stream));
    MOSAIC_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; ++i) lab[3 * i] = packed[i].x, lab[3 * i + 1] = packed[i].y, lab[3 * i + 2] = packed[i].z;
  } else {
    MOSAIC_CUDA(c, cudaMemcpyAsync(lab.data(), c->d_mat_lab, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    MOSAIC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  int rc = mosaic_similarity_tri(c, sim.data());
  if (rc != MOSAIC_OK) return rc;
  rc = mosaic_simdb_save_json(file, paths, (int)n, lab.data(), sim.data());
  if (rc != MOSAIC_OK) c->err = mosaic_last_error(nullptr);
  return rc;
}

}  // extern "C"
