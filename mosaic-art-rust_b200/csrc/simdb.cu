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
#include <cmath>
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
  char buf[64];
  std::string digits;
  int exp10 = 0;  // scientific exponent: value = d1.d2d3... * 10^exp10
  for (int prec = 0; prec < 9; ++prec) {
    std::snprintf(buf, sizeof buf, "%.*e", prec, (double)a);
    if (std::strtof(buf, nullptr) == a || prec == 8) {
      const char* e = std::strchr(buf, 'e');
      digits.assign((const char*)buf, (size_t)(e - buf));
      digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
      exp10 = std::atoi(e + 1);
      break;
    }
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

// ---- a JSON reader just large enough for the schema above -----------------------------------------------
struct Parser {
  const char* p;
  const char* end;
  std::string err;
  bool fail(const std::string& m) {
    if (err.empty()) err = m + " at byte " + std::to_string((size_t)(p - begin_));
    return false;
  }
  const char* begin_;
  Parser(const char* b, const char* e) : p(b), end(e), begin_(b) {}
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
      const char c = *p++;
      if (c == '"') return true;
      if (c != '\\') {
        out.push_back(c);
        continue;
      }
      if (p >= end) break;
      const char e = *p++;
      switch (e) {
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
          if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {  // surrogate pair
            p += 2;
            unsigned lo = 0;
            if (!hex4(lo)) return false;
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
  bool number(double& v) {
    ws();
    char* e = nullptr;
    errno = 0;
    v = std::strtod(p, &e);
    if (e == p) return fail("expected a number");
    p = e;
    return true;
  }
  bool uinteger(uint64_t& v) {
    ws();
    if (p >= end || *p < '0' || *p > '9') return fail("expected an unsigned integer");
    v = 0;
    while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
    return true;
  }
};

bool parse_db(Parser& ps, mosaic_simdb& db) {
  bool have[4] = {false, false, false, false};
  std::map<uint64_t, std::string> i2p;
  if (!ps.expect('{')) return false;
  if (!ps.peek('}')) {
    for (;;) {
      std::string key;
      if (!ps.string(key) || !ps.expect(':')) return false;
      if (key == "path_to_index") {
        have[0] = true;
        if (!ps.expect('{')) return false;
        if (!ps.peek('}'))
          for (;;) {
            std::string path;
            uint64_t idx = 0;
            if (!ps.string(path) || !ps.expect(':') || !ps.uinteger(idx)) return false;
            db.path_to_index[path] = idx;
            if (!ps.peek(',')) break;
            ps.expect(',');
          }
        if (!ps.expect('}')) return false;
      } else if (key == "index_to_path") {
        have[1] = true;
        if (!ps.expect('{')) return false;
        if (!ps.peek('}'))
          for (;;) {
            std::string k, path;
            if (!ps.string(k) || !ps.expect(':') || !ps.string(path)) return false;
            char* e = nullptr;
            const unsigned long long idx = std::strtoull(k.c_str(), &e, 10);
            if (k.empty() || *e != 0) return ps.fail("index_to_path key is not an integer");
            i2p[idx] = path;
            if (!ps.peek(',')) break;
            ps.expect(',');
          }
        if (!ps.expect('}')) return false;
      } else if (key == "lab_colors") {
        have[2] = true;
        if (!ps.expect('[')) return false;
        if (!ps.peek(']'))
          for (;;) {
            double lab[3] = {0, 0, 0};
            bool got[3] = {false, false, false};
            if (!ps.expect('{')) return false;
            for (;;) {
              std::string f;
              double v;
              if (!ps.string(f) || !ps.expect(':') || !ps.number(v)) return false;
              const int w = f == "l" ? 0 : f == "a" ? 1 : f == "b" ? 2 : -1;
              if (w < 0) return ps.fail("unknown field `" + f + "` in a lab colour");
              lab[w] = v;
              got[w] = true;
              if (!ps.peek(',')) break;
              ps.expect(',');
            }
            if (!ps.expect('}')) return false;
            if (!(got[0] && got[1] && got[2])) return ps.fail("lab colour without l, a and b");
            for (double v : lab) db.lab.push_back((float)v);  // serde: f64 parse, then `as f32`
            if (!ps.peek(',')) break;
            ps.expect(',');
          }
        if (!ps.expect(']')) return false;
      } else if (key == "similarities") {
        have[3] = true;
        if (!ps.expect('[')) return false;
        if (!ps.peek(']'))
          for (;;) {
            double v;
            if (!ps.number(v)) return false;
            db.sim.push_back((float)v);
            if (!ps.peek(',')) break;
            ps.expect(',');
          }
        if (!ps.expect(']')) return false;
      } else {
        return ps.fail("unknown field `" + key + "`");
      }
      if (!ps.peek(',')) break;
      ps.expect(',');
    }
  }
  if (!ps.expect('}')) return false;
  for (int i = 0; i < 4; ++i)
    if (!have[i]) return ps.fail("missing field");
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
