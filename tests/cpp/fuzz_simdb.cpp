// fuzz_simdb.cpp — mutation fuzzer for the similarity-database reader (csrc/simdb.cu: mosaic_simdb_load_json and the
// accessors behind it), built with -fsanitize=address,undefined by tests/test_simdb_fuzz.py.  `similarity_db.json` sits on
// disk between runs (reference src/main.rs:120-143 loads whatever is there, and the GUI edits it): a damaged file must end
// in an error code — where serde_json's `from_str` returns Err (similarity.rs:121) — never in a crash, an out-of-bounds
// access or an allocation sized by a number in the file.
//   usage: fuzz_simdb <trials> <scratch file>
// The seeds are written by the library's own writer (a fresh, a stale and an empty database, paths with escapes and
// non-ASCII bytes).  Mutations: byte flips, token-aware substitutions (digits, quotes, brackets, commas, minus signs,
// exponents, backslashes), splices of one part of the file into another, deletions, truncation, huge indices.
// A file the reader ACCEPTS is then walked through every accessor with in-range and out-of-range arguments.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "mosaic_b200.h"

// the few symbols simdb.cu takes from the rest of the library (context error plumbing and the one entry point that needs a
// device); nothing the fuzzer calls reaches them with a context
struct mosaic_ctx;
namespace mosaic {
int fail(mosaic_ctx*, int code, const std::string&) { return code; }
int cuda_fail(mosaic_ctx*, cudaError_t, const char*) { return MOSAIC_ERR_CUDA; }
}  // namespace mosaic
extern "C" int mosaic_similarity_tri(mosaic_ctx*, float*) { return MOSAIC_ERR_CUDA; }
extern "C" const char* mosaic_last_error(const mosaic_ctx*) { return ""; }

static std::string slurp(const char* file) {
  std::string s;
  if (FILE* f = std::fopen(file, "rb")) {
    char buf[65536];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, got);
    std::fclose(f);
  }
  return s;
}
static void spit(const char* file, const std::string& s) {
  FILE* f = std::fopen(file, "wb");
  if (!f) std::exit(4);
  std::fwrite(s.data(), 1, s.size(), f);
  std::fclose(f);
}

static std::vector<std::string> make_seeds(const char* scratch) {
  std::vector<std::string> seeds;
  std::mt19937 rng(7);
  auto rnd = [&](float lo, float hi) { return lo + (hi - lo) * (float)(rng() >> 8) * (1.0f / 16777216.0f); };
  for (int variant = 0; variant < 4; ++variant) {
    const int n = variant == 0 ? 0 : variant == 1 ? 1 : variant == 2 ? 5 : 9;
    std::vector<std::string> names;
    for (int i = 0; i < n; ++i) {
      std::string p = "materials/img_" + std::to_string(i) + ".png";
      if (i == 1) p = "dir with \"quotes\"\\and\\backslashes/\t\x01.jpg";
      if (i == 2) p = "\xE5\x86\x99\xE7\x9C\x9F/\xF0\x9F\x96\xBC.webp";  // multi-byte UTF-8
      if (i == 3) p = "";
      names.push_back(p);
    }
    std::vector<const char*> paths;
    for (auto& s : names) paths.push_back(s.c_str());
    std::vector<float> lab(3 * (size_t)n), sim((size_t)n * (n > 0 ? n - 1 : 0) / 2);
    for (size_t i = 0; i < lab.size(); ++i) lab[i] = i % 3 == 0 ? rnd(0.f, 100.f) : rnd(-100.f, 100.f);
    for (auto& v : sim) v = rnd(0.f, 250.f);
    if (n >= 5) lab[4] = 1e-7f, lab[5] = -0.0f, sim[0] = 3.4e38f, sim[1] = 1.17549435e-38f, sim[2] = 16777216.f;
    uint64_t n_sim = sim.size();
    if (variant == 3) n_sim = 10;  // stale list: tiles added after the last build (similarity.rs:112-116)
    if (mosaic_simdb_save_json_n(scratch, paths.data(), n, lab.data(), sim.data(), n_sim) != MOSAIC_OK) std::exit(5);
    seeds.push_back(slurp(scratch));
    mosaic_simdb* db = nullptr;
    if (mosaic_simdb_load_json(scratch, &db) != MOSAIC_OK || !db) std::exit(6);  // the writer's files must load
    mosaic_simdb_free(db);
  }
  return seeds;
}

static void walk(mosaic_simdb* db, std::mt19937& rng) {
  int rows = 0, paths = 0;
  uint64_t sims = 0;
  if (mosaic_simdb_info(db, &rows, &sims, &paths) != MOSAIC_OK) std::exit(7);
  std::vector<float> lab(3 * (size_t)rows + 1), sim(sims + 1);
  (void)mosaic_simdb_copy(db, lab.data(), sim.data());
  (void)mosaic_simdb_copy(db, nullptr, nullptr);
  std::vector<std::string> names;
  for (int i = -2; i < rows + paths + 3; ++i) {
    const char* p = mosaic_simdb_path(db, i);
    if ((p == nullptr) != (i < 0 || i >= rows)) std::exit(8);  // NULL exactly outside lab_colors' rows
    if (!p) continue;
    names.push_back(p);
    (void)mosaic_simdb_index_of(db, p);
  }
  (void)mosaic_simdb_path(db, INT32_MAX);
  (void)mosaic_simdb_path(db, INT32_MIN);
  (void)mosaic_simdb_index_of(db, "no such path");
  for (int t = 0; t < 24 && !names.empty(); ++t) {
    float out = 0.f;
    (void)mosaic_simdb_get_similarity(db, names[rng() % names.size()].c_str(), names[rng() % names.size()].c_str(), &out);
  }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const int trials = std::atoi(argv[1]);
  const char* scratch = argv[2];
  const std::vector<std::string> seeds = make_seeds(scratch);
  std::mt19937 rng(20261018);
  static const char* tokens[] = {"\"", "\\", "\\u", "\\ud800", "{", "}", "[", "]", ",", ":", "-", ".", "e", "E+", "e-999", "1e999",
                                 "0", "9", "99999999999999999999", "18446744073709551615", "4294967296", "2147483648", "-1",
                                 "null", "true", "NaN", "Infinity", " ", "\n", "\x00", "\xff", "\"l\"", "\"path_to_index\"",
                                 "\"similarities\"", "\"index_to_path\"", "\"lab_colors\""};
  const size_t n_tokens = sizeof tokens / sizeof tokens[0];
  long loaded = 0, refused = 0;
  for (size_t s = 0; s < seeds.size(); ++s) {
    for (int t = 0; t < trials; ++t) {
      std::string bad = seeds[s];
      const int edits = 1 + (int)(rng() % 3);
      for (int e = 0; e < edits && bad.size() > 2; ++e) {
        const size_t pos = rng() % bad.size();
        switch (rng() % 6) {
          case 0: bad[pos] = (char)(bad[pos] ^ (1u << (rng() % 8))); break;
          case 1: bad[pos] = (char)rng(); break;
          case 2: {  // token over the byte / inserted in front of it
            const char* tok = tokens[rng() % n_tokens];
            const size_t len = tok[0] ? std::strlen(tok) : 1;
            if (rng() & 1) bad.replace(pos, 1, tok, len); else bad.insert(pos, tok, len);
            break;
          }
          case 3: bad.erase(pos, 1 + rng() % 12); break;
          case 4: {  // splice: a stretch of the file copied somewhere else
            const size_t from = rng() % bad.size(), len = std::min<size_t>(1 + rng() % 64, bad.size() - from);
            bad.insert(pos, bad.substr(from, len));
            break;
          }
          default: {  // the digits of a number blown up (an index or a count far beyond the file)
            size_t d = pos;
            while (d < bad.size() && (bad[d] < '0' || bad[d] > '9')) ++d;
            if (d < bad.size()) bad.insert(d, std::string(1 + rng() % 18, (char)('1' + rng() % 9)));
          }
        }
      }
      if (t % 9 == 0 && bad.size() > 4) bad.resize(1 + rng() % (bad.size() - 1));
      spit(scratch, bad);
      mosaic_simdb* db = nullptr;
      const int rc = mosaic_simdb_load_json(scratch, &db);
      if (rc == MOSAIC_OK) {
        if (!db) return 9;
        walk(db, rng);
        mosaic_simdb_free(db);
        ++loaded;
      } else {
        if (db) return 10;  // an error leaves no object behind
        ++refused;
      }
    }
  }
  // a file that is not there, an empty file, and no file name
  std::remove(scratch);
  mosaic_simdb* db = nullptr;
  if (mosaic_simdb_load_json(scratch, &db) == MOSAIC_OK) return 11;
  spit(scratch, "");
  if (mosaic_simdb_load_json(scratch, &db) == MOSAIC_OK) return 12;
  std::remove(scratch);
  std::printf("fuzz_simdb: %ld loaded, %ld refused\n", loaded, refused);
  return 0;
}
