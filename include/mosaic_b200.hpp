// mosaic_b200.hpp — C++17 host-side mirror of the reference's interface for the matching path, over the C ABI
// of mosaic_b200.h (header-only; link libmosaic_b200.so).
//
// The reference is a Rust crate and there is no rustc in this image, so the host side that would live in Rust
//  is mirrored here in C++ with the reference's names, argument meaning and error behaviour:
//
//   SimilarityDatabase           src/similarity.rs:31-138   add_tile, build_similarities, get_similarity,
//                                                           get_lab_color, save_to_file, load_from_file, load_or_new
//                                src/similarity.rs:141-178  calculate_lab_distance, calculate_delta_e_2000 (free functions)
//   AdjacencyPenaltyCalculator   src/adjacency.rs:49-214    new(db, weight), calculate_penalty, calculate_total_cost(grid),
//                                                           calculate_swap_delta(grid, pos1, pos2)
//   OptimizationConfig/Result,
//   MosaicOptimizer              src/optimizer.rs:7-220           new(calculator, config), optimize(grid), optimize_greedy(grid, n)
//   MosaicGeneratorImpl          src/lib.rs:91-119          calculate_average_lab(image)
//   MosaicGenerator              src/main.rs:92-685         new(...), generate_mosaic(...)  (the CLI's private struct,
//                                                           minus directory scan and codecs: images come in memory)
//
// Not mirrored: `color_adjustment::{ColorAdjustment, calculate_optimal_adjustment}` (color_adjustment.rs:6-154) as
// free-standing items — their one caller on the path is the per-cell adjust of generate_mosaic (main.rs:551-567, 636-652),
// which is inside mosaic_render / mosaic_generate (K4's compose pass) with the strength as its only parameter; in the Rust
// crate they stay as they are (integration/rust/src/lib_compat.rs).
//
// `Option<PathBuf>` grids are std::vector<std::vector<std::optional<std::string>>>; `anyhow::Result` errors are
// exceptions (mosaic_b200::Error carries the status code and mosaic_last_error's text).  Every computation of the
// hot path is forwarded to the GPU library; what stays here is bookkeeping (path <-> index maps) and the single-cell
// calculate_penalty query (≤4 terms over GPU-built similarities; the placement kernel has its own fused copy).
// The reference's signatures carry no device handle, so SimilarityDatabase / AdjacencyPenaltyCalculator /
// MosaicOptimizer run on a lazily created process-wide scratch Context (device MOSAIC_B200_DEVICE, default 0);
// MosaicGenerator owns the Context that holds its materials.  Without an sm_100a device every compute call throws
// Error(MOSAIC_ERR_CUDA): there is no CPU fallback.
// Differences from the reference are the ones listed in INTEGRATION.md: the annealer takes a seed (the reference
// draws from the unseedable thread_rng, optimizer.rs:48), materials are uniform-size RGB8 buffers.
#ifndef MOSAIC_B200_HPP
#define MOSAIC_B200_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "mosaic_b200.h"

namespace mosaic_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

struct Lab {  // palette::Lab<D65, f32>
  float l = 0, a = 0, b = 0;
};

using PathGrid = std::vector<std::vector<std::optional<std::string>>>;  // Vec<Vec<Option<PathBuf>>>, [y][x]

// RAII mosaic_ctx
class Context {
 public:
  explicit Context(int device = 0) {
    const int rc = mosaic_ctx_create(&ctx_, device);
    if (rc != MOSAIC_OK) throw Error(rc, mosaic_last_error(nullptr));
  }
  ~Context() { mosaic_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  mosaic_ctx* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != MOSAIC_OK) throw Error(rc, mosaic_last_error(ctx_));
  }

 private:
  mosaic_ctx* ctx_ = nullptr;
};

namespace detail {
// one lazily created scratch Context per device; device < 0: MOSAIC_B200_DEVICE, default 0.  Throws Error when there is
// no sm_100a GPU.  Every helper that uses it injects ALL the state it needs (Lab means, grid, placement), so the
// helpers can be called in any order (mosaic_set_material_lab replaces the material set when the count changes).
inline Context& scratch_context(int device = -1) {
  static std::map<int, std::unique_ptr<Context>> ctxs;
  if (device < 0) {
    const char* dev = std::getenv("MOSAIC_B200_DEVICE");
    device = dev ? std::atoi(dev) : 0;
  }
  auto& ctx = ctxs[device];
  if (!ctx) ctx = std::make_unique<Context>(device);
  return *ctx;
}
}  // namespace detail

// ---- src/similarity.rs -----------------------------------------------------------------------------------
class SimilarityDatabase {
 public:
  SimilarityDatabase() = default;  // SimilarityDatabase::new()

  void add_tile(const std::string& path, Lab lab_color) {  // similarity.rs:60-66
    const size_t index = lab_.size() / 3;
    path_to_index_[path] = index;
    if (index_to_path_.size() <= index) index_to_path_.resize(index + 1);
    index_to_path_[index] = path;
    lab_.insert(lab_.end(), {lab_color.l, lab_color.a, lab_color.b});
  }

  // similarity.rs:69-82: the packed upper-triangular matrix of all current tiles, on the GPU (K3)
  void build_similarities() { build_similarities(detail::scratch_context()); }
  void build_similarities(Context& ctx) {
    const int n = (int)(lab_.size() / 3);
    sim_.assign(n > 1 ? (size_t)n * (n - 1) / 2 : 0, 0.0f);
    if (n == 0) return;  // (the reference underflows n * (n - 1) here, similarity.rs:71)
    ctx.check(mosaic_set_material_lab(ctx.get(), lab_.data(), n));
    ctx.check(mosaic_similarity_tri(ctx.get(), sim_.empty() ? nullptr : sim_.data()));
  }

  std::optional<float> get_similarity(const std::string& path1, const std::string& path2) const {  // similarity.rs:85-103
    const auto a = path_to_index_.find(path1), b = path_to_index_.find(path2);
    if (a == path_to_index_.end() || b == path_to_index_.end()) return std::nullopt;
    if (a->second == b->second) return 0.0f;
    const uint64_t i = std::min(a->second, b->second), j = std::max(a->second, b->second), n = lab_.size() / 3;
    const uint64_t pos = i * n - i * (i + 1) / 2 + j - i - 1;
    if (pos >= sim_.size()) return std::nullopt;
    return sim_[pos];
  }

  std::optional<Lab> get_lab_color(const std::string& path) const {  // similarity.rs:106-109
    const auto it = path_to_index_.find(path);
    if (it == path_to_index_.end() || it->second * 3 + 2 >= lab_.size()) return std::nullopt;
    return Lab{lab_[3 * it->second], lab_[3 * it->second + 1], lab_[3 * it->second + 2]};
  }

  void save_to_file(const std::string& file) const {  // similarity.rs:112-116
    // like the reference, whatever `similarities` holds is written — also a list that is out of date because tiles were
    // added after the last build_similarities()
    std::vector<const char*> paths;
    for (const auto& p : index_to_path_) paths.push_back(p.c_str());
    const int rc = mosaic_simdb_save_json_n(file.c_str(), paths.data(), (int)paths.size(), lab_.data(), sim_.data(),
                                            (uint64_t)sim_.size());
    if (rc != MOSAIC_OK) throw Error(rc, mosaic_last_error(nullptr));
  }

  static SimilarityDatabase load_from_file(const std::string& file) {  // similarity.rs:119-123
    mosaic_simdb* h = nullptr;
    const int rc = mosaic_simdb_load_json(file.c_str(), &h);
    if (rc != MOSAIC_OK) throw Error(rc, mosaic_last_error(nullptr));
    SimilarityDatabase db;
    int rows = 0, npaths = 0;
    uint64_t nsim = 0;
    mosaic_simdb_info(h, &rows, &nsim, &npaths);
    db.lab_.resize((size_t)rows * 3);
    db.sim_.resize(nsim);
    mosaic_simdb_copy(h, db.lab_.data(), db.sim_.data());
    db.index_to_path_.resize(rows);
    for (int i = 0; i < rows; ++i) {
      const char* p = mosaic_simdb_path(h, i);
      db.index_to_path_[i] = p ? p : "";
      if (p && *p) db.path_to_index_[p] = (size_t)mosaic_simdb_index_of(h, p);
    }
    mosaic_simdb_free(h);
    return db;
  }

  static SimilarityDatabase load_or_new(const std::string& file) {  // similarity.rs:126-137
    try {
      return load_from_file(file);
    } catch (const Error&) {
      return SimilarityDatabase();
    }
  }

  size_t len() const { return lab_.size() / 3; }
  const std::vector<float>& lab_colors() const { return lab_; }
  int64_t index_of(const std::string& path) const {
    const auto it = path_to_index_.find(path);
    return it == path_to_index_.end() ? -1 : (int64_t)it->second;
  }

 private:
  std::map<std::string, size_t> path_to_index_;
  std::vector<std::string> index_to_path_;
  std::vector<float> lab_, sim_;
};

namespace detail {
// grid of paths -> material indices of `db`; -1 = empty cell.  A path the database does not know is an OCCUPIED cell
// for the reference (the optimiser swaps it, draws random numbers for it and cools down, optimizer.rs:71-102; only its
// similarity terms are skipped because get_similarity returns None — also against itself, adjacency.rs:104-124,
// similarity.rs:85-90): `unknown` collects those cells' paths, and every such CELL gets its own pseudo index
// >= db.len(); the callers extend the Lab table with colours so far apart that every squared distance overflows to
// +inf (unknown_lab): 1/(sqrt(inf) + 1) = 0, the term the reference skips contributes +0.0, and x + 0.0 == x.
inline std::vector<int32_t> grid_ids(const SimilarityDatabase& db, const PathGrid& grid, int* gw, int* gh,
                                     std::vector<std::string>* unknown = nullptr) {
  *gh = (int)grid.size();
  *gw = *gh ? (int)grid[0].size() : 0;
  std::vector<int32_t> ids;
  ids.reserve((size_t)*gw * *gh);
  for (const auto& row : grid)
    for (const auto& cell : row) {
      if (!cell) {
        ids.push_back(-1);
        continue;
      }
      const int64_t i = db.index_of(*cell);
      if (i >= 0) {
        ids.push_back((int32_t)i);
      } else if (unknown) {
        ids.push_back((int32_t)(db.len() + unknown->size()));
        unknown->push_back(*cell);
      } else {
        ids.push_back(-1);
      }
    }
  return ids;
}
// Lab colour of pseudo material i (see grid_ids): differences of at least 3e19 in L, whose squares exceed FLT_MAX
inline void unknown_lab(size_t i, float out[3]) {
  out[0] = 3.0e19f * (float)(i + 1);
  out[1] = 0.0f;
  out[2] = 0.0f;
}
}  // namespace detail

// similarity.rs:141-146: Euclidean distance of two Lab colours, `(dl*dl + da*da + db*db).sqrt()` in f32.  One pair; the
// n(n-1)/2 pairs of build_similarities are K3's (similarity.cu), with this operation order.
inline float calculate_lab_distance(const Lab& lab1, const Lab& lab2) {
  const float dl = lab1.l - lab2.l, da = lab1.a - lab2.a, db = lab1.b - lab2.b;
  const float s = dl * dl + da * da;
  return std::sqrt(s + db * db);
}

// similarity.rs:150-178: the reference's simplified CIE2000 difference (`#[allow(dead_code)]` there: nothing on the path
// calls it; public surface, kept with its operation order)
inline float calculate_delta_e_2000(const Lab& lab1, const Lab& lab2) {
  const float kl = 1.0f, kc = 1.0f, kh = 1.0f;
  const float dl = std::fabs(lab2.l - lab1.l);
  const float da = lab2.a - lab1.a;
  const float db = lab2.b - lab1.b;
  const float c1 = std::sqrt(lab1.a * lab1.a + lab1.b * lab1.b);
  const float c2 = std::sqrt(lab2.a * lab2.a + lab2.b * lab2.b);
  const float dc = std::fabs(c2 - c1);
  const float dh2 = da * da + db * db - dc * dc;
  const float dh = dh2 > 0.0f ? std::sqrt(dh2) : 0.0f;
  const float sl = 1.0f;
  const float c_avg = (c1 + c2) / 2.0f;
  const float sc = 1.0f + 0.045f * c_avg;
  const float sh = 1.0f + 0.015f * c_avg;
  const float dl_kl_sl = dl / (kl * sl);
  const float dc_kc_sc = dc / (kc * sc);
  const float dh_kh_sh = dh / (kh * sh);
  return std::sqrt(dl_kl_sl * dl_kl_sl + dc_kc_sc * dc_kc_sc + dh_kh_sh * dh_kh_sh);
}

// ---- src/adjacency.rs -------------------------------------------------------------------------------------
struct GridPosition {  // adjacency.rs:4-47
  size_t x = 0, y = 0;
  GridPosition() = default;
  GridPosition(size_t x_, size_t y_) : x(x_), y(y_) {}
  bool operator==(const GridPosition& o) const { return x == o.x && y == o.y; }
  bool operator!=(const GridPosition& o) const { return !(*this == o); }
  // up, down, left, right — the order the penalty terms are added in (adjacency.rs:17-45)
  std::vector<GridPosition> get_adjacent_positions(size_t grid_width, size_t grid_height) const {
    std::vector<GridPosition> adjacent;
    if (y > 0) adjacent.emplace_back(x, y - 1);
    if (y + 1 < grid_height) adjacent.emplace_back(x, y + 1);
    if (x > 0) adjacent.emplace_back(x - 1, y);
    if (x + 1 < grid_width) adjacent.emplace_back(x + 1, y);
    return adjacent;
  }
};

class AdjacencyPenaltyCalculator {
 public:
  AdjacencyPenaltyCalculator(const SimilarityDatabase& similarity_db, float penalty_weight)  // adjacency.rs:55-60
      : db_(similarity_db), weight_(penalty_weight) {}

  // adjacency.rs:63-88: weight * sum over the placed up/down/left/right neighbours of 1 / (similarity + 1).  A one-cell
  // query over the GPU-built matrix (the placement kernel evaluates the same expression fused, greedy.cu).
  float calculate_penalty(const std::string& candidate_path, GridPosition position, const PathGrid& grid, size_t grid_width,
                          size_t grid_height) const {
    float penalty = 0.0f;
    for (const GridPosition& adj : position.get_adjacent_positions(grid_width, grid_height)) {
      const auto& neighbor = grid[adj.y][adj.x];
      if (!neighbor) continue;
      if (const auto similarity = db_.get_similarity(candidate_path, *neighbor)) penalty += 1.0f / (*similarity + 1.0f);
    }
    return penalty * weight_;
  }

  // adjacency.rs:91-130: sum over right / down pairs of 1 / (similarity + 1), raster order, f32
  float calculate_total_cost(const PathGrid& grid) const {
    if (grid.empty() || grid[0].empty()) return 0.0f;
    return calculate_total_cost(detail::scratch_context(), grid);
  }
  float calculate_total_cost(Context& ctx, const PathGrid& grid) const {
    int gw, gh;
    std::vector<std::string> unknown;
    const std::vector<int32_t> ids = detail::grid_ids(db_, grid, &gw, &gh, &unknown);
    if (gw == 0 || gh == 0) return 0.0f;
    load(ctx, ids, gw, gh, unknown.size());
    mosaic_opt_result res{};
    ctx.check(mosaic_anneal(ctx.get(), 0, 100.0f, 0.99995f, 0, nullptr, &res));  // 0 iterations: the cost only
    return res.initial_cost;
  }

  // adjacency.rs:133-214: the change of the (unweighted) energy if the tiles at pos1 and pos2 trade places — pos1's
  // neighbours without pos2, then pos2's without pos1, each up / down / left / right, old and new term side by side, then
  // the mutual term on both sides if the two are adjacent; 0 for an empty cell or equal paths.  A two-cell query (≤ 14
  // lookups) over the GPU-built matrix, like calculate_penalty; the annealing kernel evaluates the same expression with the
  // same operation order on the device (anneal.cu), where the 1000-iteration loop that calls it lives.
  float calculate_swap_delta(const PathGrid& grid, GridPosition pos1, GridPosition pos2) const {
    const size_t grid_height = grid.size();
    if (grid_height == 0) return 0.0f;
    const size_t grid_width = grid[0].size();
    const auto& path1 = grid[pos1.y][pos1.x];
    const auto& path2 = grid[pos2.y][pos2.x];
    if (!path1 || !path2 || *path1 == *path2) return 0.0f;
    float old_cost = 0.0f, new_cost = 0.0f;
    const std::vector<GridPosition> adj1 = pos1.get_adjacent_positions(grid_width, grid_height);
    for (const GridPosition& adj : adj1) {
      if (adj == pos2) continue;
      const auto& adj_path = grid[adj.y][adj.x];
      if (!adj_path) continue;
      if (const auto old_sim = db_.get_similarity(*path1, *adj_path)) old_cost += 1.0f / (*old_sim + 1.0f);
      if (const auto new_sim = db_.get_similarity(*path2, *adj_path)) new_cost += 1.0f / (*new_sim + 1.0f);
    }
    for (const GridPosition& adj : pos2.get_adjacent_positions(grid_width, grid_height)) {
      if (adj == pos1) continue;
      const auto& adj_path = grid[adj.y][adj.x];
      if (!adj_path) continue;
      if (const auto old_sim = db_.get_similarity(*path2, *adj_path)) old_cost += 1.0f / (*old_sim + 1.0f);
      if (const auto new_sim = db_.get_similarity(*path1, *adj_path)) new_cost += 1.0f / (*new_sim + 1.0f);
    }
    if (std::find(adj1.begin(), adj1.end(), pos2) != adj1.end()) {
      if (const auto sim = db_.get_similarity(*path1, *path2)) {
        const float cost = 1.0f / (*sim + 1.0f);
        old_cost += cost;
        new_cost += cost;
      }
    }
    return new_cost - old_cost;
  }

  const SimilarityDatabase& database() const { return db_; }
  float penalty_weight() const { return weight_; }

  // `n_unknown` pseudo materials (cells whose path the database does not know) follow the database's colours
  void load(Context& ctx, const std::vector<int32_t>& ids, int gw, int gh, size_t n_unknown = 0) const {
    std::vector<float> lab = db_.lab_colors();
    for (size_t i = 0; i < n_unknown; ++i) {
      float c[3];
      detail::unknown_lab(i, c);
      lab.insert(lab.end(), c, c + 3);
    }
    if (lab.empty()) lab.assign(3, 0.0f);  // (a grid of empty cells over an empty database: nothing is ever read)
    ctx.check(mosaic_set_material_lab(ctx.get(), lab.data(), (int)(lab.size() / 3)));
    const std::vector<float> zeros((size_t)gw * gh * 3, 0.0f);  // cell colours play no role in the energy
    ctx.check(mosaic_set_cell_lab(ctx.get(), zeros.data(), gw, gh));
    ctx.check(mosaic_set_placement(ctx.get(), ids.data()));
  }

 private:
  const SimilarityDatabase& db_;
  float weight_;
};

// ---- src/optimizer.rs -------------------------------------------------------------------------------------
struct OptimizationConfig {  // optimizer.rs:7-27
  size_t max_iterations = 1000;
  float initial_temperature = 100.0f;
  float temperature_decay = 0.99995f;
  size_t report_interval = 100;
  uint64_t seed = 0;  // extension: the reference's thread_rng cannot be seeded (optimizer.rs:48)
};

struct OptimizationResult {  // optimizer.rs:203-220
  float initial_cost = 0, final_cost = 0, best_cost = 0;
  size_t improved_count = 0, accepted_count = 0, iterations = 0;
  float improvement_percentage() const {
    return initial_cost > 0.0f ? (initial_cost - final_cost) / initial_cost * 100.0f : 0.0f;
  }
};

class MosaicOptimizer {
 public:
  MosaicOptimizer(const AdjacencyPenaltyCalculator& calculator, OptimizationConfig config)  // optimizer.rs:36-38
      : calc_(calculator), config_(config) {}

  OptimizationResult optimize(PathGrid& grid) const {
    if (grid.empty() || grid[0].empty()) return OptimizationResult();  // optimizer.rs:43-45, before any device work
    return optimize(detail::scratch_context(), grid);
  }
  OptimizationResult optimize(Context& ctx, PathGrid& grid) const {  // optimizer.rs:41-129
    return run(ctx, grid, [&](int32_t* out, mosaic_opt_result* r) {
      return mosaic_anneal(ctx.get(), config_.max_iterations, config_.initial_temperature, config_.temperature_decay, config_.seed,
                           out, r);
    });
  }
  // optimizer.rs:132-198: only improvements are accepted (the seed is the config's, an extension like optimize's)
  OptimizationResult optimize_greedy(PathGrid& grid, size_t max_iterations) const {
    if (grid.empty() || grid[0].empty()) return OptimizationResult();  // optimizer.rs:138-141
    return optimize_greedy(detail::scratch_context(), grid, max_iterations);
  }
  OptimizationResult optimize_greedy(Context& ctx, PathGrid& grid, size_t max_iterations) const {
    return run(ctx, grid, [&](int32_t* out, mosaic_opt_result* r) {
      return mosaic_optimize_greedy(ctx.get(), max_iterations, config_.seed, out, r);
    });
  }

 private:
  template <typename Call>
  OptimizationResult run(Context& ctx, PathGrid& grid, Call call) const {
    int gw, gh;
    std::vector<std::string> unknown;
    const std::vector<int32_t> ids = detail::grid_ids(calc_.database(), grid, &gw, &gh, &unknown);
    if (gh == 0 || gw == 0) return OptimizationResult();  // optimizer.rs:43-45
    calc_.load(ctx, ids, gw, gh, unknown.size());
    std::vector<int32_t> out(ids.size());
    mosaic_opt_result r{};
    ctx.check(call(out.data(), &r));
    // the kernel swaps material indices; cells whose path the database does not know carry pseudo indices behind the
    // database's (occupied cells without similarity terms, as in the reference)
    std::vector<std::string> path_of(calc_.database().len() + unknown.size());
    for (const auto& row : grid)
      for (const auto& c : row)
        if (c && calc_.database().index_of(*c) >= 0) path_of[(size_t)calc_.database().index_of(*c)] = *c;
    for (size_t i = 0; i < unknown.size(); ++i) path_of[calc_.database().len() + i] = unknown[i];
    for (int y = 0; y < gh; ++y)
      for (int x = 0; x < gw; ++x) {
        const int32_t id = out[(size_t)y * gw + x];
        if (id >= 0) grid[y][x] = path_of[(size_t)id];
      }
    OptimizationResult res;
    res.initial_cost = r.initial_cost;
    res.final_cost = r.final_cost;
    res.best_cost = r.best_cost;
    res.improved_count = (size_t)r.improved_count;
    res.accepted_count = (size_t)r.accepted_count;
    res.iterations = (size_t)r.iterations;
    return res;
  }

  const AdjacencyPenaltyCalculator& calc_;
  OptimizationConfig config_;
};

// ---- src/lib.rs ---------------------------------------------------------------------------------------------
struct RgbImage {  // image::RgbImage: interleaved RGB8, row-major
  int width = 0, height = 0;
  std::vector<uint8_t> data;
  RgbImage() = default;
  RgbImage(int w, int h) : width(w), height(h), data((size_t)w * h * 3, 0) {}
};

namespace detail {
// the C ABI takes a pointer and trusts width * height * 3 bytes behind it: an RgbImage whose buffer disagrees with its
// dimensions (image::RgbImage cannot be in that state; this struct can) is refused here instead of being read out of bounds
inline void check_image(const RgbImage& img, const char* what) {
  if (img.width <= 0 || img.height <= 0 || img.data.size() != (size_t)img.width * (size_t)img.height * 3)
    throw Error(MOSAIC_ERR_INVALID, std::string(what) + ": the pixel buffer does not hold width * height * 3 bytes");
}
}  // namespace detail

// UsageTracker (lib.rs:57-89): per-path use counts with a cap.  The placement kernel keeps its own counters in shared
// memory (u8 / u16 per material, greedy_window.cu) with exactly these semantics; this class is the public type of
// lib.rs for host code that tracks usage itself.
class UsageTracker {
 public:
  explicit UsageTracker(size_t max_usage_per_image) : max_usage_per_image_(max_usage_per_image) {}  // lib.rs:63-68
  bool can_use_image(const std::string& path) const { return get_usage_count(path) < max_usage_per_image_; }  // lib.rs:70-73
  void use_image(const std::string& path) { usage_counts_[path] = get_usage_count(path) + 1; }                // lib.rs:75-78
  size_t get_usage_count(const std::string& path) const {                                                       // lib.rs:80-83
    const auto it = usage_counts_.find(path);
    return it == usage_counts_.end() ? 0 : it->second;
  }
  void reset() { usage_counts_.clear(); }  // lib.rs:85-87

 private:
  std::map<std::string, size_t> usage_counts_;
  size_t max_usage_per_image_;
};

struct MosaicGeneratorImpl {  // impl MosaicGenerator for MosaicGeneratorImpl (trait at lib.rs:52-55)
  // MosaicGenerator::is_aspect_ratio_match (lib.rs:121-123)
  static bool is_aspect_ratio_match(float img_aspect, float target_aspect, float tolerance) {
    return std::fabs(img_aspect - target_aspect) <= tolerance;
  }
  // MosaicGenerator::calculate_average_lab (lib.rs:94-119)
  static Lab calculate_average_lab(const RgbImage& img) { return calculate_average_lab(detail::scratch_context(), img); }
  static Lab calculate_average_lab(Context& ctx, const RgbImage& img) {
    float lab[3];
    detail::check_image(img, "calculate_average_lab");
    ctx.check(mosaic_set_materials_rgb8(ctx.get(), img.data.data(), 1, img.width, img.height));
    ctx.check(mosaic_material_lab(ctx.get(), lab));
    return Lab{lab[0], lab[1], lab[2]};
  }
};

struct Tile {  // lib.rs:44-50
  std::string path;
  Lab lab_color;
  float aspect_ratio = 1.0f;
};

// ---- src/main.rs: struct MosaicGenerator (private to the CLI binary) ------------------------------------------
class MosaicGenerator {
 public:
  // MosaicGenerator::new (main.rs:105-164) with the materials already decoded and filtered (load_tiles, main.rs:166-261,
  // is the control plane): `paths[i]` names material i, `rgb` holds n uniform-size RGB8 images.  Tile Lab means on the
  // GPU (B1); the similarity database is built + saved, or loaded from `similarity_db_path` and topped up with the
  // tiles it does not know, then rebuilt (main.rs:120-143).  Similarities always come from the database's colours:
  // when a loaded file disagrees with the tiles' own means they are handed to the library separately
  // (mosaic_set_similarity_lab).  An empty `similarity_db_path` skips the file (in-memory use).
  MosaicGenerator(const std::vector<std::string>& paths, const uint8_t* rgb, int mat_w, int mat_h,
                  size_t max_usage_per_image, const std::string& similarity_db_path, bool rebuild_similarity,
                  float adjacency_penalty_weight, float color_adjustment_strength, int device = 0)
      : ctx_(device),
        device_(device),
        max_usage_(max_usage_per_image),
        weight_(adjacency_penalty_weight),
        strength_(std::min(std::max(color_adjustment_strength, 0.0f), 1.0f)) {  // main.rs:162
    ctx_.check(mosaic_set_materials_rgb8(ctx_.get(), rgb, (int)paths.size(), mat_w, mat_h));
    init(paths, similarity_db_path, rebuild_similarity);
    for (Tile& t : tiles_) t.aspect_ratio = (float)mat_w / (float)mat_h;  // main.rs:277 / 291
  }

  // Same, with materials of individual sizes — what the reference's directory of image files is (process_tile,
  // main.rs:263-301): each image keeps its own size for the Lab mean and is resized to the tile at render time.
  MosaicGenerator(const std::vector<std::string>& paths, const std::vector<RgbImage>& images, size_t max_usage_per_image,
                  const std::string& similarity_db_path, bool rebuild_similarity, float adjacency_penalty_weight,
                  float color_adjustment_strength, int device = 0)
      : ctx_(device),
        device_(device),
        max_usage_(max_usage_per_image),
        weight_(adjacency_penalty_weight),
        strength_(std::min(std::max(color_adjustment_strength, 0.0f), 1.0f)) {
    std::vector<uint8_t> blob;
    std::vector<uint64_t> offsets;
    std::vector<int32_t> widths, heights;
    for (const RgbImage& im : images) {
      detail::check_image(im, "MosaicGenerator: material");
      offsets.push_back(blob.size());
      widths.push_back(im.width);
      heights.push_back(im.height);
      blob.insert(blob.end(), im.data.begin(), im.data.end());
    }
    if (paths.size() != images.size()) throw Error(MOSAIC_ERR_INVALID, "MosaicGenerator: one path per image");
    ctx_.check(mosaic_set_materials_ragged_rgb8(ctx_.get(), blob.data(), offsets.data(), widths.data(), heights.data(),
                                                (int)images.size()));
    init(paths, similarity_db_path, rebuild_similarity);
    for (size_t i = 0; i < tiles_.size(); ++i) tiles_[i].aspect_ratio = (float)widths[i] / (float)heights[i];
  }

 private:
  void init(const std::vector<std::string>& paths, const std::string& similarity_db_path, bool rebuild_similarity) {
    const int n = (int)paths.size();
    std::vector<float> lab((size_t)n * 3);
    ctx_.check(mosaic_material_lab(ctx_.get(), lab.data()));
    for (int i = 0; i < n; ++i) tiles_.push_back(Tile{paths[i], Lab{lab[3 * i], lab[3 * i + 1], lab[3 * i + 2]}, 1.0f});

    const bool exists = !similarity_db_path.empty() && std::ifstream(similarity_db_path).good();
    if (rebuild_similarity || !exists) {  // main.rs:120-134
      for (const Tile& t : tiles_) similarity_db_.add_tile(t.path, t.lab_color);
      similarity_db_.build_similarities(ctx_);
      if (!similarity_db_path.empty()) {
        try {
          similarity_db_.save_to_file(similarity_db_path);
        } catch (const Error& e) {
          save_warning_ = std::string("Warning: Failed to save similarity database: ") + e.what();
        }
      }
    } else {
      similarity_db_ = SimilarityDatabase::load_or_new(similarity_db_path);  // main.rs:135-137
    }
    bool topped_up = false;
    for (const Tile& t : tiles_)  // main.rs:139-143
      if (!similarity_db_.get_lab_color(t.path)) {
        similarity_db_.add_tile(t.path, t.lab_color);
        topped_up = true;
      }
    if (topped_up || !(rebuild_similarity || !exists)) {
      // the database may now hold other paths too: its matrix is built on the scratch context of the generator's device,
      // the generator's own context keeps the tiles
      similarity_db_.build_similarities(detail::scratch_context(device_));
    }
    std::vector<float> db_lab((size_t)n * 3);
    bool differs = false;
    for (int i = 0; i < n; ++i) {
      const Lab c = *similarity_db_.get_lab_color(tiles_[i].path);
      db_lab[3 * i] = c.l, db_lab[3 * i + 1] = c.a, db_lab[3 * i + 2] = c.b;
      differs |= c.l != lab[3 * i] || c.a != lab[3 * i + 1] || c.b != lab[3 * i + 2];
    }
    if (differs) ctx_.check(mosaic_set_similarity_lab(ctx_.get(), db_lab.data(), n));
  }

 public:

  // generate_mosaic (main.rs:477-685) minus file I/O: returns the mosaic image; `placed` (optional) receives the grid
  RgbImage generate_mosaic(const RgbImage& target, int grid_w, int grid_h, bool enable_optimization,
                           size_t optimization_iterations, uint64_t seed = 0, PathGrid* placed = nullptr,
                           mosaic_report* report = nullptr) {
    detail::check_image(target, "generate_mosaic: target");
    ctx_.check(mosaic_set_target_rgb8(ctx_.get(), target.data.data(), target.width, target.height, grid_w, grid_h));
    mosaic_params p{};
    p.max_usage = (int)std::min<size_t>(max_usage_, (size_t)0x7fffffff);  // usize -> int: a huge cap stays "no cap"
    p.adjacency_weight = weight_;
    p.color_strength = strength_;
    p.enable_optimization = enable_optimization ? 1 : 0;
    p.optimization_iterations = optimization_iterations;
    p.seed = seed;
    p.build_similarity = 0;  // the matrix belongs to the database and was built by the constructor (main.rs:126, 143)
    int n, gw, gh, tw, th;
    mosaic_dims(ctx_.get(), &n, &gw, &gh, &tw, &th);
    RgbImage out(gw * tw, gh * th);  // main.rs:499-503
    std::vector<int32_t> ids((size_t)gw * gh);
    mosaic_report rep{};
    ctx_.check(mosaic_generate(ctx_.get(), &p, ids.data(), out.data.data(), &rep));
    if (report) *report = rep;
    if (placed) {
      placed->assign(gh, std::vector<std::optional<std::string>>(gw));
      for (int y = 0; y < gh; ++y)
        for (int x = 0; x < gw; ++x)
          if (ids[(size_t)y * gw + x] >= 0) (*placed)[y][x] = tiles_[(size_t)ids[(size_t)y * gw + x]].path;
    }
    return out;
  }

  const std::vector<Tile>& tiles() const { return tiles_; }
  SimilarityDatabase& similarity_db() { return similarity_db_; }
  Context& context() { return ctx_; }
  const std::string& save_warning() const { return save_warning_; }  // the reference prints it to stderr (main.rs:131)

 private:
  Context ctx_;
  int device_ = 0;
  std::string save_warning_;
  std::vector<Tile> tiles_;
  SimilarityDatabase similarity_db_;
  size_t max_usage_;
  float weight_, strength_;
};

}  // namespace mosaic_b200
#endif
