// The reference's own unit tests for the matching path, restated against the C++ host mirror
// (include/mosaic_b200.hpp) — same fixtures, same assertions, citing the Rust test each one follows.
//
//   test_host_mirror --cpu            tests that need no GPU (bookkeeping, file format, loud failure without a device)
//   test_host_mirror                  everything (needs an sm_100a GPU)
//   test_host_mirror --generate DIR   MosaicGenerator end to end on DIR/{materials.bin,target.bin,meta.txt}; writes
//                                     DIR/placed.bin, DIR/image.bin, DIR/report.txt for the Python parity test
//
// Built and driven by tests/test_host_mirror.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "mosaic_b200.hpp"

using namespace mosaic_b200;

static int g_failed = 0, g_run = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);            \
      ++g_failed;                                                                \
    }                                                                            \
  } while (0)
#define RUN(fn)                          \
  do {                                   \
    std::printf("%s\n", #fn);            \
    ++g_run;                             \
    try {                                \
      fn();                              \
    } catch (const std::exception& e) {  \
      std::printf("  EXCEPTION %s\n", e.what()); \
      ++g_failed;                        \
    }                                    \
  } while (0)

static std::string tmp_path(const char* name) {
  const char* d = std::getenv("TMPDIR");
  return std::string(d ? d : "/tmp") + "/mosaic_b200_" + name;
}

// ---------------------------------------------------------------- no GPU needed
static void test_similarity_database_get_lab_color() {  // similarity.rs:235-252
  SimilarityDatabase db;
  db.add_tile("test.png", Lab{50.0f, 25.0f, -15.0f});
  const auto got = db.get_lab_color("test.png");
  CHECK(got.has_value());
  CHECK(got->l == 50.0f && got->a == 25.0f && got->b == -15.0f);
  CHECK(!db.get_lab_color("nonexistent.png").has_value());
}

static void test_similarity_database_empty() {  // similarity.rs:308-318
  SimilarityDatabase db;
  CHECK(!db.get_similarity("tile1.png", "tile2.png").has_value());
  CHECK(!db.get_lab_color("tile1.png").has_value());
}

static void test_similarity_database_load_or_new() {  // similarity.rs:281-296
  CHECK(SimilarityDatabase::load_or_new("nonexistent_db.json").len() == 0);
  SimilarityDatabase original;
  original.add_tile("test.png", Lab{50.0f, 0.0f, 0.0f});
  const std::string f = tmp_path("load_or_new.json");
  original.save_to_file(f);
  CHECK(SimilarityDatabase::load_or_new(f).len() == 1);
  std::remove(f.c_str());
}

static void test_similarity_database_load_from_invalid_file() {  // similarity.rs:299-305
  const std::string f = tmp_path("invalid.json");
  std::ofstream(f) << "invalid json";
  bool threw = false;
  try {
    SimilarityDatabase::load_from_file(f);
  } catch (const Error&) {
    threw = true;
  }
  CHECK(threw);
  std::remove(f.c_str());
}

static void test_grid_position_adjacent() {  // adjacency.rs:225-241
  const auto adj = GridPosition(1, 1).get_adjacent_positions(3, 3);
  CHECK(adj.size() == 4);
  CHECK(adj[0] == GridPosition(1, 0) && adj[1] == GridPosition(1, 2) && adj[2] == GridPosition(0, 1) && adj[3] == GridPosition(2, 1));
  const auto corner = GridPosition(0, 0).get_adjacent_positions(3, 3);
  CHECK(corner.size() == 2);
  CHECK(corner[0] == GridPosition(0, 1) && corner[1] == GridPosition(1, 0));
}

static void test_optimization_result_improvement() {  // optimizer.rs:280-292, 550-586
  OptimizationResult r;
  r.initial_cost = 100.0f, r.final_cost = 75.0f, r.best_cost = 75.0f;
  CHECK(r.improvement_percentage() == 25.0f);
  OptimizationResult zero;
  CHECK(zero.improvement_percentage() == 0.0f);
  OptimizationResult worse;
  worse.initial_cost = 50.0f, worse.final_cost = 75.0f;
  CHECK(worse.improvement_percentage() == -50.0f);
  OptimizationResult perfect;
  perfect.initial_cost = 100.0f, perfect.final_cost = 0.0f;
  CHECK(perfect.improvement_percentage() == 100.0f);
}

static void test_optimization_config_defaults() {  // optimizer.rs:621-628
  OptimizationConfig c;
  CHECK(c.max_iterations == 1000 && c.initial_temperature == 100.0f && c.temperature_decay == 0.99995f && c.report_interval == 100);
}

static void test_greedy_optimization_edge_cases() {  // optimizer.rs:520-534 — returns before any device work
  PathGrid empty_grid;
  SimilarityDatabase db;
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  const OptimizationResult r = MosaicOptimizer(calc, OptimizationConfig()).optimize_greedy(empty_grid, 10);
  CHECK(r.initial_cost == 0.0f);
  CHECK(r.final_cost == 0.0f);
  CHECK(r.improved_count == 0);
}

static void test_optimization_empty_grid() {  // optimizer.rs:295-307 — returns before any device work
  PathGrid empty;
  SimilarityDatabase db;
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  MosaicOptimizer opt(calc, OptimizationConfig());
  const OptimizationResult r = opt.optimize(empty);
  CHECK(r.initial_cost == 0.0f && r.final_cost == 0.0f && r.improved_count == 0);
}

static void test_aspect_ratio_match() {  // lib.rs:160-210
  CHECK(MosaicGeneratorImpl::is_aspect_ratio_match(16.0f / 9.0f, 16.0f / 9.0f, 0.1f));
  CHECK(MosaicGeneratorImpl::is_aspect_ratio_match(1.8f, 16.0f / 9.0f, 0.1f));
  CHECK(!MosaicGeneratorImpl::is_aspect_ratio_match(4.0f / 3.0f, 16.0f / 9.0f, 0.1f));
  CHECK(!MosaicGeneratorImpl::is_aspect_ratio_match(16.0f / 9.0f, 9.0f / 16.0f, 0.1f));
}

static void test_tile_creation() {  // lib.rs:229-240
  const Tile tile{"test.png", Lab{50.0f, 0.0f, 0.0f}, 16.0f / 9.0f};
  CHECK(tile.path == "test.png");
  CHECK(tile.aspect_ratio == 16.0f / 9.0f);
  CHECK(tile.lab_color.l == 50.0f);
}

static void test_usage_tracker() {  // lib.rs:242-307
  {
    UsageTracker tracker(3);  // test_usage_tracker_creation
    CHECK(tracker.can_use_image("test.png"));
    CHECK(tracker.get_usage_count("test.png") == 0);
  }
  {
    UsageTracker tracker(2);  // test_usage_tracker_use_image
    CHECK(tracker.can_use_image("test.png"));
    tracker.use_image("test.png");
    CHECK(tracker.get_usage_count("test.png") == 1 && tracker.can_use_image("test.png"));
    tracker.use_image("test.png");
    CHECK(tracker.get_usage_count("test.png") == 2 && !tracker.can_use_image("test.png"));
  }
  {
    UsageTracker tracker(1);  // test_usage_tracker_multiple_images
    tracker.use_image("image1.png");
    CHECK(!tracker.can_use_image("image1.png") && tracker.can_use_image("image2.png"));
    tracker.use_image("image2.png");
    CHECK(!tracker.can_use_image("image2.png"));
  }
  {
    UsageTracker tracker(1);  // test_usage_tracker_reset
    tracker.use_image("test.png");
    CHECK(!tracker.can_use_image("test.png"));
    tracker.reset();
    CHECK(tracker.can_use_image("test.png") && tracker.get_usage_count("test.png") == 0);
  }
  {
    UsageTracker tracker(0);  // test_usage_tracker_max_usage_zero
    CHECK(!tracker.can_use_image("test.png"));
  }
}

static void test_save_to_file_writes_a_stale_matrix() {  // similarity.rs:112-116: whatever `similarities` holds is written
  SimilarityDatabase db;
  db.add_tile("a.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("b.png", Lab{60.0f, 10.0f, 10.0f});
  const std::string f = tmp_path("stale_db.json");
  db.save_to_file(f);  // two tiles, no build_similarities yet: an empty list, as the reference writes it
  const SimilarityDatabase back = SimilarityDatabase::load_from_file(f);
  CHECK(back.len() == 2);
  CHECK(!back.get_similarity("a.png", "b.png").has_value());  // similarity.rs:100-102: index beyond the list -> None
  std::remove(f.c_str());
}

static void test_lab_distance_and_delta_e_2000() {  // similarity.rs:363-402 (+ the fixtures of :185-232 for the distance)
  const Lab lab1{50.0f, 0.0f, 0.0f}, lab2{50.0f, 0.0f, 0.0f}, lab3{60.0f, 10.0f, 10.0f};
  CHECK(calculate_delta_e_2000(lab1, lab2) == 0.0f);
  CHECK(calculate_delta_e_2000(lab1, lab3) > 0.0f);
  CHECK(calculate_delta_e_2000(lab1, lab3) == calculate_delta_e_2000(lab3, lab1));
  const Lab a{50.0f, 20.0f, -30.0f}, b{70.0f, -10.0f, 40.0f};
  CHECK(calculate_lab_distance(a, b) > 0.0f && calculate_delta_e_2000(a, b) > 0.0f);
  CHECK(calculate_lab_distance(a, b) == calculate_lab_distance(b, a));
  CHECK(calculate_delta_e_2000(a, b) == calculate_delta_e_2000(b, a));
  CHECK(calculate_lab_distance(a, a) == 0.0f && calculate_delta_e_2000(a, a) == 0.0f);
  CHECK(calculate_lab_distance(lab1, lab3) == std::sqrt(300.0f));                      // the value build_similarities stores
  CHECK(calculate_lab_distance(lab3, Lab{40.0f, -10.0f, -10.0f}) == std::sqrt(1200.0f));
  // by hand: dl 10, c1 0, c2 sqrt(200), dc = c2, dh2 = 200 - 200 -> dh 0, sc = 1 + 0.045 * c2 / 2
  const float c2 = std::sqrt(200.0f), sc = 1.0f + 0.045f * (c2 / 2.0f);
  CHECK(std::fabs(calculate_delta_e_2000(lab1, lab3) - std::sqrt(100.0f + (c2 / sc) * (c2 / sc))) < 1e-5f);
}

// a database as the reference would have saved it, built WITHOUT a device: colours + host-side distances through the file
static SimilarityDatabase database_from_file(const std::vector<std::string>& names, const std::vector<Lab>& labs) {
  const int n = (int)names.size();
  std::vector<const char*> paths;
  std::vector<float> lab, sim;
  for (int i = 0; i < n; ++i) {
    paths.push_back(names[i].c_str());
    lab.insert(lab.end(), {labs[i].l, labs[i].a, labs[i].b});
  }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) sim.push_back(calculate_lab_distance(labs[i], labs[j]));  // packed order, similarity.rs:100
  const std::string f = tmp_path("swap_delta_db.json");
  if (mosaic_simdb_save_json(f.c_str(), paths.data(), n, lab.data(), sim.data()) != MOSAIC_OK) throw std::runtime_error("save failed");
  SimilarityDatabase db = SimilarityDatabase::load_from_file(f);
  std::remove(f.c_str());
  return db;
}

static float host_total_cost(const SimilarityDatabase& db, const PathGrid& grid) {  // adjacency.rs:91-130, for this test only
  float total = 0.0f;
  for (size_t y = 0; y < grid.size(); ++y)
    for (size_t x = 0; x < grid[y].size(); ++x) {
      if (!grid[y][x]) continue;
      if (x + 1 < grid[y].size() && grid[y][x + 1])
        if (const auto s = db.get_similarity(*grid[y][x], *grid[y][x + 1])) total += 1.0f / (*s + 1.0f);
      if (y + 1 < grid.size() && grid[y + 1][x])
        if (const auto s = db.get_similarity(*grid[y][x], *grid[y + 1][x])) total += 1.0f / (*s + 1.0f);
    }
  return total;
}

static void test_swap_delta() {  // adjacency.rs:133-214 (the optimizer's inner call, optimizer.rs:76, 164)
  std::vector<std::string> names;
  std::vector<Lab> labs;
  uint64_t state = 12345;
  auto next = [&]() {
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(state >> 33);
  };
  for (int i = 0; i < 9; ++i) {
    names.push_back("tile" + std::to_string(i) + ".png");
    labs.push_back(Lab{(float)(next() % 1000) / 10.0f, (float)(next() % 2000) / 10.0f - 100.0f, (float)(next() % 2000) / 10.0f - 100.0f});
  }
  const SimilarityDatabase db = database_from_file(names, labs);
  CHECK(db.len() == 9);
  const AdjacencyPenaltyCalculator calc(db, 0.3f);
  // the 2x2 fixture of adjacency.rs:243-290 style: neighbours, mutual term
  {
    PathGrid grid = {{names[0], names[1]}, {names[2], names[3]}};
    const float before = host_total_cost(db, grid);
    const float delta = calc.calculate_swap_delta(grid, GridPosition(0, 0), GridPosition(1, 0));  // adjacent cells
    std::swap(grid[0][0], grid[0][1]);
    CHECK(std::fabs((host_total_cost(db, grid) - before) - delta) < 1e-5f);
  }
  const size_t gw = 5, gh = 4;
  for (int trial = 0; trial < 400; ++trial) {
    PathGrid grid(gh, std::vector<std::optional<std::string>>(gw));
    for (auto& row : grid)
      for (auto& cell : row) {
        const uint32_t r = next() % 12;
        if (r < 9) cell = names[r];                 // duplicates on purpose
        else if (r == 9) cell = "unknown.png";      // a path the database does not know: its terms are skipped
      }                                             // else: an empty cell
    const GridPosition p1(next() % gw, next() % gh), p2(next() % gw, next() % gh);
    const float delta = calc.calculate_swap_delta(grid, p1, p2);
    const auto& a = grid[p1.y][p1.x];
    const auto& b = grid[p2.y][p2.x];
    if (!a || !b || *a == *b) {
      CHECK(delta == 0.0f);                         // adjacency.rs:146-158
      continue;
    }
    const float before = host_total_cost(db, grid);
    PathGrid swapped = grid;
    std::swap(swapped[p1.y][p1.x], swapped[p2.y][p2.x]);
    CHECK(std::fabs((host_total_cost(db, swapped) - before) - delta) < 2e-5f);
    CHECK(calc.calculate_swap_delta(swapped, p1, p2) == -delta || std::fabs(calc.calculate_swap_delta(swapped, p1, p2) + delta) < 1e-6f);
  }
  CHECK(calc.calculate_swap_delta(PathGrid{}, GridPosition(0, 0), GridPosition(0, 0)) == 0.0f);  // adjacency.rs:139-142
  // calculate_penalty over the same database (adjacency.rs:63-88): weight * sum of 1 / (sim + 1) over placed neighbours
  PathGrid grid(3, std::vector<std::optional<std::string>>(3));
  grid[0][1] = names[1];
  grid[1][0] = names[2];
  const float want = (1.0f / (calculate_lab_distance(labs[0], labs[1]) + 1.0f) + 1.0f / (calculate_lab_distance(labs[0], labs[2]) + 1.0f)) * 0.3f;
  CHECK(calc.calculate_penalty(names[0], GridPosition(1, 1), grid, 3, 3) == want);
}

static void test_image_buffers_are_checked() {  // an RgbImage whose buffer disagrees with its size never reaches the C ABI
  RgbImage ok(4, 3);
  bool threw = false;
  try {
    detail::check_image(ok, "test");
  } catch (const Error&) {
    threw = true;
  }
  CHECK(!threw);
  RgbImage shorter(4, 3);
  shorter.data.resize(10);
  RgbImage empty;
  RgbImage negative(4, 3);
  negative.width = -4;
  for (const RgbImage* bad : {&shorter, &empty, &negative}) {
    threw = false;
    try {
      detail::check_image(*bad, "test");
    } catch (const Error& e) {
      threw = e.code == MOSAIC_ERR_INVALID;
    }
    CHECK(threw);
  }
}

static void test_no_gpu_is_loud() {  // there is no CPU fallback behind the mirror: without a device the call throws
  SimilarityDatabase db;
  db.add_tile("a.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("b.png", Lab{60.0f, 10.0f, 10.0f});
  try {
    db.build_similarities();
    CHECK(db.get_similarity("a.png", "b.png") == std::optional<float>(std::sqrt(300.0f)));  // a GPU was there after all
    std::printf("  (device present: computed on the GPU)\n");
  } catch (const Error& e) {
    CHECK(e.code == MOSAIC_ERR_CUDA);
    CHECK(!db.get_similarity("a.png", "b.png").has_value() || *db.get_similarity("a.png", "b.png") == 0.0f);
    std::printf("  (no device: %s)\n", e.what());
  }
}

// ---------------------------------------------------------------- GPU
static SimilarityDatabase three_tiles() {
  SimilarityDatabase db;
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("tile2.png", Lab{60.0f, 10.0f, 10.0f});
  db.add_tile("tile3.png", Lab{40.0f, -10.0f, -10.0f});
  db.build_similarities();
  return db;
}

static void test_similarity_database_basic() {  // similarity.rs:186-204
  const SimilarityDatabase db = three_tiles();
  CHECK(db.get_similarity("tile1.png", "tile2.png").has_value());
  CHECK(db.get_similarity("tile1.png", "tile1.png") == std::optional<float>(0.0f));
}

static void test_similarity_database_multiple_tiles() {  // similarity.rs:336-361 + the fixture's values (√300, √1200)
  const SimilarityDatabase db = three_tiles();
  const auto s12 = db.get_similarity("tile1.png", "tile2.png"), s21 = db.get_similarity("tile2.png", "tile1.png");
  const auto s13 = db.get_similarity("tile1.png", "tile3.png"), s23 = db.get_similarity("tile2.png", "tile3.png");
  CHECK(s12 && s21 && s13 && s23);
  CHECK(s12 == s21);
  CHECK(*s12 > 0.0f && *s13 > 0.0f && *s23 > 0.0f);
  CHECK(*s12 == std::sqrt(300.0f) && *s13 == std::sqrt(300.0f) && *s23 == std::sqrt(1200.0f));
}

static void test_similarity_database_single_tile() {  // similarity.rs:321-333
  SimilarityDatabase db;
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.build_similarities();
  CHECK(db.get_similarity("tile1.png", "tile1.png") == std::optional<float>(0.0f));
  CHECK(!db.get_similarity("tile1.png", "nonexistent.png").has_value());
}

static void test_similarity_database_save_load() {  // similarity.rs:255-278
  SimilarityDatabase db;
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("tile2.png", Lab{60.0f, 10.0f, 10.0f});
  db.build_similarities();
  const std::string f = tmp_path("save_load.json");
  db.save_to_file(f);
  const SimilarityDatabase loaded = SimilarityDatabase::load_from_file(f);
  CHECK(loaded.len() == 2);
  const auto sim = loaded.get_similarity("tile1.png", "tile2.png");
  CHECK(sim && *sim > 0.0f && *sim == *db.get_similarity("tile1.png", "tile2.png"));
  std::remove(f.c_str());
}

static void test_adjacency_penalty_calculation() {  // adjacency.rs:243-271; value from SURVEY Appendix B
  const SimilarityDatabase db = three_tiles();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  PathGrid grid(3, std::vector<std::optional<std::string>>(3));
  grid[0][0] = "tile1.png";
  grid[0][1] = "tile2.png";
  const float penalty = calc.calculate_penalty("tile3.png", GridPosition(1, 0), grid, 3, 3);
  CHECK(penalty > 0.0f);
  // neighbours of (1,0): down (empty), left tile1 (√300), right (empty) -> 1 / (17.320509 + 1)
  CHECK(penalty == 1.0f / (std::sqrt(300.0f) + 1.0f));
}

static void test_total_cost_calculation() {  // adjacency.rs:273-290
  SimilarityDatabase db;
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("tile2.png", Lab{50.0f, 0.0f, 0.0f});
  db.build_similarities();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  PathGrid grid(2, std::vector<std::optional<std::string>>(2));
  grid[0][0] = "tile1.png";
  grid[0][1] = "tile2.png";
  const float total = calc.calculate_total_cost(grid);
  CHECK(total > 0.5f);
  CHECK(total == 1.0f);  // identical colours: similarity 0 -> 1 / (0 + 1)
}

static SimilarityDatabase four_tiles() {  // optimizer.rs:230-245
  SimilarityDatabase db;
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("tile2.png", Lab{60.0f, 10.0f, 10.0f});
  db.add_tile("tile3.png", Lab{40.0f, -10.0f, -10.0f});
  db.add_tile("tile4.png", Lab{55.0f, 5.0f, 5.0f});
  db.build_similarities();
  return db;
}
static PathGrid four_grid() {
  PathGrid grid(2, std::vector<std::optional<std::string>>(2));
  grid[0][0] = "tile1.png", grid[0][1] = "tile2.png", grid[1][0] = "tile3.png", grid[1][1] = "tile4.png";
  return grid;
}

static bool same_multiset(const PathGrid& a, const PathGrid& b) {
  std::vector<std::string> va, vb;
  for (const auto& r : a)
    for (const auto& c : r) va.push_back(c ? *c : "<none>");
  for (const auto& r : b)
    for (const auto& c : r) vb.push_back(c ? *c : "<none>");
  std::sort(va.begin(), va.end());
  std::sort(vb.begin(), vb.end());
  return va == vb;
}

static void test_optimization_basic() {  // optimizer.rs:247-264
  const SimilarityDatabase db = four_tiles();
  PathGrid grid = four_grid();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  OptimizationConfig cfg;
  cfg.max_iterations = 10;
  const OptimizationResult r = MosaicOptimizer(calc, cfg).optimize(grid);
  CHECK(r.iterations == 10);
  CHECK(r.final_cost >= 0.0f);
  CHECK(same_multiset(grid, four_grid()));  // swaps only permute the tiles
  // the energy the optimiser reports is the energy of the grid it returns (up to the f32 drift of summed deltas)
  CHECK(std::fabs(calc.calculate_total_cost(grid) - r.final_cost) < 1e-4f);
  // the fixture's initial energy, SURVEY Appendix B: 0x3E7FBDDA
  uint32_t bits;
  std::memcpy(&bits, &r.initial_cost, 4);
  CHECK(bits == 0x3E7FBDDAu);
}

static void test_greedy_optimization() {  // optimizer.rs:266-277
  const SimilarityDatabase db = four_tiles();
  PathGrid grid = four_grid();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  const OptimizationResult r = MosaicOptimizer(calc, OptimizationConfig()).optimize_greedy(grid, 10);
  CHECK(r.final_cost <= r.initial_cost);  // only improvements are accepted
  CHECK(r.improved_count == r.accepted_count);
  CHECK(r.iterations == 10 && r.best_cost == r.final_cost);
  CHECK(same_multiset(grid, four_grid()));
  CHECK(std::fabs(calc.calculate_total_cost(grid) - r.final_cost) < 1e-4f);
}

static void test_greedy_optimization_zero_iterations() {  // optimizer.rs:536-547
  const SimilarityDatabase db = four_tiles();
  PathGrid grid = four_grid();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  const OptimizationResult r = MosaicOptimizer(calc, OptimizationConfig()).optimize_greedy(grid, 0);
  CHECK(r.improved_count == 0);
  CHECK(r.initial_cost == r.final_cost);
  CHECK(grid == four_grid());
}

static void test_optimization_single_tile_grid() {  // optimizer.rs:310-331
  SimilarityDatabase db;
  db.add_tile("single.png", Lab{50.0f, 0.0f, 0.0f});
  db.build_similarities();
  PathGrid grid(1, std::vector<std::optional<std::string>>(1));
  grid[0][0] = "single.png";
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  OptimizationConfig cfg;
  cfg.max_iterations = 100;
  const OptimizationResult r = MosaicOptimizer(calc, cfg).optimize(grid);
  CHECK(r.initial_cost == 0.0f && r.final_cost == 0.0f && r.improved_count == 0 && r.accepted_count == 0);
  CHECK(grid[0][0] == std::optional<std::string>("single.png"));
}

static void test_optimization_grid_with_empty_cells() {  // optimizer.rs:334-358
  SimilarityDatabase db;
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("tile2.png", Lab{60.0f, 10.0f, 10.0f});
  db.build_similarities();
  PathGrid grid(3, std::vector<std::optional<std::string>>(3));
  grid[0][0] = "tile1.png";
  grid[2][2] = "tile2.png";
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  OptimizationConfig cfg;
  cfg.max_iterations = 50;
  const OptimizationResult r = MosaicOptimizer(calc, cfg).optimize(grid);
  CHECK(r.final_cost >= 0.0f);
  CHECK(r.initial_cost == 0.0f);
  int placed = 0;
  for (const auto& row : grid)
    for (const auto& c : row) placed += c.has_value();
  CHECK(placed == 2);
}

static void test_optimization_extreme_temperature_values() {  // optimizer.rs:361-397
  const SimilarityDatabase db = four_tiles();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  PathGrid g1 = four_grid(), g2 = four_grid();
  OptimizationConfig hi;
  hi.max_iterations = 20, hi.initial_temperature = 10000.0f, hi.temperature_decay = 0.95f;
  const OptimizationResult rh = MosaicOptimizer(calc, hi).optimize(g1);
  CHECK(rh.iterations == 20 && rh.final_cost >= 0.0f);
  OptimizationConfig lo;
  lo.max_iterations = 20, lo.initial_temperature = 0.001f, lo.temperature_decay = 0.99f;
  const OptimizationResult rl = MosaicOptimizer(calc, lo).optimize(g2);
  CHECK(rl.iterations == 20);
  CHECK(rl.final_cost <= rl.initial_cost || std::isfinite(rl.final_cost));
}

static void test_optimization_decay_edge_cases() {  // optimizer.rs:400-437
  const SimilarityDatabase db = four_tiles();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  PathGrid g = four_grid();
  OptimizationConfig zero;
  zero.max_iterations = 10, zero.temperature_decay = 0.0f;
  const OptimizationResult rz = MosaicOptimizer(calc, zero).optimize(g);
  CHECK(rz.iterations == 10 && std::isfinite(rz.final_cost));
  PathGrid g2 = four_grid();
  OptimizationConfig fast;
  fast.max_iterations = 50, fast.initial_temperature = 1000.0f, fast.temperature_decay = 0.5f;
  const OptimizationResult rf = MosaicOptimizer(calc, fast).optimize(g2);
  CHECK(rf.iterations == 50 && rf.final_cost >= 0.0f);
}

static void test_optimization_zero_and_single_iteration() {  // optimizer.rs:440-475
  const SimilarityDatabase db = four_tiles();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  PathGrid g = four_grid();
  OptimizationConfig c0;
  c0.max_iterations = 0;
  const OptimizationResult r0 = MosaicOptimizer(calc, c0).optimize(g);
  CHECK(r0.iterations == 0 && r0.improved_count == 0 && r0.accepted_count == 0 && r0.initial_cost == r0.final_cost);
  CHECK(g == four_grid());
  OptimizationConfig c1;
  c1.max_iterations = 1;
  const OptimizationResult r1 = MosaicOptimizer(calc, c1).optimize(g);
  CHECK(r1.iterations == 1 && std::isfinite(r1.final_cost));
}

static void test_optimization_large_grid() {  // optimizer.rs:478-518
  SimilarityDatabase db;
  for (int i = 0; i < 25; ++i)
    db.add_tile("tile_" + std::to_string(i) + ".png", Lab{50.0f + (float)i, ((float)i - 12.0f) * 2.0f, ((float)i - 12.0f) * 3.0f});
  db.build_similarities();
  PathGrid grid(5, std::vector<std::optional<std::string>>(5));
  for (int y = 0, t = 0; y < 5; ++y)
    for (int x = 0; x < 5; ++x) grid[y][x] = "tile_" + std::to_string(t++) + ".png";
  const PathGrid before = grid;
  AdjacencyPenaltyCalculator calc(db, 0.5f);
  OptimizationConfig cfg;
  cfg.max_iterations = 100, cfg.report_interval = 25;
  const OptimizationResult r = MosaicOptimizer(calc, cfg).optimize(grid);
  CHECK(r.iterations == 100 && r.final_cost >= 0.0f && r.initial_cost >= 0.0f);
  CHECK(same_multiset(grid, before));
  CHECK(r.best_cost <= r.initial_cost);
  CHECK(r.improved_count <= r.accepted_count && r.accepted_count <= 100);
}

static void test_optimization_identical_tiles() {  // optimizer.rs:589-618
  SimilarityDatabase db;
  for (int i = 1; i <= 4; ++i) db.add_tile("identical" + std::to_string(i) + ".png", Lab{50.0f, 0.0f, 0.0f});
  db.build_similarities();
  PathGrid grid(2, std::vector<std::optional<std::string>>(2));
  grid[0][0] = "identical1.png", grid[0][1] = "identical2.png", grid[1][0] = "identical3.png", grid[1][1] = "identical4.png";
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  OptimizationConfig cfg;
  cfg.max_iterations = 50;
  const OptimizationResult r = MosaicOptimizer(calc, cfg).optimize(grid);
  CHECK(r.initial_cost > 0.0f && r.iterations == 50);
  CHECK(r.initial_cost == 4.0f);  // four pairs at similarity 0
}

static void test_optimization_numerical_stability() {  // optimizer.rs:631-662
  SimilarityDatabase db;
  db.add_tile("extreme1.png", Lab{0.0f, -100.0f, -100.0f});
  db.add_tile("extreme2.png", Lab{100.0f, 100.0f, 100.0f});
  db.add_tile("extreme3.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("extreme4.png", Lab{25.0f, -50.0f, 50.0f});
  db.build_similarities();
  PathGrid grid(2, std::vector<std::optional<std::string>>(2));
  grid[0][0] = "extreme1.png", grid[0][1] = "extreme2.png", grid[1][0] = "extreme3.png", grid[1][1] = "extreme4.png";
  AdjacencyPenaltyCalculator calc(db, 2.0f);
  OptimizationConfig cfg;
  cfg.max_iterations = 30, cfg.initial_temperature = 1000.0f;
  const OptimizationResult r = MosaicOptimizer(calc, cfg).optimize(grid);
  CHECK(std::isfinite(r.initial_cost) && std::isfinite(r.final_cost) && std::isfinite(r.best_cost) && r.iterations == 30);
}

// the mirror's static helpers share one scratch context: any order of calls must work (each injects all it needs)
static void test_helpers_in_any_order() {
  RgbImage img(10, 10);
  for (size_t i = 0; i < img.data.size(); i += 3) img.data[i] = 255;
  const Lab red = MosaicGeneratorImpl::calculate_average_lab(img);  // leaves ONE material on the scratch context
  CHECK(std::fabs(red.l - 53.24f) < 1.0f);
  SimilarityDatabase db;  // ... then a three-tile database is built on it
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("tile2.png", Lab{60.0f, 10.0f, 10.0f});
  db.add_tile("tile3.png", Lab{70.0f, 20.0f, 20.0f});
  db.build_similarities();
  CHECK(db.get_similarity("tile1.png", "tile2.png").has_value());
  CHECK(std::fabs(*db.get_similarity("tile1.png", "tile2.png") - std::sqrt(300.0f)) < 1e-4f);
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  PathGrid grid = {{std::string("tile1.png"), std::string("tile2.png")}};
  const float cost = calc.calculate_total_cost(grid);
  CHECK(std::fabs(cost - 1.0f / (std::sqrt(300.0f) + 1.0f)) < 1e-6f);
  const Lab again = MosaicGeneratorImpl::calculate_average_lab(img);
  CHECK(again.l == red.l && again.a == red.a && again.b == red.b);
}

// A path the database does not know is an occupied cell without similarity terms (adjacency.rs:104-124,
// optimizer.rs:71-102): it contributes nothing to the cost — also next to itself — and the optimiser moves it.
static void test_unknown_paths_are_occupied_cells() {
  SimilarityDatabase db;
  db.add_tile("tile1.png", Lab{50.0f, 0.0f, 0.0f});
  db.add_tile("tile2.png", Lab{60.0f, 10.0f, 10.0f});
  db.build_similarities();
  AdjacencyPenaltyCalculator calc(db, 1.0f);
  const float pair = 1.0f / (std::sqrt(300.0f) + 1.0f);
  PathGrid grid = {{std::string("tile1.png"), std::string("tile2.png")}, {std::string("ghost.png"), std::string("ghost.png")}};
  CHECK(calc.calculate_total_cost(grid) == pair);  // only tile1-tile2; ghost-ghost, tile-ghost pairs are skipped
  PathGrid all_ghosts = {{std::string("g1.png"), std::string("g2.png")}, {std::string("g1.png"), std::string("g3.png")}};
  CHECK(calc.calculate_total_cost(all_ghosts) == 0.0f);
  OptimizationConfig cfg;
  cfg.max_iterations = 200;
  cfg.seed = 3;
  MosaicOptimizer opt(calc, cfg);
  const OptimizationResult r = opt.optimize(all_ghosts);  // every delta is 0: every valid swap is accepted
  CHECK(r.initial_cost == 0.0f && r.final_cost == 0.0f);
  CHECK(r.accepted_count > 0);
  size_t n1 = 0, n2 = 0, n3 = 0;
  for (const auto& row : all_ghosts)
    for (const auto& c : row) {
      CHECK(c.has_value());
      n1 += *c == "g1.png", n2 += *c == "g2.png", n3 += *c == "g3.png";
    }
  CHECK(n1 == 2 && n2 == 1 && n3 == 1);  // the same multiset of paths, only moved
}

static void test_lab_color_calculation() {  // lib.rs:133-158: solid red and gray, ±1.0
  RgbImage red(10, 10), gray(10, 10);
  for (size_t i = 0; i < red.data.size(); i += 3) red.data[i] = 255;
  std::fill(gray.data.begin(), gray.data.end(), (uint8_t)128);
  const Lab r = MosaicGeneratorImpl::calculate_average_lab(red);
  CHECK(std::fabs(r.l - 53.24f) < 1.0f && std::fabs(r.a - 80.09f) < 1.0f && std::fabs(r.b - 67.20f) < 1.0f);
  const Lab g = MosaicGeneratorImpl::calculate_average_lab(gray);
  CHECK(std::fabs(g.l - 53.59f) < 1.0f && std::fabs(g.a) < 1.0f && std::fabs(g.b) < 1.0f);
}

// ---------------------------------------------------------------- end to end (driven by the Python parity test)
static std::vector<uint8_t> read_all(const std::string& f) {
  std::ifstream in(f, std::ios::binary);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}

static int generate(const std::string& dir) {
  int n, mw, mh, tw, th, gw, gh, max_usage, optimize, rebuild;
  unsigned long long iters, seed;
  float weight, strength;
  std::string db_path;
  {
    std::ifstream meta(dir + "/meta.txt");
    meta >> n >> mw >> mh >> tw >> th >> gw >> gh >> max_usage >> weight >> strength >> optimize >> iters >> seed >> rebuild >> db_path;
    if (!meta) {
      std::printf("bad meta.txt\n");
      return 2;
    }
  }
  if (db_path == "-") db_path.clear();
  const std::vector<uint8_t> mats = read_all(dir + "/materials.bin");
  RgbImage target(tw, th);
  target.data = read_all(dir + "/target.bin");
  if (mats.size() != (size_t)n * mw * mh * 3 || target.data.size() != (size_t)tw * th * 3) {
    std::printf("input sizes do not match meta.txt\n");
    return 2;
  }
  std::vector<std::string> paths;
  for (int i = 0; i < n; ++i) paths.push_back("materials/m" + std::to_string(i) + ".png");
  MosaicGenerator gen(paths, mats.data(), mw, mh, (size_t)max_usage, db_path, rebuild != 0, weight, strength);
  PathGrid placed;
  mosaic_report rep{};
  const RgbImage out = gen.generate_mosaic(target, gw, gh, optimize != 0, iters, seed, &placed, &rep);
  std::vector<int32_t> ids;
  std::map<std::string, int32_t> index;
  for (int i = 0; i < n; ++i) index[paths[i]] = i;
  for (const auto& row : placed)
    for (const auto& c : row) ids.push_back(c ? index[*c] : -1);
  std::ofstream(dir + "/placed.bin", std::ios::binary).write((const char*)ids.data(), (std::streamsize)(ids.size() * 4));
  std::ofstream(dir + "/image.bin", std::ios::binary).write((const char*)out.data.data(), (std::streamsize)out.data.size());
  std::ofstream rpt(dir + "/report.txt");
  rpt << out.width << " " << out.height << " " << rep.n_fallbacks << " " << rep.annealed << " " << rep.opt.accepted_count << " "
      << rep.opt.improved_count << " " << gen.similarity_db().len() << "\n";
  return 0;
}

int main(int argc, char** argv) {
  const bool cpu_only = argc > 1 && std::string(argv[1]) == "--cpu";
  if (argc > 2 && std::string(argv[1]) == "--generate") {
    try {
      return generate(argv[2]);
    } catch (const std::exception& e) {
      std::printf("EXCEPTION %s\n", e.what());
      return 1;
    }
  }
  RUN(test_similarity_database_get_lab_color);
  RUN(test_similarity_database_empty);
  RUN(test_similarity_database_load_or_new);
  RUN(test_similarity_database_load_from_invalid_file);
  RUN(test_grid_position_adjacent);
  RUN(test_optimization_result_improvement);
  RUN(test_optimization_config_defaults);
  RUN(test_optimization_empty_grid);
  RUN(test_greedy_optimization_edge_cases);
  RUN(test_aspect_ratio_match);
  RUN(test_tile_creation);
  RUN(test_usage_tracker);
  RUN(test_save_to_file_writes_a_stale_matrix);
  RUN(test_lab_distance_and_delta_e_2000);
  RUN(test_swap_delta);
  RUN(test_image_buffers_are_checked);
  if (cpu_only) {
    RUN(test_no_gpu_is_loud);
  } else {
    RUN(test_similarity_database_basic);
    RUN(test_similarity_database_multiple_tiles);
    RUN(test_similarity_database_single_tile);
    RUN(test_similarity_database_save_load);
    RUN(test_adjacency_penalty_calculation);
    RUN(test_total_cost_calculation);
    RUN(test_optimization_basic);
    RUN(test_greedy_optimization);
    RUN(test_greedy_optimization_zero_iterations);
    RUN(test_optimization_single_tile_grid);
    RUN(test_optimization_grid_with_empty_cells);
    RUN(test_optimization_extreme_temperature_values);
    RUN(test_optimization_decay_edge_cases);
    RUN(test_optimization_zero_and_single_iteration);
    RUN(test_optimization_large_grid);
    RUN(test_optimization_identical_tiles);
    RUN(test_optimization_numerical_stability);
    RUN(test_lab_color_calculation);
    RUN(test_helpers_in_any_order);  // (after the Lab test on purpose: the scratch context holds one material then)
    RUN(test_unknown_paths_are_occupied_cells);
  }
  std::printf("%d tests, %d failed checks\n", g_run, g_failed);
  return g_failed ? 1 : 0;
}
