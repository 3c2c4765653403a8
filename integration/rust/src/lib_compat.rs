//! `lib.rs`-compatible surface over the B200 library.  UNVERIFIED SOURCE: this image has no
//! rustc; the same forwarding is built and tested in C++ (`include/mosaic_b200.hpp`, `tests/cpp/test_host_mirror.cpp`).
//!
//! What a maintainer does with it: `lib.rs` keeps its public items — `Tile`, `trait MosaicGenerator`, `UsageTracker`,
//! the `similarity` / `adjacency` / `optimizer` / `color_adjustment` modules — with their signatures unchanged; the
//! BODIES listed here replace the CPU loops behind a `b200` cargo feature.  Items that are pure host bookkeeping
//! (`Tile`, `UsageTracker`, `GridPosition`, `OptimizationConfig`, `OptimizationResult`, `ColorAdjustment::new`) stay
//! exactly as they are in the crate: the library has nothing to add to a `HashMap` counter, and the placement kernel keeps
//! its own usage counters with the same semantics (lib.rs:57-89 → `greedy.cu`).
//!
//! | public item (reference)                                             | body under `--features b200`             |
//! |---------------------------------------------------------------------|------------------------------------------|
//! | `MosaicGeneratorImpl::calculate_average_lab` (lib.rs:94-119)        | `average_lab` below (K1)                 |
//! | `MosaicGeneratorImpl::is_aspect_ratio_match` (lib.rs:121-123)       | unchanged (one comparison)               |
//! | `SimilarityDatabase::build_similarities` (similarity.rs:69-82)      | `build_similarities` below (K3)          |
//! | `AdjacencyPenaltyCalculator::calculate_total_cost` (adjacency.rs:91-133) | `total_cost` below (K5's cost pass) |
//! | `MosaicOptimizer::optimize` (optimizer.rs:41-129)                   | `optimize` below (K5)                    |
//! | `MosaicOptimizer::optimize_greedy` (optimizer.rs:132-198)           | `optimize_greedy` below (K5, greedy mode)|
//! | `AdjacencyPenaltyCalculator::calculate_penalty` / `calculate_swap_delta` (adjacency.rs:63-88, 133-214) | unchanged: one- and two-cell queries over `get_similarity`; the loops that call them thousands of times are `mosaic_place` and `optimize` |
//! | `calculate_lab_distance`, `calculate_delta_e_2000` (similarity.rs:141-178) | unchanged (one pair; the n(n-1)/2 pairs are `build_similarities`) |
//! | `UsageTracker`, `Tile`, `GridPosition`, config / result structs     | unchanged                                |
use std::collections::HashMap;
use std::os::raw::{c_float, c_int};
use std::path::PathBuf;

use anyhow::{anyhow, Result};

use crate::b200::{mosaic_ctx, mosaic_opt_result, Ctx};

extern "C" {
    fn mosaic_set_materials_rgb8(ctx: *mut mosaic_ctx, rgb: *const u8, n: c_int, w: c_int, h: c_int) -> c_int;
    fn mosaic_material_lab(ctx: *mut mosaic_ctx, lab_out: *mut c_float) -> c_int;
    fn mosaic_set_material_lab(ctx: *mut mosaic_ctx, lab: *const c_float, n: c_int) -> c_int;
    fn mosaic_similarity_tri(ctx: *mut mosaic_ctx, sim_out: *mut c_float) -> c_int;
    fn mosaic_set_cell_lab(ctx: *mut mosaic_ctx, lab: *const c_float, grid_w: c_int, grid_h: c_int) -> c_int;
    fn mosaic_set_placement(ctx: *mut mosaic_ctx, placed: *const i32) -> c_int;
    fn mosaic_anneal(ctx: *mut mosaic_ctx, iterations: u64, t0: c_float, decay: c_float, seed: u64, placed_out: *mut i32,
                     res: *mut mosaic_opt_result) -> c_int;
    fn mosaic_optimize_greedy(ctx: *mut mosaic_ctx, max_iterations: u64, seed: u64, placed_out: *mut i32,
                              res: *mut mosaic_opt_result) -> c_int;
}

fn ok(rc: c_int, what: &str) -> Result<()> {
    if rc == 0 { Ok(()) } else { Err(anyhow!("mosaic_b200: {what} failed ({rc})")) }
}

/// `MosaicGeneratorImpl::calculate_average_lab(img)` (lib.rs:94-119): one image is a material set of one.
/// (Batch the tiles through `Ctx::set_materials` + `Ctx::material_lab` instead wherever the caller has them all —
/// `load_tiles` does, main.rs:193-208.)
pub fn average_lab(ctx: &mut Ctx, img: &image::RgbImage) -> Result<[f32; 3]> {
    let mut lab = [0f32; 3];
    unsafe {
        ok(mosaic_set_materials_rgb8(ctx.raw(), img.as_raw().as_ptr(), 1, img.width() as c_int, img.height() as c_int), "set_materials")?;
        ok(mosaic_material_lab(ctx.raw(), lab.as_mut_ptr()), "material_lab")?;
    }
    Ok(lab)
}

/// `SimilarityDatabase::build_similarities` (similarity.rs:69-82): `lab_colors` in index order in, the packed upper
/// triangle out — `similarities[i * n - i * (i + 1) / 2 + (j - i - 1)]` for `i < j`, the layout the crate stores.
pub fn build_similarities(ctx: &mut Ctx, lab_colors: &[[f32; 3]]) -> Result<Vec<f32>> {
    let n = lab_colors.len();
    let mut sim = vec![0f32; if n > 1 { n * (n - 1) / 2 } else { 0 }];
    unsafe {
        ok(mosaic_set_material_lab(ctx.raw(), lab_colors.as_ptr() as *const c_float, n as c_int), "set_material_lab")?;
        ok(mosaic_similarity_tri(ctx.raw(), sim.as_mut_ptr()), "similarity_tri")?;
    }
    Ok(sim)
}

/// A grid of paths as the kernels want it: indices into the database (`-1` = empty cell).  A path the database does not
/// know is an OCCUPIED cell for the reference (the optimiser swaps it, draws for it and cools down, optimizer.rs:71-102);
/// only its similarity terms are skipped, because `get_similarity` returns None — also against itself (similarity.rs:85-90).
/// Every such CELL gets its own pseudo index behind the database's (two cells with the same unknown path must not look
/// like the same material: `sim(i, i)` would be 0.0 and the term 1.0), exactly as the C++ mirror does (`detail::grid_ids`).
fn grid_ids(path_to_index: &HashMap<PathBuf, usize>, n: usize, grid: &[Vec<Option<PathBuf>>]) -> (Vec<i32>, Vec<PathBuf>) {
    let mut unknown: Vec<PathBuf> = Vec::new();
    let mut ids = Vec::new();
    for row in grid {
        for cell in row {
            ids.push(match cell {
                None => -1,
                Some(p) => match path_to_index.get(p) {
                    Some(&i) => i as i32,
                    None => {
                        unknown.push(p.clone());
                        (n + unknown.len() - 1) as i32
                    }
                },
            });
        }
    }
    (ids, unknown)
}

fn load_grid(ctx: &mut Ctx, lab_colors: &[[f32; 3]], ids: &[i32], gw: usize, gh: usize, n_unknown: usize) -> Result<()> {
    // pseudo materials of the unknown cells: colours so far apart (steps of 3e19 in L) that every squared distance to
    // anything overflows to +inf: 1 / (sqrt(inf) + 1) = 0, the term the reference skips contributes +0.0 and x + 0.0 == x
    // (`detail::unknown_lab` in include/mosaic_b200.hpp; NaN would NOT do: it poisons the sums)
    let mut lab: Vec<[f32; 3]> = lab_colors.to_vec();
    lab.extend((0..n_unknown).map(|k| [3.0e19f32 * (k as f32 + 1.0), 0.0, 0.0]));
    if lab.is_empty() {
        lab.push([0.0; 3]); // a grid of empty cells over an empty database: nothing is ever read
    }
    let zeros = vec![[0f32; 3]; gw * gh];
    unsafe {
        ok(mosaic_set_material_lab(ctx.raw(), lab.as_ptr() as *const c_float, lab.len() as c_int), "set_material_lab")?;
        ok(mosaic_set_cell_lab(ctx.raw(), zeros.as_ptr() as *const c_float, gw as c_int, gh as c_int), "set_cell_lab")?;
        ok(mosaic_set_placement(ctx.raw(), ids.as_ptr()), "set_placement")?;
    }
    Ok(())
}

fn write_back(grid: &mut [Vec<Option<PathBuf>>], out: &[i32], index_to_path: &HashMap<usize, PathBuf>, n: usize, unknown: &[PathBuf]) {
    let gw = grid[0].len();
    for (y, row) in grid.iter_mut().enumerate() {
        for (x, cell) in row.iter_mut().enumerate() {
            let id = out[y * gw + x];
            if id >= 0 {
                let id = id as usize;
                *cell = Some(if id < n { index_to_path[&id].clone() } else { unknown[id - n].clone() });
            }
        }
    }
}

/// What `MosaicOptimizer::optimize` / `optimize_greedy` return (optimizer.rs:203-210): the crate's own
/// `OptimizationResult` has these six fields.
pub struct OptOutcome {
    pub initial_cost: f32,
    pub final_cost: f32,
    pub best_cost: f32,
    pub improved_count: usize,
    pub accepted_count: usize,
    pub iterations: usize,
}

impl From<mosaic_opt_result> for OptOutcome {
    fn from(r: mosaic_opt_result) -> Self {
        OptOutcome {
            initial_cost: r.initial_cost,
            final_cost: r.final_cost,
            best_cost: r.best_cost,
            improved_count: r.improved_count as usize,
            accepted_count: r.accepted_count as usize,
            iterations: r.iterations as usize,
        }
    }
}

/// `MosaicOptimizer::optimize(&self, grid)` (optimizer.rs:41-129).  `seed`: the reference draws from the OS-seeded
/// `thread_rng` (optimizer.rs:48); the library needs a seed (`StdRng::seed_from_u64(seed)` stream) — pass
/// `rand::random()` to keep the reference's behaviour of a fresh stream per call.
#[allow(clippy::too_many_arguments)]
pub fn optimize(ctx: &mut Ctx, lab_colors: &[[f32; 3]], path_to_index: &HashMap<PathBuf, usize>,
                index_to_path: &HashMap<usize, PathBuf>, grid: &mut [Vec<Option<PathBuf>>], max_iterations: usize,
                initial_temperature: f32, temperature_decay: f32, seed: u64) -> Result<OptOutcome> {
    if grid.is_empty() || grid[0].is_empty() {
        return Ok(mosaic_opt_result::default().into()); // optimizer.rs:43-45
    }
    let (gh, gw, n) = (grid.len(), grid[0].len(), lab_colors.len());
    let (ids, unknown) = grid_ids(path_to_index, n, grid);
    load_grid(ctx, lab_colors, &ids, gw, gh, unknown.len())?;
    let mut out = vec![-1i32; gw * gh];
    let mut res = mosaic_opt_result::default();
    unsafe {
        ok(mosaic_anneal(ctx.raw(), max_iterations as u64, initial_temperature, temperature_decay, seed, out.as_mut_ptr(), &mut res),
           "anneal")?;
    }
    write_back(grid, &out, index_to_path, n, &unknown);
    Ok(res.into())
}

/// `MosaicOptimizer::optimize_greedy(&self, grid, max_iterations)` (optimizer.rs:132-198).
pub fn optimize_greedy(ctx: &mut Ctx, lab_colors: &[[f32; 3]], path_to_index: &HashMap<PathBuf, usize>,
                       index_to_path: &HashMap<usize, PathBuf>, grid: &mut [Vec<Option<PathBuf>>], max_iterations: usize,
                       seed: u64) -> Result<OptOutcome> {
    if grid.is_empty() || grid[0].is_empty() {
        return Ok(mosaic_opt_result::default().into()); // optimizer.rs:138-141
    }
    let (gh, gw, n) = (grid.len(), grid[0].len(), lab_colors.len());
    let (ids, unknown) = grid_ids(path_to_index, n, grid);
    load_grid(ctx, lab_colors, &ids, gw, gh, unknown.len())?;
    let mut out = vec![-1i32; gw * gh];
    let mut res = mosaic_opt_result::default();
    unsafe {
        ok(mosaic_optimize_greedy(ctx.raw(), max_iterations as u64, seed, out.as_mut_ptr(), &mut res), "optimize_greedy")?;
    }
    write_back(grid, &out, index_to_path, n, &unknown);
    Ok(res.into())
}

/// `AdjacencyPenaltyCalculator::calculate_total_cost(grid)` (adjacency.rs:91-133) × nothing: the weight is applied by
/// the caller exactly where the reference applies it.  Zero iterations of the optimiser report the grid's energy as
/// `initial_cost` (optimizer.rs:49-51) without touching the grid.
pub fn total_cost(ctx: &mut Ctx, lab_colors: &[[f32; 3]], path_to_index: &HashMap<PathBuf, usize>,
                  grid: &[Vec<Option<PathBuf>>]) -> Result<f32> {
    if grid.is_empty() || grid[0].is_empty() {
        return Ok(0.0);
    }
    let (gh, gw, n) = (grid.len(), grid[0].len(), lab_colors.len());
    let (ids, unknown) = grid_ids(path_to_index, n, grid);
    load_grid(ctx, lab_colors, &ids, gw, gh, unknown.len())?;
    let mut res = mosaic_opt_result::default();
    unsafe {
        ok(mosaic_anneal(ctx.raw(), 0, 100.0, 0.99995, 0, std::ptr::null_mut(), &mut res), "anneal(0)")?;
    }
    Ok(res.initial_cost)
}
