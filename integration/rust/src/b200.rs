//! FFI of include/mosaic_b200.h and a safe wrapper: `mod b200;` in the reference's lib.rs / main.rs (and `mod lib_compat;`
//! beside it for the `lib.rs`-compatible forwarding bodies: Tile / MosaicGenerator / similarity / adjacency / optimizer).
//! UNVERIFIED SOURCE: there is no rustc/cargo in the image this repository was developed in.  The same sequence of
//! calls is compiled and tested in C++ (include/mosaic_b200.hpp, tests/cpp/test_host_mirror.cpp).
//!
//! One `Ctx` is used from one thread at a time — the reference runs the whole path on one thread (main.rs:529; the GUI
//! wraps the job in a single spawn_blocking, gui/app_full.rs:1282).
#![allow(non_camel_case_types, dead_code)]
use anyhow::{anyhow, Result};
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_float, c_int};
use std::path::Path;

#[repr(C)]
pub struct mosaic_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct mosaic_simdb {
    _private: [u8; 0],
}

/// mirrors OptimizationResult (optimizer.rs:203-210)
#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct mosaic_opt_result {
    pub initial_cost: c_float,
    pub final_cost: c_float,
    pub best_cost: c_float,
    pub improved_count: u64,
    pub accepted_count: u64,
    pub iterations: u64,
}

/// the CLI flags of the matching path (struct Args, main.rs:20-88)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct mosaic_params {
    pub max_usage: c_int,             // --max-usage-per-image
    pub adjacency_weight: c_float,    // --adjacency-penalty-weight
    pub color_strength: c_float,      // --color-adjustment-strength
    pub enable_optimization: c_int,   // --enable-optimization
    pub optimization_iterations: u64, // --optimization-iterations
    pub seed: u64,                    // extension: the reference's thread_rng cannot be seeded (optimizer.rs:48)
    pub build_similarity: c_int,      // materialise the packed matrix (main.rs:126, 143)
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct mosaic_report {
    pub n_fallbacks: u32,
    pub opt: mosaic_opt_result,
    pub annealed: c_int,
}

extern "C" {
    fn mosaic_ctx_create(out: *mut *mut mosaic_ctx, device: c_int) -> c_int;
    fn mosaic_ctx_destroy(ctx: *mut mosaic_ctx);
    fn mosaic_ctx_sync(ctx: *mut mosaic_ctx) -> c_int;
    fn mosaic_last_error(ctx: *const mosaic_ctx) -> *const c_char;
    fn mosaic_set_materials_rgb8(ctx: *mut mosaic_ctx, rgb: *const u8, n: c_int, w: c_int, h: c_int) -> c_int;
    fn mosaic_set_materials_ragged_rgb8(ctx: *mut mosaic_ctx, rgb: *const u8, offsets: *const u64, widths: *const i32,
                                        heights: *const i32, n: c_int) -> c_int;
    fn mosaic_set_target_rgb8(ctx: *mut mosaic_ctx, rgb: *const u8, w: c_int, h: c_int, grid_w: c_int, grid_h: c_int) -> c_int;
    fn mosaic_set_material_lab(ctx: *mut mosaic_ctx, lab: *const c_float, n: c_int) -> c_int;
    fn mosaic_set_similarity_lab(ctx: *mut mosaic_ctx, lab: *const c_float, n: c_int) -> c_int;
    fn mosaic_material_lab(ctx: *mut mosaic_ctx, lab_out: *mut c_float) -> c_int;
    fn mosaic_cell_lab(ctx: *mut mosaic_ctx, lab_out: *mut c_float) -> c_int;
    fn mosaic_similarity_tri(ctx: *mut mosaic_ctx, sim_out: *mut c_float) -> c_int;
    fn mosaic_candidates(ctx: *mut mosaic_ctx, k: c_int, idx_out: *mut u32, d2_out: *mut c_float) -> c_int;
    fn mosaic_place(ctx: *mut mosaic_ctx, max_usage: c_int, adjacency_weight: c_float, placed_out: *mut i32,
                    n_fallbacks: *mut u32) -> c_int;
    fn mosaic_set_placement(ctx: *mut mosaic_ctx, placed: *const i32) -> c_int;
    fn mosaic_anneal(ctx: *mut mosaic_ctx, iterations: u64, t0: c_float, decay: c_float, seed: u64, placed_out: *mut i32,
                     res: *mut mosaic_opt_result) -> c_int;
    fn mosaic_render(ctx: *mut mosaic_ctx, color_strength: c_float, out_rgb: *mut u8) -> c_int;
    fn mosaic_generate(ctx: *mut mosaic_ctx, params: *const mosaic_params, placed_out: *mut i32, out_rgb: *mut u8,
                       report: *mut mosaic_report) -> c_int;
    fn mosaic_dims(ctx: *const mosaic_ctx, n: *mut c_int, grid_w: *mut c_int, grid_h: *mut c_int, tile_w: *mut c_int,
                   tile_h: *mut c_int) -> c_int;
    fn mosaic_simdb_save_ctx(ctx: *mut mosaic_ctx, file: *const c_char, paths: *const *const c_char) -> c_int;
    fn mosaic_simdb_load_json(file: *const c_char, out: *mut *mut mosaic_simdb) -> c_int;
    fn mosaic_simdb_get_similarity(db: *const mosaic_simdb, p1: *const c_char, p2: *const c_char, out: *mut c_float) -> c_int;
    fn mosaic_simdb_free(db: *mut mosaic_simdb);
}

fn last_error(ctx: *const mosaic_ctx) -> String {
    unsafe { CStr::from_ptr(mosaic_last_error(ctx)).to_string_lossy().into_owned() }
}

/// One GPU.  Moved into `spawn_blocking` as a whole, never shared.
pub struct Ctx(*mut mosaic_ctx);
unsafe impl Send for Ctx {}

impl Ctx {
    /// Fails when there is no sm_100a device: the library has no CPU fallback, the caller decides what to do.
    pub fn new(device: i32) -> Result<Self> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { mosaic_ctx_create(&mut p, device) };
        if rc != 0 {
            return Err(anyhow!("mosaic_b200 ({rc}): {}", last_error(std::ptr::null())));
        }
        Ok(Ctx(p))
    }
    /// the raw handle, for the forwarding bodies of `lib_compat.rs`
    pub(crate) fn raw(&self) -> *mut mosaic_ctx {
        self.0
    }
    fn check(&self, rc: c_int) -> Result<()> {
        if rc == 0 {
            Ok(())
        } else {
            Err(anyhow!("mosaic_b200 ({rc}): {}", last_error(self.0)))
        }
    }

    /// `tiles[i]` decoded to RGB8 at its own size (process_tile, main.rs:263-301); index i == position in
    /// `self.tiles` after `truncate(max_materials)` (main.rs:255-257).
    pub fn set_materials(&mut self, images: &[image::RgbImage]) -> Result<()> {
        let (mut blob, mut off, mut w, mut h) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for im in images {
            off.push(blob.len() as u64);
            w.push(im.width() as i32);
            h.push(im.height() as i32);
            blob.extend_from_slice(im.as_raw());
        }
        self.check(unsafe {
            mosaic_set_materials_ragged_rgb8(self.0, blob.as_ptr(), off.as_ptr(), w.as_ptr(), h.as_ptr(), images.len() as c_int)
        })
    }
    /// Lab means of all materials (replaces calculate_average_lab per tile, main.rs:280, 294).
    pub fn material_lab(&mut self, n: usize) -> Result<Vec<[f32; 3]>> {
        let mut v = vec![[0f32; 3]; n];
        self.check(unsafe { mosaic_material_lab(self.0, v.as_mut_ptr() as *mut c_float) })?;
        Ok(v)
    }
    /// Colours a similarity database loaded from a file holds for the tiles, in `tiles` order (main.rs:135-143).
    pub fn set_similarity_lab(&mut self, lab: &[[f32; 3]]) -> Result<()> {
        self.check(unsafe { mosaic_set_similarity_lab(self.0, lab.as_ptr() as *const c_float, lab.len() as c_int) })
    }
    /// Packed upper triangle (replaces build_similarities, similarity.rs:69-82).
    pub fn similarity_tri(&mut self, n: usize) -> Result<Vec<f32>> {
        let mut v = vec![0f32; if n > 1 { n * (n - 1) / 2 } else { 0 }];
        self.check(unsafe { mosaic_similarity_tri(self.0, v.as_mut_ptr()) })?;
        Ok(v)
    }
    pub fn set_target(&mut self, img: &image::RgbImage, grid_w: u32, grid_h: u32) -> Result<()> {
        self.check(unsafe {
            mosaic_set_target_rgb8(self.0, img.as_raw().as_ptr(), img.width() as c_int, img.height() as c_int,
                                   grid_w as c_int, grid_h as c_int)
        })
    }
    /// generate_mosaic (main.rs:477-685) minus disk I/O: placements (indices into `tiles`) and the output image.
    pub fn generate(&mut self, p: &mosaic_params, cells: usize, out: &mut image::RgbImage) -> Result<(Vec<i32>, mosaic_report)> {
        let mut placed = vec![-1i32; cells];
        let mut rep = mosaic_report::default();
        self.check(unsafe { mosaic_generate(self.0, p, placed.as_mut_ptr(), out.as_mut_ptr(), &mut rep) })?;
        Ok((placed, rep))
    }
    /// (tile_w, tile_h) of the current grid: the output image is grid_w*tile_w x grid_h*tile_h (main.rs:495-503).
    pub fn tile_size(&self) -> (u32, u32) {
        let (mut n, mut gw, mut gh, mut tw, mut th) = (0, 0, 0, 0, 0);
        unsafe { mosaic_dims(self.0, &mut n, &mut gw, &mut gh, &mut tw, &mut th) };
        (tw as u32, th as u32)
    }
    /// SimilarityDatabase::save_to_file from the context's Lab means and matrix (main.rs:122-131).
    pub fn save_similarity_db(&mut self, file: &Path, paths: &[std::path::PathBuf]) -> Result<()> {
        let file = CString::new(file.to_string_lossy().as_bytes())?;
        let owned: Vec<CString> = paths.iter().map(|p| CString::new(p.to_string_lossy().as_bytes()).unwrap()).collect();
        let ptrs: Vec<*const c_char> = owned.iter().map(|s| s.as_ptr()).collect();
        self.check(unsafe { mosaic_simdb_save_ctx(self.0, file.as_ptr(), ptrs.as_ptr()) })
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { mosaic_ctx_destroy(self.0) }
    }
}
