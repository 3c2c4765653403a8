/*
 * mosaic_oracle.h — C API of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a scalar CPU restatement of the reference
 * (naporin0624/mosaic-art-rust) hot path, used as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing in the product (mosaic-art-rust_b200/, include/) may include, link or
 * call it.
 *
 * PARITY PINNING: the reference is Rust and cannot be built in this image (no
 * cargo/rustc, crates not vendored).  The oracle is pinned against every numeric
 * known-answer the reference's own tests hold for this path (SURVEY.md §4 / App. B:
 * Lab red/gray goldens lib.rs:133-158, avg-RGB integers main.rs:866-889, ctor clamps
 * color_adjustment.rs:169-176/444-452, packed index similarity.rs:100, adjacency /
 * optimizer fixtures adjacency.rs:243-290, optimizer.rs:230-245).  Everything the
 * reference's tests do not pin (k-NN tie order, resize arithmetic, HSV last-ulp
 * details, the RNG stream of the annealer) is "parity unpinned": the oracle DEFINES
 * it, following the published algorithms of the pinned third-party crates
 * (palette 0.7.6, kiddo 5.2.2, fast_image_resize 5.1.4, rand 0.8.5/rand_chacha 0.3.1).
 *
 * All citations are file:line under /root/reference/src/.
 */
#ifndef MOSAIC_ORACLE_H
#define MOSAIC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- Lab extraction (lib.rs:94-119; palette 0.7.6 Srgb->Lab D65) ---- */
void orc_rgb8_to_lab(uint8_t r, uint8_t g, uint8_t b, float out[3]);
/* mean Lab of the rw x rh crop at (x0,y0) of an img_w-wide RGB8 image, sequential f32 sum, row-major */
void orc_average_lab_region(const uint8_t* rgb, int img_w, int x0, int y0, int rw, int rh, float out[3]);
/* n uniform-size materials, n*h*w*3 bytes -> n*3 floats; threads>1 = Rayon-like split over materials (main.rs:193-208) */
void orc_material_lab(const uint8_t* rgb, int n, int w, int h, float* out, int threads);
/* per-cell Lab means of the target grid, raster order (main.rs:495-496, 536-541) */
void orc_cell_lab(const uint8_t* rgb, int img_w, int img_h, int grid_w, int grid_h, float* out, int threads);
/* integer mean, floor (main.rs:687-708) */
void orc_average_rgb_region(const uint8_t* rgb, int img_w, int x0, int y0, int rw, int rh, uint8_t out[3]);

/* ---- k nearest materials per cell (main.rs:348-352; kiddo SquaredEuclidean) ----
 * ascending by (d2, material index); d2 = ((dl*dl)+(da*da))+(db*db) unfused f32. k = min(n,k). */
void orc_topk(const float* mat_lab, int n, const float* cell_lab, int c, int k,
              uint32_t* out_idx, float* out_d2, int threads);

/* Same result through a bucketed k-d tree (the cost structure of the reference's kiddo search): used by the CPU arm
 * of bench.py; tests pin it to orc_topk. */
void orc_topk_kdtree(const float* mat_lab, int n, const float* cell_lab, int c, int k,
                     uint32_t* out_idx, float* out_d2, int threads);

/* ---- similarity (similarity.rs:69-82, 141-146) ---- */
float orc_lab_distance(const float a[3], const float b[3]);
void orc_similarity_tri(const float* lab, int n, float* out, int threads);
uint64_t orc_tri_index(uint64_t i, uint64_t j, uint64_t n); /* i<j, similarity.rs:100 */

/* ---- adjacency (adjacency.rs:63-214); grid = gh*gw material ids, -1 = empty ---- */
float orc_penalty(const float* mat_lab, const int32_t* grid, int gw, int gh, int x, int y,
                  int32_t candidate, float weight);
float orc_total_cost(const float* mat_lab, const int32_t* grid, int gw, int gh);
float orc_swap_delta(const float* mat_lab, const int32_t* grid, int gw, int gh,
                     int x1, int y1, int x2, int y2);

/* ---- greedy raster placement (main.rs:309-475, lib.rs:57-89) ----
 * cand_* are C x k lists from orc_topk.  fallback_cells (may be NULL) receives up to
 * fallback_cap raster indices at which fallback_tile_selection fired.  Returns #fallbacks. */
uint32_t orc_place(const uint32_t* cand_idx, const float* cand_d2, int k,
                   const float* mat_lab, int n, int gw, int gh,
                   int max_usage, float weight,
                   int32_t* placed_out, uint32_t* usage_out /* n or NULL */,
                   uint32_t* fallback_cells, uint32_t fallback_cap);

/* ---- RNG: rand 0.8.5 StdRng::seed_from_u64 == ChaCha12 (rand_chacha 0.3.1) ---- */
typedef struct orc_rng orc_rng;
orc_rng* orc_rng_new(uint64_t seed);
void orc_rng_free(orc_rng*);
uint32_t orc_rng_next_u32(orc_rng*);
uint64_t orc_rng_next_u64(orc_rng*);
uint64_t orc_rng_gen_range(orc_rng*, uint64_t n); /* gen_range(0..n) on usize */
float orc_rng_gen_f32(orc_rng*);                  /* gen::<f32>() */
void orc_chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]);

/* deterministic, (practically) correctly-rounded f32 exp used by the annealer */
float orc_expf(float x);

/* ---- simulated annealing (optimizer.rs:41-129) ---- */
typedef struct {
  float initial_cost, final_cost, best_cost;
  uint64_t improved_count, accepted_count, iterations;
} orc_opt_result;
void orc_anneal(const float* mat_lab, int32_t* grid, int gw, int gh, uint64_t iterations,
                float t0, float decay, uint64_t seed, orc_opt_result* res);
void orc_anneal_greedy(const float* mat_lab, int32_t* grid, int gw, int gh, uint64_t iterations, uint64_t seed,
                       orc_opt_result* res);

/* ---- resize (main.rs:710-738; fast_image_resize 5.1.4 Convolution(Lanczos3), U8x3) ---- */
/* coefficients of one axis: returns window size; bounds_start/size: out_size; coeffs: out_size*window i16 */
int orc_resize_window(int in_size, int out_size);
int orc_resize_coeffs(int in_size, int out_size, int32_t* bounds_start, int32_t* bounds_size,
                      int16_t* coeffs, int* precision);
void orc_resize_rgb8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);

/* ---- colour adjustment (color_adjustment.rs:28-154) ---- */
void orc_color_adjustment_new(float brightness, float contrast, float hue_shift, float saturation, float out[4]);
void orc_optimal_adjustment(const uint8_t tile_avg[3], const uint8_t target_avg[3], float strength, float out[4]);
void orc_adjust_pixel(const float adj[4], const uint8_t in[3], uint8_t out[3]);
void orc_rgb_to_hsv(float r, float g, float b, float out[3]);
void orc_hsv_to_rgb(float h, float s, float v, float out[3]);

/* ---- render: resize + adjust + composite (main.rs:547-592, 625-664) ----
 * out is (gh*th) x (gw*tw) x 3 with tw = img_w/gw, th = img_h/gh. */
void orc_render(const uint8_t* target, int img_w, int img_h, int gw, int gh,
                const uint8_t* materials, int n, int mw, int mh,
                const int32_t* placed, float strength, uint8_t* out, int threads);

#ifdef __cplusplus
}
#endif
#endif
