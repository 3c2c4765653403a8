/*
 * mosaic_b200.h — C ABI of the B200-native mosaic matching engine (libmosaic_b200.so).
 *
 * Drop-in boundary for the data-parallel matching path of naporin0624/mosaic-art-rust.
 * The reference has no FFI of its own (pure Rust); these entry points are cut at
 * the reference call sites a Rust `extern "C"` block would replace (file:line under
 * /root/reference/src/ cited per function; the Rust-side binding is in INTEGRATION.md).
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 (MOSAIC_OK) or a negative status and
 *    never unwinds across the ABI; mosaic_last_error() gives the message of the last failure.
 *  - the caller owns every host pointer for the duration of the call only.
 *  - one mosaic_ctx is used from one thread at a time (the reference runs the whole path on one
 *    thread, main.rs:529; GUI: one spawn_blocking, gui/app_full.rs:1282).
 *  - material index = position in the array given to mosaic_set_materials_rgb8 = `tiles[idx]`
 *    after truncate (main.rs:255-257) = similarity-DB index of a freshly built DB (main.rs:123-125).
 *  - there is NO CPU fallback: every compute entry point runs hand-written sm_100a CUDA kernels
 *    and fails with MOSAIC_ERR_CUDA when no usable device is present.
 *  - results, stage by stage against the CPU oracle: Lab means <= 1e-4 relative (the cube root runs on the
 *    MUFU pipe, the sums in tree order); FROM THE SAME Lab means the candidate lists (index and distance bits), the
 *    placements, the fallback cells, the annealed grid with its counters and cost bits and the similarity entries are
 *    identical; output pixels <= 1 LSB.  From PIXELS, a run and the oracle's can part ways where two candidates are tied to
 *    within the Lab tolerance (tests/test_gpu_c3.py reports the rate on the 24,000-tile configuration).
 */
#ifndef MOSAIC_B200_H
#define MOSAIC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOSAIC_OK 0
#define MOSAIC_ERR_INVALID (-1) /* bad argument */
#define MOSAIC_ERR_CUDA (-2)    /* CUDA runtime / launch failure, or no device */
#define MOSAIC_ERR_STATE (-3)   /* call order: a required input has not been set */
#define MOSAIC_ERR_NCCL (-4)    /* collective failure (sharded contexts only) */
#define MOSAIC_ERR_NOMEM (-5)

#define MOSAIC_MAX_K 100 /* main.rs:348 `self.tiles.len().min(100)` */

typedef struct mosaic_ctx mosaic_ctx;

/* mirrors OptimizationResult, optimizer.rs:203-210 */
typedef struct {
  float initial_cost, final_cost, best_cost;
  uint64_t improved_count, accepted_count, iterations;
} mosaic_opt_result;

/* stages for mosaic_stage_ms(): device time (CUDA events) of the most recent run of each kernel group */
enum {
  MOSAIC_STAGE_MATERIAL_LAB = 0, /* K1 lab_mean over materials */
  MOSAIC_STAGE_CELL_LAB = 1,     /* K1 lab_mean over target cells */
  MOSAIC_STAGE_TOPK = 2,         /* K2 nn_topk */
  MOSAIC_STAGE_PLACE = 3,        /* K2b greedy replay */
  MOSAIC_STAGE_SIMILARITY = 4,   /* K3 similarity_tri */
  MOSAIC_STAGE_ANNEAL = 5,       /* K5 anneal */
  MOSAIC_STAGE_RENDER = 6,       /* K4 resize+adjust+composite */
  MOSAIC_STAGE_COUNT = 7
};

/* ---- context ------------------------------------------------------------------------------- */
/* One context = one GPU (`device` = CUDA ordinal).  Owns device buffers, a stream and events. */
int mosaic_ctx_create(mosaic_ctx** out, int device);
/* Sharded context: rank `rank` of `world` single-GPU processes of ONE host (1/2/4/8).  Grid rows,
 * material indices and similarity entries are partitioned over ranks; Lab means, candidate lists,
 * material pixels and placements are exchanged with NCCL over NVLink.
 * `nccl_unique_id` is the 128-byte ncclUniqueId created on rank 0 (mosaic_nccl_unique_id) and
 * distributed by the host (e.g. torch.distributed broadcast / MPI / a file). */
int mosaic_ctx_create_sharded(mosaic_ctx** out, int device, int rank, int world,
                              const void* nccl_unique_id, size_t id_bytes);
int mosaic_nccl_unique_id(void* id_out, size_t id_bytes);
void mosaic_ctx_destroy(mosaic_ctx* ctx);
/* ctx may be NULL: the last failure of a call WITHOUT a context (creation, mosaic_simdb_*, the image codecs) on the CALLING
 * THREAD — such calls may run concurrently on several threads (the reference decodes tiles on rayon workers,
 * main.rs:193-208) and each thread reads its own message.  The pointer is valid until the thread's next such call. */
const char* mosaic_last_error(const mosaic_ctx* ctx);
/* Launch on a caller-owned CUDA stream (cudaStream_t as void*); NULL restores the context's own. */
int mosaic_ctx_set_stream(mosaic_ctx* ctx, void* cuda_stream);
int mosaic_ctx_sync(mosaic_ctx* ctx);
int mosaic_stage_ms(mosaic_ctx* ctx, int stage, float* ms_out);
/* number of kernels this library has launched on ctx since creation */
uint64_t mosaic_kernel_launches(const mosaic_ctx* ctx);
/* device time (ms, CUDA events on the context's stream) this rank spent in exchange steps — NCCL collectives and the
 * waits for peer-memory hand-offs — since the previous call of this function.  0 on single-GPU contexts. */
int mosaic_collective_ms(mosaic_ctx* ctx, float* ms_out);
/* How mosaic_generate exchanges data between the ranks of a sharded context: "single" (one GPU), "peer" (every rank's
 * buffers mapped into every other rank over NVLink, kernels store into the consumer's memory, flag hand-offs) or
 * "nccl: <why the peer windows could not be mapped>" (grouped NCCL broadcasts).  Valid after the first mosaic_generate. */
const char* mosaic_exchange_path(const mosaic_ctx* ctx);
/* number of SMs of the context's device (grid sizing, FP32 roofline of the matching kernel) */
int mosaic_sm_count(const mosaic_ctx* ctx);

/* ---- page-locked host memory ---------------------------------------------------------------- */
/* What makes the host-buffer path fast (DESIGN.md (e')): with page-locked buffers the uploads run on their own stream
 * under the matching, rows are downloaded under the placement chain, and the tiles annealing changed are written straight
 * into the caller's image.  A host that owns its buffers (a Rust Vec<u8>, an image::RgbImage) registers them once;
 * a host that lets the library allocate gets page-locked memory directly.  Pageable buffers work everywhere, staged. */
int mosaic_host_alloc(size_t bytes, void** out);    /* cudaHostAlloc(portable | mapped); release with mosaic_host_release */
int mosaic_host_release(void* p);
int mosaic_host_register(void* p, size_t bytes);    /* cudaHostRegister(portable | mapped) of memory the caller owns */
int mosaic_host_unregister(void* p);

/* ---- inputs -------------------------------------------------------------------------------- */
/* Host pointers belong to the caller.  Pageable buffers are staged before the call returns.  With PAGE-LOCKED buffers
 * the uploads are asynchronous (single-GPU contexts run them on their own stream — materials first, then the target in
 * bands of cell rows — and mosaic_generate matches and places each band while the following ones are still on the PCIe
 * link): keep such a buffer unchanged until the next call that returns results, or mosaic_ctx_sync. */
/* n uniform-size RGB8 materials, n*h*w*3 bytes, host memory (replaces the decoded `image::open`
 * results of process_tile, main.rs:263-301).  Copies H2D. */
int mosaic_set_materials_rgb8(mosaic_ctx* ctx, const uint8_t* rgb, int n, int w, int h);
/* n materials of INDIVIDUAL sizes — what the reference handles: it opens files of any size (process_tile,
 * main.rs:263-301: Lab mean of the image as it is) and resizes each to the tile (resize_image, main.rs:710-738).
 * Material i is widths[i] x heights[i] x 3 bytes (row-major RGB8) at rgb + offsets[i]; sizes 1..32768 per axis.
 * Copies H2D.  The Lab means, the resampling coefficients (one table per distinct width / height) and the render
 * follow each material's own size; everything else is unchanged. */
int mosaic_set_materials_ragged_rgb8(mosaic_ctx* ctx, const uint8_t* rgb, const uint64_t* offsets,
                                     const int32_t* widths, const int32_t* heights, int n);
/* Same, but `dev_rgb` is device memory on ctx's GPU that stays valid until replaced (no copy). */
int mosaic_set_materials_device(mosaic_ctx* ctx, const void* dev_rgb, int n, int w, int h);
/* Target image w*h*3 RGB8 and the grid (main.rs:492-496: tile = img / grid, u32 floor). */
int mosaic_set_target_rgb8(mosaic_ctx* ctx, const uint8_t* rgb, int w, int h, int grid_w, int grid_h);
int mosaic_set_target_device(mosaic_ctx* ctx, const void* dev_rgb, int w, int h, int grid_w, int grid_h);
/* Inject Lab means instead of computing them (similarity-database-only use, and stage-wise parity
 * tests): n*3 / (grid_w*grid_h)*3 floats, host memory. */
int mosaic_set_material_lab(mosaic_ctx* ctx, const float* lab, int n);
int mosaic_set_cell_lab(mosaic_ctx* ctx, const float* lab, int grid_w, int grid_h);
/* Lab colours the similarity database holds for the materials (SimilarityDatabase::get_lab_color,
 * similarity.rs:106-109), when they are not the tiles' own means: MosaicGenerator::new keeps the colours of a
 * database loaded from a file and only adds the tiles it does not know (main.rs:120-143), so similarities
 * (adjacency penalty, annealing energy, the matrix itself) come from the file's colours while matching uses
 * the tiles' colours.  n*3 floats in material order; NULL = back to the materials' Lab means (the state
 * after every mosaic_set_materials_*). */
int mosaic_set_similarity_lab(mosaic_ctx* ctx, const float* lab, int n);

/* ---- B1/B2: Lab means  (MosaicGeneratorImpl::calculate_average_lab, lib.rs:94-119) ---------- */
/* B1: all materials (call sites main.rs:280, 294).  lab_out: n*3 floats or NULL (keep on device). */
int mosaic_material_lab(mosaic_ctx* ctx, float* lab_out);
/* B2: all cells of the target grid, raster order (call site main.rs:540-541). */
int mosaic_cell_lab(mosaic_ctx* ctx, float* lab_out);

/* ---- B3: similarity matrix (SimilarityDatabase::build_similarities, similarity.rs:69-82) ---- */
/* packed upper triangle, n(n-1)/2 floats, entry (i<j) at i*n - i(i+1)/2 + j - i - 1
 * (similarity.rs:100).  sim_out may be NULL (matrix stays on the device). */
int mosaic_similarity_tri(mosaic_ctx* ctx, float* sim_out);
/* Sharded build without the exchange: this rank's range [tri_begin, tri_end) of the packed matrix (mosaic_shard) is
 * computed and, if range_out is not NULL, copied to the host (tri_end - tri_begin floats).  On a single-GPU context
 * the range is the whole matrix.  Row blocks of a 50,000-material database (5 GB) stay on the GPU that built them. */
int mosaic_similarity_local(mosaic_ctx* ctx, float* range_out);

/* ---- B4: matching + greedy placement -------------------------------------------------------- */
/* K nearest materials per cell (kdtree.nearest_n::<SquaredEuclidean>, main.rs:348-352), ascending
 * by (squared distance, material index); k = min(n, k) <= MOSAIC_MAX_K entries per cell.
 * idx_out / d2_out: C*k each, or NULL. */
int mosaic_candidates(mosaic_ctx* ctx, int k, uint32_t* idx_out, float* d2_out);
/* Raster-order greedy placement with usage cap, same-image-adjacent rule, adjacency penalty and the
 * usage-reset fallback (find_and_use_best_tile_with_position + fallback_tile_selection,
 * main.rs:335-475; UsageTracker lib.rs:57-89; calculate_penalty adjacency.rs:63-88).
 * placed_out: C material indices, raster order (NULL = keep on device).  n_fallbacks: number of
 * cells at which the fallback fired.  Runs mosaic_candidates first if it has not been run. */
int mosaic_place(mosaic_ctx* ctx, int max_usage, float adjacency_weight,
                 int32_t* placed_out, uint32_t* n_fallbacks);
/* Raster indices of the first `cap` cells at which the fallback fired during the last mosaic_place. */
int mosaic_fallback_cells(mosaic_ctx* ctx, uint32_t* cells_out, uint32_t cap, uint32_t* n_out);
/* Replace the context's placement (C ids, -1 = empty), e.g. to anneal or render a given grid. */
int mosaic_set_placement(mosaic_ctx* ctx, const int32_t* placed);

/* ---- B5: annealing (MosaicOptimizer::optimize, optimizer.rs:41-129) ------------------------- */
/* Works on the context's placement; placed_out (may be NULL) receives the annealed grid.
 * The RNG stream is rand 0.8.5 StdRng::seed_from_u64(seed) (ChaCha12) — the reference uses the
 * unseeded thread_rng (optimizer.rs:48), so the seed is an extension. */
int mosaic_anneal(mosaic_ctx* ctx, uint64_t iterations, float t0, float decay, uint64_t seed,
                  int32_t* placed_out, mosaic_opt_result* res);

/* MosaicOptimizer::optimize_greedy (optimizer.rs:132-198): the same random swaps, accepted only when the cost drops
 * (delta < 0.0); no temperature, no acceptance sample — four position draws per iteration.  best_cost = final_cost and
 * improved_count = accepted_count = the swaps performed, as the reference reports them. */
int mosaic_optimize_greedy(mosaic_ctx* ctx, uint64_t max_iterations, uint64_t seed, int32_t* placed_out,
                           mosaic_opt_result* res);

/* ---- B6: render (resize + colour adjustment + composite; main.rs:547-592, 625-664, 687-738;
 *          color_adjustment.rs:40-154) ---------------------------------------------------------- */
/* out_rgb: (grid_h*tile_h) x (grid_w*tile_w) x 3 bytes, or NULL (image stays on the device). */
int mosaic_render(mosaic_ctx* ctx, float color_strength, uint8_t* out_rgb);

/* ---- whole path in one call (generate_mosaic, main.rs:477-685, minus disk I/O and codecs) ---- */
typedef struct {
  int max_usage;             /* --max-usage-per-image (default 3) */
  float adjacency_weight;    /* --adjacency-penalty-weight (0.3) */
  float color_strength;      /* --color-adjustment-strength (0.3) */
  int enable_optimization;   /* --enable-optimization (true) */
  uint64_t optimization_iterations; /* --optimization-iterations (1000) */
  uint64_t seed;             /* extension: annealer seed */
  int build_similarity;      /* materialise the packed matrix (main.rs:126,143); it is never read by
                                the kernels (they recompute entries bit-identically from Lab).  A sharded
                                context computes its own range here; mosaic_similarity_tri gathers the
                                whole matrix on every rank */
} mosaic_params;
typedef struct {
  uint32_t n_fallbacks;
  mosaic_opt_result opt; /* zero if annealing did not run */
  int annealed;          /* 1 if the optimisation phase ran (main.rs:608) */
} mosaic_report;
/* Requires materials + target to be set.  placed_out (C) and out_rgb may be NULL. */
int mosaic_generate(mosaic_ctx* ctx, const mosaic_params* params, int32_t* placed_out,
                    uint8_t* out_rgb, mosaic_report* report);

/* ---- similarity-database file ------------------------------------------ */
/* The JSON file of SimilarityDatabase::save_to_file / load_from_file (similarity.rs:31-42, 112-123):
 * {"path_to_index": {path: idx}, "index_to_path": {"idx": path}, "lab_colors": [{"l","a","b"}],
 *  "similarities": [packed upper triangle]} in serde_json's pretty layout, f32 in shortest round-trip form.
 * Host-only: no CUDA device is needed for these calls (except mosaic_simdb_save_ctx). */
typedef struct mosaic_simdb mosaic_simdb;
/* paths[i] = path of material i (UTF-8; anything else is refused like serde refuses a non-UTF-8 PathBuf), lab = n*3,
 * sim = n(n-1)/2 (NULL allowed when n < 2). */
int mosaic_simdb_save_json(const char* file, const char* const* paths, int n, const float* lab, const float* sim);
/* Same with an explicit number of similarity entries: the reference serialises whatever `similarities` holds, also a
 * list that is out of date because tiles were added after the last build_similarities (similarity.rs:112-116). */
int mosaic_simdb_save_json_n(const char* file, const char* const* paths, int n, const float* lab, const float* sim,
                             uint64_t n_similarities);
/* Lab means + freshly computed packed matrix of a context -> file (main.rs:122-131). */
int mosaic_simdb_save_ctx(mosaic_ctx* ctx, const char* file, const char* const* paths);
/* SimilarityDatabase::load_from_file (similarity.rs:119-123).  Accepts what `serde_json::from_str` accepts for the struct —
 * UTF-8 only, strict JSON numbers and strings, unknown fields skipped, duplicate or missing fields refused, `usize` keys and
 * values as plain integer literals — because `load_or_new` (similarity.rs:126-139) rebuilds from scratch on ANY error:
 * what loads and what does not is behaviour.  On error *out is NULL and mosaic_last_error(NULL) says where. */
int mosaic_simdb_load_json(const char* file, mosaic_simdb** out);
void mosaic_simdb_free(mosaic_simdb* db);
/* rows of lab_colors, entries of similarities, entries of path_to_index (they differ in GUI-touched files) */
int mosaic_simdb_info(const mosaic_simdb* db, int* n_rows, uint64_t* n_similarities, int* n_paths);
const char* mosaic_simdb_path(const mosaic_simdb* db, int index);       /* index_to_path; "" if the file has no entry for
                                                                          * that row, NULL outside lab_colors' rows */
int64_t mosaic_simdb_index_of(const mosaic_simdb* db, const char* path); /* path_to_index; -1 if absent */
int mosaic_simdb_copy(const mosaic_simdb* db, float* lab_out, float* sim_out); /* either may be NULL */
/* SimilarityDatabase::get_similarity (similarity.rs:85-103); MOSAIC_ERR_STATE where the reference returns None */
int mosaic_simdb_get_similarity(const mosaic_simdb* db, const char* path1, const char* path2, float* out);

/* ---- progress surface -------------------------------------------------- */
/* The GUI's `(f32, String)` progress channel (src/gui/app_full.rs:1597-1616) and the CLI's progress lines
 * (src/main.rs:521-527, 603-606) are fed from stage events of the DEVICE: while a callback is installed,
 * mosaic_generate enqueues a host function behind every stage, and it reports as the device passes them —
 * fractions as in app_full.rs:1716-2193 (0.40 tiles loaded, 0.45 similarity database, 0.50-0.90 grid rows as the
 * placement chain commits them, 0.92 / 0.98 optimisation, 0.99 image ready; loading and saving are the host's 0.05
 * and 1.0).  `stage` is the MOSAIC_STAGE_* the event belongs to.  The callback runs on a thread of the CUDA runtime
 * and must not call into this library or the CUDA API; NULL removes it (no host functions are enqueued then). */
typedef void (*mosaic_progress_fn)(float fraction, int stage, const char* message, void* user);
int mosaic_set_progress_callback(mosaic_ctx* ctx, mosaic_progress_fn fn, void* user);

/* ---- host-side codec (no device work) ------------------------------ */
/* image::open(path)?.to_rgb8() for a PNG file (src/main.rs:263-301, 484, 548, 632): every colour type and bit depth,
 * Adam7; alpha dropped, grey replicated, 16-bit samples -> (v + 128) / 257.  *rgb_out is malloc'ed
 * (width * height * 3 bytes): release it with mosaic_host_free. */
int mosaic_png_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height);
/* RgbImage::save(path) for a ".png" path (src/main.rs:668-672): 8-bit RGB, lossless (the file bytes are this
 * writer's own choice of filters and deflate level; the pixels are the contract). */
int mosaic_png_encode_rgb8(const char* file, const uint8_t* rgb, int width, int height);
/* The same for a JPEG file (include/mosaic_jpeg.hpp): baseline, extended sequential and progressive Huffman processes, 8-bit,
 * grey / YCbCr / RGB, any sampling factors, restart intervals — with the arithmetic of the IJG / libjpeg-turbo defaults
 * (islow inverse DCT, fancy upsampling, 16-bit colour tables: bit-identical to Pillow's decode); the reference's decoder
 * (image 0.25.6 / zune-jpeg) may differ from any other conforming decoder by +-1 per channel [3P, parity unpinned]. */
int mosaic_jpeg_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height);
/* RgbImage::save(path) for a ".jpg" / ".jpeg" path (src/main.rs:668-672): baseline, YCbCr 4:4:4, Annex K tables scaled by the
 * IJG quality formula (quality <= 0: 75, the `image` crate's default) — the file libjpeg writes at that setting. */
int mosaic_jpeg_encode_rgb8(const char* file, const uint8_t* rgb, int width, int height, int quality);
/* The same for a WebP file (include/mosaic_webp.hpp; src/main.rs:181 accepts the extension): simple and extended (VP8X)
 * containers; VP8L lossless — every transform, colour cache, meta prefix codes: the file's own pixels — and VP8 lossy key frames
 * (RFC 6386) taken to RGB the way libwebp does by default (fancy chroma upsampling, its fixed-point BT.601): bit-identical to
 * Pillow's decode.  Alpha is dropped like to_rgb8() drops it; animated files are refused.  [3P image-webp: parity unpinned] */
int mosaic_webp_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height);
/* RgbImage::save(path) for a ".webp" path (src/main.rs:668-672): lossless VP8L, as the `image` crate writes WebP — the
 * pixels are the contract, the file bytes this writer's choice (subtract-green, one set of prefix codes, literals). */
int mosaic_webp_encode_rgb8(const char* file, const uint8_t* rgb, int width, int height);
/* PNG, JPEG or WebP, told apart by the file's first bytes (image::open for the formats src/main.rs:181 lists). */
int mosaic_image_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height);
/* image::open(path)?.to_rgb8() with image::open's own rule — "the image's format is determined from the path's file
 * extension" (lower-cased; png / jpg / jpeg / webp): the content is NOT sniffed, a JPEG file named x.png fails in the PNG
 * decoder like it does in the reference, where load_tiles then drops the tile with a message (src/main.rs:198-206). */
int mosaic_image_open_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height);
void mosaic_host_free(void* p);

/* ---- self test ------------------------------------------------------------------------------ */
/* Checks the branch-free IEEE sqrt / reciprocal sequences used on the serial chains against the
 * correctly rounded builtins over all 2^32 float bit patterns.  mismatches[0] = sqrt, [1] = rcp. */
int mosaic_selftest(mosaic_ctx* ctx, uint64_t mismatches[2]);

/* ---- diagnostics ------------------------------------------------------------------------------ */
/* Placement-chain profile.  `enable` != 0 makes later mosaic_place calls launch the instrumented
 * instantiation of the chain kernel (same results, clock64 reads on the consumer warp); counters (may
 * be NULL) receives the numbers of the last instrumented launch: [0..3] consumer cycles spent in cells
 * resolved by the bound-only / exact-first-chunk / generic / generic+reset path, [4..7] the cell counts
 * per path, [8] cycles waiting for producers, [9] cycles waiting for the committer, [10] total cycles. */
int mosaic_debug_chain_profile(mosaic_ctx* ctx, int enable, uint64_t counters[16]);

/* Which placement chain later calls use: 0 = the default, 1 = the 32-candidates-per-cell chain (greedy.cu), 2 = the
 * windowed chain (greedy_window.cu: one lane per cell scans, shortlists, two warps alternate windows).  Both replay
 * main.rs:335-475 bit for bit; the first is the faster one on every measured configuration. */
int mosaic_debug_set_chain(mosaic_ctx* ctx, int version);

/* ---- introspection -------------------------------------------------------------------------- */
int mosaic_dims(const mosaic_ctx* ctx, int* n_materials, int* grid_w, int* grid_h, int* tile_w, int* tile_h);
/* device pointers of resident results (NULL if absent): for zero-copy consumers */
const void* mosaic_device_output(const mosaic_ctx* ctx);
const void* mosaic_device_placement(const mosaic_ctx* ctx);
const void* mosaic_device_similarity(const mosaic_ctx* ctx);
/* sharding of this context: rows [row_begin,row_end) of the grid, materials [mat_begin,mat_end),
 * packed similarity entries [tri_begin,tri_end) */
int mosaic_shard(const mosaic_ctx* ctx, int* row_begin, int* row_end, int* mat_begin, int* mat_end,
                 uint64_t* tri_begin, uint64_t* tri_end);
/* The same partition as a pure function (no context, no GPU): rank `rank` of `world` for a grid of
 * grid_h rows and n materials.  Rows and materials are split into contiguous near-equal blocks;
 * the packed similarity triangle [0, n(n-1)/2) is split into equal-size contiguous ranges with
 * 512-entry aligned boundaries (balanced by entry count, not by row count). */
int mosaic_shard_plan(int world, int rank, int grid_h, int n_materials, int* row_begin, int* row_end,
                      int* mat_begin, int* mat_end, uint64_t* tri_begin, uint64_t* tri_end);

#ifdef __cplusplus
}
#endif
#endif
