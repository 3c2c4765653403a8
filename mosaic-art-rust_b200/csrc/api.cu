// api.cu — the extern "C" ABI of include/mosaic_b200.h: context, inputs, stage entry points and
// the fused whole-path call.  Host code only; every compute step is one of the sm_100a kernels in
// lab_mean.cu / nn_topk.cu / greedy.cu / similarity.cu / anneal.cu / render.cu.  There is no CPU
// fallback: without a usable CUDA device every entry point fails with MOSAIC_ERR_CUDA.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace mosaic {

// errors of calls that have no context (creation, the similarity-database file, the image codecs): one message PER THREAD —
// the reference decodes its tiles on rayon's workers (main.rs:193-208), and a worker that asks for the message of its own
// failed call must not read the message of its neighbour's
static thread_local std::string g_err;

void set_global_error(const std::string& msg) { g_err = msg; }

int fail(mosaic_ctx* ctx, int code, const std::string& msg) {
  if (ctx)
    ctx->err = msg;
  else
    set_global_error(msg);
  return code;
}

int cuda_fail(mosaic_ctx* ctx, cudaError_t e, const char* what) {
  std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  (void)cudaGetLastError();  // clear the sticky-less error state
  return fail(ctx, MOSAIC_ERR_CUDA, m);
}

namespace {

inline size_t cells_of(const mosaic_ctx* c) { return (size_t)c->gw * (size_t)c->gh; }

// contiguous split of [0, total) over `world` ranks, boundaries rounded down to `align`
inline void split_range(uint64_t total, int world, int rank, uint64_t align, uint64_t* b, uint64_t* e) {
  *b = split_bound(total, world, rank, align);
  *e = split_bound(total, world, rank + 1, align);
}

void shard_rows(const mosaic_ctx* c, int rank, int* b, int* e) {
  uint64_t rb, re;
  split_range((uint64_t)c->gh, c->world, rank, 1, &rb, &re);
  *b = (int)rb;
  *e = (int)re;
}
void shard_mats(const mosaic_ctx* c, int rank, int* b, int* e) {
  uint64_t rb, re;
  split_range((uint64_t)c->n, c->world, rank, 1, &rb, &re);
  *b = (int)rb;
  *e = (int)re;
}
void shard_tri(const mosaic_ctx* c, int rank, uint64_t* b, uint64_t* e) {
  const uint64_t n = (uint64_t)c->n;
  const uint64_t total = n > 1 ? n * (n - 1) / 2 : 0;
  split_range(total, c->world, rank, 512, b, e);  // 512-entry (2 KB) aligned: whole warp steps, 16-byte stores
}

// gather per-rank contiguous slices [off_r, off_r+size_r) of one device buffer on every rank
template <class F>
int allgather_slices(mosaic_ctx* c, void* buf, size_t elem_bytes, F slice_of_rank) {
  if (c->world == 1) return MOSAIC_OK;
  std::vector<size_t> offs(c->world), sizes(c->world);
  for (int r = 0; r < c->world; ++r) {
    uint64_t b, e;
    slice_of_rank(r, &b, &e);
    offs[r] = (size_t)b * elem_bytes;
    sizes[r] = (size_t)(e - b) * elem_bytes;
  }
  return dist_allgatherv_inplace(c, buf, offs.data(), sizes.data());
}

// every compute entry point starts a fresh set of stage intervals (mosaic_stage_ms reports the most recent call)
void begin_call(mosaic_ctx* c) {
  for (int s = 0; s < MOSAIC_STAGE_COUNT; ++s) c->ev_count[s] = 0;
  c->progress_events.clear();  // (every entry point synchronises before it returns: nobody reads the old events)
  c->progress_reported = 0.0f;
}

int require(mosaic_ctx* c, bool cond, const char* what) {
  if (!cond) return fail(c, MOSAIC_ERR_STATE, std::string("missing input: ") + what);
  return MOSAIC_OK;
}

int ensure_mat_lab_buffers(mosaic_ctx* c, int n) {
  const int n_pad = (n + 3) & ~3;
  if (c->mat_lab_cap < (size_t)n_pad || !c->d_mat_lab) {
    if (c->d_mat_lab) cudaFree(c->d_mat_lab);
    if (c->d_mat_lab4) cudaFree(c->d_mat_lab4);
    if (c->d_mat_soa) cudaFree(c->d_mat_soa);
    c->d_mat_lab = nullptr;
    c->d_mat_lab4 = nullptr;
    c->d_mat_soa = nullptr;
    c->mat_lab_cap = 0;
    MOSAIC_CUDA(c, cudaMalloc((void**)&c->d_mat_lab, sizeof(float) * 3 * (size_t)n_pad));
    MOSAIC_CUDA(c, cudaMalloc((void**)&c->d_mat_lab4, sizeof(float4) * (size_t)n_pad));
    MOSAIC_CUDA(c, cudaMalloc((void**)&c->d_mat_soa, sizeof(float) * 3 * (size_t)n_pad));
    c->mat_lab_cap = (size_t)n_pad;
  }
  c->n_pad = n_pad;
  return MOSAIC_OK;
}

int ensure_cell_buffers(mosaic_ctx* c) {
  const size_t cells = cells_of(c);
  if (c->cell_cap < cells || !c->d_cell_lab) {
    if (c->d_cell_lab) cudaFree(c->d_cell_lab);
    if (c->d_cell_rgbsum) cudaFree(c->d_cell_rgbsum);
    c->d_cell_lab = nullptr;
    c->d_cell_rgbsum = nullptr;
    c->cell_cap = 0;
    MOSAIC_CUDA(c, cudaMalloc((void**)&c->d_cell_lab, sizeof(float) * 3 * cells));
    MOSAIC_CUDA(c, cudaMalloc((void**)&c->d_cell_rgbsum, sizeof(unsigned long long) * 3 * cells));
    c->cell_cap = cells;
  }
  return MOSAIC_OK;
}

int set_grid(mosaic_ctx* c, int w, int h, int gw, int gh) {
  if (w <= 0 || h <= 0 || gw <= 0 || gh <= 0) return fail(c, MOSAIC_ERR_INVALID, "target/grid dimensions must be positive");
  if (w / gw == 0 || h / gh == 0) return fail(c, MOSAIC_ERR_INVALID, "grid finer than the target image (tile size 0)");
  if ((uint64_t)gw * (uint64_t)gh > 0x7fffffffull) return fail(c, MOSAIC_ERR_INVALID, "too many cells");
  c->tw_img = w;
  c->th_img = h;
  c->gw = gw;
  c->gh = gh;
  c->tile_w = w / gw;  // main.rs:495-496
  c->tile_h = h / gh;
  c->have_cell_lab = c->have_cell_rgbsum = c->have_cand = c->have_placed = c->have_out = false;
  return MOSAIC_OK;
}

int d2h(mosaic_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!dst || bytes == 0) return MOSAIC_OK;
  MOSAIC_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return MOSAIC_OK;
}

int sync(mosaic_ctx* c) {
  if (c->upload_stream) MOSAIC_CUDA(c, cudaStreamSynchronize(c->upload_stream));  // (no-op unless an upload is in flight)
  MOSAIC_CUDA(c, cudaStreamSynchronize(c->stream));
  return MOSAIC_OK;
}

// uploads run on their own stream, after whatever the context's stream is still doing with the old buffers
int begin_upload(mosaic_ctx* c) {
  MOSAIC_CUDA(c, cudaEventRecord(c->ev_main_idle, c->stream));
  MOSAIC_CUDA(c, cudaStreamWaitEvent(c->upload_stream, c->ev_main_idle, 0));
  return MOSAIC_OK;
}
// the context's stream must not touch the materials / the target before their upload has landed
int wait_materials(mosaic_ctx* c) {
  if (c->mat_upload_pending) MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_mat_up, 0));
  c->mat_upload_pending = false;
  return MOSAIC_OK;
}
int wait_target(mosaic_ctx* c) {
  if (c->tgt_upload_pending && c->tgt_bands > 0) MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_tgt_band[c->tgt_bands - 1], 0));
  c->tgt_upload_pending = false;
  return MOSAIC_OK;
}

// ---- stage drivers (no host synchronisation inside) ------------------------------------------
int run_material_lab(mosaic_ctx* c) {
  MOSAIC_TRY(require(c, c->have_materials, "materials (mosaic_set_materials_rgb8)"));
  MOSAIC_TRY(ensure_mat_lab_buffers(c, c->n));
  MOSAIC_TRY(wait_materials(c));
  if (c->ragged) {
    // every rank holds all the materials (mosaic_set_materials_ragged_rgb8 uploads everything); index-sharded means
    int b = 0, e = c->n;
    if (c->world > 1) shard_mats(c, c->rank, &b, &e);
    if (e > b) MOSAIC_TRY(launch_material_lab_ragged(c, b, e));
    MOSAIC_TRY(allgather_slices(c, c->d_mat_lab, 3 * sizeof(float), [&](int r, uint64_t* sb, uint64_t* se) {
      int ib, ie;
      shard_mats(c, r, &ib, &ie);
      *sb = (uint64_t)ib;
      *se = (uint64_t)ie;
    }));
  } else if (c->world == 1) {
    MOSAIC_TRY(launch_material_lab(c));
  } else {
    // index-sharded: this rank's slice, then allgather of 12 B x N
    int b, e;
    shard_mats(c, c->rank, &b, &e);
    const uint8_t* all = c->d_materials;
    float* lab_all = c->d_mat_lab;
    const int n_all = c->n;
    c->d_materials = all + (size_t)b * c->mw * c->mh * 3;
    c->d_mat_lab = lab_all + 3 * (size_t)b;
    c->n = e - b;
    int rc = c->n > 0 ? launch_material_lab(c) : MOSAIC_OK;
    c->d_materials = all;
    c->d_mat_lab = lab_all;
    c->n = n_all;
    MOSAIC_TRY(rc);
    MOSAIC_TRY(allgather_slices(c, c->d_mat_lab, 3 * sizeof(float), [&](int r, uint64_t* sb, uint64_t* se) {
      int ib, ie;
      shard_mats(c, r, &ib, &ie);
      *sb = (uint64_t)ib;
      *se = (uint64_t)ie;
    }));
  }
  MOSAIC_TRY(launch_pack_lab(c));
  c->have_mat_lab = true;
  c->have_cand = false;
  c->have_sim = false;
  return MOSAIC_OK;
}

int gather_cell_lab(mosaic_ctx* c) {
  MOSAIC_TRY(allgather_slices(c, c->d_cell_lab, 3 * sizeof(float) * (size_t)c->gw,
                              [&](int r, uint64_t* sb, uint64_t* se) {
                                int ib, ie;
                                shard_rows(c, r, &ib, &ie);
                                *sb = (uint64_t)ib;
                                *se = (uint64_t)ie;
                              }));
  c->cell_lab_gathered = true;
  return MOSAIC_OK;
}

int gather_candidates(mosaic_ctx* c) {
  auto rows = [&](int r, uint64_t* sb, uint64_t* se) {
    int ib, ie;
    shard_rows(c, r, &ib, &ie);
    *sb = (uint64_t)ib;
    *se = (uint64_t)ie;
  };
  MOSAIC_TRY(allgather_slices(c, c->d_cand_idx, sizeof(uint32_t) * (size_t)c->gw * c->k, rows));
  MOSAIC_TRY(allgather_slices(c, c->d_cand_d2, sizeof(float) * (size_t)c->gw * c->k, rows));
  c->cand_gathered = true;
  return MOSAIC_OK;
}

int ensure_cand_buffers(mosaic_ctx* c) {
  const size_t cells = cells_of(c);
  if (c->cand_cap < cells * (size_t)kMaxK || !c->d_cand_idx) {
    if (c->d_cand_idx) cudaFree(c->d_cand_idx);
    if (c->d_cand_d2) cudaFree(c->d_cand_d2);
    c->d_cand_idx = nullptr;
    c->d_cand_d2 = nullptr;
    c->cand_cap = 0;
    MOSAIC_CUDA(c, cudaMalloc((void**)&c->d_cand_idx, sizeof(uint32_t) * cells * kMaxK));
    MOSAIC_CUDA(c, cudaMalloc((void**)&c->d_cand_d2, sizeof(float) * cells * kMaxK));
    c->cand_cap = cells * (size_t)kMaxK;
  }
  return MOSAIC_OK;
}

// sharded contexts upload only their own slice of the material pixels; whoever needs all of them on this rank (the
// NCCL fallback of the render) gathers them here
int gather_material_pixels(mosaic_ctx* c) {
  if (!c->mats_sliced) return MOSAIC_OK;
  const size_t per = (size_t)c->mw * c->mh * 3;
  MOSAIC_TRY(allgather_slices(c, c->d_materials_owned, per, [&](int r, uint64_t* sb, uint64_t* se) {
    int ib, ie;
    shard_mats(c, r, &ib, &ie);
    *sb = (uint64_t)ib;
    *se = (uint64_t)ie;
  }));
  c->mats_sliced = false;
  return MOSAIC_OK;
}

int run_cell_lab(mosaic_ctx* c) {
  MOSAIC_TRY(require(c, c->have_target, "target (mosaic_set_target_rgb8)"));
  MOSAIC_TRY(ensure_cell_buffers(c));
  MOSAIC_TRY(wait_target(c));
  int rb, re;
  shard_rows(c, c->rank, &rb, &re);
  MOSAIC_TRY(launch_cell_lab_range(c, rb * c->gw, re * c->gw));
  MOSAIC_TRY(gather_cell_lab(c));
  c->have_cell_lab = true;
  c->have_cell_rgbsum = true;  // valid for this rank's rows (all the render kernel needs)
  c->have_cand = false;
  return MOSAIC_OK;
}

int run_candidates(mosaic_ctx* c, int k) {
  if (!c->have_mat_lab) MOSAIC_TRY(run_material_lab(c));
  if (c->have_cell_lab && !c->cell_lab_gathered) MOSAIC_TRY(gather_cell_lab(c));
  // a target still on its way up (single GPU): every band's cell Lab means and top-K start when that band has landed
  const bool banded = !c->have_cell_lab && c->tgt_upload_pending && c->world == 1 && c->tgt_bands > 1;
  if (!c->have_cell_lab && !banded) MOSAIC_TRY(run_cell_lab(c));
  if (k <= 0 || k > kMaxK) k = kMaxK;
  if (k > c->n) k = c->n;  // main.rs:348 min(len, 100)
  if (k <= 0) return fail(c, MOSAIC_ERR_STATE, "no materials");
  MOSAIC_TRY(ensure_cand_buffers(c));
  c->k = k;
  if (banded) {
    MOSAIC_TRY(require(c, c->have_target, "target (mosaic_set_target_rgb8)"));
    MOSAIC_TRY(ensure_cell_buffers(c));
    for (int b = 0; b < c->tgt_bands; ++b) {
      const int r0 = c->tgt_band_rows[b], r1 = c->tgt_band_rows[b + 1];
      if (r1 <= r0) continue;
      MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_tgt_band[b], 0));
      MOSAIC_TRY(launch_cell_lab_range(c, r0 * c->gw, r1 * c->gw));
      MOSAIC_TRY(launch_topk_range(c, k, r0 * c->gw, r1 * c->gw));
    }
    c->tgt_upload_pending = false;
    c->have_cell_lab = c->have_cell_rgbsum = true;
    c->have_cand = true;
    return MOSAIC_OK;
  }
  int rb, re;
  shard_rows(c, c->rank, &rb, &re);
  MOSAIC_TRY(launch_topk_range(c, k, rb * c->gw, re * c->gw));
  MOSAIC_TRY(gather_candidates(c));
  c->have_cand = true;
  return MOSAIC_OK;
}

// the result arena: [opt result | fallback info | placements] — one contiguous hand-off from rank 0 to the other ranks
int ensure_result(mosaic_ctx* c, size_t cells) {
  if (c->d_result && c->placed_cap >= cells) return MOSAIC_OK;
  const size_t head = mosaic_ctx::kResultHead;
  static_assert(mosaic_ctx::kResultHead % 16 == 0, "arena head keeps the placements 16-byte aligned");
  unsigned char* fresh = nullptr;
  const size_t cap = std::max<size_t>(cells, 1);
  cudaError_t e = cudaMalloc((void**)&fresh, head + cap * sizeof(int32_t));
  if (e != cudaSuccess)
    return e == cudaErrorMemoryAllocation ? fail(c, MOSAIC_ERR_NOMEM, "cudaMalloc: out of device memory") : cuda_fail(c, e, "cudaMalloc");
  MOSAIC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->d_result) {
    MOSAIC_CUDA(c, cudaMemcpy(fresh, c->d_result, head, cudaMemcpyDeviceToDevice));
    cudaFree(c->d_result);
  } else {
    MOSAIC_CUDA(c, cudaMemset(fresh, 0, head));
  }
  c->d_result = fresh;
  c->result_cap = head + cap * sizeof(int32_t);
  c->d_opt_res = fresh;
  c->d_place_info = reinterpret_cast<uint32_t*>(fresh + 256);
  c->d_placed = reinterpret_cast<int32_t*>(fresh + head);
  c->placed_cap = cap;
  c->have_placed = false;
  return MOSAIC_OK;
}
int ensure_placed(mosaic_ctx* c) { return ensure_result(c, cells_of(c)); }

__global__ void set_u32_kernel(unsigned* p, unsigned v) { *p = v; }

// A target still on its way up (single GPU, chain-kernel configuration): every band's cell Lab means, top-K and
// placement records are computed when the band has landed, and the placement chain — one CTA, serial — starts after the
// first band on its own stream and consumes the records as they appear (a device counter of prepared cells), so that
// most of its 2.6 ms at the 24,000-tile config run under the upload of the remaining bands.
int run_match_place_banded(mosaic_ctx* c, int k, int max_usage, float weight, bool with_progress) {
  const bool uploading = c->tgt_upload_pending;  // bands of an upload in flight (each with its event) — else: bands of rows
  if (!uploading) {
    const int bands = (int)std::max<size_t>(1, std::min<size_t>((size_t)mosaic_ctx::kBands, std::min<size_t>((size_t)c->gh, cells_of(c) / 3000)));
    for (int b = 0; b <= bands; ++b) c->tgt_band_rows[b] = (int)((long long)c->gh * b / bands);
    c->tgt_bands = bands;
  }
  if (k <= 0 || k > kMaxK) k = kMaxK;
  if (k > c->n) k = c->n;
  MOSAIC_TRY(ensure_cand_buffers(c));
  c->k = k;
  MOSAIC_TRY(ensure_cell_buffers(c));
  MOSAIC_TRY(ensure_placed(c));
  MOSAIC_TRY(ensure(c, &c->d_ready_cells, &c->ready_cells_cap, (size_t)1));
  MOSAIC_CUDA(c, cudaMemsetAsync(c->d_ready_cells, 0, sizeof(unsigned), c->stream));
  if (with_progress) {
    MOSAIC_TRY(ensure(c, &c->d_progress, &c->progress_cap, (size_t)1));
    MOSAIC_CUDA(c, cudaMemsetAsync(c->d_progress, 0, sizeof(unsigned), c->stream));
  }
  for (int b = 0; b < c->tgt_bands; ++b) {
    const int r0 = c->tgt_band_rows[b], r1 = c->tgt_band_rows[b + 1];
    if (r1 > r0) {
      if (uploading) MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_tgt_band[b], 0));
      MOSAIC_TRY(launch_cell_lab_range(c, r0 * c->gw, r1 * c->gw));
      // the first band is on the critical path; the others only have to keep ahead of the chain (experiment knob)
      static const int band_rounds = getenv("MOSAIC_BAND_ROUNDS") ? atoi(getenv("MOSAIC_BAND_ROUNDS")) : 0;
      c->topk_force_rounds = b == 0 ? 0 : band_rounds;
      int trc = launch_topk_range(c, k, r0 * c->gw, r1 * c->gw);
      c->topk_force_rounds = 0;
      MOSAIC_TRY(trc);
      MOSAIC_TRY(launch_place_prep_range(c, weight, r0 * c->gw, r1 * c->gw));
      set_u32_kernel<<<1, 1, 0, c->stream>>>(c->d_ready_cells, (unsigned)(r1 * c->gw));
      c->launches++;
    }
    if (b == 0) MOSAIC_CUDA(c, cudaEventRecord(c->ev_chain_go, c->stream));  // the counter is reset and covers the first band
    if (b == 0 || b + 1 == c->tgt_bands) trace_point(c, b == 0 ? "band0_prepared" : "last_band_prepared");
  }
  // The chain is ENQUEUED after every band's kernels (the host gets here long before the first band has landed) but
  // only waits for the first band's event: on the device it starts as soon as that band is prepared.  In this order a
  // tool that serialises kernel execution (ncu) runs the bands first and the chain last instead of dead-locking on a
  // chain that polls for records of kernels not launched yet.
  MOSAIC_CUDA(c, cudaStreamWaitEvent(c->chain_stream, c->ev_chain_go, 0));
  trace_point(c, "chain_starts", c->chain_stream);
  const int place_slot = stage_interval_begin(c, MOSAIC_STAGE_PLACE, c->chain_stream);
  MOSAIC_TRY(launch_place_chain_async(c, max_usage, weight, c->chain_stream, c->d_ready_cells, with_progress ? c->d_progress : nullptr));
  stage_interval_end(c, MOSAIC_STAGE_PLACE, place_slot, c->chain_stream);
  MOSAIC_CUDA(c, cudaEventRecord(c->ev_chain_done, c->chain_stream));
  MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_chain_done, 0));
  c->tgt_upload_pending = false;
  c->have_cell_lab = c->have_cell_rgbsum = true;
  c->have_cand = true;
  c->have_placed = true;
  return MOSAIC_OK;
}

int run_place(mosaic_ctx* c, int max_usage, float weight) {
  if (!c->have_cand) MOSAIC_TRY(run_candidates(c, kMaxK));
  MOSAIC_TRY(ensure_placed(c));
  if (c->rank == 0) MOSAIC_TRY(launch_place(c, max_usage, weight));  // serial chain: one GPU
  if (c->world > 1) {
    MOSAIC_TRY(dist_broadcast(c, c->d_placed, cells_of(c) * sizeof(int32_t), 0));
    MOSAIC_TRY(dist_broadcast(c, c->d_place_info, (1 + mosaic_ctx::kFallbackCap) * sizeof(uint32_t), 0));
  }
  c->have_placed = true;
  return MOSAIC_OK;
}

int run_anneal(mosaic_ctx* c, uint64_t iterations, float t0, float decay, uint64_t seed) {
  MOSAIC_TRY(require(c, c->have_placed, "placement (mosaic_place / mosaic_set_placement)"));
  MOSAIC_TRY(require(c, c->have_mat_lab, "material Lab means"));
  if (c->rank == 0) MOSAIC_TRY(launch_anneal(c, iterations, t0, decay, seed));  // one Markov chain: one GPU
  if (c->world > 1) {
    MOSAIC_TRY(dist_broadcast(c, c->d_placed, cells_of(c) * sizeof(int32_t), 0));
    MOSAIC_TRY(dist_broadcast(c, c->d_opt_res, sizeof(mosaic_opt_result), 0));
  }
  return MOSAIC_OK;
}

int run_similarity(mosaic_ctx* c, bool gather);

// ---- sharded mosaic_generate over peer windows (no collective on the data path) -----------------------------------------
//   every rank : Lab means of its material slice -> stored into every rank's table; similarity range; Lab means and
//                top-K of its grid rows, the candidate lists stored straight into RANK 0's buffers over NVLink
//   rank 0     : placement records band by band as the other ranks' lists land, the serial chain behind them (gated
//                instantiation), annealing; the result arena (placements, fallback info, optimisation result) is
//                stored into every rank's arena
//   every rank : renders its rows, reading the placed materials from their owners' memory; one barrier at the end
//                (nobody overwrites a window a peer may still be reading)
int run_render(mosaic_ctx* c, float strength, uint8_t* out_rgb);
__global__ void wait_progress_kernel(const volatile unsigned* progress, unsigned need);

int run_generate_peer(mosaic_ctx* c, const mosaic_params* p, bool anneal, uint8_t* out_rgb) {
  const int world = c->world, rank = c->rank;
  int mb, me, rb, re;
  shard_mats(c, rank, &mb, &me);
  shard_rows(c, rank, &rb, &re);
  int k = std::min(kMaxK, c->n);
  c->k = k;
  if (c->mat_signal_pending) {  // this rank's material slice has landed (the upload is ordered on this stream)
    MOSAIC_TRY(dist_peer_signal(c, FLAG_MAT + rank, c->ep_mat, -1, c->stream));
    c->mat_signal_pending = false;
  }
  // --- material Lab means: own slice, then into every rank's table
  MOSAIC_TRY(wait_materials(c));
  if (c->ragged) {
    if (me > mb) MOSAIC_TRY(launch_material_lab_ragged(c, mb, me));
  } else {
    const uint8_t* all = c->d_materials;
    float* lab_all = c->d_mat_lab;
    const int n_all = c->n;
    c->d_materials = all + (size_t)mb * c->mw * c->mh * 3;
    c->d_mat_lab = lab_all + 3 * (size_t)mb;
    c->n = me - mb;
    int rc = c->n > 0 ? launch_material_lab(c) : MOSAIC_OK;
    c->d_materials = all;
    c->d_mat_lab = lab_all;
    c->n = n_all;
    MOSAIC_TRY(rc);
  }
  const unsigned e_lab = ++c->ep_lab;
  {
    CollTimer t(c, c->stream);
    MOSAIC_TRY(dist_peer_push(c, mosaic_ctx::PW_MAT_LAB, (size_t)mb * 12, (size_t)(me - mb) * 12, -1, c->stream));
    MOSAIC_TRY(dist_peer_signal(c, FLAG_LAB + rank, e_lab, -1, c->stream));
    MOSAIC_TRY(dist_peer_wait(c, FLAG_LAB, world, e_lab, c->stream));
  }
  MOSAIC_TRY(launch_pack_lab(c));
  trace_point(c, "lab_exchanged");
  c->have_mat_lab = true;
  c->have_sim = false;
  bool sim_aside = false;
  if (p->build_similarity) {
    // this rank's range of the matrix: nothing on the placement path reads it — built on the render stream
    MOSAIC_CUDA(c, cudaEventRecord(c->ev_mat_lab, c->stream));
    MOSAIC_CUDA(c, cudaStreamWaitEvent(c->render_stream, c->ev_mat_lab, 0));
    cudaStream_t caller = c->stream;
    c->stream = c->render_stream;
    const int src = run_similarity(c, /*gather=*/false);
    c->stream = caller;
    MOSAIC_TRY(src);
    MOSAIC_CUDA(c, cudaEventRecord(c->ev_sim_done, c->render_stream));
    sim_aside = true;
  }
  trace_point(c, "sim");
  // --- own rows: cell Lab means + top-K, stored into rank 0's candidate buffers.  Rank 0 — whose chain waits for nothing
  // but its own first rows — matches them in bands of ~3,000 cells and prepares each band's records right away, so the
  // chain starts behind the first band instead of behind the whole share.
  const bool chain0 = rank == 0 && place_chain_eligible(c, p->max_usage);
  const int own_rows = re - rb;
  // (a target still on its way up arrives in bands of these very rows: mosaic_set_target_rgb8)
  const bool tgt_banded = chain0 && c->tgt_upload_pending && c->tgt_bands > 0 && c->tgt_band_rows[0] == rb && c->tgt_band_rows[c->tgt_bands] == re;
  if (!tgt_banded) MOSAIC_TRY(wait_target(c));
  const int own_bands = tgt_banded ? c->tgt_bands : (chain0 ? std::max(1, std::min(mosaic_ctx::kBands, (int)((size_t)own_rows * c->gw / 3000))) : 1);
  if (chain0) {
    MOSAIC_TRY(ensure(c, &c->d_ready_cells, &c->ready_cells_cap, (size_t)1));
    MOSAIC_CUDA(c, cudaMemsetAsync(c->d_ready_cells, 0, sizeof(unsigned), c->stream));
  }
  c->have_cell_lab = c->have_cell_rgbsum = true;
  c->cell_lab_gathered = false;
  for (int b = 0; b < own_bands; ++b) {
    const int r0 = tgt_banded ? c->tgt_band_rows[b] : rb + (int)((long long)own_rows * b / own_bands);
    const int r1 = tgt_banded ? c->tgt_band_rows[b + 1] : rb + (int)((long long)own_rows * (b + 1) / own_bands);
    if (tgt_banded) MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_tgt_band[b], 0));
    if (r1 <= r0) continue;
    MOSAIC_TRY(launch_cell_lab_range(c, r0 * c->gw, r1 * c->gw));
    c->cand_dst_idx = reinterpret_cast<uint32_t*>(c->pw_peer[mosaic_ctx::PW_CAND_IDX][0]);
    c->cand_dst_d2 = reinterpret_cast<float*>(c->pw_peer[mosaic_ctx::PW_CAND_D2][0]);
    const int trc = launch_topk_range(c, k, r0 * c->gw, r1 * c->gw);
    c->cand_dst_idx = nullptr;
    c->cand_dst_d2 = nullptr;
    MOSAIC_TRY(trc);
    if (chain0) {
      MOSAIC_TRY(launch_place_prep_range(c, p->adjacency_weight, r0 * c->gw, r1 * c->gw));
      set_u32_kernel<<<1, 1, 0, c->stream>>>(c->d_ready_cells, (unsigned)(r1 * c->gw));
      c->launches++;
      if (b == 0) MOSAIC_CUDA(c, cudaEventRecord(c->ev_chain_go, c->stream));
    }
  }
  c->tgt_upload_pending = false;
  trace_point(c, "topk");
  const unsigned e_cand = ++c->ep_cand, e_res = ++c->ep_res, e_grd = ++c->ep_grd;
  c->have_cand = true;
  c->cand_gathered = false;
  // With annealing the GREEDY grid is handed to every rank as soon as the chain ends: the ranks draw (and download) their
  // rows while rank 0 anneals a copy, and compose again only the cells of their rows that the final grid changes — the
  // sharded form of the single-GPU schedule.  (Same decision on every rank: `anneal` comes from the parameters.)
  const bool two_phase = anneal;
  const size_t arena_bytes = mosaic_ctx::kResultHead + cells_of(c) * sizeof(int32_t);
  uint8_t* host_dev = nullptr;
  {
    cudaPointerAttributes attr;
    if (out_rgb && cudaPointerGetAttributes(&attr, out_rgb) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      host_dev = reinterpret_cast<uint8_t*>(attr.devicePointer);
    (void)cudaGetLastError();
  }
  const bool early = two_phase && (out_rgb == nullptr || host_dev != nullptr);  // (a pageable destination is drawn once, at the end)
  if (two_phase) MOSAIC_TRY(ensure(c, &c->d_placed_greedy, &c->placed_greedy_cap, cells_of(c)));
  // Rows UNDER the chain: rank 0's chain counts its committed cells in a word of its flag window; every rank polls it
  // (over NVLink) and draws + downloads a band of its rows as soon as the chain has passed it, reading the placements
  // from rank 0's arena.  (Same decision on every rank: the chain's eligibility depends on the configuration only.)
  const bool under_capable = two_phase && place_chain_eligible(c, p->max_usage) && place_chain_has_progress(c);
  // (each rank decides for itself: with a host destination the download of its rows hides under the chain; an image that
  // stays on the device gains nothing from it — drawing beside the chain slows the chain — and is drawn beside the annealer)
  const bool under = under_capable && early && out_rgb != nullptr;
  unsigned* const progress0 = reinterpret_cast<unsigned*>(c->pw_peer[mosaic_ctx::PW_FLAGS][0]) + FLAG_PROGRESS;
  const int32_t* const placed0 = reinterpret_cast<const int32_t*>(reinterpret_cast<const unsigned char*>(c->pw_peer[mosaic_ctx::PW_RESULT][0]) +
                                                                  mosaic_ctx::kResultHead);
  // this rank's rows as the chain passes them, on stream `on` (copies on the copy stream); ends with the thumbnails of
  // EVERY placed material cached (the final grid may bring any of them into these rows) once `grid_complete` has fired
  auto draw_rows_under_chain = [&](cudaStream_t on) -> int {
    if (c->mats_sliced) MOSAIC_TRY(dist_peer_wait(c, FLAG_MAT, world, c->ep_mat, on));
    c->peer_render_ready = true;
    c->have_placed = true;
    const size_t out_bytes = (size_t)c->gw * c->tile_w * (size_t)c->gh * c->tile_h * 3;
    MOSAIC_TRY(ensure(c, &c->d_out, &c->out_cap, out_bytes));
    const size_t row_bytes = (size_t)c->gw * c->tile_w * 3 * (size_t)c->tile_h;
    const int rows = re - rb;
    int bands = std::min(mosaic_ctx::kBands, std::max(rows, 1));
    if ((size_t)rows * row_bytes < ((size_t)8 << 20)) bands = 1;
    int32_t* home = c->d_placed;
    cudaStream_t caller = c->stream;
    c->d_placed = const_cast<int32_t*>(placed0);
    c->stream = on;
    int rc = render_reset_cache(c);
    for (int b = 0; b < bands && rc == MOSAIC_OK; ++b) {
      const int b0 = rb + (int)((long long)rows * b / bands), b1 = rb + (int)((long long)rows * (b + 1) / bands);
      if (b1 <= b0) continue;
      wait_progress_kernel<<<1, 32, 0, on>>>(progress0, (unsigned)(b1 * c->gw));
      c->launches++;
      {
        StageTimer t(c, MOSAIC_STAGE_RENDER);
        rc = c->ragged ? launch_render_ragged(c, p->color_strength, b0, b1) : launch_render(c, p->color_strength, b0, b1);
      }
      if (rc != MOSAIC_OK) break;
      if (out_rgb) {
        cudaEventRecord(c->ev_band[b], on);
        cudaStreamWaitEvent(c->copy_stream, c->ev_band[b], 0);
        cudaMemcpyAsync(out_rgb + (size_t)b0 * row_bytes, c->d_out + (size_t)b0 * row_bytes, (size_t)(b1 - b0) * row_bytes,
                        cudaMemcpyDeviceToHost, c->copy_stream);
      }
    }
    c->stream = caller;
    c->d_placed = home;
    MOSAIC_TRY(rc);
    c->have_out = true;
    return MOSAIC_OK;
  };
  // (after the greedy grid is complete and copied) the thumbnails of every placed material: one row drawn again with the
  // whole grid marked — identical pixels, and the cache then covers whatever the final grid moves into these rows
  auto cache_all_thumbnails = [&](cudaStream_t on) -> int {
    if (re <= rb) return MOSAIC_OK;
    int32_t* home = c->d_placed;
    cudaStream_t caller = c->stream;
    c->d_placed = c->d_placed_greedy;
    c->stream = on;
    c->thumb_mark_all = true;
    const int rc = c->ragged ? launch_render_ragged(c, p->color_strength, rb, rb + 1) : launch_render(c, p->color_strength, rb, rb + 1);
    c->thumb_mark_all = false;
    c->stream = caller;
    c->d_placed = home;
    return rc;
  };
  // this rank's rows of the greedy image, from the copy of the greedy grid; every placed material's thumbnail is cached
  auto draw_greedy_rows = [&](cudaStream_t on) -> int {
    if (c->mats_sliced) MOSAIC_TRY(dist_peer_wait(c, FLAG_MAT, world, c->ep_mat, on));  // the owners' material slices have landed
    c->peer_render_ready = true;
    c->have_placed = true;
    int32_t* home = c->d_placed;
    cudaStream_t caller = c->stream;
    c->d_placed = c->d_placed_greedy;
    c->stream = on;
    c->thumb_mark_all = true;
    const int rc = run_render(c, p->color_strength, out_rgb);
    c->thumb_mark_all = false;
    c->stream = caller;
    c->d_placed = home;
    return rc;
  };
  if (rank != 0) {
    {
      CollTimer t(c, c->stream);
      MOSAIC_TRY(dist_peer_signal(c, FLAG_CAND + rank, e_cand, 0, c->stream));
    }
    if (two_phase) {
      if (under) {
        {
          CollTimer t(c, c->stream);
          MOSAIC_TRY(dist_peer_wait(c, FLAG_CHN, 1, e_grd, c->stream));  // rank 0 has reset its progress counter
        }
        MOSAIC_TRY(draw_rows_under_chain(c->stream));
      }
      {
        CollTimer t(c, c->stream);
        MOSAIC_TRY(dist_peer_wait(c, FLAG_GRD, 1, e_grd, c->stream));  // the greedy grid + fallback info have landed in the arena
      }
      MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_placed_greedy, c->d_placed, sizeof(int32_t) * cells_of(c), cudaMemcpyDeviceToDevice, c->stream));
      MOSAIC_TRY(dist_peer_signal(c, FLAG_ACK + rank, e_grd, 0, c->stream));  // rank 0 may overwrite the arena with the final grid
      if (under) {
        MOSAIC_TRY(cache_all_thumbnails(c->stream));
        if (out_rgb) {
          MOSAIC_CUDA(c, cudaEventRecord(c->ev_copy_done, c->copy_stream));
          MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
        }
      } else if (early) {
        MOSAIC_TRY(draw_greedy_rows(c->stream));
      }
    }
    CollTimer t(c, c->stream);
    MOSAIC_TRY(dist_peer_wait(c, FLAG_RES, 1, e_res, c->stream));  // placements + report from rank 0
  } else {
    const bool chain = chain0;
    const int place_slot = stage_interval_begin(c, MOSAIC_STAGE_PLACE, c->stream);
    for (int r = 1; r < world; ++r) {  // (rank 0's own rows were prepared band by band above)
      int b0, b1;
      shard_rows(c, r, &b0, &b1);
      {
        CollTimer t(c, c->stream);
        MOSAIC_TRY(dist_peer_wait(c, FLAG_CAND + r, 1, e_cand, c->stream));
      }
      if (chain && b1 > b0) {
        MOSAIC_TRY(launch_place_prep_range(c, p->adjacency_weight, b0 * c->gw, b1 * c->gw));
        set_u32_kernel<<<1, 1, 0, c->stream>>>(c->d_ready_cells, (unsigned)(b1 * c->gw));
        c->launches++;
      }
      trace_point(c, "band_prepared");
    }
    if (chain) {  // (enqueued behind every band's kernels, waits only for the first band: see run_match_place_banded)
      MOSAIC_CUDA(c, cudaStreamWaitEvent(c->chain_stream, c->ev_chain_go, 0));
      if (under_capable) {
        MOSAIC_CUDA(c, cudaMemsetAsync(progress0, 0, sizeof(unsigned), c->chain_stream));
        MOSAIC_TRY(dist_peer_signal(c, FLAG_CHN, e_grd, -1, c->chain_stream));  // the counter starts at zero: the ranks may poll it
      }
      // (row by row only when this rank itself downloads rows — the ranks are expected to agree on the kind of destination;
      // a rank that polls nevertheless is released when the chain ends)
      const bool publish_rows = under_capable && out_rgb != nullptr;
      MOSAIC_TRY(launch_place_chain_async(c, p->max_usage, p->adjacency_weight, c->chain_stream, c->d_ready_cells, publish_rows ? progress0 : nullptr));
      if (under_capable && !publish_rows) {
        set_u32_kernel<<<1, 1, 0, c->chain_stream>>>(progress0, (unsigned)cells_of(c));
        c->launches++;
      }
      MOSAIC_CUDA(c, cudaEventRecord(c->ev_chain_done, c->chain_stream));
      if (under) {  // this rank's own rows, on the render stream, as the chain passes them
        MOSAIC_CUDA(c, cudaEventRecord(c->ev_mat_lab, c->stream));  // (their cell sums are enqueued)
        MOSAIC_CUDA(c, cudaStreamWaitEvent(c->render_stream, c->ev_mat_lab, 0));
        MOSAIC_TRY(dist_peer_wait(c, FLAG_CHN, 1, e_grd, c->render_stream));
        MOSAIC_TRY(draw_rows_under_chain(c->render_stream));
      }
      MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_chain_done, 0));
      stage_interval_end(c, MOSAIC_STAGE_PLACE, place_slot, c->stream);
    } else {
      stage_interval_end(c, MOSAIC_STAGE_PLACE, place_slot, c->stream);
      MOSAIC_TRY(launch_place(c, p->max_usage, p->adjacency_weight));
    }
    c->have_placed = true;
    trace_point(c, "chain_done");
    emit_progress(c, c->stream, 0.90f, MOSAIC_STAGE_PLACE, "Grid processing completed");  // app_full.rs:2098
    if (two_phase) {
      {
        CollTimer t(c, c->stream);
        MOSAIC_TRY(dist_peer_push(c, mosaic_ctx::PW_RESULT, 0, arena_bytes, -1, c->stream));  // the greedy grid
        MOSAIC_TRY(dist_peer_signal(c, FLAG_GRD, e_grd, -1, c->stream));
      }
      MOSAIC_TRY(ensure(c, &c->d_placed_work, &c->placed_work_cap, cells_of(c)));
      MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_placed_greedy, c->d_placed, sizeof(int32_t) * cells_of(c), cudaMemcpyDeviceToDevice, c->stream));
      MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_placed_work, c->d_placed, sizeof(int32_t) * cells_of(c), cudaMemcpyDeviceToDevice, c->stream));
      if (early) {  // on the render stream, beside the annealer: this rank's own rows, or (drawn under the chain) the thumbnails
        MOSAIC_CUDA(c, cudaEventRecord(c->ev_mat_lab, c->stream));  // (the greedy grid is copied)
        MOSAIC_CUDA(c, cudaStreamWaitEvent(c->render_stream, c->ev_mat_lab, 0));
        if (under) {
          MOSAIC_TRY(cache_all_thumbnails(c->render_stream));
          if (out_rgb) {
            MOSAIC_CUDA(c, cudaEventRecord(c->ev_copy_done, c->copy_stream));
            MOSAIC_CUDA(c, cudaStreamWaitEvent(c->render_stream, c->ev_copy_done, 0));
          }
        } else {
          MOSAIC_TRY(draw_greedy_rows(c->render_stream));
        }
        MOSAIC_CUDA(c, cudaEventRecord(c->ev_render_done, c->render_stream));
      }
      {
        int32_t* home = c->d_placed;
        c->d_placed = c->d_placed_work;  // (the annealer's kernels take the grid from the context)
        const int rc = launch_anneal(c, p->optimization_iterations, 100.0f, 0.99995f, p->seed);
        c->d_placed = home;
        MOSAIC_TRY(rc);
      }
      trace_point(c, "annealed");
      CollTimer t(c, c->stream);
      if (world > 1) MOSAIC_TRY(dist_peer_wait(c, FLAG_ACK + 1, world - 1, e_grd, c->stream));  // everybody has copied the greedy grid
      MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_placed, c->d_placed_work, sizeof(int32_t) * cells_of(c), cudaMemcpyDeviceToDevice, c->stream));
      MOSAIC_TRY(dist_peer_push(c, mosaic_ctx::PW_RESULT, 0, arena_bytes, -1, c->stream));
      MOSAIC_TRY(dist_peer_signal(c, FLAG_RES, e_res, -1, c->stream));
      if (early) MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_render_done, 0));
    } else {
      if (anneal) MOSAIC_TRY(launch_anneal(c, p->optimization_iterations, 100.0f, 0.99995f, p->seed));
      trace_point(c, "annealed");
      CollTimer t(c, c->stream);
      MOSAIC_TRY(dist_peer_push(c, mosaic_ctx::PW_RESULT, 0, arena_bytes, -1, c->stream));
      MOSAIC_TRY(dist_peer_signal(c, FLAG_RES, e_res, -1, c->stream));
    }
  }
  c->peer_rendered = false;
  if (early) {
    // the final grid is in the arena: this rank's cells that differ from the greedy grid are composed again, into the
    // device image and — a page-locked destination — straight into the caller's rows
    if (render_patch_possible(c)) {
      StageTimer t(c, MOSAIC_STAGE_RENDER);
      MOSAIC_TRY(launch_render_patch(c, c->d_placed_greedy, host_dev, rb, re));
      c->peer_rendered = true;
    }
  }
  c->have_placed = true;
  // the render reads the placed materials from their owners: every owner's slice must have landed
  if (c->mats_sliced) {
    CollTimer t(c, c->stream);
    MOSAIC_TRY(dist_peer_wait(c, FLAG_MAT, world, c->ep_mat, c->stream));
  }
  c->peer_render_ready = true;
  if (sim_aside) MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_sim_done, 0));
  return MOSAIC_OK;
}

int gather_similarity(mosaic_ctx* c) {
  if (c->world == 1 || c->sim_gathered) return MOSAIC_OK;
  MOSAIC_TRY(allgather_slices(c, c->d_sim, sizeof(float), [&](int r, uint64_t* sb, uint64_t* se) { shard_tri(c, r, sb, se); }));
  c->sim_gathered = true;
  return MOSAIC_OK;
}

// gather = false (the fused path): a sharded context computes its own range and leaves the all-gather to the first
// caller that wants the whole matrix — no kernel reads it (they recompute entries from the Lab means)
int run_similarity(mosaic_ctx* c, bool gather) {
  if (!c->have_mat_lab) MOSAIC_TRY(run_material_lab(c));
  const uint64_t n = (uint64_t)c->n;
  const uint64_t total = n > 1 ? n * (n - 1) / 2 : 0;  // similarity.rs:71 (n == 0 underflows in the reference)
  MOSAIC_TRY(ensure(c, &c->d_sim, &c->sim_cap, (size_t)total));
  uint64_t b, e;
  shard_tri(c, c->rank, &b, &e);
  if (c->world == 1) {
    MOSAIC_TRY(launch_similarity(c, 0, total));
  } else {
    float* full = c->d_sim;
    c->d_sim = full + b;  // the kernel writes entry `begin` at out[0]
    int rc = launch_similarity(c, b, e);
    c->d_sim = full;
    MOSAIC_TRY(rc);
    c->sim_gathered = false;
    if (gather) MOSAIC_TRY(gather_similarity(c));
  }
  c->have_sim = true;
  return MOSAIC_OK;
}

// Renders this rank's rows.  With a host destination the rows are rendered in bands and every finished band
// is copied to the host on a second stream while the next ones render (the output image is the largest
// transfer of the path: 98.5 MB for the 24,000-tile mosaic).
int run_render(mosaic_ctx* c, float strength, uint8_t* out_rgb) {
  MOSAIC_TRY(require(c, c->have_materials, "materials"));
  MOSAIC_TRY(require(c, c->have_target, "target"));
  MOSAIC_TRY(require(c, c->have_placed, "placement"));
  MOSAIC_TRY(wait_materials(c));
  if (c->mats_sliced && !(c->peer_ok && c->peer_render_ready)) MOSAIC_TRY(gather_material_pixels(c));
  if (strength > 0.0f && !c->have_cell_rgbsum) MOSAIC_TRY(run_cell_lab(c));
  if (!c->d_cell_rgbsum) MOSAIC_TRY(ensure_cell_buffers(c));
  const size_t out_bytes = (size_t)c->gw * c->tile_w * (size_t)c->gh * c->tile_h * 3;
  MOSAIC_TRY(ensure(c, &c->d_out, &c->out_cap, out_bytes));
  int rb, re;
  shard_rows(c, c->rank, &rb, &re);  // sharded contexts deliver their own row band (other rows untouched)
  const size_t row_bytes = (size_t)c->gw * c->tile_w * 3 * (size_t)c->tile_h;
  const int rows = re - rb;
  int bands = out_rgb ? std::min(mosaic_ctx::kBands, rows) : 1;
  if ((size_t)rows * row_bytes < ((size_t)8 << 20)) bands = std::min(bands, 1);  // small images: one copy
  StageTimer t(c, MOSAIC_STAGE_RENDER);
  MOSAIC_TRY(render_reset_cache(c));
  for (int b = 0; b < std::max(bands, 1); ++b) {
    const int b0 = rb + (int)((long long)rows * b / std::max(bands, 1)), b1 = rb + (int)((long long)rows * (b + 1) / std::max(bands, 1));
    if (b1 <= b0) continue;
    MOSAIC_TRY(c->ragged ? launch_render_ragged(c, strength, b0, b1) : launch_render(c, strength, b0, b1));
    if (out_rgb) {
      MOSAIC_CUDA(c, cudaEventRecord(c->ev_band[b], c->stream));
      MOSAIC_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_band[b], 0));
      MOSAIC_CUDA(c, cudaMemcpyAsync(out_rgb + (size_t)b0 * row_bytes, c->d_out + (size_t)b0 * row_bytes,
                                     (size_t)(b1 - b0) * row_bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    }
  }
  if (out_rgb) {  // the context's stream does not run ahead of the copies: a sync on it covers them
    MOSAIC_CUDA(c, cudaEventRecord(c->ev_copy_done, c->copy_stream));
    MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
  }
  c->have_out = true;
  return MOSAIC_OK;
}

__global__ void wait_progress_kernel(const volatile unsigned* progress, unsigned need) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (*progress < need) {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000ll) __trap();  // the chain died: do not hang the GPU
    }
    __threadfence_system();  // (the counter may live on a peer: the placements it covers were written before it)
  }
}

// The image of the GREEDY placement, rendered and sent to the host band by band WHILE the placement chain is still
// walking the rows below: a band is rendered as soon as the chain's progress counter has passed it.  Kernels on
// render_stream, copies on copy_stream; the caller's stream is not involved (it waits for ev_copy_done at the end).
// The reference renders the greedy placement first as well (main.rs:533-593) and again after annealing (625-664).
int run_render_under_chain(mosaic_ctx* c, float strength, uint8_t* out_rgb) {
  MOSAIC_TRY(wait_materials(c));
  const size_t out_bytes = (size_t)c->gw * c->tile_w * (size_t)c->gh * c->tile_h * 3;
  MOSAIC_TRY(ensure(c, &c->d_out, &c->out_cap, out_bytes));
  const size_t row_bytes = (size_t)c->gw * c->tile_w * 3 * (size_t)c->tile_h;
  // What the render stream needs from the other streams: the materials (upload event), the reset of the progress
  // counter (enqueued before the first band's records, i.e. before ev_chain_go) — everything else is covered by the
  // progress counter itself (a band's cell sums and records exist before the chain can pass it).  NOT an event recorded
  // on the caller's stream now: that stream already waits for the end of the chain.
  MOSAIC_CUDA(c, cudaStreamWaitEvent(c->render_stream, c->ev_mat_up, 0));
  MOSAIC_CUDA(c, cudaStreamWaitEvent(c->render_stream, c->ev_chain_go, 0));
  cudaStream_t caller = c->stream;
  c->stream = c->render_stream;
  int rc = render_reset_cache(c);
  for (int b = 0; b < c->tgt_bands && rc == MOSAIC_OK; ++b) {
    const int b0 = c->tgt_band_rows[b], b1 = c->tgt_band_rows[b + 1];
    if (b1 <= b0) continue;
    wait_progress_kernel<<<1, 32, 0, c->render_stream>>>(c->d_progress, (unsigned)(b1 * c->gw));
    c->launches++;
    emit_progress(c, c->render_stream, 0.5f + 0.4f * (float)b1 / (float)c->gh, MOSAIC_STAGE_PLACE, "Processing grid: %.1f%%",
                  100.0 * b1 / c->gh);  // app_full.rs:1810-1827
    {
      StageTimer t(c, MOSAIC_STAGE_RENDER);  // (c->stream is the render stream here)
      rc = c->ragged ? launch_render_ragged(c, strength, b0, b1) : launch_render(c, strength, b0, b1);
    }
    if (rc != MOSAIC_OK) break;
    cudaEventRecord(c->ev_band[b], c->render_stream);
    if (b == 0 || b + 1 == c->tgt_bands) trace_point(c, b == 0 ? "band0_rendered" : "last_band_rendered", c->render_stream);
    if (out_rgb) {
      cudaStreamWaitEvent(c->copy_stream, c->ev_band[b], 0);
      cudaMemcpyAsync(out_rgb + (size_t)b0 * row_bytes, c->d_out + (size_t)b0 * row_bytes, (size_t)(b1 - b0) * row_bytes,
                      cudaMemcpyDeviceToHost, c->copy_stream);
      if (b == 0 || b + 1 == c->tgt_bands) trace_point(c, b == 0 ? "band0_copied" : "last_band_copied", c->copy_stream);
    }
  }
  c->stream = caller;
  MOSAIC_TRY(rc);
  MOSAIC_CUDA(c, cudaEventRecord(c->ev_render_done, c->render_stream));
  MOSAIC_CUDA(c, cudaEventRecord(c->ev_copy_done, c->copy_stream));
  MOSAIC_CUDA(c, cudaGetLastError());
  c->have_out = true;
  return MOSAIC_OK;
}

}  // namespace
}  // namespace mosaic

using namespace mosaic;

extern "C" {

const char* mosaic_last_error(const mosaic_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  return g_err.c_str();  // the calling thread's own last context-less failure; valid until its next such call
}

static int ctx_create_common(mosaic_ctx** out, int device) {
  if (!out) return fail(nullptr, MOSAIC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    return fail(nullptr, MOSAIC_ERR_CUDA,
                std::string("no usable CUDA device (this library has no CPU fallback): ") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  if (device < 0 || device >= count) return fail(nullptr, MOSAIC_ERR_INVALID, "device ordinal out of range");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, MOSAIC_ERR_CUDA,
                std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                    "; this library ships sm_100a code only");
  mosaic_ctx* c = new mosaic_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(nullptr, e, "cudaStreamCreate");
  }
  c->stream = c->own_stream;
  cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->chain_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->render_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ev_render_done, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_mat_lab, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_sim_done, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_chain_go, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_chain_done, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_main_idle, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_mat_up, cudaEventDisableTiming);
  for (int i = 0; i < mosaic_ctx::kBands; ++i) cudaEventCreateWithFlags(&c->ev_tgt_band[i], cudaEventDisableTiming);
  for (int i = 0; i < mosaic_ctx::kBands; ++i) cudaEventCreateWithFlags(&c->ev_band[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming);
  static_assert(sizeof(mosaic_opt_result) <= 256, "arena layout");
  if (ensure_result(c, 1) != MOSAIC_OK) {
    set_global_error(c->err);
    mosaic_ctx_destroy(c);
    return fail(nullptr, MOSAIC_ERR_CUDA, "context allocation failed");
  }
  *out = c;
  return MOSAIC_OK;
}

int mosaic_ctx_create(mosaic_ctx** out, int device) { return ctx_create_common(out, device); }

int mosaic_ctx_create_sharded(mosaic_ctx** out, int device, int rank, int world, const void* nccl_unique_id,
                              size_t id_bytes) {
  if (world < 1 || rank < 0 || rank >= world) return fail(nullptr, MOSAIC_ERR_INVALID, "bad rank/world");
  MOSAIC_TRY(ctx_create_common(out, device));
  mosaic_ctx* c = *out;
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    int rc = dist_init(c, nccl_unique_id, id_bytes);
    if (rc != MOSAIC_OK) {
      set_global_error(c->err);
      mosaic_ctx_destroy(c);
      *out = nullptr;
      return rc;
    }
  }
  return MOSAIC_OK;
}

int mosaic_nccl_unique_id(void* id_out, size_t id_bytes) { return dist_unique_id(id_out, id_bytes); }

void mosaic_ctx_destroy(mosaic_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  dist_destroy(c);
  void* bufs[] = {c->d_ready_cells, c->d_mat_off, c->d_mat_w, c->d_mat_h, c->d_rag_bounds, c->d_rag_weights, c->d_rag_desc, c->d_rag_midx, c->d_rag_scratch,
                  c->d_sim_lab4_own, c->d_sim_soa_own, c->d_materials_owned, c->d_mat_lab, c->d_mat_lab4, c->d_mat_soa, c->d_target_owned, c->d_cell_lab,
                  c->d_cell_rgbsum,     c->d_cand_idx, c->d_cand_d2, c->d_result,  c->d_usage,        c->d_flags, c->d_peer_err, c->d_place_rec,
                  c->d_patch_list, c->d_patch_count, c->d_placed_greedy, c->d_progress,
                  c->d_sim,             c->d_sim_counter, c->d_rng_ring, c->d_terms, c->d_deltas, c->d_thumbs, c->d_thumb_sums, c->d_thumb_state, c->d_out,    c->axis_h.start,   c->axis_h.size,
                  c->axis_h.w,          c->axis_v.start, c->axis_v.size, c->axis_v.w};
  for (void* p : bufs)
    if (p) cudaFree(p);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (int i = 0; i < MOSAIC_STAGE_COUNT; ++i)
    for (int j = 0; j < mosaic_ctx::kStageIntervals; ++j) {
      if (c->ev_begin[i][j]) cudaEventDestroy(c->ev_begin[i][j]);
      if (c->ev_end[i][j]) cudaEventDestroy(c->ev_end[i][j]);
    }
  if (c->upload_stream) {
    cudaStreamSynchronize(c->upload_stream);
    cudaStreamDestroy(c->upload_stream);
  }
  if (c->chain_stream) {
    cudaStreamSynchronize(c->chain_stream);
    cudaStreamDestroy(c->chain_stream);
  }
  if (c->render_stream) {
    cudaStreamSynchronize(c->render_stream);
    cudaStreamDestroy(c->render_stream);
  }
  if (c->ev_render_done) cudaEventDestroy(c->ev_render_done);
  if (c->ev_mat_lab) cudaEventDestroy(c->ev_mat_lab);
  if (c->ev_sim_done) cudaEventDestroy(c->ev_sim_done);
  if (c->ev_chain_go) cudaEventDestroy(c->ev_chain_go);
  if (c->ev_chain_done) cudaEventDestroy(c->ev_chain_done);
  if (c->ev_main_idle) cudaEventDestroy(c->ev_main_idle);
  if (c->ev_mat_up) cudaEventDestroy(c->ev_mat_up);
  for (int i = 0; i < mosaic_ctx::kBands; ++i)
    if (c->ev_tgt_band[i]) cudaEventDestroy(c->ev_tgt_band[i]);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  for (int i = 0; i < mosaic_ctx::kBands; ++i)
    if (c->ev_band[i]) cudaEventDestroy(c->ev_band[i]);
  if (c->ev_copy_done) cudaEventDestroy(c->ev_copy_done);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  for (int i = 0; i < mosaic_ctx::kCollEvents; ++i)
    for (int j = 0; j < 2; ++j)
      if (c->coll_ev[i][j]) cudaEventDestroy(c->coll_ev[i][j]);
  delete c;
}

int mosaic_ctx_set_stream(mosaic_ctx* c, void* cuda_stream) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return MOSAIC_OK;
}

int mosaic_ctx_sync(mosaic_ctx* c) {
  if (!c) return MOSAIC_ERR_INVALID;
  return sync(c);
}

int mosaic_stage_ms(mosaic_ctx* c, int stage, float* ms_out) {
  if (!c || !ms_out || stage < 0 || stage >= MOSAIC_STAGE_COUNT) return MOSAIC_ERR_INVALID;
  if (c->ev_count[stage] == 0) return fail(c, MOSAIC_ERR_STATE, "stage has not run");
  float total = 0.0f;
  for (int i = 0; i < c->ev_count[stage]; ++i) {
    float ms = 0.0f;
    MOSAIC_CUDA(c, cudaEventSynchronize(c->ev_end[stage][i]));
    MOSAIC_CUDA(c, cudaEventElapsedTime(&ms, c->ev_begin[stage][i], c->ev_end[stage][i]));
    total += ms;
  }
  *ms_out = total;
  return MOSAIC_OK;
}

uint64_t mosaic_kernel_launches(const mosaic_ctx* c) { return c ? c->launches : 0; }

int mosaic_collective_ms(mosaic_ctx* c, float* ms_out) {
  if (!c || !ms_out) return MOSAIC_ERR_INVALID;
  float total = 0.0f;
  for (int i = 0; i < c->coll_n; ++i) {
    float ms = 0.0f;
    MOSAIC_CUDA(c, cudaEventSynchronize(c->coll_ev[i][1]));
    MOSAIC_CUDA(c, cudaEventElapsedTime(&ms, c->coll_ev[i][0], c->coll_ev[i][1]));
    total += ms;
  }
  c->coll_n = 0;
  *ms_out = total;
  return MOSAIC_OK;
}

int mosaic_sm_count(const mosaic_ctx* c) { return c ? c->sm_count : 0; }

const char* mosaic_exchange_path(const mosaic_ctx* c) {
  if (!c || c->world == 1) return "single";
  if (c->peer_ok) return "peer";
  static thread_local std::string note;
  note = "nccl: " + (c->peer_failed ? (c->peer_note.empty() ? std::string("peer windows unavailable") : c->peer_note)
                                    : std::string("peer windows not set up yet"));
  return note.c_str();
}

// ---- inputs ----------------------------------------------------------------------------------
static int set_materials_common(mosaic_ctx* c, int n, int w, int h) {
  if (n <= 0 || w < 0 || h < 0 || ((w == 0) != (h == 0))) return fail(c, MOSAIC_ERR_INVALID, "materials: n, w, h must be positive");
  c->n = n;
  c->mw = w;
  c->mh = h;
  c->have_materials = true;
  c->have_mat_lab = c->have_cand = c->have_placed = c->have_sim = c->have_out = false;
  c->sim_lab_override = false;  // a new material set: the database colours of the old one do not apply
  c->ragged = false;
  c->mats_sliced = false;
  c->materials_version++;
  c->mat_upload_pending = false;  // (an upload of the previous set still in flight is ordered before the next one)
  return MOSAIC_OK;
}

int mosaic_set_materials_rgb8(mosaic_ctx* c, const uint8_t* rgb, int n, int w, int h) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!rgb) return fail(c, MOSAIC_ERR_INVALID, "materials pointer is NULL");
  if (w <= 0 || h <= 0) return fail(c, MOSAIC_ERR_INVALID, "materials: n, w, h must be positive");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  MOSAIC_TRY(set_materials_common(c, n, w, h));
  c->have_materials = false;
  const size_t per = (size_t)w * h * 3;
  MOSAIC_TRY(ensure(c, &c->d_materials_owned, &c->materials_cap, per * (size_t)n));
  c->d_materials = c->d_materials_owned;
  if (c->world == 1) {  // own stream: the target that usually follows queues behind it, compute waits for the event
    MOSAIC_TRY(begin_upload(c));
    MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_materials_owned, rgb, per * (size_t)n, cudaMemcpyHostToDevice, c->upload_stream));
    MOSAIC_CUDA(c, cudaEventRecord(c->ev_mat_up, c->upload_stream));
    c->mat_upload_pending = true;
    c->have_materials = true;
    return MOSAIC_OK;
  }
  int b = 0, e = n;
  shard_mats(c, c->rank, &b, &e);
  // Each rank uploads ITS slice over its own PCIe link and keeps only that: the Lab means need nothing else, and the
  // render reads the materials placed in this rank's rows from their owners' memory over NVLink (peer windows) — or,
  // without peer access, gathers the slices with NCCL when it first needs them (gather_material_pixels).
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_materials_owned + (size_t)b * per, rgb + (size_t)b * per, (size_t)(e - b) * per,
                                 cudaMemcpyHostToDevice, c->stream));
  c->mats_sliced = true;
  c->ep_mat++;  // (every rank calls this function the same number of times)
  c->mat_signal_pending = true;
  c->have_materials = true;
  return MOSAIC_OK;
}

int mosaic_set_materials_ragged_rgb8(mosaic_ctx* c, const uint8_t* rgb, const uint64_t* offsets, const int32_t* widths,
                                     const int32_t* heights, int n) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!rgb || !offsets || !widths || !heights || n <= 0) return fail(c, MOSAIC_ERR_INVALID, "ragged materials: bad arguments");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (widths[i] <= 0 || heights[i] <= 0 || widths[i] > 32768 || heights[i] > 32768)
      return fail(c, MOSAIC_ERR_INVALID, "ragged materials: material " + std::to_string(i) + " has a size outside 1..32768");
    total = std::max(total, (size_t)offsets[i] + (size_t)widths[i] * (size_t)heights[i] * 3);
  }
  MOSAIC_TRY(set_materials_common(c, n, 0, 0));
  c->have_materials = false;
  c->ragged = true;
  MOSAIC_TRY(ensure(c, &c->d_materials_owned, &c->materials_cap, total + 16));  // slack: the Lab kernel reads whole aligned words
  MOSAIC_TRY(ensure(c, &c->d_mat_off, &c->mat_off_cap, (size_t)n));
  MOSAIC_TRY(ensure(c, &c->d_mat_w, &c->mat_w_cap, (size_t)n));
  MOSAIC_TRY(ensure(c, &c->d_mat_h, &c->mat_h_cap, (size_t)n));
  c->d_materials = c->d_materials_owned;
  c->h_mat_off.assign(offsets, offsets + n);
  c->h_mat_w.assign(widths, widths + n);
  c->h_mat_h.assign(heights, heights + n);
  // (sharded contexts: every rank uploads everything — the per-rank PCIe split of the uniform path needs equal sizes)
  MOSAIC_TRY(begin_upload(c));
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_materials_owned, rgb, total, cudaMemcpyHostToDevice, c->upload_stream));
  MOSAIC_CUDA(c, cudaMemsetAsync(c->d_materials_owned + total, 0, 16, c->upload_stream));  // the slack the Lab kernel's last aligned word may touch
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_mat_off, c->h_mat_off.data(), sizeof(unsigned long long) * (size_t)n, cudaMemcpyHostToDevice, c->upload_stream));
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_mat_w, c->h_mat_w.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->upload_stream));
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_mat_h, c->h_mat_h.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->upload_stream));
  MOSAIC_CUDA(c, cudaEventRecord(c->ev_mat_up, c->upload_stream));
  c->mat_upload_pending = true;
  c->have_materials = true;
  return MOSAIC_OK;
}

int mosaic_set_materials_device(mosaic_ctx* c, const void* dev_rgb, int n, int w, int h) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!dev_rgb) return fail(c, MOSAIC_ERR_INVALID, "materials pointer is NULL");
  if (w <= 0 || h <= 0) return fail(c, MOSAIC_ERR_INVALID, "materials: n, w, h must be positive");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  MOSAIC_TRY(set_materials_common(c, n, w, h));
  c->d_materials = (const uint8_t*)dev_rgb;
  return MOSAIC_OK;
}

int mosaic_set_target_rgb8(mosaic_ctx* c, const uint8_t* rgb, int w, int h, int grid_w, int grid_h) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!rgb) return fail(c, MOSAIC_ERR_INVALID, "target pointer is NULL");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  MOSAIC_TRY(set_grid(c, w, h, grid_w, grid_h));
  c->have_target = false;
  MOSAIC_TRY(ensure(c, &c->d_target_owned, &c->target_cap, (size_t)w * h * 3));
  c->d_target = c->d_target_owned;
  const size_t row_bytes = (size_t)w * 3;
  if (c->world == 1) {
    // bands of cell rows on the upload stream, one event each (the last band carries the remainder pixel rows)
    MOSAIC_TRY(begin_upload(c));
    const int bands = ((size_t)w * h * 3 >= ((size_t)8 << 20)) ? std::min(mosaic_ctx::kBands, grid_h) : 1;
    for (int b = 0; b < bands; ++b) {
      const int r0 = (int)((long long)grid_h * b / bands), r1 = (int)((long long)grid_h * (b + 1) / bands);
      const size_t p0 = (size_t)r0 * c->tile_h, p1 = (b + 1 == bands) ? (size_t)h : (size_t)r1 * c->tile_h;
      c->tgt_band_rows[b] = r0;
      c->tgt_band_rows[b + 1] = r1;
      if (p1 > p0)
        MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_target_owned + p0 * row_bytes, rgb + p0 * row_bytes, (p1 - p0) * row_bytes,
                                       cudaMemcpyHostToDevice, c->upload_stream));
      MOSAIC_CUDA(c, cudaEventRecord(c->ev_tgt_band[b], c->upload_stream));
    }
    c->tgt_bands = bands;
    c->tgt_upload_pending = true;
    c->have_target = true;
    return MOSAIC_OK;
  }
  int rb = 0, re = grid_h;
  shard_rows(c, c->rank, &rb, &re);  // only this rank's band of pixel rows is uploaded
  // ... on the upload stream, in bands of ~3,000 cells with one event each: the material Lab means and their exchange run
  // under it, and rank 0 — whose placement chain starts behind its first rows — matches every band as it lands
  MOSAIC_TRY(begin_upload(c));
  const int own_rows = re - rb;
  const int bands = std::max(1, std::min(mosaic_ctx::kBands, (int)((size_t)own_rows * grid_w / 3000)));
  for (int b = 0; b < bands; ++b) {
    const int r0 = rb + (int)((long long)own_rows * b / bands), r1 = rb + (int)((long long)own_rows * (b + 1) / bands);
    const size_t p0 = (size_t)r0 * c->tile_h, p1 = (size_t)r1 * c->tile_h;
    c->tgt_band_rows[b] = r0;
    c->tgt_band_rows[b + 1] = r1;
    if (p1 > p0)
      MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_target_owned + p0 * row_bytes, rgb + p0 * row_bytes, (p1 - p0) * row_bytes,
                                     cudaMemcpyHostToDevice, c->upload_stream));
    MOSAIC_CUDA(c, cudaEventRecord(c->ev_tgt_band[b], c->upload_stream));
  }
  c->tgt_bands = bands;
  c->tgt_upload_pending = true;
  c->have_target = true;
  return MOSAIC_OK;
}

int mosaic_set_target_device(mosaic_ctx* c, const void* dev_rgb, int w, int h, int grid_w, int grid_h) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!dev_rgb) return fail(c, MOSAIC_ERR_INVALID, "target pointer is NULL");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  MOSAIC_TRY(set_grid(c, w, h, grid_w, grid_h));
  c->tgt_upload_pending = false;
  c->d_target = (const uint8_t*)dev_rgb;
  c->have_target = true;
  return MOSAIC_OK;
}

int mosaic_set_material_lab(mosaic_ctx* c, const float* lab, int n) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!lab || n <= 0) return fail(c, MOSAIC_ERR_INVALID, "material Lab: bad arguments");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  // Injected Lab means stand on their own (similarity-database-only use, stage-wise tests, the C++ mirror's static
  // helpers on a shared scratch context): a different count replaces the material set — pixels of the old one no
  // longer apply, and neither do database colours, placements or images made from it.
  if (n != c->n) {
    c->have_materials = false;
    c->mats_sliced = false;
    c->sim_lab_override = false;
    c->have_placed = c->have_out = false;
  }
  c->n = n;
  MOSAIC_TRY(ensure_mat_lab_buffers(c, n));
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_mat_lab, lab, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  MOSAIC_TRY(launch_pack_lab(c));
  MOSAIC_TRY(sync(c));
  c->have_mat_lab = true;
  c->have_cand = c->have_sim = false;
  return MOSAIC_OK;
}

int mosaic_set_similarity_lab(mosaic_ctx* c, const float* lab, int n) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  if (!lab) {  // back to the materials' own Lab means
    if (c->sim_lab_override) c->have_sim = c->have_placed = false;
    c->sim_lab_override = false;
    return MOSAIC_OK;
  }
  if (n <= 0 || n != c->n) return fail(c, MOSAIC_ERR_INVALID, "similarity Lab count differs from the material count");
  const size_t n_pad = (size_t)((n + 3) & ~3);  // == ctx->n_pad of the material tables
  MOSAIC_TRY(ensure(c, &c->d_sim_lab4_own, &c->sim_lab_cap, n_pad));
  MOSAIC_TRY(ensure(c, &c->d_sim_soa_own, &c->sim_soa_cap, 3 * n_pad));
  std::vector<float4> packed(n_pad);
  std::vector<float> soa(3 * n_pad, __builtin_inff());
  for (int i = 0; i < n; ++i) {
    packed[(size_t)i] = make_float4(lab[3 * (size_t)i], lab[3 * (size_t)i + 1], lab[3 * (size_t)i + 2], 0.0f);
    soa[(size_t)i] = lab[3 * (size_t)i], soa[n_pad + i] = lab[3 * (size_t)i + 1], soa[2 * n_pad + i] = lab[3 * (size_t)i + 2];
  }
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_sim_lab4_own, packed.data(), sizeof(float4) * n_pad, cudaMemcpyHostToDevice, c->stream));
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_sim_soa_own, soa.data(), sizeof(float) * 3 * n_pad, cudaMemcpyHostToDevice, c->stream));
  MOSAIC_TRY(sync(c));
  c->sim_lab_override = true;
  c->have_sim = c->have_placed = false;
  return MOSAIC_OK;
}

int mosaic_set_cell_lab(mosaic_ctx* c, const float* lab, int grid_w, int grid_h) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!lab || grid_w <= 0 || grid_h <= 0) return fail(c, MOSAIC_ERR_INVALID, "cell Lab: bad arguments");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  if (c->have_target) {
    if (grid_w != c->gw || grid_h != c->gh) return fail(c, MOSAIC_ERR_INVALID, "cell Lab grid differs from the target grid");
  } else {
    c->gw = grid_w;
    c->gh = grid_h;
  }
  MOSAIC_TRY(ensure_cell_buffers(c));
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_cell_lab, lab, sizeof(float) * 3 * cells_of(c), cudaMemcpyHostToDevice, c->stream));
  MOSAIC_TRY(sync(c));
  c->have_cell_lab = true;
  c->have_cand = false;
  return MOSAIC_OK;
}

// ---- stages ----------------------------------------------------------------------------------
int mosaic_material_lab(mosaic_ctx* c, float* lab_out) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_material_lab(c));
  MOSAIC_TRY(d2h(c, lab_out, c->d_mat_lab, sizeof(float) * 3 * (size_t)c->n));
  return sync(c);
}

int mosaic_cell_lab(mosaic_ctx* c, float* lab_out) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_cell_lab(c));
  MOSAIC_TRY(d2h(c, lab_out, c->d_cell_lab, sizeof(float) * 3 * cells_of(c)));
  return sync(c);
}

int mosaic_similarity_tri(mosaic_ctx* c, float* sim_out) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_similarity(c, /*gather=*/true));
  const uint64_t n = (uint64_t)c->n;
  MOSAIC_TRY(d2h(c, sim_out, c->d_sim, sizeof(float) * (size_t)(n > 1 ? n * (n - 1) / 2 : 0)));
  return sync(c);
}

int mosaic_similarity_local(mosaic_ctx* c, float* range_out) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_similarity(c, /*gather=*/false));
  uint64_t b, e;
  shard_tri(c, c->rank, &b, &e);
  MOSAIC_TRY(d2h(c, range_out, c->d_sim + b, sizeof(float) * (size_t)(e - b)));
  return sync(c);
}

int mosaic_candidates(mosaic_ctx* c, int k, uint32_t* idx_out, float* d2_out) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_candidates(c, k));
  MOSAIC_TRY(d2h(c, idx_out, c->d_cand_idx, sizeof(uint32_t) * cells_of(c) * c->k));
  MOSAIC_TRY(d2h(c, d2_out, c->d_cand_d2, sizeof(float) * cells_of(c) * c->k));
  return sync(c);
}

int mosaic_place(mosaic_ctx* c, int max_usage, float adjacency_weight, int32_t* placed_out, uint32_t* n_fallbacks) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_place(c, max_usage, adjacency_weight));
  MOSAIC_TRY(d2h(c, placed_out, c->d_placed, sizeof(int32_t) * cells_of(c)));
  MOSAIC_TRY(d2h(c, n_fallbacks, c->d_place_info, sizeof(uint32_t)));
  return sync(c);
}

int mosaic_fallback_cells(mosaic_ctx* c, uint32_t* cells_out, uint32_t cap, uint32_t* n_out) {
  if (!c || !n_out) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  MOSAIC_TRY(require(c, c->have_placed, "placement"));
  uint32_t nfb = 0;
  MOSAIC_CUDA(c, cudaMemcpyAsync(&nfb, c->d_place_info, sizeof nfb, cudaMemcpyDeviceToHost, c->stream));
  MOSAIC_TRY(sync(c));
  uint32_t m = std::min(std::min(nfb, cap), mosaic_ctx::kFallbackCap);
  if (m && cells_out) {
    MOSAIC_CUDA(c, cudaMemcpyAsync(cells_out, c->d_place_info + 1, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, c->stream));
    MOSAIC_TRY(sync(c));
  }
  *n_out = nfb;
  return MOSAIC_OK;
}

int mosaic_set_placement(mosaic_ctx* c, const int32_t* placed) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!placed) return fail(c, MOSAIC_ERR_INVALID, "placement pointer is NULL");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  if (c->gw <= 0 || c->gh <= 0) return fail(c, MOSAIC_ERR_STATE, "grid is not set");
  for (size_t i = 0, cells = cells_of(c); i < cells; ++i)  // -1 = empty cell (main.rs:628); anything else indexes the materials
    if (placed[i] < -1 || placed[i] >= c->n)
      return fail(c, MOSAIC_ERR_INVALID, "placement: cell " + std::to_string(i) + " holds material index " + std::to_string(placed[i]) +
                                             " outside [-1, " + std::to_string(c->n) + ")");
  MOSAIC_TRY(ensure_placed(c));
  MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_placed, placed, sizeof(int32_t) * cells_of(c), cudaMemcpyHostToDevice, c->stream));
  MOSAIC_TRY(sync(c));
  c->have_placed = true;
  return MOSAIC_OK;
}

int mosaic_anneal(mosaic_ctx* c, uint64_t iterations, float t0, float decay, uint64_t seed, int32_t* placed_out,
                  mosaic_opt_result* res) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_anneal(c, iterations, t0, decay, seed));
  MOSAIC_TRY(d2h(c, placed_out, c->d_placed, sizeof(int32_t) * cells_of(c)));
  MOSAIC_TRY(d2h(c, res, c->d_opt_res, sizeof(mosaic_opt_result)));
  return sync(c);
}

int mosaic_optimize_greedy(mosaic_ctx* c, uint64_t max_iterations, uint64_t seed, int32_t* placed_out, mosaic_opt_result* res) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  c->anneal_greedy = true;
  const int rc = run_anneal(c, max_iterations, 1.0f, 1.0f, seed);  // (temperature and decay are not used)
  c->anneal_greedy = false;
  MOSAIC_TRY(rc);
  MOSAIC_TRY(d2h(c, placed_out, c->d_placed, sizeof(int32_t) * cells_of(c)));
  MOSAIC_TRY(d2h(c, res, c->d_opt_res, sizeof(mosaic_opt_result)));
  return sync(c);
}

int mosaic_render(mosaic_ctx* c, float color_strength, uint8_t* out_rgb) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  begin_call(c);
  MOSAIC_TRY(run_render(c, color_strength, out_rgb));
  return sync(c);
}

int mosaic_generate(mosaic_ctx* c, const mosaic_params* p, int32_t* placed_out, uint8_t* out_rgb, mosaic_report* report) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (!p) return fail(c, MOSAIC_ERR_INVALID, "params is NULL");
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  MOSAIC_TRY(require(c, c->have_materials, "materials"));
  MOSAIC_TRY(require(c, c->have_target, "target"));
  const bool anneal = p->enable_optimization && p->adjacency_weight > 0.0f;  // main.rs:608
  begin_call(c);
  trace_point(c, "begin");
  const auto host_t0 = std::chrono::steady_clock::now();
  bool peer = false, under_chain = false, sim_aside = false;
  if (c->world > 1) {  // sharded: buffers every rank exchanges through are allocated (whole-problem sizes) and mapped
    MOSAIC_TRY(ensure_mat_lab_buffers(c, c->n));
    MOSAIC_TRY(ensure_cell_buffers(c));
    MOSAIC_TRY(ensure_cand_buffers(c));
    MOSAIC_TRY(ensure_placed(c));
    MOSAIC_TRY(dist_peer_setup(c));
    peer = c->peer_ok;
  }
  c->peer_render_ready = false;
  if (peer) {
    MOSAIC_TRY(run_generate_peer(c, p, anneal, out_rgb));
  } else {
  // main.rs:105-164 (MosaicGenerator::new) ...
  MOSAIC_TRY(run_material_lab(c));
  emit_progress(c, c->stream, 0.40f, MOSAIC_STAGE_MATERIAL_LAB, "Loaded %d tiles", c->n);  // app_full.rs:1716
  // ... main.rs:477-685 (generate_mosaic)
  static const int sched = getenv("MOSAIC_SCHED") ? atoi(getenv("MOSAIC_SCHED")) : 0;  // experiment knob (bit 0: resident inputs are not banded, bit 1: no render under the chain)
  const bool banded = c->world == 1 && place_chain_eligible(c, p->max_usage) && (c->tgt_upload_pending ? c->tgt_bands > 1 : !(sched & 1));
  if (p->build_similarity && !banded) {
    emit_progress(c, c->stream, 0.45f, MOSAIC_STAGE_SIMILARITY, "Building similarity database...");  // app_full.rs:1738
    MOSAIC_TRY(run_similarity(c, /*gather=*/false));
  }
  emit_progress(c, c->stream, 0.50f, MOSAIC_STAGE_PLACE, "Processing grid cells...");  // app_full.rs:1794
  if (banded) {
    // The serial placement chain (one CTA) is the longest kernel of the step: everything else runs around it.  Rows
    // are matched band by band — bands of an upload in flight, or plain bands of rows with resident inputs — the chain
    // starts behind the first band and consumes the records as they appear; finished rows are rendered (and sent to the
    // host, if a destination is given) under the chain as well.
    under_chain = place_chain_has_progress(c) && !(sched & 2);
    const bool uploading = c->tgt_upload_pending;
    MOSAIC_CUDA(c, cudaEventRecord(c->ev_mat_lab, c->stream));
    MOSAIC_TRY(run_match_place_banded(c, kMaxK, p->max_usage, p->adjacency_weight, under_chain));
    if (p->build_similarity) {
      // Nothing on the placement path reads the matrix (K2b / K5 recompute their entries from the Lab means): it is
      // built on the render stream, which is idle until the chain has finished the first band of rows — under the
      // target's upload when there is one, else behind the first band's records (beside them it would only delay
      // the start of the chain).
      MOSAIC_CUDA(c, cudaStreamWaitEvent(c->render_stream, uploading ? c->ev_mat_lab : c->ev_chain_go, 0));
      emit_progress(c, c->render_stream, 0.45f, MOSAIC_STAGE_SIMILARITY, "Building similarity database...");  // app_full.rs:1738
      cudaStream_t caller = c->stream;
      c->stream = c->render_stream;
      int rc = run_similarity(c, /*gather=*/false);
      c->stream = caller;
      MOSAIC_TRY(rc);
      MOSAIC_CUDA(c, cudaEventRecord(c->ev_sim_done, c->render_stream));
      sim_aside = true;
    }
    if (under_chain) MOSAIC_TRY(run_render_under_chain(c, p->color_strength, out_rgb));
  } else {
    if (!(c->tgt_upload_pending && c->world == 1)) MOSAIC_TRY(run_cell_lab(c));  // (else: band by band inside run_candidates)
    MOSAIC_TRY(run_candidates(c, kMaxK));
    MOSAIC_TRY(run_place(c, p->max_usage, p->adjacency_weight));
  }
  if (under_chain) {
    // The greedy image is on its way.  Annealing permutes materials between cells: the cells it changes (at most two
    // per accepted swap) are composed again from the thumbnail cache, straight into the device image and — when the
    // caller's buffer is page-locked, hence visible to the device — into the caller's image, after the bulk copies.
    trace_point(c, "chain_done");
    emit_progress(c, c->stream, 0.90f, MOSAIC_STAGE_PLACE, "Grid processing completed");  // app_full.rs:2098
    if (anneal) {
      uint8_t* host_dev = nullptr;
      cudaPointerAttributes attr;
      if (out_rgb && cudaPointerGetAttributes(&attr, out_rgb) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        host_dev = reinterpret_cast<uint8_t*>(attr.devicePointer);
      (void)cudaGetLastError();
      const bool patch = (out_rgb == nullptr || host_dev != nullptr) && render_patch_possible(c);
      if (patch) {
        // The last rows of the greedy image are still being drawn from the placement grid: annealing works on a copy
        // (the context's grid pointer is redirected for its kernels and for the patch pass), so it starts right behind
        // the chain; the annealed grid is copied back once the early render has finished.
        struct GridRedirect {
          mosaic_ctx* c;
          int32_t* home;
          ~GridRedirect() { c->d_placed = home; }
        } redirect{c, c->d_placed};
        MOSAIC_TRY(ensure(c, &c->d_placed_greedy, &c->placed_greedy_cap, cells_of(c)));
        MOSAIC_CUDA(c, cudaMemcpyAsync(c->d_placed_greedy, c->d_placed, sizeof(int32_t) * cells_of(c), cudaMemcpyDeviceToDevice, c->stream));
        c->d_placed = c->d_placed_greedy;
        trace_point(c, "anneal_starts");
        emit_progress(c, c->stream, 0.92f, MOSAIC_STAGE_ANNEAL, "Starting optimization phase...");  // app_full.rs:2102
        MOSAIC_TRY(run_anneal(c, p->optimization_iterations, 100.0f, 0.99995f, p->seed));  // optimizer.rs:18-27
        emit_progress(c, c->stream, 0.98f, MOSAIC_STAGE_ANNEAL, "Optimization phase completed");  // app_full.rs:2193
        trace_point(c, "annealed");
        MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_render_done, 0));
        MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));  // nothing may overwrite a patched tile afterwards
        {
          StageTimer t(c, MOSAIC_STAGE_RENDER);
          MOSAIC_TRY(launch_render_patch(c, redirect.home, host_dev));  // cells whose material differs from the greedy grid's
        }
        MOSAIC_CUDA(c, cudaMemcpyAsync(redirect.home, c->d_placed_greedy, sizeof(int32_t) * cells_of(c), cudaMemcpyDeviceToDevice, c->stream));
      } else {
        MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_render_done, 0));  // the early render reads the greedy grid
        trace_point(c, "anneal_starts");
        emit_progress(c, c->stream, 0.92f, MOSAIC_STAGE_ANNEAL, "Starting optimization phase...");  // app_full.rs:2102
        MOSAIC_TRY(run_anneal(c, p->optimization_iterations, 100.0f, 0.99995f, p->seed));  // optimizer.rs:18-27
        emit_progress(c, c->stream, 0.98f, MOSAIC_STAGE_ANNEAL, "Optimization phase completed");  // app_full.rs:2193
        trace_point(c, "annealed");
        MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
        MOSAIC_TRY(run_render(c, p->color_strength, out_rgb));  // pageable destination / per-cell render path: once more, whole
      }
    } else {
      MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_render_done, 0));
      MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
    }
    MOSAIC_TRY(d2h(c, placed_out, c->d_placed, sizeof(int32_t) * cells_of(c)));
    trace_point(c, "rendered");
  } else {
  emit_progress(c, c->stream, 0.90f, MOSAIC_STAGE_PLACE, "Grid processing completed");  // app_full.rs:2098
  if (anneal) {
    emit_progress(c, c->stream, 0.92f, MOSAIC_STAGE_ANNEAL, "Starting optimization phase...");  // app_full.rs:2102
    MOSAIC_TRY(run_anneal(c, p->optimization_iterations, 100.0f, 0.99995f, p->seed));  // optimizer.rs:18-27
    emit_progress(c, c->stream, 0.98f, MOSAIC_STAGE_ANNEAL, "Optimization phase completed");  // app_full.rs:2193
  }
  }
  }
  if (!under_chain) {
  trace_point(c, "placed");
  MOSAIC_TRY(d2h(c, placed_out, c->d_placed, sizeof(int32_t) * cells_of(c)));
  if (!(peer && c->peer_rendered)) MOSAIC_TRY(run_render(c, p->color_strength, out_rgb));
  trace_point(c, "rendered");
  }
  if (peer) {
    CollTimer t(c, c->stream);
    MOSAIC_TRY(dist_peer_barrier(c, c->stream));
    c->peer_render_ready = false;  // (a later stand-alone mosaic_render has no hand-off protecting the peers' slices)
  }
  if (sim_aside) MOSAIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_sim_done, 0));
  if (report) {
    std::memset(report, 0, sizeof *report);
    MOSAIC_TRY(d2h(c, &report->n_fallbacks, c->d_place_info, sizeof(uint32_t)));
    if (anneal) MOSAIC_TRY(d2h(c, &report->opt, c->d_opt_res, sizeof(mosaic_opt_result)));
    report->annealed = anneal ? 1 : 0;
  }
  emit_progress(c, c->stream, 0.99f, MOSAIC_STAGE_RENDER, "Mosaic image ready");  // (app_full.rs:2201: saving is the host's)
  trace_point(c, "end");
  const auto host_t1 = std::chrono::steady_clock::now();
  MOSAIC_TRY(sync(c));
  if (!c->trace.empty()) {
    static thread_local std::chrono::steady_clock::time_point last_exit;
    const auto host_t2 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::duration d) { return std::chrono::duration<double, std::milli>(d).count(); };
    char hb[160];
    snprintf(hb, sizeof hb, " | host: since_last_exit=%.3f enqueue=%.3f sync=%.3f", ms(host_t0 - last_exit), ms(host_t1 - host_t0), ms(host_t2 - host_t1));
    last_exit = host_t2;
    std::string line = "[mosaic trace rank " + std::to_string(c->rank) + "]" + hb;
    for (size_t i = 0; i < c->trace.size(); ++i) {
      float ms = 0.0f;
      cudaEventSynchronize(c->trace[i].second);
      cudaEventElapsedTime(&ms, c->trace[0].second, c->trace[i].second);
      char buf[96];
      snprintf(buf, sizeof buf, " %s=%.3f", c->trace[i].first, ms);
      line += buf;
    }
    fprintf(stderr, "%s\n", line.c_str());
    for (auto& t : c->trace) cudaEventDestroy(t.second);
    c->trace.clear();
  }
  return dist_peer_check(c);
}

int mosaic_selftest(mosaic_ctx* c, uint64_t mismatches[2]) {
  if (!c || !mismatches) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  unsigned long long m[2] = {0, 0};
  MOSAIC_TRY(launch_selftest(c, m));
  mismatches[0] = m[0];
  mismatches[1] = m[1];
  return MOSAIC_OK;
}

int mosaic_debug_chain_profile(mosaic_ctx* c, int enable, uint64_t counters[16]) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  if (counters) {
    unsigned long long v[16] = {};
    MOSAIC_TRY(read_chain_profile(c, v));
    for (int i = 0; i < 16; ++i) counters[i] = v[i];
  }
  c->profile_chain = enable != 0;
  return MOSAIC_OK;
}

int mosaic_host_alloc(size_t bytes, void** out) {
  if (!out) return MOSAIC_ERR_INVALID;
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped);
  return e == cudaSuccess ? MOSAIC_OK : cuda_fail(nullptr, e, "cudaHostAlloc");
}
int mosaic_host_release(void* p) {
  if (!p) return MOSAIC_OK;
  const cudaError_t e = cudaFreeHost(p);
  return e == cudaSuccess ? MOSAIC_OK : cuda_fail(nullptr, e, "cudaFreeHost");
}
int mosaic_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return MOSAIC_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  return e == cudaSuccess ? MOSAIC_OK : cuda_fail(nullptr, e, "cudaHostRegister");
}
int mosaic_host_unregister(void* p) {
  if (!p) return MOSAIC_ERR_INVALID;
  const cudaError_t e = cudaHostUnregister(p);
  return e == cudaSuccess ? MOSAIC_OK : cuda_fail(nullptr, e, "cudaHostUnregister");
}

int mosaic_set_progress_callback(mosaic_ctx* c, mosaic_progress_fn fn, void* user) {
  if (!c) return MOSAIC_ERR_INVALID;
  MOSAIC_CUDA(c, cudaSetDevice(c->device));
  MOSAIC_TRY(sync(c));
  c->progress_fn = fn;
  c->progress_user = user;
  return MOSAIC_OK;
}

int mosaic_debug_set_chain(mosaic_ctx* c, int version) {
  if (!c || version < 0 || version > 3) return MOSAIC_ERR_INVALID;
  c->chain_version = version;
  c->have_placed = false;
  return MOSAIC_OK;
}

int mosaic_dims(const mosaic_ctx* c, int* n_materials, int* grid_w, int* grid_h, int* tile_w, int* tile_h) {
  if (!c) return MOSAIC_ERR_INVALID;
  if (n_materials) *n_materials = c->n;
  if (grid_w) *grid_w = c->gw;
  if (grid_h) *grid_h = c->gh;
  if (tile_w) *tile_w = c->tile_w;
  if (tile_h) *tile_h = c->tile_h;
  return MOSAIC_OK;
}

const void* mosaic_device_output(const mosaic_ctx* c) { return c && c->have_out ? c->d_out : nullptr; }
const void* mosaic_device_placement(const mosaic_ctx* c) { return c && c->have_placed ? c->d_placed : nullptr; }
const void* mosaic_device_similarity(const mosaic_ctx* c) {
  return c && c->have_sim && (c->world == 1 || c->sim_gathered) ? c->d_sim : nullptr;
}

int mosaic_shard(const mosaic_ctx* c, int* row_begin, int* row_end, int* mat_begin, int* mat_end, uint64_t* tri_begin,
                 uint64_t* tri_end) {
  if (!c) return MOSAIC_ERR_INVALID;
  int rb, re, mb, me;
  uint64_t tb, te;
  shard_rows(c, c->rank, &rb, &re);
  shard_mats(c, c->rank, &mb, &me);
  shard_tri(c, c->rank, &tb, &te);
  if (row_begin) *row_begin = rb;
  if (row_end) *row_end = re;
  if (mat_begin) *mat_begin = mb;
  if (mat_end) *mat_end = me;
  if (tri_begin) *tri_begin = tb;
  if (tri_end) *tri_end = te;
  return MOSAIC_OK;
}

int mosaic_shard_plan(int world, int rank, int grid_h, int n_materials, int* row_begin, int* row_end, int* mat_begin,
                      int* mat_end, uint64_t* tri_begin, uint64_t* tri_end) {
  if (world < 1 || rank < 0 || rank >= world || grid_h < 0 || n_materials < 0) return MOSAIC_ERR_INVALID;
  mosaic_ctx tmp;  // plain host object: only the shape fields are read
  tmp.world = world;
  tmp.rank = rank;
  tmp.gh = grid_h;
  tmp.n = n_materials;
  return mosaic_shard(&tmp, row_begin, row_end, mat_begin, mat_end, tri_begin, tri_end);
}

}  // extern "C"
