// greedy.cu — K2b: raster-order greedy placement replayed on one GPU.
//
// Replaces the body of the reference's cell loop (src/main.rs:529-580):
//   find_and_use_best_tile_with_position (main.rs:335-418), fallback_tile_selection (main.rs:420-475),
//   can_place_at_position (main.rs:309-333), UsageTracker (lib.rs:57-89),
//   AdjacencyPenaltyCalculator::calculate_penalty (adjacency.rs:63-88) with
//   SimilarityDatabase::get_similarity (similarity.rs:85-103).
//
// The chain is inherently serial (global usage counts + Left/Up neighbours), so it runs as ONE warp:
// lanes evaluate 32 candidates of the current cell at a time (usage test, same-image-adjacent test,
// penalty = weight * (1/(sim(c,Up)+1) + 1/(sim(c,Left)+1)) in the reference's Up,Left order,
// total = d2 + penalty), a warp min-reduction on the (non-negative) float bits picks the strict-<
// winner with ties going to the earliest candidate, and — because candidates arrive in ascending d2
// and penalty >= 0 — the scan stops as soon as the next candidate's d2 >= best total (exact early
// exit).  Similarity entries are recomputed from the Lab means with the same IEEE
// operations as the packed matrix (similarity.rs:141-146), so no 5 GB matrix is read.
// State-independent data (candidate ids, d2, candidate Lab) of the NEXT cells is prefetched into
// registers while the current cell is resolved.
#include "common.cuh"

namespace mosaic {

namespace {

struct CellPrefetch {
  uint32_t id;
  float d2;
  float4 lab;
};

template <bool USAGE_SMEM>
__global__ void __launch_bounds__(32, 1) greedy_place_kernel(const uint32_t* __restrict__ cand_idx,
                                                             const float* __restrict__ cand_d2, int k,
                                                             const float4* __restrict__ lab4, int n, int gw, int gh,
                                                             unsigned max_usage, float weight,
                                                             int32_t* __restrict__ placed, uint32_t* usage_gmem,
                                                             uint32_t* __restrict__ info, uint32_t fallback_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // smem: rowbuf[gw] (material id of the row above / current row), then usage[n] when it fits
  int32_t* rowbuf = reinterpret_cast<int32_t*>(smem_raw);
  uint32_t* usage = USAGE_SMEM ? reinterpret_cast<uint32_t*>(smem_raw + (((size_t)gw * 4 + 15) & ~(size_t)15))
                               : usage_gmem;
  const int lane = threadIdx.x;
  for (int i = lane; i < n; i += 32) usage[i] = 0u;
  for (int i = lane; i < gw; i += 32) rowbuf[i] = -1;
  __syncwarp();

  const int cells = gw * gh;
  const bool use_penalty = weight > 0.0f;  // main.rs:355-362
  uint32_t n_fallbacks = 0;

  // prefetch of chunk 0 of the next cell
  uint32_t nx_id = 0;
  float nx_d2 = 0.f, nx_d2_32 = 0.f;  // *_d2_32: d2 of candidate 32 (first of the second chunk), for the early exit
  float4 nx_lab = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cells > 0 && lane < k) {
    nx_id = cand_idx[lane];
    nx_d2 = cand_d2[lane];
    nx_lab = __ldg(&lab4[nx_id]);
  }
  if (cells > 0 && k > 32) nx_d2_32 = cand_d2[32];
  uint32_t nn_id = 0;  // ids of cell+2 (loaded one cell ahead of their Lab gather)
  float nn_d2 = 0.f, nn_d2_32 = 0.f;
  if (cells > 1 && lane < k) {
    nn_id = cand_idx[(size_t)k + lane];
    nn_d2 = cand_d2[(size_t)k + lane];
  }
  if (cells > 1 && k > 32) nn_d2_32 = cand_d2[(size_t)k + 32];

  int32_t left = -1;
  float4 left_lab = make_float4(0.f, 0.f, 0.f, 0.f);
  int x = 0, y = 0;
  for (int cell = 0; cell < cells; ++cell) {
    // rotate the prefetch registers
    uint32_t cid = nx_id;
    float cd2 = nx_d2;
    float4 clab = nx_lab;
    const float d2_32 = nx_d2_32;
    nx_id = nn_id;
    nx_d2 = nn_d2;
    nx_d2_32 = nn_d2_32;
    if (cell + 1 < cells && lane < k) nx_lab = __ldg(&lab4[nx_id]);
    if (cell + 2 < cells) {
      if (lane < k) {
        nn_id = cand_idx[(size_t)(cell + 2) * k + lane];
        nn_d2 = cand_d2[(size_t)(cell + 2) * k + lane];
      }
      if (k > 32) nn_d2_32 = cand_d2[(size_t)(cell + 2) * k + 32];
    }

    const int32_t up = (y > 0) ? rowbuf[x] : -1;
    if (x == 0) left = -1;
    float4 up_lab = make_float4(0.f, 0.f, 0.f, 0.f);
    if (up >= 0) up_lab = __ldg(&lab4[up]);

    unsigned best_bits = 0xFFFFFFFFu;
    int32_t best_id = -1;
    float4 best_lab = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t row = (size_t)cell * k;

    for (int base = 0; base < k; base += 32) {
      const int j = base + lane;
      if (base > 0) {  // slow path: chunks beyond the first are fetched on demand
        if (j < k) {
          cid = cand_idx[row + j];
          cd2 = cand_d2[row + j];
          clab = __ldg(&lab4[cid]);
        }
      }
      bool ok = j < k;
      if (ok) ok = usage[cid] < max_usage;                        // lib.rs:71-74, main.rs:375
      ok = ok && (int32_t)cid != left && (int32_t)cid != up;      // main.rs:309-333, 380 (Right/Down are empty)
      float pen = 0.0f;
      if (use_penalty) {                                          // adjacency.rs:71-87: Up, (Down), Left, (Right)
        if (up >= 0)
          pen = __fadd_rn(pen, adjacency_term(lab_distance_rn(clab.x, clab.y, clab.z, up_lab.x, up_lab.y, up_lab.z)));
        if (left >= 0)
          pen = __fadd_rn(pen, adjacency_term(lab_distance_rn(clab.x, clab.y, clab.z, left_lab.x, left_lab.y,
                                                              left_lab.z)));
        pen = __fmul_rn(pen, weight);
      }
      const float total = __fadd_rn(cd2, pen);                     // main.rs:398
      const unsigned tb = ok ? __float_as_uint(total) : 0xFFFFFFFFu;
      const unsigned mn = __reduce_min_sync(0xffffffffu, tb);
      if (mn < best_bits) {                                        // strict <, earlier chunks win ties (main.rs:401-407)
        const unsigned who = __ffs(__ballot_sync(0xffffffffu, tb == mn)) - 1;  // first in order wins ties
        best_bits = mn;
        best_id = (int32_t)__shfl_sync(0xffffffffu, cid, who);
        best_lab.x = __shfl_sync(0xffffffffu, clab.x, who);
        best_lab.y = __shfl_sync(0xffffffffu, clab.y, who);
        best_lab.z = __shfl_sync(0xffffffffu, clab.z, who);
      }
      // exact early exit: candidates are ascending in d2 and total >= d2
      if (base + 32 < k) {
        const float next_d2 = base == 0 ? d2_32 : cand_d2[row + base + 32];
        if (best_id >= 0 && __float_as_uint(next_d2) >= best_bits) break;
      }
    }

    int32_t chosen = best_id;
    float4 chosen_lab = best_lab;
    if (chosen < 0) {
      // fallback_tile_selection (main.rs:420-475): reset ALL usage counts, then first placeable
      if (lane == 0 && n_fallbacks < fallback_cap) info[1 + n_fallbacks] = (uint32_t)cell;
      ++n_fallbacks;
      for (int i = lane; i < n; i += 32) usage[i] = 0u;            // lib.rs:86-88
      __syncwarp();
      for (int base = 0; base < k && chosen < 0; base += 32) {
        const int j = base + lane;
        uint32_t fid = 0;
        bool ok = false;
        if (j < k) {
          fid = cand_idx[row + j];
          ok = (int32_t)fid != left && (int32_t)fid != up;         // main.rs:448
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) chosen = (int32_t)__shfl_sync(0xffffffffu, fid, __ffs(m) - 1);
      }
      if (chosen < 0) chosen = (int32_t)cand_idx[row];             // main.rs:456-474 nearest_one
      chosen_lab = __ldg(&lab4[chosen]);
    }
    if (lane == 0) {
      usage[chosen] += 1u;                                         // main.rs:410-411 / 449 / 470
      placed[cell] = chosen;                                       // main.rs:412
      rowbuf[x] = chosen;
    }
    __syncwarp();
    left = chosen;
    left_lab = chosen_lab;
    if (++x == gw) {
      x = 0;
      ++y;
    }
  }
  if (lane == 0) info[0] = n_fallbacks;
}

}  // namespace

int launch_place(mosaic_ctx* ctx, int max_usage, float weight) {
  const int n = ctx->n, gw = ctx->gw, gh = ctx->gh;
  const size_t row_bytes = ((size_t)gw * 4 + 15) & ~(size_t)15;
  const size_t smem_with_usage = row_bytes + (size_t)n * 4;
  const size_t smem_limit = 200 * 1024;
  const unsigned cap = max_usage < 0 ? 0u : (unsigned)max_usage;
  if (row_bytes > smem_limit) return fail(ctx, MOSAIC_ERR_INVALID, "grid_w too large for the placement kernel");
  StageTimer t(ctx, MOSAIC_STAGE_PLACE);
  if (smem_with_usage <= smem_limit) {
    MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_place_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_with_usage));
    greedy_place_kernel<true><<<1, 32, smem_with_usage, ctx->stream>>>(
        ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->d_mat_lab4, n, gw, gh, cap, weight, ctx->d_placed, nullptr,
        ctx->d_place_info, mosaic_ctx::kFallbackCap);
  } else {
    MOSAIC_TRY(ensure(ctx, &ctx->d_usage, &ctx->usage_cap, (size_t)n));
    MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_place_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)row_bytes));
    greedy_place_kernel<false><<<1, 32, row_bytes, ctx->stream>>>(
        ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->d_mat_lab4, n, gw, gh, cap, weight, ctx->d_placed, ctx->d_usage,
        ctx->d_place_info, mosaic_ctx::kFallbackCap);
  }
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace mosaic
