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

template <bool USAGE_SMEM>
__global__ void __launch_bounds__(32, 1) greedy_generic_kernel(const uint32_t* __restrict__ cand_idx,
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


// ------------------------------------------------------------------------------------------------
// Fast path: producer warps + one consumer warp.
//
// Everything the serial chain needs that does NOT depend on earlier placements (candidate ids, d2,
// candidate Lab of the first 32 candidates of each cell, d2 of candidate 32) is staged by 7 producer
// warps into a shared-memory ring, 64 cells deep, so the consumer never waits on HBM/L2 latency.
// Usage counters (u8/u16), the previous row's material ids and their Lab values live in shared
// memory too.  The consumer's per-cell critical path is: Lab of the Left winner (3 shuffles) ->
// sim -> 1/(sim+1) -> (t_up + t_left) * w -> d2 + pen -> warp min -> ballot.  The term with the Up
// neighbour and the usage lookups of the NEXT cell are issued in the shadow of that chain.
// Cells whose first 32 candidates do not decide the winner (exhausted or early-exit bound not met)
// take the generic multi-chunk path against the global candidate lists.
constexpr int RING_DEPTH = 64;
constexpr int RING_BATCH = 4;        // cells a producer warp keeps in flight
constexpr int RING_PRODUCERS = 7;    // producer warps (CTA = 8 warps)

struct RingSlot {
  uint32_t id[32];
  float d2[32];
  float l[32], a[32], b[32];
};

template <typename UsageT>
__global__ void __launch_bounds__((RING_PRODUCERS + 1) * 32, 1)
    greedy_ring_kernel(const uint32_t* __restrict__ cand_idx, const float* __restrict__ cand_d2, int k,
                       const float4* __restrict__ lab4, int n, int gw, int gh, unsigned max_usage, float weight,
                       int32_t* __restrict__ placed, uint32_t* __restrict__ info, uint32_t fallback_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RingSlot* ring = reinterpret_cast<RingSlot*>(smem_raw);
  float* ring_d2_32 = reinterpret_cast<float*>(smem_raw + sizeof(RingSlot) * RING_DEPTH);
  volatile uint32_t* ready = reinterpret_cast<volatile uint32_t*>(ring_d2_32 + RING_DEPTH);
  volatile uint32_t* consumed = ready + RING_DEPTH;  // [0]: number of cells whose slot has been read
  float4* rowlab = reinterpret_cast<float4*>(const_cast<uint32_t*>(consumed) + 4);
  int32_t* rowbuf = reinterpret_cast<int32_t*>(rowlab + gw);
  UsageT* usage = reinterpret_cast<UsageT*>(rowbuf + ((gw + 3) & ~3));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cells = gw * gh;
  for (int i = tid; i < n; i += blockDim.x) usage[i] = 0;
  for (int i = tid; i < gw; i += blockDim.x) rowbuf[i] = -1;
  for (int i = tid; i < RING_DEPTH; i += blockDim.x) ready[i] = 0u;
  if (tid == 0) consumed[0] = 0u;
  __syncthreads();

  if (warp > 0) {
    // ---------------- producers ----------------
    const int pw = warp - 1;
    for (int c0 = pw * RING_BATCH; c0 < cells; c0 += RING_PRODUCERS * RING_BATCH) {
      const int last = min(c0 + RING_BATCH, cells) - 1;
      while ((int)consumed[0] + RING_DEPTH <= last) __nanosleep(40);  // wait for free slots
      uint32_t id[RING_BATCH];
      float d2[RING_BATCH], d32[RING_BATCH];
      float4 lab[RING_BATCH];
#pragma unroll
      for (int b = 0; b < RING_BATCH; ++b) {
        const int c = min(c0 + b, cells - 1);
        id[b] = 0u;
        d2[b] = 0.f;
        if (lane < k) {
          id[b] = __ldg(cand_idx + (size_t)c * k + lane);
          d2[b] = __ldg(cand_d2 + (size_t)c * k + lane);
        }
        d32[b] = k > 32 ? __ldg(cand_d2 + (size_t)c * k + 32) : __int_as_float(0x7f800000);
      }
#pragma unroll
      for (int b = 0; b < RING_BATCH; ++b) lab[b] = __ldg(&lab4[id[b]]);
#pragma unroll
      for (int b = 0; b < RING_BATCH; ++b) {
        const int c = c0 + b;
        if (c < cells) {
          RingSlot& s = ring[c % RING_DEPTH];
          s.id[lane] = id[b];
          s.d2[lane] = d2[b];
          s.l[lane] = lab[b].x;
          s.a[lane] = lab[b].y;
          s.b[lane] = lab[b].z;
          if (lane == 0) ring_d2_32[c % RING_DEPTH] = d32[b];
        }
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < RING_BATCH; ++b)
          if (c0 + b < cells) ready[(c0 + b) % RING_DEPTH] = (uint32_t)(c0 + b) + 1u;
      }
    }
    return;
  }

  // ---------------- consumer (warp 0) ----------------
  // Per-lane registers of one cell: candidate `lane` of its first chunk + what is known ahead of time.
  struct CellRegs {
    uint32_t id;
    float d2, l, a, b, d2_32;
    float t_up;    // 1/(sim(candidate, Up)+1), 0 when there is no Up neighbour
    int32_t up;    // material above, -1 if none
    bool flag;     // t_up left the fast-math domain -> resolve this cell on the exact generic path
  };
  const bool use_penalty = weight > 0.0f;  // main.rs:355-362
  uint32_t n_fallbacks = 0;
  int32_t left = -1;
  float left_l = 0.f, left_a = 0.f, left_b = 0.f;
  const volatile RingSlot* vring = ring;  // volatile: ptxas must not hoist slot reads above the flag poll
  const volatile float* vd2_32 = ring_d2_32;

  auto wait_slot = [&](int c, uint32_t flag) {
    const int s = c % RING_DEPTH;
    while (flag != (uint32_t)c + 1u) flag = ready[s];
    asm volatile("" ::: "memory");
  };
  auto load_slot = [&](int c, CellRegs& r) {
    const int s = c % RING_DEPTH;
    r.id = vring[s].id[lane];
    r.d2 = vring[s].d2[lane];
    r.l = vring[s].l[lane];
    r.a = vring[s].a[lane];
    r.b = vring[s].b[lane];
    r.d2_32 = vd2_32[s];
  };

  // exact multi-chunk resolution of one cell (first chunk from registers, the rest from the global lists),
  // then the reset fallback: the reference's loop, bit for bit (main.rs:335-475)
  auto resolve_generic = [&](int cell, const CellRegs& A, int32_t& chosen, float& ch_l, float& ch_a, float& ch_b) {
    const int32_t up = A.up;
    float4 up_lab = make_float4(0.f, 0.f, 0.f, 0.f);
    const int x = cell % gw;
    if (up >= 0) up_lab = rowlab[x];
    unsigned best_bits = 0xFFFFFFFFu;
    int32_t best_id = -1;
    float4 best_lab = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t row = (size_t)cell * k;
    for (int base = 0; base < k; base += 32) {
      if (base > 0 && best_id >= 0 && __float_as_uint(cand_d2[row + base]) >= best_bits) break;  // exact early exit
      const int j = base + lane;
      uint32_t gid = A.id;
      float gd2 = A.d2;
      float4 gl = make_float4(A.l, A.a, A.b, 0.f);
      bool gok = j < k;
      if (base > 0 && gok) {
        gid = cand_idx[row + j];
        gd2 = cand_d2[row + j];
        gl = __ldg(&lab4[gid]);
      }
      if (gok) gok = (unsigned)usage[gid] < max_usage && (int32_t)gid != left && (int32_t)gid != up;
      float gpen = 0.0f;
      if (use_penalty) {  // adjacency.rs:71-87: Up, (Down), Left, (Right)
        if (up >= 0) gpen = __fadd_rn(gpen, adjacency_term(lab_distance_rn(gl.x, gl.y, gl.z, up_lab.x, up_lab.y, up_lab.z)));
        if (left >= 0) gpen = __fadd_rn(gpen, adjacency_term(lab_distance_rn(gl.x, gl.y, gl.z, left_l, left_a, left_b)));
        gpen = __fmul_rn(gpen, weight);
      }
      const unsigned gtb = gok ? __float_as_uint(__fadd_rn(gd2, gpen)) : 0xFFFFFFFFu;
      const unsigned gmn = __reduce_min_sync(0xffffffffu, gtb);
      if (gmn < best_bits) {  // strict <: earlier chunks / earlier lanes win ties (main.rs:401-407)
        const int who = __ffs(__ballot_sync(0xffffffffu, gtb == gmn)) - 1;
        best_bits = gmn;
        best_id = (int32_t)__shfl_sync(0xffffffffu, gid, who);
        best_lab.x = __shfl_sync(0xffffffffu, gl.x, who);
        best_lab.y = __shfl_sync(0xffffffffu, gl.y, who);
        best_lab.z = __shfl_sync(0xffffffffu, gl.z, who);
      }
    }
    if (best_id < 0) {
      // fallback_tile_selection (main.rs:420-475)
      if (lane == 0 && n_fallbacks < fallback_cap) info[1 + n_fallbacks] = (uint32_t)cell;
      ++n_fallbacks;
      for (int i = lane; i < n; i += 32) usage[i] = 0;  // lib.rs:86-88
      __syncwarp();
      for (int base = 0; base < k && best_id < 0; base += 32) {
        const int j = base + lane;
        uint32_t fid = 0;
        bool fok = false;
        if (j < k) {
          fid = cand_idx[row + j];
          fok = (int32_t)fid != left && (int32_t)fid != up;  // main.rs:448
        }
        const unsigned m = __ballot_sync(0xffffffffu, fok);
        if (m) best_id = (int32_t)__shfl_sync(0xffffffffu, fid, __ffs(m) - 1);
      }
      if (best_id < 0) best_id = (int32_t)cand_idx[row];  // main.rs:456-474 nearest_one
      best_lab = __ldg(&lab4[best_id]);
    }
    chosen = best_id;
    ch_l = best_lab.x;
    ch_a = best_lab.y;
    ch_b = best_lab.z;
  };

  int x = 0, y = 0;
  // one cell: A = registers of `cell` (its t_up/up already known), B = registers of cell + 1 (slot loaded,
  // t_up/up filled in here, in the shadow of A's Left chain).  At the end A is refilled with cell + 2.
  auto step = [&](int cell, CellRegs& A, CellRegs& B) {
    const uint32_t flag2 = (cell + 2 < cells) ? ready[(cell + 2) % RING_DEPTH] : 0u;
    // ---- next cell's Up term (independent of this cell's outcome as long as gw >= 2)
    int xn = x + 1, yn = y;
    if (xn == gw) {
      xn = 0;
      yn = y + 1;
    }
    const int32_t up_n = (yn > 0) ? rowbuf[xn] : -1;
    const float4 upl_n = rowlab[xn];
    const unsigned u = (unsigned)usage[A.id];
    bool flag_n = false;
    float t_up_n = 0.0f;
    if (use_penalty) {
      t_up_n = adjacency_term_fast(B.l, B.a, B.b, upl_n.x, upl_n.y, upl_n.z, flag_n);
      t_up_n = up_n >= 0 ? t_up_n : 0.0f;
      flag_n = flag_n && up_n >= 0;
    }
    // ---- this cell: the chain through the Left neighbour
    if (x == 0) left = -1;
    bool flag = A.flag;
    float pen = 0.0f;
    if (use_penalty) {
      float t_left = adjacency_term_fast(A.l, A.a, A.b, left_l, left_a, left_b, flag);
      t_left = left >= 0 ? t_left : 0.0f;                 // absent terms are +0.0: x + 0.0 == x exactly
      pen = __fmul_rn(__fadd_rn(A.t_up, t_left), weight);  // adjacency.rs:71-87 order Up, Left; then * weight
    }
    const float total = __fadd_rn(A.d2, pen);              // main.rs:398
    const bool ok = lane < k && u < max_usage && (int32_t)A.id != A.up && (int32_t)A.id != left;  // main.rs:375, 380
    unsigned tb = ok ? __float_as_uint(total) : 0xFFFFFFFFu;
    tb = flag ? 0u : tb;  // out-of-domain arithmetic forces mn == 0 -> generic path
    const unsigned mn = __reduce_min_sync(0xffffffffu, tb);

    int32_t chosen;
    float ch_l, ch_a, ch_b;
    if (mn - 1u < 0xFFFFFFFEu && __float_as_uint(A.d2_32) >= mn) {
      // decided by the first chunk (exact early exit: later candidates have d2 >= d2_32 >= best total);
      // strict-< winner, first in candidate order on ties (main.rs:401-407)
      const int who = __ffs(__ballot_sync(0xffffffffu, tb == mn)) - 1;
      chosen = (int32_t)__shfl_sync(0xffffffffu, A.id, who);
      ch_l = __shfl_sync(0xffffffffu, A.l, who);
      ch_a = __shfl_sync(0xffffffffu, A.a, who);
      ch_b = __shfl_sync(0xffffffffu, A.b, who);
    } else {
      resolve_generic(cell, A, chosen, ch_l, ch_a, ch_b);
    }
    if (lane == 0) {
      usage[chosen] = (UsageT)(usage[chosen] + 1);  // main.rs:410-411 / 449 / 470 (never exceeds max_usage)
      placed[cell] = chosen;                        // main.rs:412
      rowbuf[x] = chosen;
      rowlab[x] = make_float4(ch_l, ch_a, ch_b, 0.f);
    }
    __syncwarp();
    left = chosen;
    left_l = ch_l;
    left_a = ch_a;
    left_b = ch_b;
    x = xn;
    y = yn;
    B.t_up = t_up_n;
    B.up = up_n;
    B.flag = flag_n;
    if (cell + 2 < cells) {
      wait_slot(cell + 2, flag2);
      load_slot(cell + 2, A);
      if (lane == 0) consumed[0] = (uint32_t)cell + 3u;
    }
  };

  CellRegs R0, R1;
  R0.t_up = 0.f; R0.up = -1; R0.flag = false;
  R1 = R0;
  R0.id = R1.id = 0u;
  R0.d2 = R0.l = R0.a = R0.b = R0.d2_32 = 0.f;
  R1.d2 = R1.l = R1.a = R1.b = R1.d2_32 = 0.f;
  if (cells > 0) {
    wait_slot(0, ready[0]);
    load_slot(0, R0);
    if (cells > 1) {
      wait_slot(1, ready[1 % RING_DEPTH]);
      load_slot(1, R1);
    }
    __syncwarp();
    if (lane == 0) consumed[0] = (uint32_t)min(2, cells);
  }
  for (int cell = 0; cell < cells; cell += 2) {
    step(cell, R0, R1);
    if (cell + 1 < cells) step(cell + 1, R1, R0);
  }
  if (lane == 0) info[0] = n_fallbacks;
}

}  // namespace

template <typename UsageT>
static int launch_ring(mosaic_ctx* ctx, unsigned cap, float weight, size_t smem) {
  MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_ring_kernel<UsageT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  greedy_ring_kernel<UsageT><<<1, (RING_PRODUCERS + 1) * 32, smem, ctx->stream>>>(
      ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->d_mat_lab4, ctx->n, ctx->gw, ctx->gh, cap, weight, ctx->d_placed,
      ctx->d_place_info, mosaic_ctx::kFallbackCap);
  return MOSAIC_OK;
}

int launch_place(mosaic_ctx* ctx, int max_usage, float weight) {
  const int n = ctx->n, gw = ctx->gw, gh = ctx->gh;
  const unsigned cap = max_usage < 0 ? 0u : (unsigned)max_usage;
  const size_t smem_limit = 220 * 1024;
  // fast path: ring + row state + usage counters in shared memory
  const size_t ring_fixed = sizeof(RingSlot) * RING_DEPTH + sizeof(float) * RING_DEPTH + sizeof(uint32_t) * RING_DEPTH + 16 +
                            sizeof(float4) * (size_t)gw + sizeof(int32_t) * (size_t)((gw + 3) & ~3);
  const size_t usage_bytes = cap <= 255u ? (size_t)n : (cap <= 65535u ? (size_t)n * 2 : (size_t)-1);
  StageTimer t(ctx, MOSAIC_STAGE_PLACE);
  if (gw >= 2 && usage_bytes != (size_t)-1 && ring_fixed + usage_bytes + 16 <= smem_limit) {
    if (cap <= 255u)
      MOSAIC_TRY(launch_ring<uint8_t>(ctx, cap, weight, ring_fixed + usage_bytes + 16));
    else
      MOSAIC_TRY(launch_ring<uint16_t>(ctx, cap, weight, ring_fixed + usage_bytes + 16));
  } else {
    // generic path: single warp, usage counters in shared memory when they fit, else in HBM
    const size_t row_bytes = ((size_t)gw * 4 + 15) & ~(size_t)15;
    const size_t smem_with_usage = row_bytes + (size_t)n * 4;
    if (row_bytes > smem_limit) return fail(ctx, MOSAIC_ERR_INVALID, "grid_w too large for the placement kernel");
    if (smem_with_usage <= smem_limit) {
      MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem_with_usage));
      greedy_generic_kernel<true><<<1, 32, smem_with_usage, ctx->stream>>>(
          ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->d_mat_lab4, n, gw, gh, cap, weight, ctx->d_placed, nullptr,
          ctx->d_place_info, mosaic_ctx::kFallbackCap);
    } else {
      MOSAIC_TRY(ensure(ctx, &ctx->d_usage, &ctx->usage_cap, (size_t)n));
      MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)row_bytes));
      greedy_generic_kernel<false><<<1, 32, row_bytes, ctx->stream>>>(
          ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->d_mat_lab4, n, gw, gh, cap, weight, ctx->d_placed, ctx->d_usage,
          ctx->d_place_info, mosaic_ctx::kFallbackCap);
    }
  }
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace mosaic
