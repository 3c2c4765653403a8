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
// Fast path: one CTA, three warp roles around shared-memory rings.
//
//   producers (6 warps)  stream the state-independent data of each cell — ids, d2 and Lab of its first 32
//                        candidates, d2 of candidate 32 — from HBM/L2 into a 64-slot ring, far ahead, and
//                        derive per candidate j the CONTENDER MASK
//                            cont[j] = { j' > j : d2[j'] < hi[j] },  hi[j] = fl(d2[j] + fl(2 * w))
//                        (all ones if d2 of candidate 32 < hi[j]).  hi[j] bounds the candidate's total from
//                        above: total = fl(d2 + fl(fl(t_up + t_left) * w)) with t_up, t_left in [0, 1] and
//                        IEEE rounding monotone; and every total is bounded from below by its d2.
//   consumer   (1 warp)   walks the cells in raster order.  Its loop-carried chain is integer-only in the
//                        common case: valid = (id != Up) & (id != Left) & (usage < cap); j0 = first valid
//                        lane; if no other valid lane is a contender of j0 then no candidate can be
//                        STRICTLY better than j0 (total_j' >= d2_j' >= hi_j0 >= total_j0), so j0 is exactly
//                        the reference's winner — without evaluating a single penalty (91 % of the cells
//                        of the 24,000-tile config).  Otherwise the exact totals of the chunk are evaluated
//                        (sqrt / reciprocal sequences bit-identical to IEEE), and cells the first chunk
//                        cannot decide (all 32 rejected, or the early-exit bound fails) take the generic
//                        multi-chunk + reset-fallback path.  Every shortcut is an exact bound: placements
//                        are bit-identical to the serial loop of main.rs:335-475.
//   committer  (1 warp)   drains the consumer's decisions, up to 32 cells at a time (one per lane): row
//                        buffers (material id + Lab of the row above), the placement array in HBM, and
//                        the `decided` counter that frees ring slots.
constexpr int CH_DEPTH = 64;
constexpr int CH_GUARD = 3;   // slots of the last decided cells stay intact (the exact path re-reads them)
constexpr int CH_BATCH = 4;
constexpr int CH_PWARPS = 6;
constexpr int CH_MIN_GW = 33;  // the committer handles up to 32 consecutive cells at once: one per column

struct CSlot {
  uint32_t id[32];  // material id, or the dummy id `n` (usage pinned at the cap) for lanes >= k
  float d2[32], l[32], a[32], b[32];
  uint32_t cont[32];
};

template <typename UsageT>
__global__ void __launch_bounds__((2 + CH_PWARPS) * 32, 1)
    greedy_chain_kernel(const uint32_t* __restrict__ cand_idx, const float* __restrict__ cand_d2, int k,
                        const float4* __restrict__ lab4, int n, int gw, int gh, unsigned max_usage, float weight,
                        int32_t* __restrict__ placed, uint32_t* __restrict__ info, uint32_t fallback_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CSlot* ring = reinterpret_cast<CSlot*>(smem_raw);
  float* s_d2_32 = reinterpret_cast<float*>(smem_raw + sizeof(CSlot) * CH_DEPTH);
  volatile unsigned long long* queue = reinterpret_cast<volatile unsigned long long*>(s_d2_32 + CH_DEPTH);
  volatile uint32_t* readyP = reinterpret_cast<volatile uint32_t*>(const_cast<unsigned long long*>(queue) + CH_DEPTH);
  volatile uint32_t* decided = readyP + CH_DEPTH;  // [0] = number of cells committed
  float4* rowlab = reinterpret_cast<float4*>(const_cast<uint32_t*>(decided) + 4);
  int32_t* rowbuf = reinterpret_cast<int32_t*>(rowlab + gw);
  UsageT* usage = reinterpret_cast<UsageT*>(rowbuf + ((gw + 3) & ~3));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cells = gw * gh;
  const bool use_penalty = weight > 0.0f;  // main.rs:355-362
  for (int i = tid; i < n; i += blockDim.x) usage[i] = 0;
  for (int i = tid; i < gw; i += blockDim.x) rowbuf[i] = -1;
  for (int i = tid; i < CH_DEPTH; i += blockDim.x) {
    readyP[i] = 0u;
    queue[i] = 0ull;
  }
  if (tid == 0) {
    decided[0] = 0u;
    usage[n] = (UsageT)~(UsageT)0;  // dummy material: never usable
  }
  __syncthreads();

  if (warp >= 2) {
    // ---------------- producers ----------------
    const int pw = warp - 2;
    const float two_w = __fmul_rn(2.0f, weight);
    for (int c0 = pw * CH_BATCH; c0 < cells; c0 += CH_PWARPS * CH_BATCH) {
      const int last = min(c0 + CH_BATCH, cells) - 1;
      while ((int)decided[0] + (CH_DEPTH - CH_GUARD) <= last) __nanosleep(64);  // slot still in use
      uint32_t id[CH_BATCH];
      float d2[CH_BATCH], d32[CH_BATCH];
      float4 lab[CH_BATCH];
#pragma unroll
      for (int b = 0; b < CH_BATCH; ++b) {
        const int c = min(c0 + b, cells - 1);
        id[b] = (uint32_t)n;
        d2[b] = __int_as_float(0x7f800000);
        if (lane < k) {
          id[b] = __ldg(cand_idx + (size_t)c * k + lane);
          d2[b] = __ldg(cand_d2 + (size_t)c * k + lane);
        }
        d32[b] = k > 32 ? __ldg(cand_d2 + (size_t)c * k + 32) : __int_as_float(0x7f800000);
      }
#pragma unroll
      for (int b = 0; b < CH_BATCH; ++b)
        lab[b] = lane < k ? __ldg(&lab4[id[b]]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < CH_BATCH; ++b) {
        const int c = c0 + b;
        if (c < cells) {  // warp-uniform
          CSlot& s = ring[c % CH_DEPTH];
          s.id[lane] = id[b];
          s.d2[lane] = d2[b];
          s.l[lane] = lab[b].x;
          s.a[lane] = lab[b].y;
          s.b[lane] = lab[b].z;
          __syncwarp();
          const float hi = use_penalty ? __fadd_rn(d2[b], two_w) : d2[b];
          int r = lane + 1;
          while (r < 32 && s.d2[r] < hi) ++r;  // d2 ascending: contenders are the lanes (lane, r)
          uint32_t cont = (r >= 32 ? 0xFFFFFFFFu : ((1u << r) - 1u)) & ~((2u << lane) - 1u);
          if (r >= 32 && d32[b] < hi) cont = 0xFFFFFFFFu;  // the next chunk may hold a contender
          s.cont[lane] = cont;
          if (lane == 0) s_d2_32[c % CH_DEPTH] = d32[b];
        }
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < CH_BATCH; ++b)
          if (c0 + b < cells) readyP[(c0 + b) % CH_DEPTH] = (uint32_t)(c0 + b) + 1u;
      }
    }
    return;
  }

  if (warp == 1) {
    // ---------------- committer ----------------
    const volatile CSlot* vring = ring;
    int base = 0;
    while (base < cells) {
      const int cell = base + lane;
      const int s = cell % CH_DEPTH;
      const unsigned long long e = queue[s];
      const bool have = cell < cells && (uint32_t)(e >> 32) == (uint32_t)cell + 1u;
      const unsigned ready_mask = __ballot_sync(0xffffffffu, have);
      const int cnt = __ffs(~ready_mask) - 1 < 0 ? 32 : __ffs(~ready_mask) - 1;  // consecutive from lane 0
      if (cnt == 0) {
        __nanosleep(20);
        continue;
      }
      if (lane < cnt) {
        const int32_t chosen = (int32_t)((uint32_t)e & 0x03FFFFFFu);
        const unsigned who = ((uint32_t)e >> 26) & 63u;
        float4 lab;
        if (who < 32u)
          lab = make_float4(vring[s].l[who], vring[s].a[who], vring[s].b[who], 0.f);
        else
          lab = __ldg(&lab4[chosen]);
        const int x = cell % gw;
        rowbuf[x] = chosen;
        rowlab[x] = lab;
        placed[cell] = chosen;  // main.rs:412
      }
      __threadfence_block();
      __syncwarp();
      base += cnt;
      if (lane == 0) decided[0] = (uint32_t)base;
    }
    return;
  }

  // ---------------- consumer (warp 0) ----------------
  const volatile CSlot* vring = ring;
  const volatile float* vd2_32 = s_d2_32;
  uint32_t n_fallbacks = 0;
  int32_t prev = -1;      // material placed in the previous cell (raster order)
  int32_t prev_u = -1;    // == prev while its usage increment is not yet reflected in the preloaded `u`
  unsigned prev_who = 63; // lane of prev in the previous cell's slot (63: not in the slot)
  int x = 0, y = 0;

  struct CellRegs {
    uint32_t id, cont;
    unsigned u;   // usage[id] read BEFORE the previous cell's increment (fixed up through prev_u)
    int32_t up;   // material above (-1: first row); warp-uniform
  };
  auto wait_slot = [&](int c, uint32_t flag) {
    const int s = c % CH_DEPTH;
    while (flag != (uint32_t)c + 1u) flag = readyP[s];
    asm volatile("" ::: "memory");
  };
  auto load_slot = [&](int c, CellRegs& r) {
    const int s = c % CH_DEPTH;
    r.id = vring[s].id[lane];
    r.cont = vring[s].cont[lane];
  };
  auto prev_lab = [&](int cell) -> float4 {  // Lab of the material placed in cell - 1
    if (prev_who < 32u) {
      const int ps = (cell - 1) % CH_DEPTH;
      return make_float4(vring[ps].l[prev_who], vring[ps].a[prev_who], vring[ps].b[prev_who], 0.f);
    }
    return __ldg(&lab4[prev]);
  };

  // exact multi-chunk resolution + reset fallback: the reference's loop, bit for bit (main.rs:335-475)
  auto resolve_generic = [&](int cell, int32_t left, int32_t up, float4 left_lab) -> int32_t {
    float4 up_lab = make_float4(0.f, 0.f, 0.f, 0.f);
    if (up >= 0) {
      const volatile float4* ul = rowlab + x;
      up_lab = make_float4(ul->x, ul->y, ul->z, 0.f);
    }
    unsigned best_bits = 0xFFFFFFFFu;
    int32_t best_id = -1;
    const size_t row = (size_t)cell * k;
    for (int base = 0; base < k; base += 32) {
      if (base > 0 && best_id >= 0 && __float_as_uint(cand_d2[row + base]) >= best_bits) break;  // exact early exit
      const int j = base + lane;
      uint32_t gid = 0;
      float gd2 = 0.f;
      float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
      bool gok = j < k;
      if (gok) {
        gid = cand_idx[row + j];
        gd2 = cand_d2[row + j];
        gl = __ldg(&lab4[gid]);
        gok = (unsigned)usage[gid] < max_usage && (int32_t)gid != left && (int32_t)gid != up;
      }
      float gpen = 0.0f;
      if (use_penalty) {  // adjacency.rs:71-87: Up, (Down), Left, (Right)
        if (up >= 0) gpen = __fadd_rn(gpen, adjacency_term(lab_distance_rn(gl.x, gl.y, gl.z, up_lab.x, up_lab.y, up_lab.z)));
        if (left >= 0)
          gpen = __fadd_rn(gpen, adjacency_term(lab_distance_rn(gl.x, gl.y, gl.z, left_lab.x, left_lab.y, left_lab.z)));
        gpen = __fmul_rn(gpen, weight);
      }
      const unsigned gtb = gok ? __float_as_uint(__fadd_rn(gd2, gpen)) : 0xFFFFFFFFu;
      const unsigned gmn = __reduce_min_sync(0xffffffffu, gtb);
      if (gmn < best_bits) {  // strict <: earlier chunks / earlier lanes win ties (main.rs:401-407)
        const int who = __ffs(__ballot_sync(0xffffffffu, gtb == gmn)) - 1;
        best_bits = gmn;
        best_id = (int32_t)__shfl_sync(0xffffffffu, gid, who);
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
    }
    usage[best_id] = (UsageT)(usage[best_id] + 1);  // main.rs:410-411 / 449 / 470
    __syncwarp();
    return best_id;
  };

  // exact totals of the first chunk (main.rs:386-407); returns the winning lane, or 63 if the chunk cannot decide
  auto resolve_exact = [&](int cell, int32_t left, int32_t up, bool valid) -> unsigned {
    const int sA = cell % CH_DEPTH;
    const float d2 = vring[sA].d2[lane];
    const float l = vring[sA].l[lane], a = vring[sA].a[lane], b = vring[sA].b[lane];
    const float d2_32 = vd2_32[sA];
    const volatile float4* ul = rowlab + x;
    const float ux = ul->x, uy = ul->y, uz = ul->z;
    float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (left >= 0) pl = prev_lab(cell);
    bool oor_u = false, oor_l = false;
    float pen = 0.0f;
    if (use_penalty) {
      float t_up = adjacency_term_fast(l, a, b, ux, uy, uz, oor_u);
      float t_left = adjacency_term_fast(l, a, b, pl.x, pl.y, pl.z, oor_l);
      t_up = up >= 0 ? t_up : 0.0f;  // absent terms are +0.0: x + 0.0 == x exactly
      t_left = left >= 0 ? t_left : 0.0f;
      pen = __fmul_rn(__fadd_rn(t_up, t_left), weight);  // adjacency.rs:71-87 order Up, Left; then * weight
    }
    const float total = __fadd_rn(d2, pen);  // main.rs:398
    unsigned tb = valid ? __float_as_uint(total) : 0xFFFFFFFFu;
    tb = ((oor_u && up >= 0) || (oor_l && left >= 0)) ? 0u : tb;  // out-of-domain arithmetic -> generic path
    const unsigned mn = __reduce_min_sync(0xffffffffu, tb);
    if (mn - 1u < 0xFFFFFFFEu && __float_as_uint(d2_32) >= mn)
      return (unsigned)(__ffs(__ballot_sync(0xffffffffu, tb == mn)) - 1);  // strict <, first in order on ties
    return 63u;
  };

  // one cell: A = registers of `cell`, B = registers of cell + 1; A is refilled with cell + 2 at the end.
  // Everything cell + 1 needs from shared memory is requested at the top, so that the loop-carried chain
  // (prev -> valid -> vote -> first lane -> shuffle -> prev) never waits on a load.
  auto step = [&](int cell, CellRegs& A, CellRegs& B) {
    const int sA = cell % CH_DEPTH;
    const uint32_t flag2 = (cell + 2 < cells) ? readyP[(cell + 2) % CH_DEPTH] : 0u;
    B.u = (unsigned)usage[B.id];  // before this cell's increment; fixed up through prev_u next iteration
    int xn = x + 1, yn = y;
    if (xn == gw) {
      xn = 0;
      yn = y + 1;
    }
    uint32_t dec = decided[0];
    int32_t up_n = *reinterpret_cast<volatile int32_t*>(rowbuf + xn);  // meaningful once cell + 1 - gw is committed

    const int32_t left = (x == 0) ? -1 : prev;
    const unsigned ucur = A.u + ((int32_t)A.id == prev_u ? 1u : 0u);
    const bool valid = (int32_t)A.id != A.up && (int32_t)A.id != left && ucur < max_usage;  // main.rs:375, 380
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int j0 = (__ffs(m) - 1) & 31;
    const unsigned cont0 = __shfl_sync(0xffffffffu, A.cont, j0);
    int32_t chosen = (int32_t)__shfl_sync(0xffffffffu, A.id, j0);
    const unsigned cnt0 = __shfl_sync(0xffffffffu, ucur, j0);
    unsigned who = (unsigned)j0;
    bool generic = false;
    if (m == 0u || (m & cont0) != 0u) {  // cold
      who = m != 0u ? resolve_exact(cell, left, A.up, valid) : 63u;
      if (who < 32u) {
        chosen = (int32_t)__shfl_sync(0xffffffffu, A.id, who);
        usage[chosen] = (UsageT)(__shfl_sync(0xffffffffu, ucur, who) + 1u);
      } else {
        float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (left >= 0) pl = prev_lab(cell);
        chosen = resolve_generic(cell, left, A.up, pl);
        generic = true;
      }
    } else {
      usage[chosen] = (UsageT)(cnt0 + 1u);  // main.rs:410-411; all lanes store the same value
    }
    queue[sA] = ((unsigned long long)((uint32_t)cell + 1u) << 32) | ((unsigned long long)who << 26) | (uint32_t)chosen;
    prev_u = chosen;
    if (generic) {  // B.u predates a possible usage reset: re-read it (it then includes this cell's increment)
      B.u = (unsigned)usage[B.id];
      prev_u = -1;
    }
    prev = chosen;
    prev_who = who;
    // Up neighbour of cell + 1
    if (yn > 0) {
      if ((int)dec < cell + 2 - gw) {  // cold: the committer lags by a whole row
        while ((int)decided[0] < cell + 2 - gw) {
        }
        asm volatile("" ::: "memory");
        up_n = *reinterpret_cast<volatile int32_t*>(rowbuf + xn);
      }
      B.up = up_n;
    } else {
      B.up = -1;
    }
    x = xn;
    y = yn;
    if (cell + 2 < cells) {
      wait_slot(cell + 2, flag2);
      load_slot(cell + 2, A);
    }
  };

  CellRegs R0, R1;
  R0.id = R1.id = (uint32_t)n;
  R0.cont = R1.cont = 0u;
  R0.u = R1.u = 0u;
  R0.up = R1.up = -1;
  if (cells > 0) {
    wait_slot(0, readyP[0]);
    load_slot(0, R0);
    if (cells > 1) {
      wait_slot(1, readyP[1 % CH_DEPTH]);
      load_slot(1, R1);
    }
  }
  for (int cell = 0; cell < cells; cell += 2) {
    step(cell, R0, R1);
    if (cell + 1 < cells) step(cell + 1, R1, R0);
  }
  if (lane == 0) info[0] = n_fallbacks;
}

}  // namespace

template <typename UsageT>
static int launch_chain(mosaic_ctx* ctx, unsigned cap, float weight, size_t smem) {
  MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_chain_kernel<UsageT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  greedy_chain_kernel<UsageT><<<1, (2 + CH_PWARPS) * 32, smem, ctx->stream>>>(
      ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->d_mat_lab4, ctx->n, ctx->gw, ctx->gh, cap, weight, ctx->d_placed,
      ctx->d_place_info, mosaic_ctx::kFallbackCap);
  return MOSAIC_OK;
}

int launch_place(mosaic_ctx* ctx, int max_usage, float weight) {
  const int n = ctx->n, gw = ctx->gw, gh = ctx->gh;
  const unsigned cap = max_usage < 0 ? 0u : (unsigned)max_usage;
  const size_t smem_limit = 220 * 1024;
  // fast path: rings + row state + usage counters in shared memory
  const size_t ring_fixed = sizeof(CSlot) * CH_DEPTH + (4 + 8 + 4) * CH_DEPTH + 16 + sizeof(float4) * (size_t)gw +
                            sizeof(int32_t) * (size_t)((gw + 3) & ~3);
  const size_t usage_bytes = cap <= 255u ? (size_t)n + 1 : (cap <= 65535u ? ((size_t)n + 1) * 2 : (size_t)-1);
  StageTimer t(ctx, MOSAIC_STAGE_PLACE);
  if (gw >= CH_MIN_GW && ctx->n < (1 << 26) && usage_bytes != (size_t)-1 && ring_fixed + usage_bytes + 16 <= smem_limit) {
    if (cap <= 255u)
      MOSAIC_TRY(launch_chain<uint8_t>(ctx, cap, weight, ring_fixed + usage_bytes + 16));
    else
      MOSAIC_TRY(launch_chain<uint16_t>(ctx, cap, weight, ring_fixed + usage_bytes + 16));
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
