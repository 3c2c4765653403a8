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
#include "place_records.cuh"

// MOSAIC_CHAIN_PAIRS=1 (measured experiment of round 2, off): take the reductions off the loop-carried chain by finding a
// cell's first AND second valid candidate one step early (vote + two shuffles) and selecting between them once the
// winner to its left is known.  Bit-exact, and slower: 3.07 ms instead of 2.59 on the 24,000-tile config — a hot block
// went from ~600 to ~800 cycles.  The step is not bound by the REDUX latency on its recurrence but by the ~45
// instructions a lone in-order warp issues per cell; a dozen more cost more than the shorter recurrence saves.
#ifndef MOSAIC_CHAIN_PAIRS
#define MOSAIC_CHAIN_PAIRS 0
#endif

namespace mosaic {

namespace {

template <bool USAGE_SMEM>
__global__ void __launch_bounds__(32, 1) greedy_generic_kernel(const uint32_t* __restrict__ cand_idx,
                                                             const float* __restrict__ cand_d2, int k,
                                                             const float4* __restrict__ lab4, int n, int gw, int gh,
                                                             unsigned max_usage, float weight,
                                                             int32_t* __restrict__ placed, uint32_t* usage_gmem,
                                                             uint32_t* __restrict__ info, uint32_t fallback_cap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
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
// Fast path: a fully parallel preparation kernel + ONE CTA with two warps around a TMA-fed ring.
//
//   place_prep_kernel     (whole GPU, one warp per cell) gathers the state-independent data of the cell's
//                        first 32 candidates — id, d2, Lab — into two contiguous records (a 256-byte hot one,
//                        a 528-byte cold one) and derives per candidate i its THREAT WINDOW: the set of earlier
//                        candidates j < i with  d2[i] < hi[j],  where
//                            hi[j] = fl(d2[j] + fl(fl(M_up[j] + M_left[j]) * w))
//                        bounds candidate j's total from above: M_x[j] is the adjacency term of j against the
//                        nearest OTHER material among the (<= 100) candidates of the neighbour cell — the
//                        neighbour's material is always one of them, IEEE rounding is monotone, and every total
//                        is bounded from below by its d2.  So if j0 is the first valid candidate, a later valid
//                        candidate i can be STRICTLY better than j0 only if bit j0 of its window is set.  The
//                        window is a 32-bit mask.  Lane 0 (which can never threaten anybody) carries the window
//                        of candidate 32, the first one beyond the record — and all ones for the first cell of
//                        every grid row, which sends those cells (the only ones where the material of the
//                        previous cell is NOT a neighbour) to the full step.
//   consumer  (warp 0)    walks the cells in raster order; records arrive in a 64-slot shared-memory ring by
//                        1-D TMA bulk copies (8 records per copy, one mbarrier per batch).  Its loop-carried
//                        chain is four instructions: compare (id == previous winner) -> select -> REDUX.MIN ->
//                        mask.  key = valid ? (lane << 26 | id) : ~0, so the minimum IS the first valid
//                        candidate and its id; validity (id != Up, usage < cap) is precomputed two cells ahead.
//                        A second reduction, REDUX.OR over the window masks of the valid lanes, is issued next to
//                        it — it does not depend on the winner — and one funnel shift by j0 tells whether any
//                        valid candidate threatens the first valid one.  If none does, the first valid candidate
//                        is exactly the reference's winner and no penalty is evaluated (91 % of the cells of the
//                        24,000-tile config).  Otherwise the exact totals of the first 32 are evaluated (sqrt /
//                        reciprocal sequences bit-identical to IEEE), and cells the record cannot decide (all 32
//                        rejected, or the early-exit bound fails) take the generic multi-chunk + reset-fallback
//                        path.  Every shortcut is an exact bound: placements are bit-identical to
//                        main.rs:335-475.
//   committer (warp 1)    drains the consumer's decisions, up to 32 cells at a time (one per lane): Lab of the
//                        row above, the placement array in HBM, the `decided` counter; lane 0 re-arms the
//                        mbarrier of every ring batch that has been fully consumed and issues its next TMA copies.
// Ring slots (records): a template parameter of the chain kernel.  A ring position is refilled when its eight cells have
// been consumed, so the copy of a batch has DEPTH - 8 cells (~100 cycles each) to arrive.  Alone on the GPU 64 slots are
// plenty (the copy takes ~1.5 us); under the kernels that run beside the chain in a step — matching of later bands, the
// render of finished rows — the copies take longer and the consumer met missing batches: 3.77 ms -> 3.64 ms per step at the
// 24,000-tile config with 128 slots (place 2.81 -> 2.68 ms).  The deepest ring that fits beside the usage counters is used.
constexpr int CH_DEPTH_MAX = 256;
constexpr int CH_DEPTH_MIN = 64;
constexpr int CH_BATCH = 8;    // records per TMA copy / mbarrier
constexpr int CH_SHADOW = 4;   // hot slots 64..67 mirror slots 0..3: a hot block reads five consecutive slots
constexpr int CH_GUARD = 3;    // slots of the last decided cells stay intact (the exact path re-reads them)
constexpr int CH_MIN_GW = 33;  // the committer handles up to 32 consecutive cells at once: one per column



__global__ void __launch_bounds__(256) place_prep_kernel(const uint32_t* __restrict__ cand_idx,
                                                         const float* __restrict__ cand_d2, int k,
                                                         const float4* __restrict__ lab4, int n, int cell_begin,
                                                         int cells /* end of the range */, int gw, float weight,
                                                         HRec* __restrict__ hrec, XRec* __restrict__ xrec,
                                                         float* __restrict__ glab /* cells x 3 x 128 */) {
  __shared__ float4 nb_s[8][2][kMaxK];  // per warp: (Lab, id) of every candidate of the cell above / to the left
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const bool use_penalty = weight > 0.0f;
  const float inf = __int_as_float(0x7f800000);
  float4(*nb)[kMaxK] = nb_s[threadIdx.x >> 5];
  for (int c = cell_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); c < cells; c += warps) {
    uint32_t id = (uint32_t)n;
    float d2 = inf;
    float4 lab = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < k) {
      id = __ldg(cand_idx + (size_t)c * k + lane);
      d2 = __ldg(cand_d2 + (size_t)c * k + lane);
      lab = __ldg(&lab4[id]);
    }
    const float d32 = k > 32 ? __ldg(cand_d2 + (size_t)c * k + 32) : inf;
    // Lab of ALL k candidates, gathered once here (in parallel, off the chain) so that the rare cells that need the
    // candidates behind the first 32 fetch ids, distances and colours in ONE round trip instead of two dependent ones
    {
      float* g = glab + (size_t)c * (3 * kKPad);
      g[lane] = lab.x;
      g[kKPad + lane] = lab.y;
      g[2 * kKPad + lane] = lab.z;
      for (int r = 32 + lane; r < k; r += 32) {
        const float4 q = __ldg(&lab4[__ldg(cand_idx + (size_t)c * k + r)]);
        g[r] = q.x;
        g[kKPad + r] = q.y;
        g[2 * kKPad + r] = q.z;
      }
    }
    // Upper bound of each candidate's penalty.  Whatever the replay decides, the material above / to the left is
    // one of the <= 100 candidates of that cell (main.rs:349-413; the fallback picks among the same list,
    // main.rs:436-470), and it is not this candidate (a candidate equal to a neighbour is rejected,
    // main.rs:380).  1/(sqrt(d2)+1) decreases with d2 and every rounding is monotone, so the term of the nearest
    // OTHER candidate of the neighbour cell bounds the real one from above.
    float m_sum = 0.0f;
    if (use_penalty) {
      const int nbc[2] = {c >= gw ? c - gw : -1, c % gw != 0 ? c - 1 : -1};  // adjacency.rs:71-87 order: Up, Left
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (nbc[s] >= 0)
          for (int u = lane; u < k; u += 32) {
            const uint32_t idu = __ldg(cand_idx + (size_t)nbc[s] * k + u);
            const float4 q = __ldg(&lab4[idu]);
            nb[s][u] = make_float4(q.x, q.y, q.z, __uint_as_float(idu));
          }
      __syncwarp();
      float m[2] = {0.0f, 0.0f};
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (nbc[s] >= 0) {
          float dmin = inf;
          for (int u = 0; u < k; ++u) {
            const float4 q = nb[s][u];
            const float dl = __fsub_rn(lab.x, q.x), da = __fsub_rn(lab.y, q.y), db = __fsub_rn(lab.z, q.z);
            const float dd = __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db));
            dmin = (__float_as_uint(q.w) != id && dd < dmin) ? dd : dmin;
          }
          m[s] = dmin < inf ? adjacency_term(__fsqrt_rn(dmin)) : 0.0f;
        }
      m_sum = __fadd_rn(m[0], m[1]);
    }
    // hi >= the candidate's total  fl(d2 + fl(fl(t_up + t_left) * w))  (main.rs:386-398); every total >= its d2
    const float hi = use_penalty ? __fadd_rn(d2, __fmul_rn(m_sum, weight)) : d2;
    uint32_t win = 0u, win32 = 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float hj = __shfl_sync(0xffffffffu, hi, j);
      win |= (j < lane && d2 < hj) ? (1u << j) : 0u;   // this candidate can be STRICTLY better than candidate j
      win32 |= d32 < hj ? (1u << j) : 0u;              // so can candidate 32 (and only then anyone behind it)
    }
    if (lane == 0) win = (c % gw == 0) ? 0xFFFFFFFFu : win32;
    hrec[c].wm[lane] = make_uint2(((uint32_t)lane << 26) | id, win);
    XRec& x = xrec[c];
    x.d2[lane] = d2;
    x.l[lane] = lab.x;
    x.a[lane] = lab.y;
    x.b[lane] = lab.z;
    if (lane == 0) {
      x.d2_32 = d32;
      x.win32 = win32;
    }
  }
}

// diagnostics (mosaic_debug_chain_profile): consumer-warp cycles and cell counts per path of the last
// PROFILED launch: [0..3] cycles fast / exact / generic / generic+fallback, [4..7] the cell counts,
// [8] cycles waiting for a TMA batch, [9] cycles waiting for the committer (row above), [10] total cycles
__device__ unsigned long long g_chain_prof[16];

template <typename UsageT, bool PROF, bool GATED, int DEPTH>
__global__ void __launch_bounds__(64, 1)
    greedy_chain_kernel(const HRec* __restrict__ hrec, const XRec* __restrict__ xrec, const float* __restrict__ glab,
                        const uint32_t* __restrict__ cand_idx, const float* __restrict__ cand_d2, int k,
                        const float4* __restrict__ lab4, int n, int gw, int gh,
                        unsigned max_usage, float weight, int32_t* __restrict__ placed, uint32_t* __restrict__ info,
                        uint32_t fallback_cap,
                        const volatile unsigned* ready /* cells whose records exist so far; NULL = all of them */,
                        volatile unsigned* progress /* GATED: cells committed so far, for whoever renders rows under the chain */,
                        bool progress_sys /* the counter is read by other GPUs: publish at system scope */) {
  constexpr int CH_DEPTH = DEPTH;
  constexpr int CH_NBATCH = DEPTH / CH_BATCH;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  HRec* hring = reinterpret_cast<HRec*>(smem_raw);
  XRec* xring = reinterpret_cast<XRec*>(smem_raw + sizeof(HRec) * (CH_DEPTH + CH_SHADOW));
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(xring + CH_DEPTH);
  volatile unsigned long long* queue = bars + CH_NBATCH;
  volatile uint32_t* decided = reinterpret_cast<volatile uint32_t*>(const_cast<unsigned long long*>(queue) + CH_DEPTH);
  float4* rowlab = reinterpret_cast<float4*>(const_cast<uint32_t*>(decided) + 4);
  int32_t* rowbuf = reinterpret_cast<int32_t*>(rowlab + gw);
  UsageT* usage = reinterpret_cast<UsageT*>(rowbuf + ((gw + 3) & ~3));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cells = gw * gh;
  const int nbatches = (cells + CH_BATCH - 1) / CH_BATCH;
  const bool use_penalty = weight > 0.0f;  // main.rs:355-362
  for (int i = tid; i < n; i += blockDim.x) usage[i] = 0;
  for (int i = tid; i < gw; i += blockDim.x) rowbuf[i] = -1;
  for (int i = tid; i < CH_DEPTH; i += blockDim.x) queue[i] = 0ull;
  if (tid == 0) {
    decided[0] = 0u;
    usage[n] = (UsageT)~(UsageT)0;  // dummy material: never usable
    for (int b = 0; b < CH_NBATCH; ++b) mbar_init(&bars[b], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue_batch = [&](int batch) {  // records [batch*8, batch*8+8) -> ring position batch % 8
    const int first = batch * CH_BATCH;
    const int cnt = min(CH_BATCH, cells - first);
    const int slot = first % CH_DEPTH;
    const int shadow = slot == 0 ? min(cnt, CH_SHADOW) : 0;
    unsigned long long* bar = &bars[batch % CH_NBATCH];
    mbar_expect_tx(bar, (unsigned)((cnt + shadow) * (int)sizeof(HRec) + cnt * (int)sizeof(XRec)));
    tma_load_1d(hring + slot, hrec + first, (unsigned)(cnt * (int)sizeof(HRec)), bar);
    tma_load_1d(xring + slot, xrec + first, (unsigned)(cnt * (int)sizeof(XRec)), bar);
    if (shadow) tma_load_1d(hring + CH_DEPTH, hrec + first, (unsigned)(shadow * (int)sizeof(HRec)), bar);
  };

  if (warp == 1) {
    // ---------------- committer ----------------
    const volatile XRec* vx = xring;
    int next_batch = 0;
    int base = 0;
    // GATED: the records may still be in production (the target image arrives in bands, api.cu): `ready` counts the
    // cells whose records a finished preparation kernel has written; a batch is fetched only when all its cells are
    // there.  (A template flag: the ungated instantiation is the tuned kernel, instruction for instruction.)
    unsigned ready_seen = 0u;
    auto batch_ready = [&](int batch) {
      const unsigned need = (unsigned)min(cells, (batch + 1) * CH_BATCH);
      if (need > ready_seen) {
        ready_seen = *ready;
        __threadfence();  // the records were written before the counter moved
      }
      return need <= ready_seen;
    };
    // ring position free (the first CH_NBATCH batches; later: all 8 cells older than the guard window) and records there
    auto try_issue = [&]() {
      while (next_batch < nbatches &&
             (next_batch < CH_NBATCH || (next_batch - CH_NBATCH + 1) * CH_BATCH + CH_GUARD <= base) &&
             batch_ready(next_batch)) {
        issue_batch(next_batch);
        ++next_batch;
      }
    };
    if (GATED) {
      if (lane == 0) try_issue();
    } else {
      if (lane == 0)
        for (; next_batch < min(CH_NBATCH, nbatches); ++next_batch) issue_batch(next_batch);
    }
    int next_pub = gw;
    while (base < cells) {
      const int cell = base + lane;
      const int s = cell % CH_DEPTH;
      const unsigned long long e = queue[s];
      const bool have = cell < cells && (uint32_t)(e >> 32) == (uint32_t)cell + 1u;
      const unsigned ready_mask = __ballot_sync(0xffffffffu, have);
      const int cnt = ready_mask == 0xFFFFFFFFu ? 32 : __ffs(~ready_mask) - 1;  // consecutive from lane 0
      if (cnt == 0) {
        if (GATED && lane == 0) try_issue();  // (a batch that was not there when its ring position became free)
        __nanosleep(20);
        continue;
      }
      if (lane < cnt) {
        const int32_t chosen = (int32_t)((uint32_t)e & CH_IDMASK);
        const unsigned who = ((uint32_t)e >> 26) & 63u;
        float4 lab;
        if (who < 32u)
          lab = make_float4(vx[s].l[who], vx[s].a[who], vx[s].b[who], 0.f);
        else
          lab = __ldg(&lab4[chosen]);
        rowlab[cell % gw] = lab;
        placed[cell] = chosen;  // main.rs:412
      }
      __threadfence_block();
      __syncwarp();
      base += cnt;
      if (lane == 0) {
        decided[0] = (uint32_t)base;
        if (GATED && progress && (base >= next_pub || base >= cells)) {  // the placements of the cells before `base` are in HBM
          // (once per grid row: the fence is a round trip to L2 in the loop that also refills the ring, and whoever
          // renders under the chain waits for whole rows anyway)
          if (progress_sys)
            __threadfence_system();  // (a sharded context's peers poll this counter and read the placements over NVLink)
          else
            __threadfence();
          *progress = (unsigned)base;
          next_pub = (base / gw + 1) * gw;
        }
        // ring batches whose 8 cells are all older than the guard window are free: fetch their successors
        if (GATED) {
          try_issue();
        } else {
          while (next_batch < nbatches && (next_batch - CH_NBATCH + 1) * CH_BATCH + CH_GUARD <= base) {
            issue_batch(next_batch);
            ++next_batch;
          }
        }
      }
    }
    return;
  }

  // ---------------- consumer (warp 0) ----------------
  // Software pipeline over five register sets (cells c .. c+4):
  //   c+4  record words requested from the ring
  //   c+3  usage counter + material above requested (AFTER the commit of c, so the counter covers cells <= c)
  //   c+2  key + window mask computed (counter += [id == winner of c])
  //   c+1  REDUX.MIN / REDUX.OR issued with the presumed winner of c, BEFORE the threat test of c resolves
  //   c    threat test -> (cold path) -> commit
  // so that on the hot path no instruction consumes a shared-memory load issued in the same step, and the
  // test/branch of cell c overlaps the reductions of cell c+1.
  const volatile HRec* vh = hring;
  const volatile XRec* vx = xring;
  constexpr uint32_t NOKEY = 0xFFFFFFFFu;
  const bool l0 = lane == 0;
  uint32_t n_fallbacks = 0;
  unsigned long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // [11] cycles from a block's exit to full_step, [12] exits, [13] cycles of full steps of uncontended cells, [14] their count, [15] hot blocks run

  struct CellSet {
    uint32_t w, m;  // record words: lane << 26 | id, window mask
    uint32_t id;
    unsigned u;     // usage[id]; see the pipeline comment for which cells it covers at each stage
    int32_t up;     // material above (-1: first row); warp-uniform
    int x;          // column
    uint32_t k0;    // redux key if id is not the winner of the previous cell (else: NOKEY, except in column 0)
    uint32_t km;    // window mask that goes into the OR reduction in that case
  };
  auto wait_batch = [&](int c) {  // the batch of cell c
    unsigned long long* bar = &bars[(c / CH_BATCH) % CH_NBATCH];
    const unsigned parity = (unsigned)((c / CH_DEPTH) & 1);
    if (!mbar_test(bar, parity)) {
      long long tw = 0;
      if (PROF) tw = clock64();
      mbar_wait(bar, parity);
      if (PROF) prof[8] += (unsigned long long)(clock64() - tw);
    }
  };
  auto load_rec = [&](int c, CellSet& s) {
    const int slot = c % CH_DEPTH;
    s.w = vh[slot].wm[lane].x;
    s.m = vh[slot].wm[lane].y;
  };
  auto make_keys = [&](CellSet& s) {
    const bool ok = (int32_t)s.id != s.up && s.u < max_usage;  // main.rs:380 (Up), main.rs:375
    s.k0 = ok ? s.w : NOKEY;
    s.km = (ok || l0) ? s.m : 0u;
  };
  auto row_above_lab = [&](int cell, int x) -> float4 {  // rowlab[x] holds the Lab of cell - gw once it is committed
    const volatile float4* ul = rowlab + x;
    float4 v = make_float4(ul->x, ul->y, ul->z, 0.f);
    if ((int)decided[0] <= cell - gw) {  // cold: the committer lags by a whole row
      long long tw = 0;
      if (PROF) tw = clock64();
      while ((int)decided[0] <= cell - gw) {
      }
      if (PROF) prof[9] += (unsigned long long)(clock64() - tw);
      asm volatile("" ::: "memory");
      v = make_float4(ul->x, ul->y, ul->z, 0.f);
    }
    return v;
  };

  uint32_t r = NOKEY;      // REDUX.MIN of the current cell: first valid lane << 26 | its material id
  uint32_t T = NOKEY;      // REDUX.OR of the current cell: bit j set <=> some valid candidate threatens candidate j
  uint32_t rprev = NOKEY;  // r of the previous cell after its resolution (lane 63 = not in the record)

  auto prev_lab = [&](int cell) -> float4 {  // Lab of the material placed in cell - 1
    const unsigned prev_who = rprev >> 26;
    if (prev_who < 32u) {
      const int ps = (cell - 1) % CH_DEPTH;
      return make_float4(vx[ps].l[prev_who], vx[ps].a[prev_who], vx[ps].b[prev_who], 0.f);
    }
    return __ldg(&lab4[rprev & CH_IDMASK]);
  };

  // exact multi-chunk resolution + reset fallback: the reference's loop, bit for bit (main.rs:335-475)
  auto resolve_generic = [&](int cell, int x, int32_t left, int32_t up, float4 left_lab) -> int32_t {
    const size_t row = (size_t)cell * k;
    // all chunks are requested up front, colours from the table the preparation kernel gathered: ONE global round trip
    uint32_t gid[4];
    float gd2[4];
    float4 gl[4];
    const float* g = glab + (size_t)cell * (3 * kKPad);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int j = ch * 32 + lane;
      gid[ch] = 0u;
      gd2[ch] = 0.f;
      gl[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < k) {
        gid[ch] = __ldg(cand_idx + row + j);
        gd2[ch] = __ldg(cand_d2 + row + j);
        gl[ch] = make_float4(__ldg(g + j), __ldg(g + kKPad + j), __ldg(g + 2 * kKPad + j), 0.f);
      }
    }
    float4 up_lab = make_float4(0.f, 0.f, 0.f, 0.f);
    if (up >= 0) up_lab = row_above_lab(cell, x);
    unsigned best_bits = 0xFFFFFFFFu;
    int32_t best_id = -1;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int base = ch * 32;
      if (base >= k) break;
      if (base > 0 && best_id >= 0 && __float_as_uint(__shfl_sync(0xffffffffu, gd2[ch], 0)) >= best_bits) break;  // exact early exit
      bool gok = base + lane < k;
      if (gok) gok = (unsigned)usage[gid[ch]] < max_usage && (int32_t)gid[ch] != left && (int32_t)gid[ch] != up;
      float gpen = 0.0f;
      if (use_penalty) {  // adjacency.rs:71-87: Up, (Down), Left, (Right); absent terms are +0.0 (x + 0.0 == x)
        bool oor_u = false, oor_l = false;
        float t_up = adjacency_term_fast(gl[ch].x, gl[ch].y, gl[ch].z, up_lab.x, up_lab.y, up_lab.z, oor_u);
        float t_left = adjacency_term_fast(gl[ch].x, gl[ch].y, gl[ch].z, left_lab.x, left_lab.y, left_lab.z, oor_l);
        if (__any_sync(0xffffffffu, (oor_u && up >= 0) || (oor_l && left >= 0))) {  // outside the fast domain
          t_up = adjacency_term(lab_distance_rn(gl[ch].x, gl[ch].y, gl[ch].z, up_lab.x, up_lab.y, up_lab.z));
          t_left = adjacency_term(lab_distance_rn(gl[ch].x, gl[ch].y, gl[ch].z, left_lab.x, left_lab.y, left_lab.z));
        }
        t_up = up >= 0 ? t_up : 0.0f;
        t_left = left >= 0 ? t_left : 0.0f;
        gpen = __fmul_rn(__fadd_rn(t_up, t_left), weight);
      }
      const unsigned gtb = gok ? __float_as_uint(__fadd_rn(gd2[ch], gpen)) : 0xFFFFFFFFu;
      const unsigned gmn = __reduce_min_sync(0xffffffffu, gtb);
      if (gmn < best_bits) {  // strict <: earlier chunks / earlier lanes win ties (main.rs:401-407)
        const int who = __ffs(__ballot_sync(0xffffffffu, gtb == gmn)) - 1;
        best_bits = gmn;
        best_id = (int32_t)__shfl_sync(0xffffffffu, gid[ch], who);
      }
    }
    if (best_id < 0) {
      // fallback_tile_selection (main.rs:420-475)
      if (lane == 0 && n_fallbacks < fallback_cap) info[1 + n_fallbacks] = (uint32_t)cell;
      ++n_fallbacks;
      for (int i = lane; i < n; i += 32) usage[i] = 0;  // lib.rs:86-88
      __syncwarp();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch * 32 >= k || best_id >= 0) break;
        const bool fok = ch * 32 + lane < k && (int32_t)gid[ch] != left && (int32_t)gid[ch] != up;  // main.rs:448
        const unsigned m = __ballot_sync(0xffffffffu, fok);
        if (m) best_id = (int32_t)__shfl_sync(0xffffffffu, gid[ch], __ffs(m) - 1);
      }
      if (best_id < 0) best_id = (int32_t)__shfl_sync(0xffffffffu, gid[0], 0);  // main.rs:456-474 nearest_one
    }
    __syncwarp();
    if (lane == 0) usage[best_id] = (UsageT)(usage[best_id] + 1);  // main.rs:410-411 / 449 / 470
    __syncwarp();
    return best_id;
  };

  // exact totals of the record's 32 candidates (main.rs:386-407); returns the winning lane, or 63 if they cannot decide
  auto resolve_exact = [&](int cell, int x, int32_t left, int32_t up, bool valid) -> unsigned {
    const int sA = cell % CH_DEPTH;
    // every load is requested before the first one is consumed
    const float d2 = vx[sA].d2[lane];
    const float l = vx[sA].l[lane], a = vx[sA].a[lane], b = vx[sA].b[lane];
    const float d2_32 = vx[sA].d2_32;
    float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (left >= 0) pl = prev_lab(cell);
    float4 ul = make_float4(0.f, 0.f, 0.f, 0.f);
    if (up >= 0) ul = row_above_lab(cell, x);
    bool oor_u = false, oor_l = false;
    float pen = 0.0f;
    if (use_penalty) {
      float t_up = adjacency_term_fast(l, a, b, ul.x, ul.y, ul.z, oor_u);
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

  // One cell with every path (threat test -> exact totals -> generic multi-chunk / reset fallback).  ONE copy of
  // this code exists: the hot loop below leaves through it whenever a cell is (possibly) contended or starts a
  // grid row, and the last cells of the grid run through it.
  auto full_step = [&](int cell, CellSet& A, CellSet& B, CellSet& C, CellSet& D, CellSet& E) {
    long long t_begin = 0;
    unsigned long long w0 = 0;
    if (PROF) {
      t_begin = clock64();
      w0 = prof[8] + prof[9];
    }
    int path = 0;
    // (1) the key of this cell, for real: in column 0 the previous winner is not a neighbour (main.rs:309-333)
    const int32_t prev = cell > 0 ? (int32_t)(rprev & CH_IDMASK) : -1;
    const bool same = cell > 0 && A.id == (uint32_t)prev;
    const unsigned ucur = A.u + (same ? 1u : 0u);  // usage[id] over every earlier cell
    uint32_t key = same ? NOKEY : A.k0;
    uint32_t tm = (same && !l0) ? 0u : A.km;
    if (A.x == 0) {
      const bool ok = (int32_t)A.id != A.up && ucur < max_usage;
      key = ok ? A.w : NOKEY;
      tm = l0 ? vx[cell % CH_DEPTH].win32 : (ok ? A.m : 0u);  // (lane 0 of the hot record carries all ones here)
    }
    r = __reduce_min_sync(0xffffffffu, key);  // (a dead hot block leaves r and T undefined)
    T = __reduce_or_sync(0xffffffffu, tm);
    const uint32_t j0 = r >> 26;
    if (__funnelshift_rc(T, 0xFFFFFFFFu, j0) & 1u) {  // (possibly) more than one contender, or no valid candidate
      const bool valid = key != NOKEY;
      const int32_t left = (A.x == 0) ? -1 : prev;
      unsigned who = 63u;
      path = 1;
      if (j0 < 32u) who = resolve_exact(cell, A.x, left, A.up, valid);
      if (who < 32u) {
        r = __shfl_sync(0xffffffffu, A.w, who);
      } else {
        float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (left >= 0) pl = prev_lab(cell);
        const uint32_t nfb0 = n_fallbacks;
        const int32_t g = resolve_generic(cell, A.x, left, A.up, pl);
        path = n_fallbacks != nfb0 ? 3 : 2;
        r = (63u << 26) | (uint32_t)g;
        // the generic path updated (possibly reset) the counters itself: re-read the counters in flight so
        // that they cover exactly the cells they are supposed to cover (all of them exclude this cell)
        B.u = (unsigned)usage[B.id] - (B.id == (uint32_t)g ? 1u : 0u);
        make_keys(B);
        C.u = (unsigned)usage[C.id] - (C.id == (uint32_t)g ? 1u : 0u);
      }
    }
    // (2) commit
    const uint32_t chosen = r & CH_IDMASK;
    if ((uint32_t)lane == (r >> 26)) usage[A.id] = (UsageT)(ucur + 1u);  // main.rs:410-411
    if (lane == 0) {
      *reinterpret_cast<volatile int32_t*>(rowbuf + A.x) = (int32_t)chosen;
      queue[cell % CH_DEPTH] = ((unsigned long long)((uint32_t)cell + 1u) << 32) | r;
    }
    __syncwarp();
    // (3) reductions of cell + 1
    const bool same_b = B.id == chosen;
    const uint32_t r_next = __reduce_min_sync(0xffffffffu, same_b ? NOKEY : B.k0);
    const uint32_t T_next = __reduce_or_sync(0xffffffffu, (same_b && !l0) ? 0u : B.km);
    // (4) cell + 3: unpack the record requested one step ago; usage counter (covers cells <= cell), material above
    D.id = D.w & CH_IDMASK;
    D.u = (unsigned)usage[D.id];
    D.up = *reinterpret_cast<volatile int32_t*>(rowbuf + D.x);
    // (5) cell + 2: key + mask (counter now covers cells <= cell)
    C.u += C.id == chosen ? 1u : 0u;
    make_keys(C);
    // (6) cell + 4: record
    E.x = D.x + 1 == gw ? 0 : D.x + 1;
    if (cell + 4 < cells) {
      if (((cell + 4) % CH_BATCH) == 0) wait_batch(cell + 4);
      load_rec(cell + 4, E);
    }
    rprev = r;
    r = r_next;
    T = T_next;
    if (PROF) {
      prof[path] += (unsigned long long)(clock64() - t_begin) - (prof[8] + prof[9] - w0);
      prof[4 + path] += 1ull;
      if (path == 0) prof[13] += (unsigned long long)(clock64() - t_begin), prof[14]++;
    }
  };

  // The hot block: five cells whose first valid candidates are provably the reference's winners.  It is ONE
  // basic block — no branch between the steps — so that ptxas can park the reductions of cell c+1 behind the
  // bookkeeping of cell c: once a cell is (possibly) contended the block goes `dead` and every later state
  // update in it is predicated off; the caller re-runs that cell through full_step.  No cold code inside: the
  // unrolled steps must stay inside the ~6 KB L0 instruction cache of the SM sub-partition (a far jump per cell
  // to out-of-line code costs more than the whole dependency chain).  Each decision is parked in lane STEP's
  // `myr`; the block's decisions go to the row buffer and the committer's queue with one store per lane.
  // `hb`: this lane's pair in the hot record of the block's first cell + 4 (five consecutive slots).
  uint32_t myr = 0u;
  bool dead = false;
  int done = 0;
#if MOSAIC_CHAIN_PAIRS
  // first and second valid candidate of a cell (record words: lane << 26 | id; NOKEY if there is none) and the OR of
  // the valid candidates' windows — everything the next-but-one step needs from this cell except who wins the cell
  // in between
  uint32_t p1 = NOKEY, p2 = NOKEY, TB = 0xFFFFFFFFu;
  auto find_pair = [&](const CellSet& Cs, uint32_t& q1, uint32_t& q2, uint32_t& tt) {
    const unsigned v1 = __ballot_sync(0xffffffffu, Cs.k0 != NOKEY);
    const unsigned v2 = v1 & (v1 - 1u);
    const uint32_t a = __shfl_sync(0xffffffffu, Cs.w, (__ffs(v1) - 1) & 31), b = __shfl_sync(0xffffffffu, Cs.w, (__ffs(v2) - 1) & 31);
    q1 = v1 ? a : NOKEY;
    q2 = v2 ? b : NOKEY;
    tt = __reduce_or_sync(0xffffffffu, Cs.km);
  };
#endif
  auto hot_step = [&](const int STEP, const uint2* hb, CellSet& A, CellSet& B, CellSet& C, CellSet& D, CellSet& E) {
    const uint32_t rc = r;
    const uint32_t chosen = rc & CH_IDMASK;
    const uint32_t j0 = rc >> 26;
    dead = dead || (__funnelshift_rc(T, 0xFFFFFFFFu, j0) & 1u) != 0u;  // j0 = 63 (no valid candidate) shifts the ones in
    const bool live = !dead;
#if MOSAIC_CHAIN_PAIRS
    // The winner of the next cell without a reduction on the loop-carried chain: its first and second valid candidates
    // (p1, p2: valid by usage and the material above; only the rule of the material to the LEFT — the winner of this
    // cell — is open) were found a step ago.  If this cell's winner is the first one, the second one wins.
    r = (p1 & CH_IDMASK) == chosen ? p2 : p1;
    T = TB;  // (may still count the excluded candidate's window: a superset of the threats, so at worst a cold cell more)
#else
    const bool same_b = B.id == chosen;
    r = __reduce_min_sync(0xffffffffu, same_b ? NOKEY : B.k0);  // the winner of this cell is presumed
    T = __reduce_or_sync(0xffffffffu, (same_b && !l0) ? 0u : B.km);
#endif
    // commit: the winner is not the previous winner (that lane's key is NOKEY), so A.u covers every earlier cell
    // (a predicated store, not a branch: the step must stay one basic block.  The warp is converged and its
    // shared-memory accesses execute in program order, so the counter reads below see the store.)
    st_shared_if(live && (uint32_t)lane == j0, &usage[A.id], (UsageT)(A.u + 1u));  // main.rs:410-411
    myr = (live && lane == STEP) ? rc : myr;
    done += live ? 1 : 0;
    // cell + 3: usage counter (covers cells <= cell), material above
    const uint32_t did = D.w & CH_IDMASK;
    const unsigned du = (unsigned)usage[did];
    const int32_t dup = rowbuf[D.x];
    D.id = live ? did : D.id;
    D.u = live ? du : D.u;
    D.up = live ? dup : D.up;
    // cell + 2: key + mask
    C.u += (live && C.id == chosen) ? 1u : 0u;
    const bool ok = (int32_t)C.id != C.up && C.u < max_usage;  // main.rs:380 (Up), main.rs:375
    C.k0 = live ? (ok ? C.w : NOKEY) : C.k0;
    C.km = live ? ((ok || l0) ? C.m : 0u) : C.km;
#if MOSAIC_CHAIN_PAIRS
    find_pair(C, p1, p2, TB);  // for the step after the next one: two steps of slack for a vote, two shuffles and a REDUX.OR
#endif
    // cell + 4: record (its batch has been waited for)
    const uint2 e = hb[STEP * 32];
    const int ex = D.x + 1 == gw ? 0 : D.x + 1;
    E.x = live ? ex : E.x;
    E.w = live ? e.x : E.w;
    E.m = live ? e.y : E.m;
  };

  CellSet S[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    S[t].w = (uint32_t)n;  // (a dead hot step may look up the counter of a set that has not been loaded yet)
    S[t].m = 0u;
    S[t].id = (uint32_t)n;
    S[t].u = 0u;
    S[t].up = -1;
    S[t].x = t;  // gw >= CH_MIN_GW: the first cells are in the first row
    S[t].k0 = NOKEY;
    S[t].km = 0u;
  }
  // prologue (gw >= 33 => cells >= 33): records of cells 0..3; cells 0..2 see the initial counters and no row above
  wait_batch(0);
#pragma unroll
  for (int t = 0; t < 4; ++t) load_rec(t, S[t]);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    S[t].id = S[t].w & CH_IDMASK;
    S[t].u = (unsigned)usage[S[t].id];  // zero, except for the dummy material of lanes >= k (pinned at the cap)
  }
  make_keys(S[0]);
  make_keys(S[1]);
  r = __reduce_min_sync(0xffffffffu, S[0].k0);
  T = __reduce_or_sync(0xffffffffu, S[0].km);
  long long t_all = 0;
  if (PROF) t_all = clock64();
  auto rotate = [&](int by) {  // S[i] <- S[(i + by) % 5]: the set of the current cell moves to S[0] (cold)
    for (int t = 0; t < by; ++t) {
      const CellSet s0 = S[0];
      S[0] = S[1];
      S[1] = S[2];
      S[2] = S[3];
      S[3] = S[4];
      S[4] = s0;
    }
  };
  const uint32_t ucells = (uint32_t)cells;
  uint32_t cell = 0;
  while (cell < ucells) {
    if (cell + 9u <= ucells) {  // hot blocks: records up to cell + 8 are requested
      wait_batch((int)cell + 8);  // a block reads records cell + 4 .. cell + 8: this batch, or one already waited for
      bool ran_block = false;
#if MOSAIC_CHAIN_PAIRS
      find_pair(S[1], p1, p2, TB);  // (after a cold cell or at the start: the pair of the block's second cell)
#endif
      for (;;) {
        // probe the batch the NEXT block needs now, look at the answer after this block (near the end of the grid
        // the probe looks at a batch that is certainly there: branches are what this loop cannot afford)
        const uint32_t c13 = min(cell + 13u, ucells - 1u);
        const bool next_ready = mbar_test(&bars[(c13 / CH_BATCH) % CH_NBATCH], (c13 / CH_DEPTH) & 1u);
        const uint2* hb = reinterpret_cast<const uint2*>(hring) + ((cell + 4u) % CH_DEPTH) * 32u + lane;
        const int x0 = S[0].x;
        dead = false;
        done = 0;
        hot_step(0, hb, S[0], S[1], S[2], S[3], S[4]);
        hot_step(1, hb, S[1], S[2], S[3], S[4], S[0]);
        hot_step(2, hb, S[2], S[3], S[4], S[0], S[1]);
        hot_step(3, hb, S[3], S[4], S[0], S[1], S[2]);
        hot_step(4, hb, S[4], S[0], S[1], S[2], S[3]);
        {  // flush the parked decisions (predicated stores): lane t holds the one of cell + t
          int xs = x0 + lane;
          xs = xs >= gw ? xs - gw : xs;
          xs = lane < done ? xs : 0;
          st_shared_if(lane < done, reinterpret_cast<unsigned*>(rowbuf) + xs, myr & CH_IDMASK);
          st_shared_if(lane < done, const_cast<unsigned long long*>(queue) + ((cell + (uint32_t)lane) % CH_DEPTH),
                       ((unsigned long long)(cell + (uint32_t)lane + 1u) << 32) | myr);
        }
        cell += (uint32_t)done;
        if (PROF) prof[15]++;
        // one exit test per block: a contended cell, the end of the grid, or records that have not arrived yet
        if (done < 5 || cell + 9u > ucells || !next_ready) break;
        ran_block = true;
      }
      long long t_exit = 0;
      if (PROF) t_exit = clock64();
      __syncwarp();
      // the decision of the last cell that went through: lane done-1 of this block, else lane 4 of the one before
      if (done > 0 || ran_block) rprev = __shfl_sync(0xffffffffu, myr, done > 0 ? done - 1 : 4);
      rotate(done % 5);
      if (PROF) prof[11] += (unsigned long long)(clock64() - t_exit), prof[12]++;
      if (cell >= ucells) break;
      if (done == 5) continue;  // only the next records were missing (or the grid is about to end): wait and go on
    }
    full_step((int)cell, S[0], S[1], S[2], S[3], S[4]);
    rotate(1);
    cell += 1u;
  }
  if (lane == 0) info[0] = n_fallbacks;
  if (PROF && lane == 0) {
    prof[10] = (unsigned long long)(clock64() - t_all);
    // the hot steps carry no instrumentation: their share is what the other paths and the waits leave
    prof[4] = (unsigned long long)cells - prof[5] - prof[6] - prof[7];
    prof[0] = prof[10] - prof[1] - prof[2] - prof[3] - prof[8] - prof[9];
    for (int i = 0; i < 16; ++i) g_chain_prof[i] = prof[i];
  }
}

}  // namespace

static size_t chain_smem_at(const mosaic_ctx* ctx, unsigned cap, int depth);
static int chain_depth(const mosaic_ctx* ctx, unsigned cap);

template <typename UsageT, bool PROF, bool GATED, int DEPTH>
static int launch_chain_depth(mosaic_ctx* ctx, unsigned cap, float weight, cudaStream_t stream, const unsigned* ready, unsigned* progress) {
  size_t smem = chain_smem_at(ctx, cap, DEPTH);
  const size_t cells = (size_t)ctx->gw * ctx->gh;
  // under the chain other kernels run (matching of later bands, the render of finished rows): claiming all of the SM's
  // shared memory keeps their CTAs off the SM of the one warp everything waits for
  if (GATED) smem = std::max(smem, (size_t)226 * 1024);
  MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_chain_kernel<UsageT, PROF, GATED, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  greedy_chain_kernel<UsageT, PROF, GATED, DEPTH><<<1, 64, smem, stream>>>(
      reinterpret_cast<const HRec*>(ctx->d_place_rec), reinterpret_cast<const XRec*>(ctx->d_place_rec + cells * sizeof(HRec)),
      reinterpret_cast<const float*>(ctx->d_place_rec + cells * (sizeof(HRec) + sizeof(XRec))), ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->sim_lab4(), ctx->n, ctx->gw, ctx->gh, cap, weight, ctx->d_placed,
      ctx->d_place_info, mosaic_ctx::kFallbackCap, ready, progress, ctx->world > 1);
  return MOSAIC_OK;
}

template <typename UsageT, bool PROF, bool GATED = false>
static int launch_chain_impl(mosaic_ctx* ctx, unsigned cap, float weight, size_t /*smem: follows from the depth*/, cudaStream_t stream,
                             const unsigned* ready, unsigned* progress = nullptr) {
  switch (chain_depth(ctx, cap)) {
    case 256: return launch_chain_depth<UsageT, PROF, GATED, 256>(ctx, cap, weight, stream, ready, progress);
    case 128: return launch_chain_depth<UsageT, PROF, GATED, 128>(ctx, cap, weight, stream, ready, progress);
    default: return launch_chain_depth<UsageT, PROF, GATED, 64>(ctx, cap, weight, stream, ready, progress);
  }
}

static int ensure_place_records(mosaic_ctx* ctx) {
  const int cells = ctx->gw * ctx->gh;
  return ensure(ctx, &ctx->d_place_rec, &ctx->place_rec_cap, (size_t)cells * (sizeof(HRec) + sizeof(XRec) + 3 * kKPad * sizeof(float)));
}

// ---- which chain ------------------------------------------------------------------------------------------------------
//   1  first generation (this file): one consumer warp behind a TMA ring, 108 ns per cell on the 24,000-tile mosaic
//   2  windowed (greedy_window.cu): lane = cell scans + shortlists, 200 ns per cell — a measured experiment
//      (profiles/r02_chain_window_experiment.json)
//   3  wave (greedy_wave.cu): a window of 32 cells iterated to its fixed point, one warp per cell
// mosaic_debug_set_chain(ctx, v) or MOSAIC_CHAIN=v1|window|wave in the environment select one (all parity-tested);
// configurations a chain does not cover fall back to the first generation / the generic kernel.
bool window_chain_eligible(const mosaic_ctx* ctx, int max_usage);
int launch_window_prep_range(mosaic_ctx* ctx, float weight, int cell_begin, int cell_end);
int launch_window_chain(mosaic_ctx* ctx, int max_usage, float weight, cudaStream_t stream, const unsigned* ready, bool prof);
int read_window_profile(mosaic_ctx* ctx, unsigned long long* out16);
bool wave_chain_eligible(const mosaic_ctx* ctx, int max_usage);
int launch_wave_chain(mosaic_ctx* ctx, int max_usage, float weight, cudaStream_t stream, const unsigned* ready, unsigned* progress,
                      bool prof);
int read_wave_profile(mosaic_ctx* ctx, unsigned long long* out16);
constexpr int kDefaultChain = 1;
static int chain_choice(const mosaic_ctx* ctx) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("MOSAIC_CHAIN");
    env = !e ? kDefaultChain : (e[0] == 'w' && e[1] == 'a') ? 3 : (e[0] == 'w') ? 2 : (e[0] == 'v') ? 1 : kDefaultChain;
  }
  return ctx->chain_version ? ctx->chain_version : env;
}
static bool use_v1_chain(const mosaic_ctx* ctx) { return chain_choice(ctx) != 2; }  // (the wave chain reads the first generation's records)
static bool use_window(const mosaic_ctx* ctx, int max_usage) { return chain_choice(ctx) == 2 && window_chain_eligible(ctx, max_usage); }
static bool use_wave(const mosaic_ctx* ctx, int max_usage) { return chain_choice(ctx) == 3 && wave_chain_eligible(ctx, max_usage); }

// candidate records of cells [cell_begin, cell_end) (needs the candidate lists of those cells and of the row above)
int launch_place_prep_range(mosaic_ctx* ctx, float weight, int cell_begin, int cell_end) {
  if (!use_v1_chain(ctx)) return launch_window_prep_range(ctx, weight, cell_begin, cell_end);
  if (cell_end <= cell_begin) return MOSAIC_OK;
  const int cells = ctx->gw * ctx->gh;
  MOSAIC_TRY(ensure_place_records(ctx));
  int blocks = (cell_end - cell_begin + 7) / 8;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  place_prep_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->sim_lab4(), ctx->n, cell_begin,
                                                     cell_end, ctx->gw, weight, reinterpret_cast<HRec*>(ctx->d_place_rec),
                                                     reinterpret_cast<XRec*>(ctx->d_place_rec + (size_t)cells * sizeof(HRec)),
                                                     reinterpret_cast<float*>(ctx->d_place_rec + (size_t)cells * (sizeof(HRec) + sizeof(XRec))));
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

// shared-memory footprint of the chain kernel, or 0 when the grid / cap / material count need the generic kernel
static size_t chain_smem_at(const mosaic_ctx* ctx, unsigned cap, int depth) {
  const size_t smem_limit = 220 * 1024;
  const int n = ctx->n, gw = ctx->gw;
  const size_t ring_fixed = sizeof(HRec) * (size_t)(depth + CH_SHADOW) + sizeof(XRec) * (size_t)depth + 8 * (size_t)(depth / CH_BATCH) +
                            8 * (size_t)depth + 16 + sizeof(float4) * (size_t)gw + sizeof(int32_t) * (size_t)((gw + 3) & ~3);
  const size_t usage_bytes = cap <= 255u ? (size_t)n + 1 : (cap <= 65535u ? ((size_t)n + 1) * 2 : (size_t)-1);
  if (gw >= CH_MIN_GW && n < (1 << 26) && usage_bytes != (size_t)-1 && ring_fixed + usage_bytes + 16 <= smem_limit)
    return ring_fixed + usage_bytes + 16;
  return 0;
}
// the deepest ring that fits (MOSAIC_CHAIN_DEPTH caps it: a measurement knob); 0 when not even the shallowest does
static int chain_depth(const mosaic_ctx* ctx, unsigned cap) {
  static const int limit = getenv("MOSAIC_CHAIN_DEPTH") ? atoi(getenv("MOSAIC_CHAIN_DEPTH")) : CH_DEPTH_MAX;
  for (int d = CH_DEPTH_MAX; d >= CH_DEPTH_MIN; d >>= 1)
    if (d <= std::max(limit, CH_DEPTH_MIN) && chain_smem_at(ctx, cap, d)) return d;
  return 0;
}
static size_t chain_smem(const mosaic_ctx* ctx, unsigned cap) {
  const int d = chain_depth(ctx, cap);
  return d ? chain_smem_at(ctx, cap, d) : 0;
}
bool place_chain_has_progress(const mosaic_ctx* ctx) { return use_v1_chain(ctx); }

bool place_chain_eligible(const mosaic_ctx* ctx, int max_usage) {
  if (!use_v1_chain(ctx)) return window_chain_eligible(ctx, max_usage) && !ctx->profile_chain;
  if (use_wave(ctx, max_usage)) return !ctx->profile_chain;
  return chain_smem(ctx, max_usage < 0 ? 0u : (unsigned)max_usage) != 0 && !ctx->profile_chain;
}

// the chain kernel alone, on `stream`, consuming records as `ready` (a device counter of prepared cells) advances
int launch_place_chain_async(mosaic_ctx* ctx, int max_usage, float weight, cudaStream_t stream, const unsigned* ready, unsigned* progress) {
  if (!use_v1_chain(ctx)) {
    if (progress) return fail(ctx, MOSAIC_ERR_STATE, "the windowed chain publishes no progress counter");
    return launch_window_chain(ctx, max_usage, weight, stream, ready, false);
  }
  const unsigned cap = max_usage < 0 ? 0u : (unsigned)max_usage;
  if (use_wave(ctx, max_usage)) {
    MOSAIC_TRY(ensure_place_records(ctx));
    MOSAIC_TRY(launch_wave_chain(ctx, max_usage, weight, stream, ready, progress, false));
    ctx->launches++;
    return MOSAIC_OK;
  }
  const size_t smem = chain_smem(ctx, cap);
  if (!smem) return fail(ctx, MOSAIC_ERR_STATE, "placement: the chain kernel does not cover this configuration");
  MOSAIC_TRY(ensure_place_records(ctx));
  if (cap <= 255u)
    MOSAIC_TRY((launch_chain_impl<uint8_t, false, true>(ctx, cap, weight, smem, stream, ready, progress)));
  else
    MOSAIC_TRY((launch_chain_impl<uint16_t, false, true>(ctx, cap, weight, smem, stream, ready, progress)));
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

template <typename UsageT>
static int launch_chain(mosaic_ctx* ctx, unsigned cap, float weight, size_t smem) {
  MOSAIC_TRY(launch_place_prep_range(ctx, weight, 0, ctx->gw * ctx->gh));
  return ctx->profile_chain ? launch_chain_impl<UsageT, true>(ctx, cap, weight, smem, ctx->stream, nullptr)
                            : launch_chain_impl<UsageT, false>(ctx, cap, weight, smem, ctx->stream, nullptr);
}

int read_chain_profile(mosaic_ctx* ctx, unsigned long long* out16) {
  if (!use_v1_chain(ctx)) return read_window_profile(ctx, out16);
  if (chain_choice(ctx) == 3) return read_wave_profile(ctx, out16);
  MOSAIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MOSAIC_CUDA(ctx, cudaMemcpyFromSymbol(out16, g_chain_prof, sizeof(unsigned long long) * 16));
  return MOSAIC_OK;
}

int launch_place(mosaic_ctx* ctx, int max_usage, float weight) {
  const int n = ctx->n, gw = ctx->gw, gh = ctx->gh;
  const unsigned cap = max_usage < 0 ? 0u : (unsigned)max_usage;
  const size_t smem_limit = 220 * 1024;
  // fast path: record ring + row state + usage counters in shared memory
  const size_t fast_smem = chain_smem(ctx, cap);
  StageTimer t(ctx, MOSAIC_STAGE_PLACE);
  if (use_window(ctx, max_usage)) {
    MOSAIC_TRY(launch_window_prep_range(ctx, weight, 0, gw * gh));
    return launch_window_chain(ctx, max_usage, weight, ctx->stream, nullptr, ctx->profile_chain);
  }
  if (use_wave(ctx, max_usage)) {
    MOSAIC_TRY(launch_place_prep_range(ctx, weight, 0, gw * gh));
    MOSAIC_TRY(launch_wave_chain(ctx, max_usage, weight, ctx->stream, nullptr, nullptr, ctx->profile_chain));
  } else if (use_v1_chain(ctx) && fast_smem) {
    if (cap <= 255u)
      MOSAIC_TRY(launch_chain<uint8_t>(ctx, cap, weight, fast_smem));
    else
      MOSAIC_TRY(launch_chain<uint16_t>(ctx, cap, weight, fast_smem));
  } else {
    // generic path: single warp, usage counters in shared memory when they fit, else in HBM
    const size_t row_bytes = ((size_t)gw * 4 + 15) & ~(size_t)15;
    const size_t smem_with_usage = row_bytes + (size_t)n * 4;
    if (row_bytes > smem_limit) return fail(ctx, MOSAIC_ERR_INVALID, "grid_w too large for the placement kernel");
    if (smem_with_usage <= smem_limit) {
      MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem_with_usage));
      greedy_generic_kernel<true><<<1, 32, smem_with_usage, ctx->stream>>>(
          ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->sim_lab4(), n, gw, gh, cap, weight, ctx->d_placed, nullptr,
          ctx->d_place_info, mosaic_ctx::kFallbackCap);
    } else {
      MOSAIC_TRY(ensure(ctx, &ctx->d_usage, &ctx->usage_cap, (size_t)n));
      MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)row_bytes));
      greedy_generic_kernel<false><<<1, 32, row_bytes, ctx->stream>>>(
          ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->sim_lab4(), n, gw, gh, cap, weight, ctx->d_placed, ctx->d_usage,
          ctx->d_place_info, mosaic_ctx::kFallbackCap);
    }
  }
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace mosaic
