// greedy_window.cu — K2b, second generation: the raster-order greedy placement as a WINDOWED chain.
//
// Same reference semantics as greedy.cu (main.rs:335-475, lib.rs:57-89, adjacency.rs:63-88; the bound argument is the
// one documented there: a candidate's total is at least its d2 and at most hi = fl(d2 + fl(fl(M_up + M_left) * w)), so
// the first VALID candidate wins unless a later valid one has d2 < hi of the first).  What changes is who does what:
//
//   lane = cell      A window is 32 consecutive raster cells.  One warp SCANS a window with one lane per cell: every
//                    lane walks its own candidate list (rank-major records: a rank is one conflict-free row) against the
//                    usage counters and the material above, and keeps the first four valid candidates — a shortlist of
//                    (id, remaining uses, threat bits).  The median winner has rank 1, the 99th percentile rank 37: the
//                    32-candidates-per-cell reduction of the first generation spent 90 % of its lanes on candidates
//                    that never matter; here 32 cells are scanned for the price of one.
//   lane = (cell, shortlist entry)
//                    The CHAIN runs over 8 cells x 4 entries per register set.  Per cell it is: compare the entry ids
//                    with the previous winner (left-neighbour rule / usage), select the pre-computed key variant,
//                    REDUX.MIN over the cell's four lanes (first valid entry), a parallel REDUX.OR of threat bits.
//                    Decisions are published in a shared-memory log; every in-flight shortlist is patched by comparing
//                    its ids with the decision (remaining uses - 1).  Usage counters are only written at window
//                    boundaries, so a scan sees a state that is exact for "all windows before the previous one" and
//                    listens to the previous window's decisions through the log.
//   two warps        alternate windows: while one walks the chain of window w, the other has committed its own window
//                    w-1, scans window w+1 and applies the decisions of window w as they appear.  The hand-off costs
//                    one shared-memory round trip per 32 cells.
//   slow path        a cell whose shortlist is exhausted or whose first valid entry is threatened is resolved by the
//                    whole warp with the reference's loop verbatim (exact totals, multi-chunk scan, usage reset +
//                    fallback); a reset invalidates every shortlist: scans are redone.
//
// Records arrive in a shared-memory ring by 1-D TMA bulk copies (three 4 KB rows blocks per window: ids, d2, hi of the
// first 32 ranks); deeper ranks are read from global memory by the scanning warp (off the critical path).
#include "common.cuh"

namespace mosaic {

namespace {

constexpr int WN = 32;                     // cells per window
constexpr int WV = 4;                      // shortlist entries per cell
constexpr int WROWS = kMaxK;               // ranks per record (k <= 100)
// one window record in global memory (words; every field is [rank][32 cells], cell position rotated by the rank):
//   ids of all ranks | d2 of all ranks | hi, mleft, L, a, b of ranks 0..31
//   hi    = upper bound of the candidate's total from both neighbour cells' candidate lists (state independent)
//   mleft = the left neighbour's share of that bound (the scan replaces the upper one by the exact term)
constexpr int WR_D2 = WROWS * WN;
constexpr int WR_HI = 2 * WROWS * WN;
constexpr int WREC_WORDS = WR_HI + 5 * 32 * WN;
// ring slot (words): ids of ranks 0..31 | d2 of ranks 0..32 | hi, mleft, L, a, b of ranks 0..31 — three bulk copies
constexpr int WS_D2 = 32 * WN;
constexpr int WS_HI = WS_D2 + 33 * WN;
constexpr int WS_ML = WS_HI + 32 * WN;
constexpr int WS_LAB = WS_ML + 32 * WN;
constexpr int WSLOT_BYTES = (WS_HI + 5 * 32 * WN) * 4;
#ifndef WCHAIN_UNROLL
#define WCHAIN_UNROLL 4
#endif
#ifndef WSPIN_NS
#define WSPIN_NS 40
#endif
constexpr uint32_t W_IDMASK = 0x03FFFFFFu;
constexpr uint32_t W_NOKEY = 0xFFFFFFFFu;
constexpr uint32_t W_NOID = 0xFFFFFFFFu;

// element (rank j, cell l of the window) of a field: rows are ranks, the cell position is rotated by the rank so that
// both a row (scan: 32 cells, one rank) and a column (slow path: one cell, 32 ranks) touch 32 different banks
__device__ __forceinline__ int wpos(int j, int l) { return j * WN + ((l + j) & 31); }

// ---- preparation: candidate lists -> window records ------------------------------------------------------------
// One warp per cell.  For every candidate: id, d2 and the upper bound hi of its total (see greedy.cu for the
// argument); ids and d2 of all k candidates, the bound for the 32 the shortlists are built from.
__global__ void __launch_bounds__(256) place_prep_window_kernel(const uint32_t* __restrict__ cand_idx,
                                                                const float* __restrict__ cand_d2, int k,
                                                                const float4* __restrict__ lab4, int n, int cell_begin,
                                                                int cell_end, int gw, float weight,
                                                                uint32_t* __restrict__ wrec) {
  __shared__ float4 nb_s[8][2][kMaxK];  // per warp: (Lab, id) of every candidate of the cell above / to the left
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const bool use_penalty = weight > 0.0f;
  const float inf = __int_as_float(0x7f800000);
  float4(*nb)[kMaxK] = nb_s[threadIdx.x >> 5];
  for (int c = cell_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); c < cell_end; c += warps) {
    const int w = c >> 5, l = c & 31;
    uint32_t* rec = wrec + (size_t)w * WREC_WORDS;
    const int nbc[2] = {c >= gw ? c - gw : -1, c % gw != 0 ? c - 1 : -1};  // adjacency.rs:71-87 order: Up, Left
    if (use_penalty) {
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
    }
    for (int base = 0; base < WROWS; base += 32) {
      const int j = base + lane;
      uint32_t id = (uint32_t)n;
      float d2 = inf, hi = inf, mleft = 0.0f;
      float4 lab = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < k) {
        id = __ldg(cand_idx + (size_t)c * k + j);
        d2 = __ldg(cand_d2 + (size_t)c * k + j);
        hi = d2;
        if (base == 0) lab = __ldg(&lab4[id]);
        if (use_penalty && base == 0) {  // (the shortlists are built from the first 32 ranks: deeper ones need no bound)
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
          // hi >= fl(d2 + fl(fl(t_up + t_left) * w))  (main.rs:386-398): monotone roundings
          hi = __fadd_rn(d2, __fmul_rn(__fadd_rn(m[0], m[1]), weight));
          mleft = m[1];
        }
      }
      if (j < WROWS) {
        rec[wpos(j, l)] = id;
        rec[WR_D2 + wpos(j, l)] = __float_as_uint(d2);
      }
      if (base == 0) {  // bound and colour of the 32 candidates the chain works with
        const int p = wpos(lane, l);
        rec[WR_HI + p] = __float_as_uint(hi);
        rec[WR_HI + 32 * WN + p] = __float_as_uint(mleft);
        rec[WR_HI + 2 * 32 * WN + p] = __float_as_uint(lab.x);
        rec[WR_HI + 3 * 32 * WN + p] = __float_as_uint(lab.y);
        rec[WR_HI + 4 * 32 * WN + p] = __float_as_uint(lab.z);
      }
    }
  }
}

// diagnostics of the last PROFILED launch (mosaic_debug_chain_profile): see read_window_profile
__device__ unsigned long long g_window_prof[16];

struct WinCtl {  // shared-memory control block
  volatile unsigned unused0;
  volatile unsigned scan_done[2];  // per warp: 1 + index of the last window whose scan is complete
  volatile unsigned committed;     // windows [0, committed) are in the usage counters
  volatile unsigned reset_epoch;   // bumped by every usage reset (fallback_tile_selection)
  volatile unsigned fallbacks;     // cells at which fallback_tile_selection fired so far
  unsigned pad[2];
};

// everything the out-of-line parts (scan, slow path) need; lives in local memory, touched off the hot loop only
struct WinEnv {
  const uint32_t* ring;   // shared: nslot x (ids | d2 | hi) of the first 32 ranks
  const uint32_t* wrec;   // global: all ranks
  unsigned long long* bars;
  WinCtl* ctl;
  uint2* stage;           // this warp's 32 x 4 entries
  float4* uplab;          // this warp's window: colour of the material above each cell (slow path)
  volatile uint32_t* log;  // decision of cell c at [c & lmask]: material id | phase bit 31 = (c >> lshift) & 1
  unsigned* usage_w;
  const float4* lab4;
  uint32_t* info;
  int k, n, gw, cells, nwin, nslot, warp, lane;
  unsigned cap, lmask, lshift, fallback_cap;
  float weight;
  int commit_from;        // decisions [commit_from, ...) of the current window are not in the counters yet
  unsigned long long prof[12];
};

// conditional decrement in two instructions (ISETP + predicated IADD): one decision reaching one shortlist entry
__device__ __forceinline__ void dec_if_eq(int& sl, uint32_t id, uint32_t chosen) {
  asm("{\n.reg .pred p;\nsetp.eq.u32 p, %1, %2;\n@p add.s32 %0, %0, -1;\n}" : "+r"(sl) : "r"(id), "r"(chosen));
}

// Every wait of one warp for the other goes through here.  The sleep matters: a warp that polls shared memory back to
// back slows the OTHER warp's chain several times over (measured: 200 -> see DESIGN cycles per cell) — the two share
// the SM's memory pipeline.  A hand-off that never comes is a bug: fail loudly instead of hanging the GPU.
__device__ __forceinline__ void spin_guard(long long t0) {
  __nanosleep(WSPIN_NS);
  if (clock64() - t0 > 4000000000ll) __trap();
}

template <typename UsageT>
__device__ __forceinline__ void usage_add(unsigned* usage_w, uint32_t id) {  // concurrent adds of a warp are atomic
  constexpr int USHIFT = sizeof(UsageT) == 1 ? 2 : 1;
  constexpr int UBITS = sizeof(UsageT) * 8;
  atomicAdd(&usage_w[id >> USHIFT], 1u << ((id & ((1u << USHIFT) - 1u)) * UBITS));
}

// ---- SCAN: lanes = cells [32w + from, 32w + 32) -> shortlists (first four valid candidates) in this warp's stage area.
// The counters must not change while this runs (scan_done / committed protocol in the kernel).
template <typename UsageT, bool PROF>
__device__ __noinline__ void scan_window(WinEnv& e, int w, int from) {
  long long ts = 0;
  if (PROF) ts = clock64();
  const int lane = e.lane;
  const UsageT* usage = reinterpret_cast<const UsageT*>(e.usage_w);
  const int base = w * WN;
  const int c = base + lane;
  const bool in = c < e.cells && lane >= from;
  // the material above: decided long ago — unless the grid is narrower than two windows and this scan runs while the
  // window before is still being walked; then the slot still holds the cell one ring length earlier (phase bit) and the
  // rule is applied when that decision is listened to (kernel, listen loop)
  int32_t up = -1;
  if (in && c >= e.gw) {
    const unsigned uc = (unsigned)(c - e.gw);
    const uint32_t v = e.log[uc & e.lmask];
    if ((v >> 31) == ((uc >> e.lshift) & 1u)) up = (int32_t)(v & W_IDMASK);
  }
  // its colour, for the cells of this window that end in the slow path (one gather per lane, in flight during the scan)
  float4 ul = make_float4(0.f, 0.f, 0.f, 0.f);
  if (up >= 0 && e.weight > 0.0f) ul = __ldg(&e.lab4[up]);
  {
    unsigned long long* bar = &e.bars[w % e.nslot];
    const unsigned parity = (unsigned)((w / e.nslot) & 1);
    if (!mbar_test(bar, parity)) {
      long long tw = 0;
      if (PROF) tw = clock64();
      mbar_wait(bar, parity);
      if (PROF) e.prof[6] += (unsigned long long)(clock64() - tw);
    }
  }
  const uint32_t* rs = e.ring + (size_t)(w % e.nslot) * (WSLOT_BYTES / 4);
  const float inf = __int_as_float(0x7f800000);
  const unsigned cap = e.cap;
  const uint32_t nn = (uint32_t)e.n;
  const int k = e.k;
  const int k32 = min(k, 32);
  // pass 1: which of the first 32 ranks are valid for this cell (lib.rs:71-74, main.rs:375, 380 (Up)) — one bit per rank.
  // Eight rows per step, loads first.  The first 32 ranks only (they are in the ring): a cell whose candidates behind
  // rank 31 matter ends in the slow path, which reads them itself.
  uint32_t vm = 0u;
  for (int jb = 0; jb < k32; jb += 8) {
    uint32_t bi[8];
    unsigned bu[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) bi[t] = min(rs[wpos(min(jb + t, k32 - 1), lane)], nn);  // (lanes outside the grid read anything)
#pragma unroll
    for (int t = 0; t < 8; ++t) bu[t] = (unsigned)usage[bi[t]];
#pragma unroll
    for (int t = 0; t < 8; ++t)
      vm |= (jb + t < k32 && bu[t] < cap && (int32_t)bi[t] != up) ? (1u << (jb + t)) : 0u;
    if (__all_sync(0xffffffffu, !in || __popc(vm) > WV)) break;
  }
  if (!in) vm = 0u;
  // pass 2: the first four valid ranks become the shortlist.  Their totals are bracketed with the EXACT term of the
  // material above (it is known here; adjacency.rs:71-87 adds Up before Left):
  //   lo = fl(d2 + fl(t_up * w))  <=  total = fl(d2 + fl(fl(t_up + t_left) * w))  <=  hi = fl(d2 + fl(fl(t_up + mleft) * w))
  // (every rounding is monotone, 0 <= t_left <= mleft).  A cell whose material above is not decided yet (narrow grids)
  // keeps the state-independent bound of the preparation kernel.
  uint32_t e_id[WV];
  int e_sl[WV];
  float e_lo[WV], e_hi[WV];
  const float weight = e.weight;
  const bool exact_up = weight > 0.0f && (up >= 0 || c < e.gw);
#pragma unroll
  for (int v = 0; v < WV; ++v) {
    const bool have = vm != 0u;
    const int j = have ? __ffs(vm) - 1 : 0;
    vm &= vm - 1u;
    const int p = wpos(j, lane);
    const uint32_t id = min(rs[p], nn);
    const float d2 = __uint_as_float(rs[WS_D2 + p]);
    float lo = d2, hi = __uint_as_float(rs[WS_HI + p]);
    if (exact_up) {
      float t_up = 0.0f;
      bool oor = false;
      if (up >= 0)
        t_up = adjacency_term_fast(__uint_as_float(rs[WS_LAB + p]), __uint_as_float(rs[WS_LAB + 32 * WN + p]),
                                   __uint_as_float(rs[WS_LAB + 2 * 32 * WN + p]), ul.x, ul.y, ul.z, oor);
      if (!oor) {  // (outside the fast arithmetic's domain: keep the state-independent bound)
        lo = __fadd_rn(d2, __fmul_rn(t_up, weight));
        hi = __fadd_rn(d2, __fmul_rn(__fadd_rn(t_up, __uint_as_float(rs[WS_ML + p])), weight));
      }
    }
    e_id[v] = have ? id : nn;
    e_sl[v] = have ? (int)(cap - (unsigned)usage[id]) : 0;
    e_lo[v] = have ? lo : inf;
    e_hi[v] = have ? hi : inf;
  }
  // d2 of the first candidate behind the shortlist: the fifth valid rank if one was seen, else whatever follows the
  // scanned ranks (rank 32, or nothing)
  float beyond = inf;
  if (in) {
    if (vm != 0u)
      beyond = __uint_as_float(rs[WS_D2 + wpos(__ffs(vm) - 1, lane)]);
    else if (k > 32)
      beyond = __uint_as_float(rs[WS_D2 + wpos(32, lane)]);
  }
  // threat bits (key word bits 26..29): bit u < v — entry v can be STRICTLY better than the earlier entry u (lo[v] <
  // hi[u]); bit v itself — so can the first candidate behind the shortlist (its d2 < hi[v]: the shortlist cannot decide)
  uint2 packed[WV];
#pragma unroll
  for (int v = 0; v < WV; ++v) {
    uint32_t tm = (beyond < e_hi[v]) ? (1u << v) : 0u;
#pragma unroll
    for (int u = 0; u < v; ++u) tm |= (e_lo[v] < e_hi[u]) ? (1u << u) : 0u;
    packed[v] = make_uint2(((uint32_t)v << 30) | (tm << 26) | e_id[v], (uint32_t)e_sl[v]);
  }
  uint4* st4 = reinterpret_cast<uint4*>(e.stage + (size_t)lane * WV);
  st4[0] = make_uint4(packed[0].x, packed[0].y, packed[1].x, packed[1].y);
  st4[1] = make_uint4(packed[2].x, packed[2].y, packed[3].x, packed[3].y);
  if (lane >= from) e.uplab[lane] = ul;
  __syncwarp();
  if (PROF) e.prof[4] += (unsigned long long)(clock64() - ts);
}

// ---- the counters hold every window before this one (the other warp puts the previous one in right after walking it)
__device__ __forceinline__ void wait_previous_committed(WinEnv& e, int w) {
  const long long t0 = clock64();
  while (e.ctl->committed < (unsigned)w) spin_guard(t0);
}

// ---- window end (or a rescan in the middle): the decisions [commit_from, d) of this window go into the counters.  The
// other warp's scan reads them: it must be done (scan_done) before they change.
template <typename UsageT>
__device__ __forceinline__ void commit_decisions(WinEnv& e, int w, int d) {
  const long long t0 = clock64();
  if (w + 1 < e.nwin)
    while (e.ctl->scan_done[e.warp ^ 1] < (unsigned)(w + 2)) spin_guard(t0);
  wait_previous_committed(e, w);
  if (e.lane >= e.commit_from && e.lane < d) usage_add<UsageT>(e.usage_w, e.log[(unsigned)(w * WN + e.lane) & e.lmask] & W_IDMASK);
  e.commit_from = max(e.commit_from, d);
  __syncwarp();
}

// ---- SLOW PATH: the reference's loop, bit for bit (main.rs:335-475), for cell c = 32w + d.  Returns the material id;
// bit 31: the usage counters were reset (fallback_tile_selection) — every shortlist is stale.
// Exact use counts without touching the counters: counter (all windows before this one) + this window's decisions so
// far, which sit in the log.
template <typename UsageT, bool PROF>
__device__ __noinline__ uint32_t resolve_slow(WinEnv& e, int w, int d, uint32_t left_id) {
  long long ts = 0;
  if (PROF) ts = clock64();
  wait_previous_committed(e, w);
  const int lane = e.lane, k = e.k, n = e.n;
  const UsageT* usage = reinterpret_cast<const UsageT*>(e.usage_w);
  const unsigned cap = e.cap;
  const float weight = e.weight;
  const bool use_penalty = weight > 0.0f;  // main.rs:355-362
  const int base = w * WN;
  const int c = base + d;
  const int x = c % e.gw;
  const int32_t left = (x == 0 || left_id == W_NOID) ? -1 : (int32_t)left_id;
  const int32_t up = c >= e.gw ? (int32_t)(e.log[(unsigned)(c - e.gw) & e.lmask] & W_IDMASK) : -1;
  const uint32_t* rs = e.ring + (size_t)(w % e.nslot) * (WSLOT_BYTES / 4);
  const uint32_t* rg = e.wrec + (size_t)w * WREC_WORDS;
  const float inf = __int_as_float(0x7f800000);
  // this window's decisions that are not in the counters yet: lane t holds decision commit_from + t
  const int npend = d - e.commit_from;
  const uint32_t pend = lane < npend ? (e.log[(unsigned)(base + e.commit_from + lane) & e.lmask] & W_IDMASK) : W_NOID;
  auto uses = [&](uint32_t id) -> unsigned {  // uniform control flow: every lane asks about its own candidate
    unsigned u = (unsigned)usage[min(id, (uint32_t)n)];
    for (int t = 0; t < npend; ++t) u += (__shfl_sync(0xffffffffu, pend, t) == id) ? 1u : 0u;
    return u;
  };
  // colours: the candidates' first 32 from the ring, the material above from the scan's prefetch, the left one from
  // the ring too when it is one of the previous cell's first 32 candidates
  float4 up_lab = make_float4(0.f, 0.f, 0.f, 0.f), left_lab = up_lab;
  if (use_penalty) {
    // (a grid narrower than two windows is scanned before the row above is complete: no prefetch to rely on)
    if (up >= 0) up_lab = e.gw < 2 * WN ? __ldg(&e.lab4[up]) : e.uplab[d];
    if (left >= 0) {
      bool found = false;
      if (d > 0) {
        const int p = wpos(lane, d - 1);
        const unsigned m = __ballot_sync(0xffffffffu, lane < k && rs[p] == (uint32_t)left);
        if (m) {
          const int q = wpos(__ffs(m) - 1, d - 1);
          left_lab = make_float4(__uint_as_float(rs[WS_LAB + q]), __uint_as_float(rs[WS_LAB + 32 * WN + q]),
                                 __uint_as_float(rs[WS_LAB + 2 * 32 * WN + q]), 0.f);
          found = true;
        }
      }
      if (!found) left_lab = __ldg(&e.lab4[left]);
    }
  }
  unsigned best_bits = 0xFFFFFFFFu;
  int32_t best_id = -1;
  uint32_t first_ok = W_NOID;  // first candidate the fallback may take (main.rs:441-453)
  uint32_t first_id = W_NOID;
  for (int ch = 0; ch * 32 < k; ++ch) {
    const int j = ch * 32 + lane;
    uint32_t gid = (uint32_t)n;
    float gd2 = inf;
    float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < k) {
      const int p = wpos(j, d);
      if (ch == 0) {
        gid = rs[p];
        gd2 = __uint_as_float(rs[WS_D2 + p]);
        gl = make_float4(__uint_as_float(rs[WS_LAB + p]), __uint_as_float(rs[WS_LAB + 32 * WN + p]),
                         __uint_as_float(rs[WS_LAB + 2 * 32 * WN + p]), 0.f);
      } else {
        gid = __ldg(rg + p);
        gd2 = __uint_as_float(__ldg(rg + WR_D2 + p));
      }
    }
    if (ch > 0 && best_id >= 0 && __float_as_uint(__shfl_sync(0xffffffffu, gd2, 0)) >= best_bits) break;  // exact early exit
    if (ch > 0 && use_penalty && j < k) gl = __ldg(&e.lab4[gid]);
    if (ch == 0) first_id = __shfl_sync(0xffffffffu, gid, 0);
    const bool placeable = j < k && (int32_t)gid != left && (int32_t)gid != up;  // main.rs:309-333 (Right/Down are empty)
    if (first_ok == W_NOID) {
      const unsigned m = __ballot_sync(0xffffffffu, placeable);
      if (m) first_ok = __shfl_sync(0xffffffffu, gid, __ffs(m) - 1);
    }
    const unsigned u = uses(gid);
    const bool ok = placeable && u < cap;  // lib.rs:71-74, main.rs:375, 380
    float pen = 0.0f;
    if (use_penalty) {  // adjacency.rs:71-87: Up, (Down), Left, (Right); absent terms are +0.0 (x + 0.0 == x)
      bool oor_u = false, oor_l = false;  // branch-free IEEE sequences (common.cuh); outside their domain: the builtins
      float t_up = adjacency_term_fast(gl.x, gl.y, gl.z, up_lab.x, up_lab.y, up_lab.z, oor_u);
      float t_left = adjacency_term_fast(gl.x, gl.y, gl.z, left_lab.x, left_lab.y, left_lab.z, oor_l);
      if (__any_sync(0xffffffffu, (oor_u && up >= 0) || (oor_l && left >= 0))) {
        t_up = adjacency_term(lab_distance_rn(gl.x, gl.y, gl.z, up_lab.x, up_lab.y, up_lab.z));
        t_left = adjacency_term(lab_distance_rn(gl.x, gl.y, gl.z, left_lab.x, left_lab.y, left_lab.z));
      }
      t_up = up >= 0 ? t_up : 0.0f;
      t_left = left >= 0 ? t_left : 0.0f;
      pen = __fmul_rn(__fadd_rn(t_up, t_left), weight);
    }
    const unsigned tb = ok ? __float_as_uint(__fadd_rn(gd2, pen)) : 0xFFFFFFFFu;  // main.rs:398
    const unsigned mn = __reduce_min_sync(0xffffffffu, tb);
    if (mn < best_bits) {  // strict <: earlier chunks / earlier lanes win ties (main.rs:401-407)
      const int who = __ffs(__ballot_sync(0xffffffffu, tb == mn)) - 1;
      best_bits = mn;
      best_id = (int32_t)__shfl_sync(0xffffffffu, gid, who);
    }
  }
  bool was_reset = false;
  if (best_id < 0) {
    // fallback_tile_selection (main.rs:420-475): reset ALL usage counts, then the first placeable candidate, else the
    // nearest one.  (The early exit above never fires without a best: every chunk has been looked at.)
    if (lane == 0) {  // (the chain is sequential: whichever warp walks it appends to one list, in raster order)
      const unsigned nf = e.ctl->fallbacks;
      if (nf < e.fallback_cap) e.info[1 + nf] = (uint32_t)c;
      e.ctl->fallbacks = nf + 1u;
    }
    constexpr int USHIFT = sizeof(UsageT) == 1 ? 2 : 1;
    for (int i = lane; i < ((n + 1) >> USHIFT) + 1; i += 32) e.usage_w[i] = 0u;  // lib.rs:86-88
    __syncwarp();
    best_id = (int32_t)(first_ok != W_NOID ? first_ok : first_id);  // main.rs:441-453, 456-474 (nearest_one)
    e.commit_from = d;  // what this window decided before is wiped with the rest; this decision is the first one pending
    was_reset = true;
    if (lane == 0) {
      __threadfence_block();
      e.ctl->reset_epoch = e.ctl->reset_epoch + 1u;
    }
    __syncwarp();
  }
  if (PROF) {
    e.prof[was_reset ? 3 : 2] += (unsigned long long)(clock64() - ts);
    e.prof[was_reset ? 9 : 8] += 1ull;
  }
  return (uint32_t)best_id | (was_reset ? 0x80000000u : 0u);
}

template <typename UsageT, bool PROF, bool GATED>
__global__ void __launch_bounds__(64, 1)
    greedy_window_kernel(const uint32_t* __restrict__ wrec, int k, const float4* __restrict__ lab4, int n, int gw, int gh,
                         unsigned cap, float weight, int32_t* __restrict__ placed, uint32_t* __restrict__ info,
                         uint32_t fallback_cap, const volatile unsigned* ready, int nslot, unsigned lshift) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const unsigned lmask = (1u << lshift) - 1u;
  // layout: ring slots | mbarriers | control | stage | log | usage
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)nslot * WSLOT_BYTES);
  WinCtl* ctl = reinterpret_cast<WinCtl*>(bars + 8);
  uint2* stage_all = reinterpret_cast<uint2*>(ctl + 1);  // [2 warps][32 cells][4 entries]
  float4* uplab_all = reinterpret_cast<float4*>(stage_all + 2 * WN * WV);  // [2 warps][32 cells]
  volatile uint32_t* log = reinterpret_cast<volatile uint32_t*>(uplab_all + 2 * WN);  // chosen id per cell (ring)
  unsigned* usage_w = const_cast<unsigned*>(reinterpret_cast<volatile unsigned*>(log + (lmask + 1)));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cells = gw * gh;
  const int nwin = (cells + WN - 1) / WN;
  constexpr int USHIFT = sizeof(UsageT) == 1 ? 2 : 1;  // counters per 32-bit word: 4 (u8) or 2 (u16)

  for (int i = tid; i < ((n + 1) >> USHIFT) + 1; i += blockDim.x) usage_w[i] = 0u;
  for (int i = tid; i <= (int)lmask; i += blockDim.x) log[i] = W_NOID;
  if (tid == 0) {
    ctl->unused0 = 0u;
    ctl->scan_done[0] = ctl->scan_done[1] = 0u;
    ctl->committed = 0u;
    ctl->reset_epoch = 0u;
    ctl->fallbacks = 0u;
    for (int b = 0; b < nslot; ++b) mbar_init(&bars[b], 1);
    mbar_fence_init();
  }
  __syncthreads();

  WinEnv e;
  e.ring = ring;
  e.wrec = wrec;
  e.bars = bars;
  e.ctl = ctl;
  e.stage = stage_all + (size_t)warp * WN * WV;
  e.uplab = uplab_all + (size_t)warp * WN;
  e.log = log;
  e.usage_w = usage_w;
  e.lab4 = lab4;
  e.info = info;
  e.k = k;
  e.n = n;
  e.gw = gw;
  e.cells = cells;
  e.nwin = nwin;
  e.nslot = nslot;
  e.warp = warp;
  e.lane = lane;
  e.cap = cap;
  e.lmask = lmask;
  e.lshift = lshift;
  e.fallback_cap = fallback_cap;
  e.weight = weight;
  e.commit_from = 0;
  for (int i = 0; i < 12; ++i) e.prof[i] = 0ull;
  long long t_all = 0;
  if (PROF) t_all = clock64();

  // ---- records: window u -> ring slot u % nslot (one elected lane issues; the slot must be free)
  auto issue_window = [&](int u) {
    if (u >= nwin) return;
    if (GATED) {  // the records may still be in production: `ready` counts the cells whose records exist
      const unsigned need = (unsigned)min(cells, (u + 1) * WN);
      const long long t0 = clock64();
      while (*ready < need) spin_guard(t0);
      __threadfence();
    }
    const int slot = u % nslot;
    unsigned long long* bar = &bars[slot];
    uint32_t* dst = ring + (size_t)slot * (WSLOT_BYTES / 4);
    const uint32_t* src = wrec + (size_t)u * WREC_WORDS;
    mbar_expect_tx(bar, (unsigned)WSLOT_BYTES);
    tma_load_1d(dst, src, 32 * WN * 4, bar);
    tma_load_1d(dst + WS_D2, src + WR_D2, 33 * WN * 4, bar);  // (row 32: the distance the shortlists cannot see past)
    tma_load_1d(dst + WS_HI, src + WR_HI, 5 * 32 * WN * 4, bar);
  };
  if (tid == 0)
    for (int u = 0; u < nslot; ++u) issue_window(u);

  // ---- per-warp state: four register sets of (cell, entry) lanes
  uint32_t W0[4], ID[4];  // key word: entry << 30 | beyond << 29 | threat bits << 26 | id; the id alone
  int SL[4];              // remaining uses of the entry's material (<= 0: not a candidate any more)
  bool C0[4];             // the lane's cell is in column 0 (no left neighbour, main.rs:309-333)
  const int grp = lane >> 2, ent = lane & 3;
  unsigned epoch_seen = 0;  // reset epoch when the other warp started the window before this warp's next one

  // shortlists of window w from the stage area -> register sets.  `prev` != W_NOID: the counters the scan saw already
  // hold the decision before the first cell to decide; the chain applies that decision itself (left-neighbour rule +
  // one use), so it is taken out again here.
  auto load_sets = [&](int w, uint32_t prev) {
    const int x0 = (w * WN) % gw;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint2 en = e.stage[(size_t)(8 * s + grp) * WV + ent];
      W0[s] = en.x;
      ID[s] = en.x & W_IDMASK;
      SL[s] = (int)en.y + ((en.x & W_IDMASK) == prev ? 1 : 0);
      int x = x0 + 8 * s + grp;
      x = x >= gw ? x - gw : x;
      C0[s] = x == 0;
    }
    __syncwarp();
  };

  // one store publishes a decision: the phase bit tells a reader whether the slot already holds cell c or still the
  // cell one ring length before it
  auto publish = [&](int c, uint32_t chosen) {
    if (lane == 0) log[(unsigned)c & lmask] = chosen | ((((unsigned)c >> lshift) & 1u) << 31);
  };
  auto wait_decision = [&](unsigned c, long long t0) -> uint32_t {
    for (;;) {
      const uint32_t v = log[c & lmask];
      if ((v >> 31) == ((c >> lshift) & 1u)) return v & W_IDMASK;
      spin_guard(t0);
    }
  };

  // ---- the windows of this warp
  for (int w = warp; w < nwin; w += 2) {
    const int base = w * WN;
    const int wcells = min(WN, cells - base);
    // scan under the counters as they are (windows < w-1), then listen to window w-1 through the log
    scan_window<UsageT, PROF>(e, w, 0);
    if (lane == 0) {
      __threadfence_block();
      ctl->scan_done[warp] = (unsigned)(w + 1);
    }
    load_sets(w, W_NOID);
    uint32_t chosen = W_NOID;  // the decision before the next cell to decide
    if (w > 0) {
      long long tl = 0;
      if (PROF) tl = clock64();
      // every decision of the previous window but the last one goes into the shortlists here; the last one is the
      // `chosen` the chain starts from
      const long long t0 = clock64();
      const bool late_up = gw < 2 * WN;  // cells of this window sit below cells of the previous one
      for (unsigned seen = (unsigned)(base - WN); seen < (unsigned)(base - 1); ++seen) {
        const uint32_t ch = wait_decision(seen, t0);
#pragma unroll
        for (int s = 0; s < 4; ++s) dec_if_eq(SL[s], ID[s], ch);
        if (late_up) {  // main.rs:380: the cell below `seen` may not take its material
#pragma unroll
          for (int s = 0; s < 4; ++s)
            if (base + 8 * s + grp == (int)seen + gw && ID[s] == ch) SL[s] = 0;
        }
      }
      chosen = wait_decision((unsigned)(base - 1), t0);
      if (PROF) e.prof[5] += (unsigned long long)(clock64() - tl);
      if (ctl->reset_epoch != epoch_seen) {
        // the usage counters were reset while the previous window was walked: every shortlist is stale.  Wait until
        // that window is in the counters, then scan again (nothing left to listen to).
        while (ctl->committed < (unsigned)w) spin_guard(t0);
        scan_window<UsageT, PROF>(e, w, 0);
        load_sets(w, chosen);
        if (PROF) e.prof[10] += 1ull;
      }
    }
    // a window never wraps in the log ring (its length is a multiple of 32): pointer and phase bit of the whole window
    volatile uint32_t* const wlog = log + ((unsigned)base & lmask);
    const uint32_t wphase = (((unsigned)base >> lshift) & 1u) << 31;
    int from = 0;  // first cell of the window that still has to be decided
    while (from < wcells) {
      bool rescan = false;
      if ((from & 7) == 0 && from + 8 <= wcells) {
        // whole sets of 8 cells through the unrolled chain; the set to be processed sits at index 0
        for (int r = 0; r < (from >> 3); ++r) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            W0[s] = W0[s + 1];
            ID[s] = ID[s + 1];
            SL[s] = SL[s + 1];
            C0[s] = C0[s + 1];
          }
        }
        for (; from + 8 <= wcells && !rescan;) {
          // INVARIANT at the top of a step for cell c: SL holds every decision up to c-2, `chosen` is decision c-1.
          long long tset = 0;
          unsigned long long slow0 = 0;
          if (PROF) {
            tset = clock64();
            slow0 = e.prof[2] + e.prof[3];
          }
          uint32_t KN, KH, TN, TH;
          auto variants = [&](int g) {
            const bool ing = grp == g;
            const bool vn = ing && SL[0] > 0;           // decision c-1 is some other material
            const bool vh = ing && SL[0] > 1 && C0[0];  // decision c-1 is this material: one use less, and only a cell in
                                                        // column 0 may keep it (main.rs:380: it is the left neighbour)
            KN = vn ? W0[0] : W_NOKEY;
            KH = vh ? W0[0] : W_NOKEY;
            const uint32_t tm = (W0[0] >> 26) & 15u;
            TN = vn ? tm : 0u;
            TH = vh ? tm : 0u;
          };
          variants(0);
          // WCHAIN_UNROLL steps per loop iteration: the unrolled steps must stay inside the L0 instruction cache of the
          // SM sub-partition (~6 KB), and a taken branch is expensive for a lone warp
          for (int g0 = 0; g0 < 8 && !rescan; g0 += WCHAIN_UNROLL) {
#pragma unroll
          for (int gt = 0; gt < WCHAIN_UNROLL; ++gt) {
            const int g = g0 + gt;
            const int d = from + g;
            // on the chain: compare -> select -> REDUX.MIN
            const bool hit = ID[0] == chosen;
            uint32_t r = __reduce_min_sync(0xffffffffu, hit ? KH : KN);
            const uint32_t T = __reduce_or_sync(0xffffffffu, hit ? TH : TN);
            // off the chain: decision c-1 reaches every shortlist in flight, then the key variants of cell c+1
#pragma unroll
            for (int s = 0; s < 4; ++s) dec_if_eq(SL[s], ID[s], chosen);
            variants(g + 1);  // (g + 1 == 8: no lane is in that group — the next set computes its own)
            // bit v of T: some valid entry (or whatever follows the shortlist) can beat entry v; no valid entry: W_NOKEY
            const bool slow = r == W_NOKEY || ((T >> (r >> 30)) & 1u) != 0u;
            if (slow) {
              r = resolve_slow<UsageT, PROF>(e, w, d, chosen);
              rescan = (r & 0x80000000u) != 0u;
            } else if (PROF) {
              e.prof[7] += 1ull;
            }
            chosen = r & W_IDMASK;
            if (lane == 0) wlog[d] = chosen | wphase;  // publish (one store: see `publish`)
            if (rescan) {
              from = d + 1;
              break;
            }
          }
          }
          if (PROF) {  // cycles of this set's fast steps (slow-path time taken out), number of sets
            e.prof[0] += (unsigned long long)(clock64() - tset) - (e.prof[2] + e.prof[3] - slow0);
            e.prof[1] += 1ull;
          }
          if (!rescan) {
            from += 8;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              W0[s] = W0[s + 1];
              ID[s] = ID[s + 1];
              SL[s] = SL[s + 1];
              C0[s] = C0[s + 1];
            }
          }
        }
      } else {
        // a partial set (end of the grid, or the cells of a set behind a reset): the slow path decides
        const int stop = min(wcells, (from | 7) + 1);
        while (from < stop) {
          const uint32_t r = resolve_slow<UsageT, PROF>(e, w, from, chosen);
          rescan = rescan || (r & 0x80000000u) != 0u;
          chosen = r & W_IDMASK;
          publish(base + from, chosen);
          ++from;
        }
        rescan = true;  // (the sets behind it were patched with nothing while the slow path walked: scan them again)
      }
      if (rescan && from < wcells) {
        if ((from & 7) != 0) continue;  // finish the split set the slow way first
        // a scan reads the counters: everything before cell `from` goes in first
        commit_decisions<UsageT>(e, w, from);
        scan_window<UsageT, PROF>(e, w, from);
        load_sets(w, chosen);
        if (PROF) e.prof[10] += 1ull;
      }
    }
    epoch_seen = ctl->reset_epoch;  // the other warp walks the next window from here on: resets from now on are its own
    // ---- window done: placements out, decisions into the counters (once the other warp's scan no longer reads them)
    __syncwarp();
    if (lane < wcells) placed[base + lane] = (int32_t)(log[(unsigned)(base + lane) & lmask] & W_IDMASK);  // main.rs:412
    commit_decisions<UsageT>(e, w, wcells);
    e.commit_from = 0;
    if (lane == 0) {
      __threadfence_block();
      ctl->committed = (unsigned)(w + 1);
      issue_window(w + nslot);  // this window's ring slot is free
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0) info[0] = ctl->fallbacks;
  if (PROF && lane == 0) {
    e.prof[11] = (unsigned long long)(clock64() - t_all);
    for (int i = 0; i < 12; ++i) atomicAdd(&g_window_prof[i], e.prof[i]);
  }
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------------------------------
static size_t window_rec_bytes(const mosaic_ctx* ctx) {
  const size_t cells = (size_t)ctx->gw * ctx->gh;
  return ((cells + WN - 1) / WN) * (size_t)WREC_WORDS * 4;
}

// shared-memory plan; 0 = this configuration needs the generic kernel
static size_t window_smem(const mosaic_ctx* ctx, unsigned cap, int* nslot_out, unsigned* lshift_out) {
  const size_t limit = 220 * 1024;
  const int n = ctx->n, gw = ctx->gw;
  if (gw <= WN || n >= (1 << 26)) return 0;  // (the material above must be decided before its window is walked)
  if (cap > 65535u) return 0;
  const size_t usage_bytes = (cap <= 255u ? ((size_t)n + 1 + 8) : ((size_t)n + 1 + 4) * 2 + 8) & ~(size_t)3;
  unsigned l = 64, lshift = 6;
  while (l < (unsigned)gw + 2 * WN) l <<= 1, ++lshift;
  const size_t fixed = 8 * 8 + sizeof(WinCtl) + sizeof(uint2) * 2 * WN * WV + sizeof(float4) * 2 * WN + (size_t)l * 4 + usage_bytes + 16;
  for (int nslot = 6; nslot >= 3; --nslot) {
    const size_t total = (size_t)nslot * WSLOT_BYTES + fixed;
    if (total <= limit) {
      if (nslot_out) *nslot_out = nslot;
      if (lshift_out) *lshift_out = lshift;
      return total;
    }
  }
  return 0;
}

bool window_chain_eligible(const mosaic_ctx* ctx, int max_usage) {
  return window_smem(ctx, max_usage < 0 ? 0u : (unsigned)max_usage, nullptr, nullptr) != 0;
}

int launch_window_prep_range(mosaic_ctx* ctx, float weight, int cell_begin, int cell_end) {
  if (cell_end <= cell_begin) return MOSAIC_OK;
  MOSAIC_TRY(ensure(ctx, &ctx->d_place_rec, &ctx->place_rec_cap, window_rec_bytes(ctx)));
  int blocks = (cell_end - cell_begin + 7) / 8;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  place_prep_window_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_cand_idx, ctx->d_cand_d2, ctx->k, ctx->sim_lab4(), ctx->n,
                                                            cell_begin, cell_end, ctx->gw, weight,
                                                            reinterpret_cast<uint32_t*>(ctx->d_place_rec));
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

template <typename UsageT, bool PROF, bool GATED>
static int launch_window_impl(mosaic_ctx* ctx, unsigned cap, float weight, cudaStream_t stream, const unsigned* ready) {
  int nslot = 0;
  unsigned lshift = 0;
  const size_t smem = window_smem(ctx, cap, &nslot, &lshift);
  if (!smem) return fail(ctx, MOSAIC_ERR_STATE, "placement: the window kernel does not cover this configuration");
  MOSAIC_CUDA(ctx, cudaFuncSetAttribute(greedy_window_kernel<UsageT, PROF, GATED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (PROF) {
    unsigned long long zero[16] = {};
    MOSAIC_CUDA(ctx, cudaMemcpyToSymbolAsync(g_window_prof, zero, sizeof zero, 0, cudaMemcpyHostToDevice, stream));
  }
  greedy_window_kernel<UsageT, PROF, GATED><<<1, 64, smem, stream>>>(
      reinterpret_cast<const uint32_t*>(ctx->d_place_rec), ctx->k, ctx->sim_lab4(), ctx->n, ctx->gw, ctx->gh, cap, weight,
      ctx->d_placed, ctx->d_place_info, mosaic_ctx::kFallbackCap, ready, nslot, lshift);
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

// the chain alone, on `stream`; `ready` (may be NULL) = device counter of cells whose records exist
int launch_window_chain(mosaic_ctx* ctx, int max_usage, float weight, cudaStream_t stream, const unsigned* ready, bool prof) {
  const unsigned cap = max_usage < 0 ? 0u : (unsigned)max_usage;
  if (ready) {
    if (cap <= 255u) return launch_window_impl<uint8_t, false, true>(ctx, cap, weight, stream, ready);
    return launch_window_impl<uint16_t, false, true>(ctx, cap, weight, stream, ready);
  }
  if (prof) {
    if (cap <= 255u) return launch_window_impl<uint8_t, true, false>(ctx, cap, weight, stream, nullptr);
    return launch_window_impl<uint16_t, true, false>(ctx, cap, weight, stream, nullptr);
  }
  if (cap <= 255u) return launch_window_impl<uint8_t, false, false>(ctx, cap, weight, stream, nullptr);
  return launch_window_impl<uint16_t, false, false>(ctx, cap, weight, stream, nullptr);
}

int read_window_profile(mosaic_ctx* ctx, unsigned long long* out16) {
  MOSAIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MOSAIC_CUDA(ctx, cudaMemcpyFromSymbol(out16, g_window_prof, sizeof(unsigned long long) * 16));
  return MOSAIC_OK;
}

}  // namespace mosaic
