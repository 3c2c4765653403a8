// greedy_wave.cu — K2b, third generation: the raster-order greedy placement as a WINDOW ITERATED TO ITS FIXED POINT.
//
// Same reference semantics as greedy.cu (main.rs:335-475, lib.rs:57-89, adjacency.rs:63-88) and the same records
// (place_prep_kernel: candidate words, threat windows, d2 + Lab of the first 32 candidates).  What changes is the shape of
// the computation.  The reference's loop is a chain — cell i needs the usage counters and the Left / Up materials that
// cells 0..i-1 left behind — and the first two generations walk it with one warp, one cell after the other.  Here a
// window of W consecutive cells is evaluated AT ONCE, one warp per cell, every sweep:
//
//   sweep      cell i decides from (a) the committed state (usage counters, the rows above, the cell before the window)
//              and (b) the TENTATIVE choices cells p..i-1 of the window made in the previous sweep: a candidate is usable
//              if  usage[m] + #{earlier window cells currently choosing m} < cap  and it is not the material of the
//              cell above (committed) or to the left (tentative).  The decision itself is the first generation's: first
//              valid candidate unless a later valid one threatens it, then the exact totals of the 32.
//   commit     let t be the first window cell whose choice CHANGED in this sweep.  Cells p..t are final: by induction
//              over the raster order every one of them decided from inputs that are final (the cells before it did not
//              change, so what it read last sweep is what they end up with).  They are committed — counters, placement,
//              row tables — the window slides by t-p+1 cells, the warps that became free take the cells behind it.
//              The head cell p is always final, so every sweep commits at least one cell: the result is the serial
//              one, bit for bit, in at most as many sweeps as there are cells and in practice a sixth of that
//              (scratch/jacobi_sim2.py replays this on the CPU: 5.7 cells per sweep on the 24,000-tile config).
//   markers    a cell the first 32 candidates cannot decide (all rejected, or the early-exit bound fails) makes no
//              tentative choice; it stops the commits in front of it and is resolved with the reference's whole loop
//              (every candidate, usage reset + fallback) once it is the head of the window, by its warp alone.
//
// MEASURED (profiles/r02_chain_wave_experiment.json, ncu: profiles/r02_ncu_wave_chain.txt): bit-exact on every parity
// case, and exactly as many sweeps as the CPU replay predicts (4,230 for 24,000 cells) — but a sweep costs ~4,000
// cycles, not the ~400 the design hoped for: the 32 warps execute ~250 instructions each per sweep (the window
// arithmetic, the bucket lookups, the commit decision every warp repeats), and 32 warps share four issue ports, so the
// kernel is ISSUE-bound (53 % of the SM's issue slots) where the serial chains are latency-bound.  60 % of the sweeps
// commit a single cell (a change ripples from a cell to its right neighbour, one sweep per cell), and there the serial
// chain's ~95 cycles per cell cannot be beaten by any CTA-wide sweep.  8.7 ms on the 24,000-tile config against 2.6 ms
// for the first generation: this file is a parity-green, measured experiment (mosaic_debug_set_chain(ctx, 3) or
// MOSAIC_CHAIN=wave), not the default.
//
// Tentative choosers are found without a search: a table of 2,048 buckets holds, per hash of a material id, the window
// slots whose current choice hashes there; a lane ANDs its candidate's bucket with the mask of the slots before its
// cell and compares ids only for the (rare) bits that remain.
#include "common.cuh"
#include "place_records.cuh"

namespace mosaic {

namespace {

constexpr uint32_t WV_NOKEY = 0xFFFFFFFFu;  // no valid candidate / no choice (as a neighbour: -1, "no material")
constexpr uint32_t WV_FRESH = 0xFFFFFFFEu;  // the cell has not been evaluated yet
constexpr uint32_t WV_MARK = 0xFFFFFFFDu;   // the record cannot decide this cell: resolved at the head of the window
constexpr int WV_NB = 2048;                 // buckets of the tentative-chooser table
__device__ __forceinline__ bool wv_real(uint32_t res) { return res < 0xF0000000u; }
__device__ __forceinline__ unsigned wv_hash(uint32_t id) { return (id * 2654435761u) >> 21; }

template <int W>
__device__ __forceinline__ uint32_t rotr_w(uint32_t x, int r) {
  if constexpr (W == 32) return __funnelshift_r(x, x, r);
  else
  return ((x >> r) | (x << (W - r))) & ((1u << W) - 1u);
}
template <int W>
__device__ __forceinline__ uint32_t rotl_w(uint32_t x, int r) {
  if constexpr (W == 32) return __funnelshift_l(x, x, r);
  else
  return ((x << r) | (x >> (W - r))) & ((1u << W) - 1u);
}

template <typename UsageT>
__device__ __forceinline__ void usage_add_atomic(UsageT* usage, uint32_t id) {  // several warps may commit the same material
  unsigned* word = reinterpret_cast<unsigned*>(usage) + (sizeof(UsageT) == 1 ? (id >> 2) : (id >> 1));
  const unsigned shift = sizeof(UsageT) == 1 ? 8u * (id & 3u) : 16u * (id & 1u);
  atomicAdd(word, 1u << shift);
}

// diagnostics (mosaic_debug_chain_profile), thread 0: [0] sweeps, [1] its serial resolutions, [2] its exact evaluations,
// [3] cycles in phase 1, [4] at the first barrier, [5] in phase 2, [6] its evaluations, [7] cycles at the second barrier,
// [12] cycles total; summed over the other warps: [8] evaluations, [9] exact evaluations, [10] serial resolutions,
// [11] cycles in phase 1, [13] cycles in phase 2
__device__ unsigned long long g_wave_prof[16];

template <typename UsageT, int NW, bool GATED, bool PROF>
__global__ void __launch_bounds__(NW * 32, 1)
    greedy_wave_kernel(const HRec* __restrict__ hrec, const XRec* __restrict__ xrec, const float* __restrict__ glab,
                       const uint32_t* __restrict__ cand_idx, const float* __restrict__ cand_d2, int k,
                       const float4* __restrict__ lab4, int n, int gw, int gh, unsigned max_usage, float weight,
                       int32_t* __restrict__ placed, uint32_t* __restrict__ info, uint32_t fallback_cap,
                       const volatile unsigned* ready, volatile unsigned* progress) {
  constexpr int W = NW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  volatile unsigned* ctl = reinterpret_cast<volatile unsigned*>(smem_raw);  // [6] fallbacks
  volatile uint32_t* wch = reinterpret_cast<volatile uint32_t*>(smem_raw + 32);  // tentative choice per window slot
  volatile uint32_t* wstate = wch + W;  // per slot and sweep parity: bit 0 changed, bit 1 marker, bit 2 record not there yet
  float4* wlab = reinterpret_cast<float4*>(smem_raw + 32 + 12 * W);              // its Lab
  unsigned* buckets = reinterpret_cast<unsigned*>(wlab + W);
  float4* rowlab = reinterpret_cast<float4*>(buckets + WV_NB);  // Lab / id of the last committed cell of every column
  int32_t* rowbuf = reinterpret_cast<int32_t*>(rowlab + gw);
  UsageT* usage = reinterpret_cast<UsageT*>(rowbuf + ((gw + 3) & ~3));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cells = gw * gh;
  const bool use_penalty = weight > 0.0f;  // main.rs:355-362
  for (int i = tid; i < ((n + 4) & ~3); i += NW * 32) usage[i] = 0;
  for (int i = tid; i < gw; i += NW * 32) rowbuf[i] = -1;
  for (int i = tid; i < WV_NB; i += NW * 32) buckets[i] = 0u;
  if (tid < W) wch[tid] = WV_NOKEY, wstate[tid] = 0u, wstate[W + tid] = 0u;
  if (tid < 8) ctl[tid] = 0u;
  __syncthreads();
  if (tid == 0) usage[n] = (UsageT)~(UsageT)0;  // dummy material of lanes >= k: never usable
  __syncthreads();

  // ---- per-warp state: the cell in evaluation and the prefetched record of the next one (cell + W) ----
  int cell = warp;  // this warp's cells are those = warp (mod W): the window slot of a cell is its warp
  int p = 0;        // head of the window (every warp tracks it)
  bool active = false, nloaded = false;
  uint32_t w = 0u, m = 0u, nw = 0u, nm = 0u;
  float d2 = 0.f, l = 0.f, a = 0.f, b = 0.f, d32 = 0.f;
  float nd2 = 0.f, nl = 0.f, na = 0.f, nb = 0.f, nd32 = 0.f;
  uint32_t win = 0u, nwin = 0u;
  int x = 0;
  uint32_t cur = WV_FRESH, sig_ok = 0u;
  int32_t sig_left = -2;
  unsigned ready_seen = (GATED && ready) ? 0u : (unsigned)cells;
  unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  auto load_record = [&](int c, uint32_t& rw, uint32_t& rm, float& rd2, float& rl, float& ra, float& rb, float& rd32, uint32_t& rwin) {
    // (L2 loads: with GATED the records are written by kernels that run beside this one)
    const uint2 hm = __ldcg(&hrec[c].wm[lane]);
    rw = hm.x;
    rm = hm.y;
    const XRec* xr = xrec + c;
    rd2 = __ldcg(&xr->d2[lane]);
    rl = __ldcg(&xr->l[lane]);
    ra = __ldcg(&xr->a[lane]);
    rb = __ldcg(&xr->b[lane]);
    rd32 = __ldcg(&xr->d2_32);
    // lane 0 carries the window of candidate 32 (in the first column the record holds all ones there: the serial
    // chains send those cells to their full step; here the left neighbour is explicit)
    rwin = __ldcg(&xr->win32);  // (its own register: overwriting rm would wait for the load that is filling it)
  };
  auto refresh_ready = [&](int c) {
    if (GATED && ready && (unsigned)c >= ready_seen) {
      unsigned rs = 0u;
      if (lane == 0) rs = *ready;
      rs = __shfl_sync(0xffffffffu, rs, 0);
      __threadfence();  // the records were written before the counter moved
      ready_seen = rs;
    }
  };
  auto enter = [&]() {  // take `cell` (if its record exists), request the record of cell + W
    active = false;
    cur = WV_FRESH;
    if (cell >= cells) return;
    refresh_ready(cell);
    if ((unsigned)cell >= ready_seen) return;
    if (nloaded) {
      w = nw, m = nm, d2 = nd2, l = nl, a = na, b = nb, d32 = nd32, win = nwin;
    } else {
      load_record(cell, w, m, d2, l, a, b, d32, win);
    }
    x = cell % gw;
    active = true;
    nloaded = false;
    if (cell + W < cells && (unsigned)(cell + W) < ready_seen) {
      load_record(cell + W, nw, nm, nd2, nl, na, nb, nd32, nwin);
      nloaded = true;
    }
  };

  // the reference's whole loop for the head cell (every candidate, usage reset + fallback): main.rs:335-475
  auto resolve_serial = [&]() {
    const size_t row = (size_t)cell * k;
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
        gid[ch] = __ldcg(cand_idx + row + j);
        gd2[ch] = __ldcg(cand_d2 + row + j);
        gl[ch] = make_float4(__ldcg(g + j), __ldcg(g + kKPad + j), __ldcg(g + 2 * kKPad + j), 0.f);
      }
    }
    const int32_t up = cell >= gw ? rowbuf[x] : -1;
    const int32_t left = x > 0 ? rowbuf[x - 1] : -1;
    float4 up_lab = make_float4(0.f, 0.f, 0.f, 0.f), left_lab = make_float4(0.f, 0.f, 0.f, 0.f);
    if (up >= 0) up_lab = rowlab[x];
    if (left >= 0) left_lab = rowlab[x - 1];
    unsigned best_bits = 0xFFFFFFFFu;
    int32_t best_id = -1;
    float4 best_lab = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int base = ch * 32;
      if (base >= k) break;
      if (base > 0 && best_id >= 0 && __float_as_uint(__shfl_sync(0xffffffffu, gd2[ch], 0)) >= best_bits) break;  // exact early exit
      bool gok = base + lane < k;
      if (gok) gok = (unsigned)usage[gid[ch]] < max_usage && (int32_t)gid[ch] != left && (int32_t)gid[ch] != up;
      float gpen = 0.0f;
      if (use_penalty) {  // adjacency.rs:71-87: Up, (Down), Left, (Right); absent terms are +0.0
        float t_up = adjacency_term(lab_distance_rn(gl[ch].x, gl[ch].y, gl[ch].z, up_lab.x, up_lab.y, up_lab.z));
        float t_left = adjacency_term(lab_distance_rn(gl[ch].x, gl[ch].y, gl[ch].z, left_lab.x, left_lab.y, left_lab.z));
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
        best_lab.x = __shfl_sync(0xffffffffu, gl[ch].x, who);
        best_lab.y = __shfl_sync(0xffffffffu, gl[ch].y, who);
        best_lab.z = __shfl_sync(0xffffffffu, gl[ch].z, who);
      }
    }
    if (best_id < 0) {
      // fallback_tile_selection (main.rs:420-475)
      const unsigned nfb = ctl[6];
      if (lane == 0) {
        if (nfb < fallback_cap) info[1 + nfb] = (uint32_t)cell;
        ctl[6] = nfb + 1u;
      }
      for (int i = lane; i < n; i += 32) usage[i] = 0;  // lib.rs:86-88
      __syncwarp();
      int who = -1, fch = 0;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch * 32 >= k || best_id >= 0) break;
        const bool fok = ch * 32 + lane < k && (int32_t)gid[ch] != left && (int32_t)gid[ch] != up;  // main.rs:448
        const unsigned mm = __ballot_sync(0xffffffffu, fok);
        if (mm) {
          who = __ffs(mm) - 1;
          fch = ch;
          best_id = (int32_t)__shfl_sync(0xffffffffu, gid[ch], who);
        }
      }
      if (best_id < 0) {  // main.rs:456-474 nearest_one
        who = 0;
        fch = 0;
        best_id = (int32_t)__shfl_sync(0xffffffffu, gid[0], 0);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        if (ch == fch) {
          best_lab.x = __shfl_sync(0xffffffffu, gl[ch].x, who);
          best_lab.y = __shfl_sync(0xffffffffu, gl[ch].y, who);
          best_lab.z = __shfl_sync(0xffffffffu, gl[ch].z, who);
        }
    }
    __syncwarp();
    if (lane == 0) {
      usage[best_id] = (UsageT)(usage[best_id] + 1);  // main.rs:410-411 / 449 / 470 (nobody else commits in this sweep)
      placed[cell] = best_id;                         // main.rs:412
      rowbuf[x] = best_id;
      rowlab[x] = best_lab;
    }
    __syncwarp();
  };

  enter();
  const long long t_all = clock64();
  unsigned next_pub = (unsigned)gw;
  for (unsigned sweep = 0;; ++sweep) {
    const int par = (int)(sweep & 1u);
    long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0;
    if (PROF) tp0 = clock64();
    // ---------------- phase 1: every window cell decides from the committed state + last sweep's choices ----------------
    if (active) {
      const int rel = cell - p;  // 0 .. W-1
      const uint32_t before = rotl_w<W>((1u << rel) - 1u, p % W);  // slots of the window cells in front of this one
      // The tentative choice of the cell to the left changes under this warp's feet (its warp publishes it in this same
      // phase): it is read ONCE, by one lane, and broadcast — lanes that left the divergent loop below at different
      // times would otherwise see different neighbours and part ways in front of the warp-wide reductions.
      // (The neighbour's id is read BEFORE its Lab and written after it: see the publication below.)
      int32_t left = -1;
      if (x > 0) left = rel == 0 ? *reinterpret_cast<volatile int32_t*>(rowbuf + x - 1) : (int32_t)wch[(warp + W - 1) % W];
      left = __shfl_sync(0xffffffffu, left, 0);
      const int32_t up = cell >= gw ? *reinterpret_cast<volatile int32_t*>(rowbuf + x) : -1;  // (committed: stable)
      const uint32_t id = w & CH_IDMASK;
      const unsigned u = (unsigned)usage[id];
      uint32_t bm = buckets[wv_hash(id)] & before;
      unsigned cnt = 0u;
      while (bm) {  // (rare: some earlier window cell's choice hashes like this candidate)
        const int s = __ffs(bm) - 1;
        bm &= bm - 1u;
        cnt += wch[s] == id ? 1u : 0u;
      }
      const bool ok = (int32_t)id != up && (int32_t)id != left && u + cnt < max_usage;  // main.rs:375, 380
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      uint32_t res = cur;
      if (cur == WV_FRESH || okmask != sig_ok || left != sig_left) {  // (Up is committed: it cannot change)
        sig_ok = okmask;
        sig_left = left;
        if (PROF) prof[6]++;
        const uint32_t r = __reduce_min_sync(0xffffffffu, ok ? w : WV_NOKEY);       // first valid candidate
        const uint32_t T = __reduce_or_sync(0xffffffffu, lane == 0 ? win : (ok ? m : 0u));  // who is threatened by a valid one
        const uint32_t j0 = r >> 26;
        res = r;
        if (__funnelshift_rc(T, 0xFFFFFFFFu, j0) & 1u) {  // (possibly) more than one contender, or no valid candidate
          res = WV_MARK;
          if (j0 < 32u) {  // exact totals of the record's 32 candidates (main.rs:386-407)
            if (PROF) prof[2]++;
            float4 pl = make_float4(0.f, 0.f, 0.f, 0.f), ul = make_float4(0.f, 0.f, 0.f, 0.f);
            if (left >= 0) {
              const volatile float4* q = rel == 0 ? rowlab + (x - 1) : wlab + ((warp + W - 1) % W);
              pl = make_float4(q->x, q->y, q->z, 0.f);
              pl.x = __shfl_sync(0xffffffffu, pl.x, 0);
              pl.y = __shfl_sync(0xffffffffu, pl.y, 0);
              pl.z = __shfl_sync(0xffffffffu, pl.z, 0);
              // id, Lab, id again: if the neighbour published in between, this evaluation may have paired the old id
              // with the new colour — it is not final anyway (the neighbour changed), and it must not be reused
              if (rel != 0 && __shfl_sync(0xffffffffu, (int32_t)wch[(warp + W - 1) % W], 0) != left) sig_left = -2;
            }
            if (up >= 0) {
              const volatile float4* q = rowlab + x;
              ul = make_float4(q->x, q->y, q->z, 0.f);
            }
            bool oor_u = false, oor_l = false;
            float pen = 0.0f;
            if (use_penalty) {
              float t_up = adjacency_term_fast(l, a, b, ul.x, ul.y, ul.z, oor_u);
              float t_left = adjacency_term_fast(l, a, b, pl.x, pl.y, pl.z, oor_l);
              t_up = up >= 0 ? t_up : 0.0f;  // absent terms are +0.0: x + 0.0 == x exactly
              t_left = left >= 0 ? t_left : 0.0f;
              pen = __fmul_rn(__fadd_rn(t_up, t_left), weight);  // adjacency.rs:71-87 order Up, Left; then * weight
            }
            unsigned tb = ok ? __float_as_uint(__fadd_rn(d2, pen)) : 0xFFFFFFFFu;  // main.rs:398
            tb = ((oor_u && up >= 0) || (oor_l && left >= 0)) ? 0u : tb;           // out-of-domain arithmetic -> serial path
            const unsigned mn = __reduce_min_sync(0xffffffffu, tb);
            if (mn - 1u < 0xFFFFFFFEu && __float_as_uint(d32) >= mn) {
              const int who = __ffs(__ballot_sync(0xffffffffu, tb == mn)) - 1;  // strict <, first in order on ties
              res = __shfl_sync(0xffffffffu, w, who);
            }
          }
        }
      }
      uint32_t st = res == WV_MARK ? 2u : 0u;
      if (res != cur) {
        // publication order: Lab, then id — a reader that sees the new id also sees the new Lab (it reads the id first)
        if (wv_real(res) && (uint32_t)lane == (res >> 26)) {
          volatile float4* q = wlab + warp;
          q->x = l, q->y = a, q->z = b;
        }
        __syncwarp();
        if (lane == 0) {
          if (wv_real(cur)) atomicAnd(&buckets[wv_hash(cur & CH_IDMASK)], ~(1u << warp));
          if (wv_real(res)) atomicOr(&buckets[wv_hash(res & CH_IDMASK)], 1u << warp);
          wch[warp] = wv_real(res) ? (res & CH_IDMASK) : WV_NOKEY;
        }
        cur = res;
        st |= 1u;
      }
      if (lane == 0) wstate[par * W + warp] = st;
    } else {
      if (lane == 0) wstate[par * W + warp] = cell < cells ? 4u : 0u;  // its record does not exist yet / behind the grid
    }
    if (PROF) tp1 = clock64();
    __syncthreads();
    if (PROF) tp2 = clock64();
    // ---------------- phase 2: how far is the window final? ----------------
    const int rot = p % W;
    const uint32_t stw = lane < W ? wstate[par * W + lane] : 0u;
    const uint32_t rc = rotr_w<W>(__ballot_sync(0xffffffffu, stw & 1u), rot), rm = rotr_w<W>(__ballot_sync(0xffffffffu, stw & 2u), rot),
                   ri = rotr_w<W>(__ballot_sync(0xffffffffu, stw & 4u), rot);
    int t = min(W, cells - p);
    if (rc) t = min(t, __ffs(rc));      // the first cell that changed is final too: everything before it was
    if (rm) t = min(t, __ffs(rm) - 1);  // a marker is resolved at the head, alone
    if (ri) t = min(t, __ffs(ri) - 1);
    const bool serial = t == 0 && (rm & 1u);
    if (PROF && tid == 0) {
      prof[0]++;
    }
    if (active && cell - p < t) {
      const uint32_t id = cur & CH_IDMASK;
      if ((uint32_t)lane == (cur >> 26)) rowlab[x] = make_float4(l, a, b, 0.f);
      if (lane == 0) {
        usage_add_atomic<UsageT>(usage, id);  // main.rs:410-411
        placed[cell] = (int32_t)id;           // main.rs:412
        rowbuf[x] = (int32_t)id;
        atomicAnd(&buckets[wv_hash(id)], ~(1u << warp));
        wch[warp] = WV_NOKEY;
      }
      cell += W;
      enter();
    } else if (serial && cell == p) {
      resolve_serial();
      if (PROF) prof[1]++;
      cell += W;
      enter();
    } else if (!active && cell < cells) {
      enter();  // (its record may have arrived)
    }
    p += serial ? 1 : t;
    if (t == 0 && !serial) __nanosleep(100);  // the head's record is not there yet
    if (PROF) tp3 = clock64();
    __syncthreads();
    if (PROF) prof[7] += (unsigned long long)(clock64() - tp3), prof[3] += (unsigned long long)(tp1 - tp0), prof[4] += (unsigned long long)(tp2 - tp1), prof[5] += (unsigned long long)(tp3 - tp2);
    if (GATED && progress && tid == 0 && ((unsigned)p >= next_pub || p >= cells)) {
      __threadfence();  // the placements of the cells before p are in HBM (the barrier ordered the other warps' stores)
      *progress = (unsigned)p;
      next_pub = ((unsigned)p / (unsigned)gw + 1u) * (unsigned)gw;
    }
    if (p >= cells) break;
    if ((sweep & 1023u) == 1023u && clock64() - t_all > 40000000000ll) __trap();  // (20 s: the producer died; do not hang the GPU)
  }
  if (tid == 0) info[0] = ctl[6];
  if (PROF && tid == 0) {
    for (int i = 0; i < 8; ++i) g_wave_prof[i] = prof[i];
    g_wave_prof[12] = (unsigned long long)(clock64() - t_all);
  }
  if (PROF && lane == 0 && tid != 0) {  // evaluations are counted per warp
    atomicAdd(&g_wave_prof[8], prof[6]);
    atomicAdd(&g_wave_prof[9], prof[2]);
    atomicAdd(&g_wave_prof[10], prof[1]);
    atomicAdd(&g_wave_prof[11], prof[3]);
    atomicAdd(&g_wave_prof[13], prof[5]);
  }
}

constexpr int WV_WARPS = 32;

size_t wave_smem(const mosaic_ctx* ctx, unsigned cap) {
  const size_t usage_bytes = cap <= 254u ? (size_t)ctx->n + 4 : (cap <= 65534u ? ((size_t)ctx->n + 4) * 2 : (size_t)-1);
  if (usage_bytes == (size_t)-1) return 0;
  const size_t fixed = 32 + 12 * WV_WARPS + 16 * WV_WARPS + 4 * WV_NB + 16 * (size_t)ctx->gw + 4 * (size_t)((ctx->gw + 3) & ~3);
  const size_t total = fixed + ((usage_bytes + 15) & ~(size_t)15) + 16;
  return total <= 220 * 1024 ? total : 0;
}

template <typename UsageT, bool GATED, bool PROF>
int launch_wave_impl(mosaic_ctx* ctx, unsigned cap, float weight, size_t smem, cudaStream_t stream, const unsigned* ready,
                     unsigned* progress) {
  const size_t cells = (size_t)ctx->gw * ctx->gh;
  if (GATED) smem = std::max(smem, (size_t)226 * 1024);  // other kernels run beside the chain: keep their CTAs off this SM
  auto kern = greedy_wave_kernel<UsageT, WV_WARPS, GATED, PROF>;
  MOSAIC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<1, WV_WARPS * 32, smem, stream>>>(
      reinterpret_cast<const HRec*>(ctx->d_place_rec), reinterpret_cast<const XRec*>(ctx->d_place_rec + cells * sizeof(HRec)),
      reinterpret_cast<const float*>(ctx->d_place_rec + cells * (sizeof(HRec) + sizeof(XRec))), ctx->d_cand_idx, ctx->d_cand_d2,
      ctx->k, ctx->sim_lab4(), ctx->n, ctx->gw, ctx->gh, cap, weight, ctx->d_placed, ctx->d_place_info, mosaic_ctx::kFallbackCap,
      ready, progress);
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace

// the wave chain covers: a grid at least as wide as the window (the cell above is always committed), at least one use
// per material (cap 0 makes every cell a fallback: the serial chains handle that), counters of 8 or 16 bits
bool wave_chain_eligible(const mosaic_ctx* ctx, int max_usage) {
  const unsigned cap = max_usage < 0 ? 0u : (unsigned)max_usage;
  return cap >= 1u && ctx->gw >= WV_WARPS && ctx->n < (1 << 26) && wave_smem(ctx, cap) != 0;
}

// records: place_prep_kernel's (greedy.cu); `ready` / `progress` as for the first generation (both may be NULL)
int launch_wave_chain(mosaic_ctx* ctx, int max_usage, float weight, cudaStream_t stream, const unsigned* ready, unsigned* progress,
                      bool prof) {
  const unsigned cap = (unsigned)max_usage;
  const size_t smem = wave_smem(ctx, cap);
  if (!smem) return fail(ctx, MOSAIC_ERR_STATE, "placement: the wave chain does not cover this configuration");
  const bool gated = ready != nullptr || progress != nullptr;
  if (prof) {
    static const unsigned long long zeros[16] = {};
    MOSAIC_CUDA(ctx, cudaMemcpyToSymbolAsync(g_wave_prof, zeros, sizeof zeros, 0, cudaMemcpyHostToDevice, stream));
    if (cap <= 254u) return launch_wave_impl<uint8_t, false, true>(ctx, cap, weight, smem, stream, nullptr, nullptr);
    return launch_wave_impl<uint16_t, false, true>(ctx, cap, weight, smem, stream, nullptr, nullptr);
  }
  if (cap <= 254u)
    return gated ? launch_wave_impl<uint8_t, true, false>(ctx, cap, weight, smem, stream, ready, progress)
                 : launch_wave_impl<uint8_t, false, false>(ctx, cap, weight, smem, stream, nullptr, nullptr);
  return gated ? launch_wave_impl<uint16_t, true, false>(ctx, cap, weight, smem, stream, ready, progress)
               : launch_wave_impl<uint16_t, false, false>(ctx, cap, weight, smem, stream, nullptr, nullptr);
}

int read_wave_profile(mosaic_ctx* ctx, unsigned long long* out16) {
  MOSAIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MOSAIC_CUDA(ctx, cudaMemcpyFromSymbol(out16, g_wave_prof, sizeof(unsigned long long) * 16));
  return MOSAIC_OK;
}

}  // namespace mosaic
