// nn_topk.cu — K2: brute-force FP32 Lab nearest-neighbour scoring + per-cell ordered top-K.
//
// Replaces the kiddo k-d tree queries of the reference (src/main.rs:90, 145-151 build;
// nearest_n::<SquaredEuclidean> at main.rs:348-352 / 435-439; nearest_one at main.rs:456-459).
// Result per cell: the K = min(N, 100) materials with the smallest
//     d2 = ((0 + dl*dl) + da*da) + db*db          (kiddo SquaredEuclidean: unfused f32 left fold)
// ascending by (d2, material index).  kiddo's order among equal distances is unspecified; the
// (d2, index) order is the one the oracle defines.
//
// Layout: material Lab means as SoA l[], a[], b[] in HBM; a CTA streams them through shared memory in 1024-material
// stages with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered).
// Each cell warp owns Q=2 cells whose Lab values live in registers; lanes map to materials, so one stage read (3
// conflict-free LDS) serves 2x32 pairs.  A pair costs 8 FP32 ops (3 sub, 3 mul, 2 add; FMA contraction would change the
// bits) + compare + vote.  Selection: a 64-bit key (d2 bits << 32 | index) is appended to a per-cell shared-memory POOL
// of 256 slots when d2 <= tau (an upper bound of the K-th smallest d2 seen so far).  When the pool holds more than 224
// keys it is PRUNED: the warp bisects on the distance bits for a threshold that keeps between K and K+16 keys (each
// step counts the keys below the probe: 8 compares per lane + one REDUX.ADD), drops the rest and lowers tau.  Only the
// END of a cell sorts: one more bisection down to <= 128 keys, one bitonic network.  Pools the bisection cannot thin out
// (more equal distances than slots) take the exact path: two bitonic sorts and a merge by the full (d2, index) key.
// A warp handles up to four consecutive groups of raster cells; from the second group on, tau starts at the largest
// distance from the new cell to the k materials of the previous cell's answer (k materials within that distance exist,
// so it bounds the k-th smallest from above): raster neighbours have almost the same colour, which removes most of the
// k*ln(n/k) early insertions.
//
// Round 2 (ncu, C3): 317 M warp instructions — 56 % the scan itself (47 instructions per 2 x 32 pairs before, ~38 after
// keeping tau and the counters in registers: passing them by reference to the out-of-line prune had pinned them to the
// stack), 23 % the appends, 8 % the final sorts, 3 % the prunes.  0.584 -> 0.432 ms.  Measured and dropped: Q=4 cells
// per warp with 4 warps per CTA (0.58 ms: the kernel wants its warps more than it wants fewer loads per pair); the slack
// of a prune (2 .. 90 keys: 0.456 .. 0.468 ms, no effect); a producer warp with four 512-material stages and full / empty
// mbarriers instead of the CTA barrier per stage (0.435 ms against 0.432: the barrier was not the cost).
#include "common.cuh"

namespace mosaic {

namespace {

constexpr int TK_WARPS = 8;
constexpr int TK_Q = 2;
constexpr int TK_CHUNK = 1024;  // materials per stage (two stages, double buffered)
constexpr int TK_SLOTS = 128;
constexpr int TK_MAX_ROUNDS = 4;  // consecutive cell groups per warp: every group after the first starts from its predecessor's answer

constexpr int TK_POOL = 2 * TK_SLOTS;   // keys a cell's pool holds
constexpr int TK_PRUNE = TK_POOL - 32;  // prune when more than this many keys are pooled (one more step adds <= 32)
constexpr int TK_SLACK = 16;            // a prune keeps between k and k + TK_SLACK keys

struct __align__(16) QueryState {
  unsigned long long pool[TK_POOL];  // unsorted while a cell is scanned; its first k slots hold the sorted answer afterwards
};

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return ((unsigned long long)hi << 32) | lo;
}

// Bitonic network over 128 keys, element e = lane*4 + r.  Runs phases kbegin..128.
template <int KBEGIN>
__device__ __forceinline__ void bitonic128(unsigned long long (&v)[4], int lane) {
#pragma unroll
  for (int k = KBEGIN; k <= 128; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 4) {
        const int lm = j >> 2;
        const bool lower = (lane & lm) == 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = lane * 4 + r;
          const bool asc = (e & k) == 0;
          unsigned long long o = shfl_xor_u64(v[r], lm);
          unsigned long long mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
          v[r] = (lower == asc) ? mn : mx;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if ((r & j) == 0) {
            const int e = lane * 4 + r;
            const bool asc = (e & k) == 0;
            unsigned long long a = v[r], b = v[r | j];
            bool sw = (a > b) == asc;
            v[r] = sw ? b : a;
            v[r | j] = sw ? a : b;
          }
        }
      }
    }
  }
}

// The 128 smallest keys (by the full (d2, index) key) of a pool of <= 256, sorted, into slots 0..127: two bitonic sorts
// and a merge.  The exact path: end of a cell whose pool the bisection could not bring under 128 keys, and pools with
// more equal distances than the bisection can separate.
__device__ __noinline__ void exact_top128(QueryState& st, int cnt, int lane) {
  __syncwarp();
  unsigned long long v[4], u[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = lane * 4 + r;
    v[r] = e < cnt ? st.pool[e] : ~0ull;
    u[r] = TK_SLOTS + e < cnt ? st.pool[TK_SLOTS + e] : ~0ull;
  }
  bitonic128<2>(v, lane);  // ascending
  bitonic128<2>(u, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) st.pool[TK_SLOTS + lane * 4 + r] = u[r];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) {  // 128 smallest of the two sorted halves as a bitonic sequence
    const unsigned long long t = st.pool[TK_SLOTS + 127 - (lane * 4 + r)];
    v[r] = v[r] < t ? v[r] : t;
  }
  bitonic128<128>(v, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) st.pool[lane * 4 + r] = v[r];
  __syncwarp();
}

// Thin the pool out to at most `limit` keys without losing any of the k smallest: bisection on the distance bits for a
// threshold hi with  k <= #{d2 <= hi} <= limit  (exact when it has to be: the k-th smallest distance itself), then
// compaction in place.  Slot e = lane + 32 r belongs to lane `lane`.  Returns false — pool untouched — when even the exact
// k-th distance keeps more than `limit` keys (ties).  On success tau = hi: every later key above it cannot be among the
// k smallest.
// (cnt and tau travel by value and come back packed — new count << 32 | new tau, or ~0 for "pool untouched": a reference
// to them would pin both to the stack for the whole scan loop)
__device__ __noinline__ unsigned long long bisect_prune(QueryState& st, int cnt, unsigned tau, int k, int limit, int lane) {
  __syncwarp();
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long key[8];
  unsigned d[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int e = lane + 32 * r;
    key[r] = e < cnt ? st.pool[e] : ~0ull;
    d[r] = (unsigned)(key[r] >> 32);  // (empty slots: 0xFFFFFFFF, above every probe)
  }
  // invariant: #{d2 <= lo} < k (lo = -1: none), #{d2 <= hi} = chi >= k
  long long lo = -1, hi = (long long)tau;
  int chi = cnt;
  while (chi > limit && hi - lo > 1) {
    const long long mid = lo + ((hi - lo) >> 1);
    const unsigned m = (unsigned)mid;
    int c = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) c += d[r] <= m ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) {
      hi = mid;
      chi = c;
    } else {
      lo = mid;
    }
  }
  if (chi > limit) return ~0ull;
  const unsigned h = (unsigned)hi;
  int base = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {  // in place: a key never moves to a higher slot, and a round's reads precede its writes
    const bool keep = d[r] <= h;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) st.pool[base + __popc(mask & lt_mask)] = key[r];
    base += __popc(mask);
  }
  __syncwarp();
  return ((unsigned long long)(unsigned)base << 32) | min(tau, h);
}

// mid-scan prune: down to [k, k + slack] keys
__device__ __forceinline__ void prune_pool(QueryState& st, int& cnt, unsigned& tau, int k, int lane, int slack) {
  if (cnt <= k) return;  // (fewer keys than the answer needs: nothing can go)
  const unsigned long long r = bisect_prune(st, cnt, tau, k, k + slack, lane);
  if (r != ~0ull) {
    cnt = (int)(r >> 32);
    tau = (unsigned)r;
  } else {
    exact_top128(st, cnt, lane);  // ties: the 128 smallest by the full key; k <= 100 of them are the answer so far
    cnt = min(cnt, TK_SLOTS);
    tau = min(tau, (unsigned)(st.pool[k - 1] >> 32));
  }
}

// end of a cell: slots 0..k-1 <- the k smallest keys, ascending by (d2, index)
__device__ __forceinline__ void finish_pool(QueryState& st, int cnt, unsigned tau, int k, int lane) {
  if (cnt > TK_SLOTS) {
    const unsigned long long r = bisect_prune(st, cnt, tau, k, TK_SLOTS, lane);
    if (r == ~0ull) {
      exact_top128(st, cnt, lane);
      return;
    }
    cnt = (int)(r >> 32);
  }
  __syncwarp();
  unsigned long long v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = lane * 4 + r;
    v[r] = e < cnt ? st.pool[e] : ~0ull;
  }
  bitonic128<2>(v, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) st.pool[lane * 4 + r] = v[r];
  __syncwarp();
}

__global__ void __launch_bounds__(TK_WARPS * 32, 3) nn_topk_kernel(const float* __restrict__ soa, const float4* __restrict__ lab4,
                                                                int n, int n_pad, const float* __restrict__ cell_lab,
                                                                int cell_begin, int cell_end, int k, int rounds, int slack,
                                                                uint32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* mats = reinterpret_cast<float*>(smem_raw);  // [2][3][TK_CHUNK]
  QueryState* states = reinterpret_cast<QueryState*>(smem_raw + 2 * 3 * TK_CHUNK * sizeof(float));
  unsigned long long* bars =
      reinterpret_cast<unsigned long long*>(smem_raw + 2 * 3 * TK_CHUNK * sizeof(float) +
                                            TK_WARPS * TK_Q * sizeof(QueryState));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = (n + TK_CHUNK - 1) / TK_CHUNK;
  const int total_chunks = nchunks * rounds;  // the materials are streamed once per round

  auto issue = [&](int g) {
    const int s = g & 1, c = g % nchunks;
    const int cnt = min(TK_CHUNK, n_pad - c * TK_CHUNK);  // n_pad is a multiple of 4 floats -> 16-byte sizes
    const unsigned bytes = (unsigned)cnt * sizeof(float);
    mbar_expect_tx(&bars[s], 3 * bytes);
#pragma unroll
    for (int d = 0; d < 3; ++d)
      tma_load_1d(mats + (s * 3 + d) * TK_CHUNK, soa + (size_t)d * n_pad + (size_t)c * TK_CHUNK, bytes, &bars[s]);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    issue(0);
    if (total_chunks > 1) issue(1);
  }

  QueryState* st = states + warp * TK_Q;
  const unsigned lt_mask = (1u << lane) - 1u;
  int g = 0;  // running chunk number across the rounds
  for (int round = 0; round < rounds; ++round) {
    // this warp's cells of this round: rounds * TK_Q consecutive raster cells per warp
    const int first_cell = cell_begin + ((blockIdx.x * TK_WARPS + warp) * rounds + round) * TK_Q;
    const bool warp_active = first_cell < cell_end;  // warp-uniform
    float ql[TK_Q], qa[TK_Q], qb[TK_Q];
    int cnt[TK_Q];
    unsigned tau[TK_Q];
#pragma unroll
    for (int qi = 0; qi < TK_Q; ++qi) {
      const int cell = min(first_cell + qi, cell_end - 1);
      ql[qi] = __ldg(cell_lab + 3 * (size_t)cell);
      qa[qi] = __ldg(cell_lab + 3 * (size_t)cell + 1);
      qb[qi] = __ldg(cell_lab + 3 * (size_t)cell + 2);
      cnt[qi] = 0;
      tau[qi] = 0xFFFFFFFFu;
    }
    if (round > 0 && warp_active) {
      // Seed the threshold from the answer of the previous cell (raster neighbour, almost the same colour): its
      // k materials are k materials whose distance to THIS cell is <= the largest of them, so that value bounds
      // this cell's k-th smallest distance from above — same arithmetic as the scan, hence an exact bound.
      float far_[TK_Q];
#pragma unroll
      for (int qi = 0; qi < TK_Q; ++qi) far_[qi] = 0.0f;
      for (int e = lane; e < k; e += 32) {
        const float4 m = __ldg(&lab4[(uint32_t)st[TK_Q - 1].pool[e]]);
#pragma unroll
        for (int qi = 0; qi < TK_Q; ++qi) {
          const float dl = __fsub_rn(ql[qi], m.x), da = __fsub_rn(qa[qi], m.y), db = __fsub_rn(qb[qi], m.z);
          far_[qi] = fmaxf(far_[qi], __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db)));
        }
      }
#pragma unroll
      for (int qi = 0; qi < TK_Q; ++qi) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) far_[qi] = fmaxf(far_[qi], __shfl_xor_sync(0xffffffffu, far_[qi], o));
        tau[qi] = __float_as_uint(far_[qi]);  // d2 >= 0: the bit patterns order like the values
      }
    }
    __syncwarp();  // (the previous answer has been read: the pools start empty)

    for (int c = 0; c < nchunks; ++c, ++g) {
      const int s = g & 1;
      mbar_wait(&bars[s], (unsigned)((g >> 1) & 1));
      if (warp_active) {
        const int count = min(TK_CHUNK, n - c * TK_CHUNK);
        const float* pl = mats + (s * 3) * TK_CHUNK + lane;  // l; a and b sit TK_CHUNK and 2 * TK_CHUNK floats behind
        unsigned midx = (unsigned)(c * TK_CHUNK) + (unsigned)lane;
        // 32 materials x TK_Q cells; `in_range` is false only for lanes behind the last material of the last chunk
        auto step = [&](bool in_range) {
          const float ml = pl[0], ma = pl[TK_CHUNK], mb = pl[2 * TK_CHUNK];
#pragma unroll
          for (int qi = 0; qi < TK_Q; ++qi) {
            const float dl = __fsub_rn(ql[qi], ml), da = __fsub_rn(qa[qi], ma), db = __fsub_rn(qb[qi], mb);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db));
            const unsigned bits = __float_as_uint(d2);
            const bool pass = in_range && bits <= tau[qi];
            const unsigned mask = __ballot_sync(0xffffffffu, pass);
            if (mask) {
              if (pass) st[qi].pool[cnt[qi] + __popc(mask & lt_mask)] = ((unsigned long long)bits << 32) | midx;
              cnt[qi] += __popc(mask);
              if (cnt[qi] > TK_PRUNE) prune_pool(st[qi], cnt[qi], tau[qi], k, lane, slack);
            }
          }
          pl += 32;
          midx += 32u;
        };
        const int whole = count & ~31;
        for (int j0 = 0; j0 < whole; j0 += 32) step(true);
        if (whole < count) step(lane < count - whole);
      }
      __syncthreads();  // every warp is done with stage s
      if (tid == 0 && g + 2 < total_chunks) issue(g + 2);
    }

    if (warp_active) {
#pragma unroll
      for (int qi = 0; qi < TK_Q; ++qi) {
        finish_pool(st[qi], cnt[qi], tau[qi], k, lane);
        const int cell = first_cell + qi;
        if (cell < cell_end) {
          for (int e = lane; e < k; e += 32) {
            unsigned long long key = st[qi].pool[e];
            out_idx[(size_t)cell * k + e] = (uint32_t)key;
            out_d2[(size_t)cell * k + e] = __uint_as_float((unsigned)(key >> 32));
          }
        }
      }
    }
    __syncwarp();
  }
}

constexpr size_t kTopkSmem =
    2 * 3 * TK_CHUNK * sizeof(float) + TK_WARPS * TK_Q * sizeof(QueryState) + 2 * sizeof(unsigned long long);

}  // namespace

int launch_topk_range(mosaic_ctx* ctx, int k, int cell_begin, int cell_end) {
  const int cells = cell_end - cell_begin;
  if (cells <= 0) return MOSAIC_OK;
  static bool attr_set[64] = {};
  if (ctx->device >= 64 || !attr_set[ctx->device]) {
    MOSAIC_CUDA(ctx, cudaFuncSetAttribute(nn_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopkSmem));
    if (ctx->device < 64) attr_set[ctx->device] = true;
  }
  // rounds per warp: as many as keep at least two CTAs per SM in flight (a sharded context has fewer cells per GPU)
  int rounds = TK_MAX_ROUNDS;
  while (rounds > 1 && (cells + TK_WARPS * TK_Q * rounds - 1) / (TK_WARPS * TK_Q * rounds) < 2 * ctx->sm_count) rounds >>= 1;
  // A band that only has to stay ahead of the serial placement chain (api.cu) wants the least WORK, not the least
  // latency: every group after a warp's first starts from its predecessor's answer (a third of the selection work).
  if (ctx->topk_force_rounds > 0) rounds = ctx->topk_force_rounds;
  static const int tk_slack = getenv("MOSAIC_TK_SLACK") ? atoi(getenv("MOSAIC_TK_SLACK")) : TK_SLACK;  // (tuning knob)
  const int per_cta = TK_WARPS * TK_Q * rounds;
  const int blocks = (cells + per_cta - 1) / per_cta;
  {
    StageTimer t(ctx, MOSAIC_STAGE_TOPK);
    nn_topk_kernel<<<blocks, TK_WARPS * 32, kTopkSmem, ctx->stream>>>(ctx->d_mat_soa, ctx->d_mat_lab4, ctx->n, ctx->n_pad,
                                                                      ctx->d_cell_lab, cell_begin, cell_end, k, rounds, tk_slack,
                                                                      ctx->cand_dst_idx ? ctx->cand_dst_idx : ctx->d_cand_idx,
                                                                      ctx->cand_dst_d2 ? ctx->cand_dst_d2 : ctx->d_cand_d2);
    ctx->launches++;
  }
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace mosaic
