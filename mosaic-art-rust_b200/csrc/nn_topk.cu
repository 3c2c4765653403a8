// nn_topk.cu — K2: brute-force FP32 Lab nearest-neighbour scoring + per-cell ordered top-K.
//
// Replaces the kiddo k-d tree queries of the reference (src/main.rs:90, 145-151 build;
// nearest_n::<SquaredEuclidean> at main.rs:348-352 / 435-439; nearest_one at main.rs:456-459).
// Result per cell: the K = min(N, 100) materials with the smallest
//     d2 = ((0 + dl*dl) + da*da) + db*db          (kiddo SquaredEuclidean: unfused f32 left fold)
// ascending by (d2, material index).  kiddo's order among equal distances is unspecified; the
// (d2, index) order is the one the oracle defines.
//
// Layout: material Lab means as SoA l[], a[], b[] in HBM; a CTA streams them through shared memory
// in 2048-material stages with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered).
// Each warp owns Q=4 cells whose Lab values live in registers; lanes map to materials, so one
// stage read (3 conflict-free LDS) serves 4x32 pairs.  A pair costs 8 FP32 ops (3 sub, 3 mul, 2 add;
// FMA contraction would change the bits) + compare + vote.  Selection: a 64-bit key
// (d2 bits << 32 | index) is appended to a per-cell shared-memory buffer when d2 <= tau (the current
// K-th smallest d2); when the buffer holds > 96 keys the warp bitonic-sorts it in registers and
// merges it with the sorted top list (128 slots), tightening tau.  A warp handles up to four consecutive groups
// of raster cells; from the second group on, tau starts at the largest distance from the new cell to the k
// materials of the previous cell's answer (k materials within that distance exist, so it bounds the k-th
// smallest from above): raster neighbours have almost the same colour, which removes most of the k*ln(n/k)
// early insertions and their sorts.
#include "common.cuh"

namespace mosaic {

namespace {

constexpr int TK_WARPS = 8;
constexpr int TK_Q = 2;
constexpr int TK_CHUNK = 1024;
constexpr int TK_SLOTS = 128;
constexpr int TK_FLUSH = 96;  // merge when more than this many keys are buffered (<= 128 - 32)
constexpr int TK_MAX_ROUNDS = 4;  // consecutive cell groups per warp: every group after the first starts from its predecessor's answer

struct __align__(16) QueryState {
  unsigned long long topk[TK_SLOTS];
  unsigned long long buf[TK_SLOTS];
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

__device__ __forceinline__ void merge_buffer(QueryState& st, int& cnt, unsigned& tau, int k, int lane) {
  __syncwarp();
  unsigned long long v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int e = lane * 4 + r;
    v[r] = e < cnt ? st.buf[e] : ~0ull;
  }
  bitonic128<2>(v, lane);  // ascending
#pragma unroll
  for (int r = 0; r < 4; ++r) {  // 128 smallest of (sorted buf) U (sorted topk) as a bitonic sequence
    unsigned long long t = st.topk[127 - (lane * 4 + r)];
    v[r] = v[r] < t ? v[r] : t;
  }
  bitonic128<128>(v, lane);
  __syncwarp();  // every lane has read its (mirrored) slots of the old list before anybody overwrites them
#pragma unroll
  for (int r = 0; r < 4; ++r) st.topk[lane * 4 + r] = v[r];
  __syncwarp();
  tau = min(tau, (unsigned)(st.topk[k - 1] >> 32));  // (the list holds ~0 until k keys have been seen)
  cnt = 0;
}

__global__ void __launch_bounds__(TK_WARPS * 32) nn_topk_kernel(const float* __restrict__ soa, const float4* __restrict__ lab4,
                                                                int n, int n_pad, const float* __restrict__ cell_lab,
                                                                int cell_begin, int cell_end, int k, int rounds,
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
        const float4 m = __ldg(&lab4[(uint32_t)st[TK_Q - 1].topk[e]]);
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
    __syncwarp();
#pragma unroll
    for (int qi = 0; qi < TK_Q; ++qi) {
#pragma unroll
      for (int r = 0; r < 4; ++r) st[qi].topk[lane * 4 + r] = ~0ull;
    }
    __syncwarp();

    for (int c = 0; c < nchunks; ++c, ++g) {
      const int s = g & 1;
      mbar_wait(&bars[s], (unsigned)((g >> 1) & 1));
      if (warp_active) {
        const float* sl = mats + (s * 3 + 0) * TK_CHUNK;
        const float* sa = mats + (s * 3 + 1) * TK_CHUNK;
        const float* sb = mats + (s * 3 + 2) * TK_CHUNK;
        const int count = min(TK_CHUNK, n - c * TK_CHUNK);
        const unsigned base_idx = (unsigned)(c * TK_CHUNK);
        for (int j0 = 0; j0 < count; j0 += 32) {
          const int j = j0 + lane;
          const bool in_range = j < count;
          const float ml = sl[j], ma = sa[j], mb = sb[j];
#pragma unroll
          for (int qi = 0; qi < TK_Q; ++qi) {
            const float dl = __fsub_rn(ql[qi], ml), da = __fsub_rn(qa[qi], ma), db = __fsub_rn(qb[qi], mb);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db));
            const unsigned bits = __float_as_uint(d2);
            const bool pass = in_range && bits <= tau[qi];
            const unsigned mask = __ballot_sync(0xffffffffu, pass);
            if (mask) {
              if (pass)
                st[qi].buf[cnt[qi] + __popc(mask & lt_mask)] = ((unsigned long long)bits << 32) | (base_idx + j);
              cnt[qi] += __popc(mask);
              if (cnt[qi] > TK_FLUSH) merge_buffer(st[qi], cnt[qi], tau[qi], k, lane);
            }
          }
        }
      }
      __syncthreads();  // every warp is done with stage s
      if (tid == 0 && g + 2 < total_chunks) issue(g + 2);
    }

    if (warp_active) {
#pragma unroll
      for (int qi = 0; qi < TK_Q; ++qi) {
        if (cnt[qi] > 0) merge_buffer(st[qi], cnt[qi], tau[qi], k, lane);
        const int cell = first_cell + qi;
        if (cell < cell_end) {
          for (int e = lane; e < k; e += 32) {
            unsigned long long key = st[qi].topk[e];
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
  const int per_cta = TK_WARPS * TK_Q * rounds;
  const int blocks = (cells + per_cta - 1) / per_cta;
  {
    StageTimer t(ctx, MOSAIC_STAGE_TOPK);
    nn_topk_kernel<<<blocks, TK_WARPS * 32, kTopkSmem, ctx->stream>>>(ctx->d_mat_soa, ctx->d_mat_lab4, ctx->n, ctx->n_pad,
                                                                      ctx->d_cell_lab, cell_begin, cell_end, k, rounds,
                                                                      ctx->cand_dst_idx ? ctx->cand_dst_idx : ctx->d_cand_idx,
                                                                      ctx->cand_dst_d2 ? ctx->cand_dst_d2 : ctx->d_cand_d2);
    ctx->launches++;
  }
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace mosaic
