// anneal.cu — K5: simulated annealing of the placement as ONE persistent kernel.
//
// Replaces MosaicOptimizer::optimize (reference src/optimizer.rs:41-129) with
// AdjacencyPenaltyCalculator::calculate_total_cost / calculate_swap_delta (adjacency.rs:91-214).
// The Markov chain is serial, so it runs as a single warp that never leaves the SM:
//   * RNG: the reference draws from rand::thread_rng() (optimizer.rs:48), which is OS-seeded; the
//     stream is therefore DEFINED as rand 0.8.5 StdRng::seed_from_u64(seed) = ChaCha12 (rand_chacha
//     0.3.1: 64-bit block counter, stream 0, key = PCG32 expansion of the seed).  The 32 lanes each
//     compute one ChaCha12 block (32 blocks = 512 words per refill) into a shared-memory ring; the
//     chain consumes words exactly like BlockRng: next_u64 = two consecutive words (low first),
//     gen_range(0..n) = widening-multiply rejection sampling, gen::<f32>() = (next_u32 >> 8) * 2^-24.
//   * swap delta: the <= 8 neighbour terms 1/(sim+1) for the old and the new assignment are
//     evaluated in parallel across lanes, then summed in the reference's order (pos1: Up, Down,
//     Left, Right skipping pos2; pos2 likewise skipping pos1; then the mutual term) so the f32 sums
//     are bit-identical.  sim is recomputed from the Lab means exactly like the packed matrix entry.
//   * acceptance: exp(-delta / T) uses a deterministic double-precision exp (argument reduction +
//     degree-13 Horner, unfused) rounded once to f32 — the same operation sequence as the oracle.
//   * initial cost: calculate_total_cost's raster-order f32 sum is evaluated in 32-pair batches
//     (terms in parallel, accumulation in order).
#include <algorithm>

#include "common.cuh"

namespace mosaic {

namespace {

constexpr int RNG_HALF_WORDS = 512;   // 32 ChaCha blocks: one refill
constexpr int RNG_RING_WORDS = 1024;  // two halves
constexpr int RNG_SLACK = 16;         // words one iteration may look ahead (8 for the positions + 1 for the f32)
constexpr int RNG_SMEM_WORDS = RNG_RING_WORDS + 16;  // + the key (8 words), padded

__device__ __forceinline__ unsigned rotl32(unsigned v, int c) { return __funnelshift_l(v, v, c); }

#define MOSAIC_QR(a, b, c, d) \
  a += b; d ^= a; d = rotl32(d, 16); \
  c += d; b ^= c; b = rotl32(b, 12); \
  a += b; d ^= a; d = rotl32(d, 8);  \
  c += d; b ^= c; b = rotl32(b, 7);

// 32 ChaCha12 blocks (one per lane) -> ring[0 .. 512); the key sits behind the ring (ring[512 .. 520)).
// Out of line on purpose: inlined at every draw it made the kernel 190 KB of code, and the chain's loop
// has to live in the instruction caches (L0 ~6 KB, L1.5 32 KB).
__device__ __noinline__ void chacha12_refill(unsigned* half, const unsigned* key, unsigned long long first_block, int lane) {
  const unsigned long long counter = first_block + (unsigned long long)lane;
  unsigned* out = half + 16 * lane;
  unsigned s0 = 0x61707865u, s1 = 0x3320646eu, s2 = 0x79622d32u, s3 = 0x6b206574u;
  unsigned s4 = key[0], s5 = key[1], s6 = key[2], s7 = key[3], s8 = key[4], s9 = key[5], s10 = key[6], s11 = key[7];
  unsigned s12 = (unsigned)counter, s13 = (unsigned)(counter >> 32), s14 = 0u, s15 = 0u;
  unsigned x0 = s0, x1 = s1, x2 = s2, x3 = s3, x4 = s4, x5 = s5, x6 = s6, x7 = s7, x8 = s8, x9 = s9, x10 = s10,
           x11 = s11, x12 = s12, x13 = s13, x14 = s14, x15 = s15;
  __syncwarp();
#pragma unroll 1
  for (int r = 0; r < 6; ++r) {
    MOSAIC_QR(x0, x4, x8, x12) MOSAIC_QR(x1, x5, x9, x13) MOSAIC_QR(x2, x6, x10, x14) MOSAIC_QR(x3, x7, x11, x15)
    MOSAIC_QR(x0, x5, x10, x15) MOSAIC_QR(x1, x6, x11, x12) MOSAIC_QR(x2, x7, x8, x13) MOSAIC_QR(x3, x4, x9, x14)
  }
  out[0] = x0 + s0;   out[1] = x1 + s1;   out[2] = x2 + s2;    out[3] = x3 + s3;
  out[4] = x4 + s4;   out[5] = x5 + s5;   out[6] = x6 + s6;    out[7] = x7 + s7;
  out[8] = x8 + s8;   out[9] = x9 + s9;   out[10] = x10 + s10; out[11] = x11 + s11;
  out[12] = x12 + s12; out[13] = x13 + s13; out[14] = x14 + s14; out[15] = x15 + s15;
  __syncwarp();
}

// The word stream of rand 0.8.5 StdRng (BlockRng over ChaCha12): words are consumed strictly in order; a u64 is
// two consecutive words, low first.  The ring holds two halves of 32 blocks; `pos` / `filled` are stream word
// indices (mod 2^32), and a half is regenerated only when every word in it lies >= 512 - RNG_SLACK words behind `pos`.
struct Rng {
  unsigned long long next_block;  // counter of the first block of the next refill
  unsigned* ring;                 // shared memory: RNG_RING_WORDS words, then the key (8 words)
  unsigned pos, filled;

  __device__ __forceinline__ void ensure(int words, int lane) {  // words <= RNG_HALF_WORDS
    if ((int)(filled - pos) < words) {
      chacha12_refill(ring + (filled & (RNG_RING_WORDS - 1)), ring + RNG_RING_WORDS, next_block, lane);
      next_block += 32ull;
      filled += RNG_HALF_WORDS;
    }
  }
  __device__ __forceinline__ unsigned peek(int i) const { return ring[(pos + (unsigned)i) & (RNG_RING_WORDS - 1)]; }
  // the general (sequential, rejection-aware) draws: only the cold paths use them
  __device__ __forceinline__ unsigned next_u32(int lane) {
    ensure(1, lane);
    const unsigned v = peek(0);
    pos += 1u;
    return v;
  }
  __device__ __forceinline__ unsigned long long next_u64(int lane) {
    const unsigned lo = next_u32(lane);
    const unsigned hi = next_u32(lane);
    return ((unsigned long long)hi << 32) | lo;
  }
  __device__ __forceinline__ unsigned long long gen_range_slow(unsigned long long n, unsigned long long zone, int lane) {
    // UniformInt<usize>::sample_single: zone = (n << n.leading_zeros()) - 1, hoisted by the caller
    for (;;) {
      const unsigned long long v = next_u64(lane);
      const unsigned long long hi = __umul64hi(v, n), lo = v * n;
      if (lo <= zone) return hi;
    }
  }
};

// deterministic f32 exp: identical operation sequence to the oracle (all IEEE double, unfused)
__device__ __forceinline__ float exp_det(float xf) {
  if (xf != xf) return xf;
  if (xf > 88.8f) return __int_as_float(0x7f800000);
  if (xf < -104.0f) return 0.0f;
  const double x = (double)xf;
  const double kd = floor(__dadd_rn(__dmul_rn(x, 1.4426950408889634), 0.5));
  double r = __dsub_rn(x, __dmul_rn(kd, 0.693147180369123816490e+00));
  r = __dsub_rn(r, __dmul_rn(kd, 1.90821492927058770002e-10));
  const double c[14] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
                        1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
                        1.0 / 479001600, 1.0 / 6227020800.0};
  double p = c[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = __dadd_rn(__dmul_rn(p, r), c[i]);
  const int k = (int)kd;
  const double scale = __longlong_as_double((long long)(k + 1023) << 52);
  return __double2float_rn(__dmul_rn(p, scale));
}

// ---- cold paths, out of line so that the chain's loop stays contiguous in the instruction cache ----------
struct SlowPositions {
  int x1, y1, x2, y2;
  unsigned pos, filled;
  unsigned long long next_block;
};
// the four position draws when one of them is rejected (UniformInt<usize>::sample_single's loop, optimizer.rs:60-63)
__device__ __noinline__ SlowPositions positions_slow(unsigned* ring, unsigned pos, unsigned filled,
                                                     unsigned long long next_block, unsigned long long gw,
                                                     unsigned long long gh, unsigned long long zone_w,
                                                     unsigned long long zone_h, int lane) {
  Rng rng;
  rng.next_block = next_block;
  rng.ring = ring;
  rng.pos = pos;
  rng.filled = filled;
  SlowPositions o;
  o.x1 = (int)rng.gen_range_slow(gw, zone_w, lane);
  o.y1 = (int)rng.gen_range_slow(gh, zone_h, lane);
  o.x2 = (int)rng.gen_range_slow(gw, zone_w, lane);
  o.y2 = (int)rng.gen_range_slow(gh, zone_h, lane);
  rng.ensure(1, lane);  // the f32 the caller may draw next
  o.pos = rng.pos;
  o.filled = rng.filled;
  o.next_block = rng.next_block;
  return o;
}
// u < exp(-delta / T) with the deterministic exponential (optimizer.rs:82-83)
__device__ __noinline__ bool accept_exact(float delta, float temperature, float u) {
  return u < exp_det(__fdiv_rn(-delta, temperature));
}

// exact builtins for inputs outside the domain of the branch-free sequences (out of line: cold)
__device__ __noinline__ float adjacency_term_exact(float l1, float a1, float b1, float l2, float a2, float b2) {
  return adjacency_term(lab_distance_rn(l1, a1, b1, l2, a2, b2));
}

// ---- initial cost: all pair terms in parallel, summed later in the reference's order ----------
// terms[2c] = right pair of cell c, terms[2c+1] = down pair (adjacency.rs:104-124); absent pairs are +0.0,
// which the sequential f32 sum absorbs exactly (x + 0.0 == x).
__global__ void __launch_bounds__(256) cost_terms_kernel(const float4* __restrict__ lab4, const int32_t* __restrict__ grid,
                                                         int gw, int gh, float* __restrict__ terms) {
  const long long pairs = 2ll * gw * gh;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < pairs; s += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(s >> 1), x = c % gw, y = c / gw;
    const int cur = grid[c];
    int other = -1;
    if (cur >= 0) {
      if ((s & 1) == 0) {
        if (x < gw - 1) other = grid[c + 1];
      } else if (y < gh - 1) {
        other = grid[c + gw];
      }
    }
    float t = 0.0f;
    if (other >= 0) {
      // get_similarity (similarity.rs:85-103): same index -> 0.0, else the packed entry of (min, max)
      float sim = 0.0f;
      if (other != cur) {
        const float4 p = __ldg(&lab4[min(cur, other)]), q = __ldg(&lab4[max(cur, other)]);
        sim = lab_distance_rn(p.x, p.y, p.z, q.x, q.y, q.z);
      }
      t = adjacency_term(sim);
    }
    terms[s] = t;
  }
}

// Where the chain's working set lives: the grid as u16 / i32 in shared memory (or in HBM when it does not
// fit), the material Lab means as SoA in shared memory (or the float4 table in HBM/L2).
template <typename GridT, bool GRID_SMEM, bool LAB_SMEM>
struct AnnealMem {
  volatile GridT* grid_s;
  volatile int32_t* grid_g;
  const float* lab_s;  // l[n], a[n], b[n]
  const float4* lab_g;
  int n;
  __device__ __forceinline__ int get(int i) const {
    if (GRID_SMEM) {
      const GridT v = grid_s[i];
      return (sizeof(GridT) == 2 && v == (GridT)0xFFFF) ? -1 : (int)v;
    }
    return grid_g[i];
  }
  __device__ __forceinline__ void set(int i, int v) const {
    if (GRID_SMEM)
      grid_s[i] = (GridT)v;
    else
      grid_g[i] = v;
  }
  __device__ __forceinline__ void lab(int m, float& l, float& a, float& b) const {
    if (LAB_SMEM) {
      l = lab_s[m];
      a = lab_s[n + m];
      b = lab_s[2 * n + m];
    } else {
      const float4 v = __ldg(&lab_g[m]);
      l = v.x;
      a = v.y;
      b = v.z;
    }
  }
};

template <typename GridT, bool GRID_SMEM, bool LAB_SMEM>
__global__ void __launch_bounds__(256, 1)
    anneal_kernel(const float4* __restrict__ lab4, int n, int32_t* grid_gmem, const float* __restrict__ terms, int gw,
                  int gh, unsigned long long iterations, float t0, float decay, unsigned long long seed,
                  mosaic_opt_result* res, float* __restrict__ deltas /* one slot per iteration, or NULL */) {
  __shared__ float s_initial;
  __shared__ volatile int s_initial_ready;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* ring = reinterpret_cast<unsigned*>(smem_raw);
  float* lab_s = reinterpret_cast<float*>(smem_raw + RNG_SMEM_WORDS * 4);
  GridT* grid_s = reinterpret_cast<GridT*>(smem_raw + RNG_SMEM_WORDS * 4 + (LAB_SMEM ? (size_t)n * 12 : 0));
  const int tid = threadIdx.x, lane = tid & 31;
  const int cells = gw * gh;
  if (GRID_SMEM)
    for (int i = tid; i < cells; i += blockDim.x) grid_s[i] = (GridT)grid_gmem[i];  // -1 -> 0xFFFF for u16
  if (LAB_SMEM)
    for (int i = tid; i < n; i += blockDim.x) {
      const float4 v = __ldg(&lab4[i]);
      lab_s[i] = v.x;
      lab_s[n + i] = v.y;
      lab_s[2 * n + i] = v.z;
    }
  if (tid == 0) s_initial_ready = 0;
  __syncthreads();

  // ---- calculate_total_cost (adjacency.rs:91-130): raster order, right then down, sequential f32 adds ----
  // The chain is one FADD per term (4 cycles).  Each lane fetches four consecutive terms (LDG.128, two batches
  // of 128 terms in flight to cover the L2 latency); they are handed to the chain in order by shuffles.
  auto total_cost = [&]() -> float {
    float acc = 0.0f;
    const long long pairs = 2ll * cells;
    const long long nvec = pairs >> 2;  // `terms` is 16-byte aligned (cudaMalloc)
    const float4* t4 = reinterpret_cast<const float4*>(terms);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);  // absent terms are +0.0: x + 0.0 == x exactly
    float4 n0 = lane < nvec ? __ldg(t4 + lane) : zero4;
    float4 n1 = 32 + lane < nvec ? __ldg(t4 + 32 + lane) : zero4;
    for (long long base = 0; base < nvec; base += 32) {
      const float4 cur = n0;
      n0 = n1;
      n1 = base + 64 + lane < nvec ? __ldg(t4 + base + 64 + lane) : zero4;
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, cur.x, l));
        acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, cur.y, l));
        acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, cur.z, l));
        acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, cur.w, l));
      }
    }
    for (long long s = nvec << 2; s < pairs; ++s) acc = __fadd_rn(acc, __ldg(terms + s));
    return acc;
  };
  // The acceptance rule never looks at the running cost (optimizer.rs:80-84), only the bookkeeping does
  // (optimizer.rs:86-99).  So, given a delta log, the 2*cells-term serial sum runs on warp 1 WHILE warp 0 walks the
  // Markov chain; warp 0 replays `current += delta` / best / improved afterwards, in the original order.
  const bool deferred = deltas != nullptr;
  if (tid >= 32) {
    if (deferred && tid < 64) {
      const float v = total_cost();
      if (lane == 0) {
        s_initial = v;
        __threadfence_block();
        s_initial_ready = 1;
      }
    }
    return;  // the Markov chain is one warp
  }

  AnnealMem<GridT, GRID_SMEM, LAB_SMEM> mem;
  mem.grid_s = grid_s;
  mem.grid_g = grid_gmem;
  mem.lab_s = lab_s;
  mem.lab_g = lab4;
  mem.n = n;

  float current = deferred ? 0.0f : total_cost();
  float initial = current;
  float best = current;
  unsigned long long improved = 0, accepted = 0;
  float temperature = t0;

  Rng rng;
  {
    unsigned long long state = seed;  // rand_core seed_from_u64: PCG32 stream fills the 32-byte key
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      const unsigned xorshifted = (unsigned)(((state >> 18) ^ state) >> 27);
      const unsigned rot = (unsigned)(state >> 59);
      ring[RNG_RING_WORDS + i] = __funnelshift_r(xorshifted, xorshifted, rot);  // (every lane writes the same words)
    }
    __syncwarp();
    rng.next_block = 0ull;
    rng.ring = ring;
    rng.pos = 0u;
    rng.filled = 0u;
  }
  const unsigned long long ugw = (unsigned long long)gw, ugh = (unsigned long long)gh;
  const unsigned long long zone_w = (ugw << __clzll((long long)gw)) - 1ull;
  const unsigned long long zone_h = (ugh << __clzll((long long)gh)) - 1ull;

  // The loop is written for ONE warp on its own: a branch costs it about as much as the arithmetic of a whole
  // swap delta, so everything that can be a select is a select and the rare cases leave through cold paths.
  for (unsigned long long it = 0; it < iterations; ++it) {
    rng.ensure(RNG_SLACK, lane);
    // optimizer.rs:60-63: x1, y1, x2, y2 — four u64 draws, no rejection in all but ~n/2^64 of the cases.
    // gen_range(0..n): (hi, lo) = v * n (128-bit), accepted iff lo <= zone, result hi.  n < 2^31, so the product
    // is two 32x32->64 multiplies.
    int xy[4];
    bool no_reject = true;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned nq = (q & 1) ? (unsigned)gh : (unsigned)gw;
      const unsigned long long zq = (q & 1) ? zone_h : zone_w;
      const unsigned long long t0 = (unsigned long long)rng.peek(2 * q) * nq;
      const unsigned long long t1 = (unsigned long long)rng.peek(2 * q + 1) * nq + (t0 >> 32);
      const unsigned long long lo = (t1 << 32) | (t0 & 0xFFFFFFFFull);
      xy[q] = (int)(t1 >> 32);
      no_reject = no_reject && lo <= zq;
    }
    int x1 = xy[0], y1 = xy[1], x2 = xy[2], y2 = xy[3];
    if (__builtin_expect(!no_reject, 0)) {
      const SlowPositions sp = positions_slow(ring, rng.pos, rng.filled, rng.next_block, ugw, ugh, zone_w, zone_h, lane);
      x1 = sp.x1, y1 = sp.y1, x2 = sp.x2, y2 = sp.y2;
      rng.pos = sp.pos;
      rng.filled = sp.filled;
      rng.next_block = sp.next_block;
    } else {
      rng.pos += 8u;
    }
    const int i1 = y1 * gw + x1, i2 = y2 * gw + x2;
    const int p1 = mem.get(i1), p2 = mem.get(i2);
    // lanes 0-3: neighbours of pos1 (Up, Down, Left, Right); lanes 4-7: neighbours of pos2
    const int side = (lane >> 2) & 1, dir = lane & 3;
    const int cx = side ? x2 : x1, cy = side ? y2 : y1;
    const int ox = side ? x1 : x2, oy = side ? y1 : y2;  // the other swap position (skipped)
    const int nx = cx + (dir == 3) - (dir == 2), ny = cy + (dir == 1) - (dir == 0);
    const bool inside = lane < 8 && nx >= 0 && nx < gw && ny >= 0 && ny < gh && !(nx == ox && ny == oy);
    const int nb = mem.get(inside ? ny * gw + nx : i1);  // (a valid address either way; ignored when !inside)
    // optimizer.rs:66-68 (same position: no temperature decay) and 71-73 (an empty cell)
    if (__builtin_expect((x1 == x2 && y1 == y2) || p1 < 0 || p2 < 0, 0)) continue;

    const bool present = inside && nb >= 0;
    float l1, a1, b1, l2, a2, b2, ln, an, bn;
    mem.lab(p1, l1, a1, b1);
    mem.lab(p2, l2, a2, b2);
    mem.lab(present ? nb : p1, ln, an, bn);
    const float ml = side ? l2 : l1, ma = side ? a2 : a1, mb = side ? b2 : b1;  // material now at this position
    const float ol = side ? l1 : l2, oa = side ? a1 : a2, ob = side ? b1 : b2;  // material after the swap
    bool oor = false, oor_m = false;
    float t_old = adjacency_term_fast(ml, ma, mb, ln, an, bn, oor);
    float t_new = adjacency_term_fast(ol, oa, ob, ln, an, bn, oor);
    float mutual = adjacency_term_fast(l1, a1, b1, l2, a2, b2, oor_m);  // warp-uniform
    const bool adjacent12 = (x1 == x2 && (y1 - y2 == 1 || y2 - y1 == 1)) || (y1 == y2 && (x1 - x2 == 1 || x2 - x1 == 1));
    if (__builtin_expect(__any_sync(0xffffffffu, (oor && present) || (oor_m && adjacent12)), 0)) {
      // outside the domain of the branch-free sequences: the exact builtins
      t_old = adjacency_term_exact(ml, ma, mb, ln, an, bn);
      t_new = adjacency_term_exact(ol, oa, ob, ln, an, bn);
      mutual = adjacency_term_exact(l1, a1, b1, l2, a2, b2);
    }
    // same material on both sides: get_similarity returns 0.0 -> term 1.0; sqrt(0) gives the same
    t_old = present ? t_old : 0.0f;  // absent neighbours contribute +0.0 (exact identity in the sums below)
    t_new = present ? t_new : 0.0f;
    float old_cost = 0.0f, new_cost = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) {  // reference order: pos1 U,D,L,R then pos2 U,D,L,R (adjacency.rs:165-200)
      old_cost = __fadd_rn(old_cost, __shfl_sync(0xffffffffu, t_old, l));
      new_cost = __fadd_rn(new_cost, __shfl_sync(0xffffffffu, t_new, l));
    }
    mutual = adjacent12 ? mutual : 0.0f;       // adjacency.rs:203-211 (x + 0.0 == x: the sums are >= +0.0)
    old_cost = __fadd_rn(old_cost, mutual);
    new_cost = __fadd_rn(new_cost, mutual);
    const float delta = p1 != p2 ? __fsub_rn(new_cost, old_cost) : 0.0f;  // adjacency.rs:156-158

    // accept <=> delta < 0 || u < exp(-delta / T), u drawn only when delta >= 0 (optimizer.rs:80-84).
    // 1 - q <= exp(-q) <= 1 / (1 + q) (q >= 0) decide almost every draw without the division and the
    // exponential; the margins (2^-18) dwarf every rounding error involved.
    const bool downhill = delta < 0.0f;
    const float u = __fmul_rn((float)(rng.peek(0) >> 8), 1.0f / 16777216.0f);
    rng.pos += downhill ? 0u : 1u;
    float rT, hi;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rT) : "f"(temperature));
    const float q = delta * rT;
    const float lo = 1.0f - q * 1.000004f - 3.8146973e-06f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(hi) : "f"(1.0f + q * 0.999996f));
    hi += 3.8146973e-06f;
    const bool bounds_ok = q >= 0.0f;  // the bounds hold for q >= 0 (T > 0); NaN / negative -> exact path
    bool accept = downhill || (bounds_ok && u < lo);
    if (__builtin_expect(!accept && !(bounds_ok && u > hi), 0)) {
      accept = accept_exact(delta, temperature, u);
    }
    // (every lane stores the same values: no branch, no divergence)
    mem.set(i1, accept ? p2 : p1);
    mem.set(i2, accept ? p1 : p2);
    __syncwarp();
    if (deferred) {
      if (accept) deltas[accepted] = delta;  // (every lane: same address, same value)
    } else {
      current = accept ? __fadd_rn(current, delta) : current;
      const bool better = accept && current < best;
      best = better ? current : best;
      improved += better ? 1ull : 0ull;
    }
    accepted += accept ? 1ull : 0ull;
    temperature = __fmul_rn(temperature, decay);     // optimizer.rs:102
  }

  if (GRID_SMEM) {
    __syncwarp();
    for (int i = lane; i < cells; i += 32) grid_gmem[i] = mem.get(i);
  }
  if (deferred) {  // optimizer.rs:86-99 replayed over the logged deltas
    while (s_initial_ready == 0) {
    }
    __threadfence_block();
    initial = s_initial;
    current = initial;
    best = initial;
    __syncwarp();
    for (unsigned long long base = 0; base < accepted; base += 32) {
      const float d = base + lane < accepted ? deltas[base + lane] : 0.0f;
      const int cnt = (int)min(32ull, accepted - base);
      for (int l = 0; l < cnt; ++l) {
        current = __fadd_rn(current, __shfl_sync(0xffffffffu, d, l));
        if (current < best) {
          best = current;
          improved += 1ull;
        }
      }
    }
  }
  if (lane == 0) {
    res->initial_cost = initial;
    res->final_cost = current;
    res->best_cost = best;
    res->improved_count = improved;
    res->accepted_count = accepted;
    res->iterations = iterations;  // optimizer.rs:127
  }
}

}  // namespace

template <typename GridT, bool GS, bool LS>
static int launch_anneal_variant(mosaic_ctx* ctx, size_t smem, uint64_t iterations, float t0, float decay, uint64_t seed) {
  MOSAIC_CUDA(ctx, cudaFuncSetAttribute(anneal_kernel<GridT, GS, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // delta log: one float per iteration (up to 256 MB); beyond that the bookkeeping stays inside the loop
  float* deltas = nullptr;
  if (iterations > 0 && iterations <= (1ull << 26)) {
    MOSAIC_TRY(ensure(ctx, &ctx->d_deltas, &ctx->deltas_cap, (size_t)iterations));
    deltas = ctx->d_deltas;
  }
  anneal_kernel<GridT, GS, LS><<<1, 256, smem, ctx->stream>>>(ctx->d_mat_lab4, ctx->n, ctx->d_placed, ctx->d_terms,
                                                              ctx->gw, ctx->gh, iterations, t0, decay, seed,
                                                              (mosaic_opt_result*)ctx->d_opt_res, deltas);
  return MOSAIC_OK;
}

int launch_anneal(mosaic_ctx* ctx, uint64_t iterations, float t0, float decay, uint64_t seed) {
  const size_t cells = (size_t)ctx->gw * ctx->gh;
  const size_t n = (size_t)ctx->n;
  const size_t ring_bytes = RNG_SMEM_WORDS * 4;
  const size_t limit = 220 * 1024;
  MOSAIC_TRY(ensure(ctx, &ctx->d_terms, &ctx->terms_cap, 2 * cells));
  StageTimer t(ctx, MOSAIC_STAGE_ANNEAL);
  {
    int blocks = (int)std::min<size_t>((2 * cells + 255) / 256, (size_t)ctx->sm_count * 8);
    if (blocks < 1) blocks = 1;
    cost_terms_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_mat_lab4, ctx->d_placed, ctx->gw, ctx->gh, ctx->d_terms);
    ctx->launches++;
    MOSAIC_CUDA(ctx, cudaGetLastError());
  }
  const bool g16 = n <= 65534;
  const size_t grid_bytes = cells * (g16 ? 2 : 4);
  const size_t lab_bytes = n * 12;
  // preference: grid + Lab in shared memory; then grid only; then everything in HBM/L2
  if (ring_bytes + grid_bytes + lab_bytes <= limit) {
    if (g16)
      MOSAIC_TRY((launch_anneal_variant<uint16_t, true, true>(ctx, ring_bytes + grid_bytes + lab_bytes, iterations, t0, decay, seed)));
    else
      MOSAIC_TRY((launch_anneal_variant<int32_t, true, true>(ctx, ring_bytes + grid_bytes + lab_bytes, iterations, t0, decay, seed)));
  } else if (ring_bytes + grid_bytes <= limit) {
    if (g16)
      MOSAIC_TRY((launch_anneal_variant<uint16_t, true, false>(ctx, ring_bytes + grid_bytes, iterations, t0, decay, seed)));
    else
      MOSAIC_TRY((launch_anneal_variant<int32_t, true, false>(ctx, ring_bytes + grid_bytes, iterations, t0, decay, seed)));
  } else {
    MOSAIC_TRY((launch_anneal_variant<int32_t, false, false>(ctx, ring_bytes, iterations, t0, decay, seed)));
  }
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace mosaic
