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
constexpr int RNG_SLACK = 128;        // words one round may look ahead (a 65-word sampling window + an acceptance sample behind each iteration)
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
  // predicated store (no branch): the loop of the chain must stay free of them
  __device__ __forceinline__ void set_if(bool pred, int i, int v) const {
    if (GRID_SMEM) {
      if (sizeof(GridT) == 2)
        st_shared_if(pred, reinterpret_cast<unsigned short*>(const_cast<GridT*>(grid_s)) + i, (unsigned short)v);
      else
        st_shared_if(pred, reinterpret_cast<unsigned*>(const_cast<GridT*>(grid_s)) + i, (unsigned)v);
    } else {
      st_global_if(pred, reinterpret_cast<unsigned*>(const_cast<int32_t*>(grid_g)) + i, (unsigned)v);
    }
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
                  mosaic_opt_result* res, float* __restrict__ deltas /* one slot per iteration, or NULL */,
                  bool greedy /* optimize_greedy (optimizer.rs:132-198): only improvements, no acceptance sample */) {
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
  //
  // A swap delta only occupies 8 lanes (one per neighbour term), and the chain is pure latency.  So the warp runs
  // THREE slots of 8 lanes per round: slot 0 evaluates the next iteration at its known stream position; slots 1 and
  // 2 evaluate the iteration after it at the two positions it can start from (right behind this one's position
  // draws if this one draws no acceptance sample, one word later if it does).  The round then retires the two
  // iterations strictly in order: each one's acceptance uses the real temperature and the real stream position,
  // and the speculated delta is used only if no swap accepted by the first iteration touched a cell the second one
  // read (the two swapped cells against its two cells and their neighbours).  Every retired iteration is therefore
  // computed from exactly the state the serial loop would have seen — same arithmetic, same order — and a round
  // retires two iterations for the latency of one.  The retirement is straight-line code (selects and predicated
  // stores): a branch costs this lone warp ~50 cycles.
  //
  // Position draws (optimizer.rs:60-63): gen_range(0..n) = Lemire's widening multiply with rejection, (hi, lo) =
  // v * n accepted iff lo <= zone — for n = 200 more than a fifth of the samples are rejected, so rejection is the
  // normal case, not a cold path.  Each lane takes the u64 at an even and at an odd word offset of a 65-word
  // window and evaluates both for both ranges (n < 2^31: two 32x32->64 multiplies each); four ballots later, the
  // sequential x1, y1, x2, y2 sampling of this iteration and of both candidates for the next one is a walk over
  // bit masks.  A window that runs out (never, in practice) takes the general sampler.
  const int slot = lane >> 3, sl = lane & 7, slot_base = lane & 24;
  const int side = (sl >> 2) & 1, dir = sl & 3;  // lanes 0-3 of a slot: neighbours of pos1 (U, D, L, R); 4-7: of pos2
  struct Walk {
    int j0, j1, j2, j3;
    bool ok;
  };
  auto walk = [](unsigned mw, unsigned mh, unsigned startbit) -> Walk {  // x1 (w), y1 (h), x2 (w), y2 (h), in stream order
    Walk k;
    unsigned m = startbit < 32u ? mw & ~((1u << startbit) - 1u) : 0u;
    k.ok = m != 0u;
    k.j0 = m ? __ffs(m) - 1 : 0;
    m = mh & ~((2u << k.j0) - 1u);
    k.ok = k.ok && m != 0u;
    k.j1 = m ? __ffs(m) - 1 : 0;
    m = mw & ~((2u << k.j1) - 1u);
    k.ok = k.ok && m != 0u;
    k.j2 = m ? __ffs(m) - 1 : 0;
    m = mh & ~((2u << k.j2) - 1u);
    k.ok = k.ok && m != 0u;
    k.j3 = m ? __ffs(m) - 1 : 0;
    return k;
  };
  auto near = [](int ax, int ay, int bx, int by) { return abs(ax - bx) + abs(ay - by) <= 1; };
  auto accept_bounds = [](float delta, float temperature, float u, bool& undecided) -> bool {
    // accept <=> delta < 0 || u < exp(-delta / T) (optimizer.rs:80-84).  1 - q <= exp(-q) <= 1 / (1 + q) (q >= 0)
    // decide almost every draw without the division and the exponential; the margins (2^-18) dwarf every
    // rounding error involved.  `undecided`: the sliver in between (or q < 0 / NaN) needs the exact exponential.
    float rT, hi;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rT) : "f"(temperature));
    const float q = delta * rT;
    const float lo = 1.0f - q * 1.000004f - 3.8146973e-06f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(hi) : "f"(1.0f + q * 0.999996f));
    hi += 3.8146973e-06f;
    const bool bounds_ok = q >= 0.0f;
    const bool acc = delta < 0.0f || (bounds_ok && u < lo);
    undecided = !acc && !(bounds_ok && u > hi);
    return acc;
  };
  unsigned long long it = 0;
  while (it < iterations) {
    rng.ensure(RNG_SLACK, lane);
    const unsigned basepos = rng.pos;
    // ---- candidate samples of the window: lane j holds the u64 at words 2j (even set) and 2j+1 (odd set) ----
    const unsigned w0 = ring[(basepos + 2u * lane) & (RNG_RING_WORDS - 1)], w1 = ring[(basepos + 2u * lane + 1u) & (RNG_RING_WORDS - 1)],
                   w2 = ring[(basepos + 2u * lane + 2u) & (RNG_RING_WORDS - 1)];
    int hiw_e, hih_e, hiw_o, hih_o;
    bool okw_e, okh_e, okw_o, okh_o;
    {
      const unsigned long long a0 = (unsigned long long)w0 * (unsigned)gw, a1 = (unsigned long long)w1 * (unsigned)gw + (a0 >> 32);
      const unsigned long long b0 = (unsigned long long)w0 * (unsigned)gh, b1 = (unsigned long long)w1 * (unsigned)gh + (b0 >> 32);
      okw_e = ((a1 << 32) | (a0 & 0xFFFFFFFFull)) <= zone_w;
      okh_e = ((b1 << 32) | (b0 & 0xFFFFFFFFull)) <= zone_h;
      hiw_e = (int)(a1 >> 32);
      hih_e = (int)(b1 >> 32);
      const unsigned long long c0 = (unsigned long long)w1 * (unsigned)gw, c1 = (unsigned long long)w2 * (unsigned)gw + (c0 >> 32);
      const unsigned long long d0 = (unsigned long long)w1 * (unsigned)gh, d1 = (unsigned long long)w2 * (unsigned)gh + (d0 >> 32);
      okw_o = ((c1 << 32) | (c0 & 0xFFFFFFFFull)) <= zone_w;
      okh_o = ((d1 << 32) | (d0 & 0xFFFFFFFFull)) <= zone_h;
      hiw_o = (int)(c1 >> 32);
      hih_o = (int)(d1 >> 32);
    }
    const unsigned EW = __ballot_sync(0xffffffffu, okw_e), EH = __ballot_sync(0xffffffffu, okh_e);
    const unsigned OW = __ballot_sync(0xffffffffu, okw_o), OH = __ballot_sync(0xffffffffu, okh_o);
    const Walk k0 = walk(EW, EH, 0u);                       // this iteration
    unsigned adv0 = 2u * (unsigned)(k0.j3 + 1);              // words its position draws consume
    const unsigned sb = (unsigned)(k0.j3 + 1);               // the next iteration starts at word adv0 (even set, bit sb) ...
    const Walk ka = walk(EW, EH, sb), kb = walk(OW, OH, sb);  // ... or at word adv0 + 1 (odd set, bit sb)
    // every lane fetches the positions of ITS slot: slot 0 <- k0 (even set), slot 1 <- ka (even), slot 2 <- kb (odd)
    const Walk& mine = slot == 0 ? k0 : (slot == 1 ? ka : kb);
    const int e1 = __shfl_sync(0xffffffffu, hiw_e, mine.j0), o1 = __shfl_sync(0xffffffffu, hiw_o, mine.j0);
    const int e2 = __shfl_sync(0xffffffffu, hih_e, mine.j1), o2 = __shfl_sync(0xffffffffu, hih_o, mine.j1);
    const int e3 = __shfl_sync(0xffffffffu, hiw_e, mine.j2), o3 = __shfl_sync(0xffffffffu, hiw_o, mine.j2);
    const int e4 = __shfl_sync(0xffffffffu, hih_e, mine.j3), o4 = __shfl_sync(0xffffffffu, hih_o, mine.j3);
    int x1 = slot == 2 ? o1 : e1, y1 = slot == 2 ? o2 : e2, x2 = slot == 2 ? o3 : e3, y2 = slot == 2 ? o4 : e4;
    unsigned my_adv = slot == 0 ? adv0 : 2u * (unsigned)(mine.j3 + 1) - adv0;
    bool inval = slot == 3 || (slot != 0 && !mine.ok);  // this slot's speculation cannot be used (slot 3 is idle)
    if (__builtin_expect(!k0.ok, 0)) {  // the window ran out: the general sampler for this iteration, no speculation
      const SlowPositions sp = positions_slow(ring, basepos, rng.filled, rng.next_block, ugw, ugh, zone_w, zone_h, lane);
      if (slot == 0) x1 = sp.x1, y1 = sp.y1, x2 = sp.x2, y2 = sp.y2;
      adv0 = sp.pos - basepos;
      if (slot == 0) my_adv = adv0;
      inval = slot != 0;
      rng.filled = sp.filled;
      rng.next_block = sp.next_block;
      rng.pos = sp.pos;
      rng.ensure(RNG_SLACK, lane);
      rng.pos = basepos;
    }

    // ---- the swap deltas of the three slots ----
    const int i1 = y1 * gw + x1, i2 = y2 * gw + x2;
    const int p1 = mem.get(i1), p2 = mem.get(i2);
    const int cx = side ? x2 : x1, cy = side ? y2 : y1;
    const int ox = side ? x1 : x2, oy = side ? y1 : y2;  // the other swap position (skipped)
    const int nx = cx + (dir == 3) - (dir == 2), ny = cy + (dir == 1) - (dir == 0);
    const bool inside = nx >= 0 && nx < gw && ny >= 0 && ny < gh && !(nx == ox && ny == oy);
    const int nb = mem.get(inside ? ny * gw + nx : i1);  // (a valid address either way; ignored when !inside)
    // optimizer.rs:66-68 (same position: no temperature decay) and 71-73 (an empty cell)
    const bool skip = (x1 == x2 && y1 == y2) || p1 < 0 || p2 < 0;
    const int q1 = skip ? 0 : p1, q2 = skip ? 0 : p2;  // (a skipped iteration evaluates nothing; keep the loads in range)
    const bool present = inside && nb >= 0;
    float l1, a1, b1, l2, a2, b2, ln, an, bn;
    mem.lab(q1, l1, a1, b1);
    mem.lab(q2, l2, a2, b2);
    mem.lab(present ? nb : q1, ln, an, bn);
    const float ml = side ? l2 : l1, ma = side ? a2 : a1, mb = side ? b2 : b1;  // material now at this position
    const float ol = side ? l1 : l2, oa = side ? a1 : a2, ob = side ? b1 : b2;  // material after the swap
    bool oor = false, oor_m = false;
    float t_old = adjacency_term_fast(ml, ma, mb, ln, an, bn, oor);
    float t_new = adjacency_term_fast(ol, oa, ob, ln, an, bn, oor);
    float mutual = adjacency_term_fast(l1, a1, b1, l2, a2, b2, oor_m);  // slot-uniform
    const bool adjacent12 = (x1 == x2 && (y1 - y2 == 1 || y2 - y1 == 1)) || (y1 == y2 && (x1 - x2 == 1 || x2 - x1 == 1));
    if (__builtin_expect(__any_sync(0xffffffffu, !skip && ((oor && present) || (oor_m && adjacent12))), 0)) {
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
      old_cost = __fadd_rn(old_cost, __shfl_sync(0xffffffffu, t_old, slot_base + l));
      new_cost = __fadd_rn(new_cost, __shfl_sync(0xffffffffu, t_new, slot_base + l));
    }
    mutual = adjacent12 ? mutual : 0.0f;       // adjacency.rs:203-211 (x + 0.0 == x: the sums are >= +0.0)
    old_cost = __fadd_rn(old_cost, mutual);
    new_cost = __fadd_rn(new_cost, mutual);
    const float my_delta = p1 != p2 ? __fsub_rn(new_cost, old_cost) : 0.0f;  // adjacency.rs:156-158

    // ---- retire iteration `it` (slot 0) ----
    const int ax1 = __shfl_sync(0xffffffffu, x1, 0), ay1 = __shfl_sync(0xffffffffu, y1, 0);
    const int ax2 = __shfl_sync(0xffffffffu, x2, 0), ay2 = __shfl_sync(0xffffffffu, y2, 0);
    const int ap1 = __shfl_sync(0xffffffffu, p1, 0), ap2 = __shfl_sync(0xffffffffu, p2, 0);
    const float delta0 = __shfl_sync(0xffffffffu, my_delta, 0);
    const bool skip0 = __shfl_sync(0xffffffffu, skip, 0);
    const float u0 = __fmul_rn((float)(ring[(basepos + adv0) & (RNG_RING_WORDS - 1)] >> 8), 1.0f / 16777216.0f);
    bool und0, und1;
    bool accept0 = accept_bounds(delta0, temperature, u0, und0) && !skip0;
    und0 = und0 && !skip0 && !greedy;
    accept0 = greedy ? (delta0 < 0.0f && !skip0) : accept0;   // optimizer.rs:165 (delta < 0.0 only)
    const bool nodraw0 = greedy || skip0 || delta0 < 0.0f;  // u drawn only when delta >= 0 (optimizer.rs:80-84); never by optimize_greedy
    const float temperature1 = skip0 ? temperature : __fmul_rn(temperature, decay);  // optimizer.rs:102 (not on skips)
    const unsigned cur1 = basepos + adv0 + (nodraw0 ? 0u : 1u);
    // ---- iteration `it + 1`: the slot that sits at cur1 ----
    const int src1 = nodraw0 ? 8 : 16;
    const int bx1 = __shfl_sync(0xffffffffu, x1, src1), by1 = __shfl_sync(0xffffffffu, y1, src1);
    const int bx2 = __shfl_sync(0xffffffffu, x2, src1), by2 = __shfl_sync(0xffffffffu, y2, src1);
    const int bp1 = __shfl_sync(0xffffffffu, p1, src1), bp2 = __shfl_sync(0xffffffffu, p2, src1);
    const float delta1 = __shfl_sync(0xffffffffu, my_delta, src1);
    const bool skip1 = __shfl_sync(0xffffffffu, skip, src1);
    const unsigned adv1 = __shfl_sync(0xffffffffu, my_adv, src1);
    const bool usable1 = !__shfl_sync(0xffffffffu, inval, src1) && it + 1ull < iterations;
    // stale if the first swap changed a cell the second one read (its two cells or one of their neighbours)
    const bool touch01 = near(bx1, by1, ax1, ay1) || near(bx1, by1, ax2, ay2) || near(bx2, by2, ax1, ay1) || near(bx2, by2, ax2, ay2);
    const float u1 = __fmul_rn((float)(ring[(cur1 + adv1) & (RNG_RING_WORDS - 1)] >> 8), 1.0f / 16777216.0f);
    bool accept1 = accept_bounds(delta1, temperature1, u1, und1) && !skip1;
    und1 = und1 && !skip1 && !greedy;
    accept1 = greedy ? (delta1 < 0.0f && !skip1) : accept1;
    if (__builtin_expect(und0 || (usable1 && und1), 0)) {  // the exact exponential, in order
      if (und0) accept0 = accept_exact(delta0, temperature, u0);
      if (usable1 && !(accept0 && touch01) && und1) accept1 = accept_exact(delta1, temperature1, u1);
    }
    const bool go1 = usable1 && !(accept0 && touch01);
    accept1 = accept1 && go1;
    // apply (every lane stores the same values; a rejected or skipped first iteration stores the cells' own values back)
    const int ai1 = ay1 * gw + ax1, ai2 = ay2 * gw + ax2, bi1 = by1 * gw + bx1, bi2 = by2 * gw + bx2;
    mem.set(ai1, accept0 ? ap2 : ap1);
    mem.set(ai2, accept0 ? ap1 : ap2);
    mem.set_if(accept1, bi1, bp2);
    mem.set_if(accept1, bi2, bp1);
    __syncwarp();
    if (deferred) {
      st_global_if(accept0, reinterpret_cast<unsigned*>(deltas + accepted), __float_as_uint(delta0));
      st_global_if(accept1, reinterpret_cast<unsigned*>(deltas + accepted + (accept0 ? 1ull : 0ull)), __float_as_uint(delta1));
    } else {
      current = accept0 ? __fadd_rn(current, delta0) : current;
      const bool better0 = accept0 && current < best;
      best = better0 ? current : best;
      improved += better0 ? 1ull : 0ull;
      current = accept1 ? __fadd_rn(current, delta1) : current;
      const bool better1 = accept1 && current < best;
      best = better1 ? current : best;
      improved += better1 ? 1ull : 0ull;
    }
    accepted += (accept0 ? 1ull : 0ull) + (accept1 ? 1ull : 0ull);
    const bool decay1 = go1 && !skip1;
    temperature = decay1 ? __fmul_rn(temperature1, decay) : temperature1;
    const bool nodraw1 = greedy || skip1 || delta1 < 0.0f;
    rng.pos = go1 ? cur1 + adv1 + (nodraw1 ? 0u : 1u) : cur1;
    it += go1 ? 2ull : 1ull;
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
  if (greedy) {  // optimizer.rs:190-197: best_cost = the final cost, both counters = the swaps performed
    best = current;
    improved = accepted;
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
  anneal_kernel<GridT, GS, LS><<<1, 256, smem, ctx->stream>>>(ctx->sim_lab4(), ctx->n, ctx->d_placed, ctx->d_terms,
                                                              ctx->gw, ctx->gh, iterations, t0, decay, seed,
                                                              (mosaic_opt_result*)ctx->d_opt_res, deltas, ctx->anneal_greedy);
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
    cost_terms_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->sim_lab4(), ctx->d_placed, ctx->gw, ctx->gh, ctx->d_terms);
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
