/*
 * mosaic_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see mosaic_oracle.h).
 *
 * Scalar restatement of the reference hot path.  Compile with strict f32 semantics:
 *   g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared
 * so that no multiply-add is fused and no reduction is re-associated: every function
 * below evaluates its floating-point expressions in the order of the cited Rust code.
 *
 * Citations are file:line under /root/reference/src/.  [3P] marks arithmetic that lives
 * in a third-party crate whose source is not in /root/reference (pinned in Cargo.lock):
 * the algorithm is restated from the published crate and is "parity unpinned" beyond the
 * reference's own test vectors.  What this image allows beyond those vectors is in
 * tests/test_oracle_golden.py: the ChaCha block function against OpenSSL's ChaCha20 for arbitrary
 * keys and counters, HSV against Python's colorsys, sRGB->Lab against the formula in float64 and
 * OpenCV, the resize against Pillow's LANCZOS, top-K against scipy's k-d tree — and
 * tests/golden/pipeline_small.npz pins every stage's output bits against drift.
 */
#include "mosaic_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace {

/* ------------------------------------------------------------------ helpers */
template <class F>
void parallel_for(int64_t n, int threads, F f) {
  if (threads <= 1 || n <= 1) {
    f(0, n);
    return;
  }
  if ((int64_t)threads > n) threads = (int)n;
  std::vector<std::thread> pool;
  int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back([=] { f(b, e); });
  }
  for (auto& th : pool) th.join();
}

/* ------------------------------------------------------------------ sRGB -> Lab
 * [3P] palette 0.7.6.  lib.rs:102-107 builds Srgb<f32>(u8/255.0) and calls
 * Lab::from_color: Srgb -> linear (transfer fn) -> XYZ (D65 matrix) -> / white -> f() -> Lab. */
struct LinLut {
  float v[256];
  LinLut() {
    for (int i = 0; i < 256; ++i) {
      float x = (float)i / 255.0f; /* lib.rs:103 `pixel[0] as f32 / 255.0` */
      /* palette Srgb::into_linear: x<=0.04045 ? x/12.92 : ((x+0.055)/1.055)^2.4, written as
         1/12.92*x and x.mul_add(1/1.055, 0.055/1.055).powf(2.4) */
      if (x <= 0.04045f)
        v[i] = (float)(1.0 / 12.92) * x;
      else
        v[i] = powf(fmaf(x, (float)(1.0 / 1.055), (float)(0.055 / 1.055)), 2.4f);
    }
  }
};
const LinLut kLin;

inline float lab_f(float c) {
  /* palette lab.rs: epsilon = (6/29)^3, kappa = 841/108, delta = 4/29 */
  const float e0 = (float)(6.0 / 29.0);
  const float epsilon = e0 * e0 * e0;
  const float kappa = (float)(841.0 / 108.0);
  const float delta = (float)(4.0 / 29.0);
  return c > epsilon ? cbrtf(c) : (kappa * c) + delta;
}

inline void rgb8_to_lab(uint8_t r8, uint8_t g8, uint8_t b8, float& L, float& A, float& B) {
  float r = kLin.v[r8], g = kLin.v[g8], b = kLin.v[b8];
  /* palette multiply_rgb_to_xyz, sRGB/D65 matrix (Lindbloom), each row (m0*r + m1*g) + m2*b */
  float x = (0.4124564f * r + 0.3575761f * g) + 0.1804375f * b;
  float y = (0.2126729f * r + 0.7151522f * g) + 0.0721750f * b;
  float z = (0.0193339f * r + 0.1191920f * g) + 0.9503041f * b;
  /* Xyz / D65 white point (0.95047, 1.0, 1.08883) */
  x = x / 0.95047f;
  y = y / 1.0f;
  z = z / 1.08883f;
  float fx = lab_f(x), fy = lab_f(y), fz = lab_f(z);
  L = (fy * 116.0f) - 16.0f;
  A = (fx - fy) * 500.0f;
  B = (fy - fz) * 200.0f;
}

/* lib.rs:94-119: sequential f32 fold over pixels in row-major order, then / (w*h) as f32 */
void average_lab_region(const uint8_t* rgb, int img_w, int x0, int y0, int rw, int rh, float out[3]) {
  float sl = 0.0f, sa = 0.0f, sb = 0.0f;
  for (int y = 0; y < rh; ++y) {
    const uint8_t* row = rgb + ((size_t)(y0 + y) * img_w + x0) * 3;
    for (int x = 0; x < rw; ++x) {
      float L, A, B;
      rgb8_to_lab(row[3 * x], row[3 * x + 1], row[3 * x + 2], L, A, B);
      sl = sl + L;
      sa = sa + A;
      sb = sb + B;
    }
  }
  float total = (float)((uint32_t)rw * (uint32_t)rh); /* lib.rs:97 `(width * height) as f32` */
  out[0] = sl / total;
  out[1] = sa / total;
  out[2] = sb / total;
}

/* ------------------------------------------------------------------ distances */
/* kiddo 5.2.2 SquaredEuclidean::dist [3P]: fold over dims of (a-b)*(a-b) from 0.0, unfused */
inline float sq_dist(const float* a, const float* b) {
  float dl = a[0] - b[0], da = a[1] - b[1], db = a[2] - b[2];
  return ((0.0f + dl * dl) + da * da) + db * db;
}
/* similarity.rs:141-146 */
inline float lab_distance(const float* a, const float* b) {
  float dl = a[0] - b[0], da = a[1] - b[1], db = a[2] - b[2];
  return sqrtf(dl * dl + da * da + db * db);
}
/* similarity.rs:85-103 get_similarity by index: equal -> 0.0, else packed entry of (min,max) */
inline float sim_idx(const float* mat_lab, int32_t i, int32_t j) {
  if (i == j) return 0.0f;
  int32_t lo = std::min(i, j), hi = std::max(i, j);
  return lab_distance(mat_lab + 3 * (size_t)lo, mat_lab + 3 * (size_t)hi); /* similarity.rs:75-78 */
}

/* adjacency.rs:17-45: neighbours in order Up, Down, Left, Right */
inline int adjacent(int x, int y, int gw, int gh, int nx[4], int ny[4]) {
  int c = 0;
  if (y > 0) { nx[c] = x; ny[c] = y - 1; ++c; }
  if (y < gh - 1) { nx[c] = x; ny[c] = y + 1; ++c; }
  if (x > 0) { nx[c] = x - 1; ny[c] = y; ++c; }
  if (x < gw - 1) { nx[c] = x + 1; ny[c] = y; ++c; }
  return c;
}

/* adjacency.rs:63-88 */
float penalty(const float* mat_lab, const int32_t* grid, int gw, int gh, int x, int y,
              int32_t cand, float weight) {
  int nx[4], ny[4];
  int c = adjacent(x, y, gw, gh, nx, ny);
  float p = 0.0f;
  for (int i = 0; i < c; ++i) {
    int32_t nb = grid[(size_t)ny[i] * gw + nx[i]];
    if (nb >= 0) p = p + 1.0f / (sim_idx(mat_lab, cand, nb) + 1.0f);
  }
  return p * weight;
}

/* main.rs:309-333: reject a candidate equal to the material at Left/Right/Up/Down */
inline bool can_place(const int32_t* grid, int gw, int gh, int x, int y, int32_t cand) {
  if (x > 0 && grid[(size_t)y * gw + x - 1] == cand) return false;
  if (x + 1 < gw && grid[(size_t)y * gw + x + 1] == cand) return false;
  if (y > 0 && grid[(size_t)(y - 1) * gw + x] == cand) return false;
  if (y + 1 < gh && grid[(size_t)(y + 1) * gw + x] == cand) return false;
  return true;
}

/* adjacency.rs:91-130 (penalty_weight is NOT applied here) */
float total_cost(const float* mat_lab, const int32_t* grid, int gw, int gh) {
  if (gh == 0) return 0.0f;
  float total = 0.0f;
  for (int y = 0; y < gh; ++y)
    for (int x = 0; x < gw; ++x) {
      int32_t cur = grid[(size_t)y * gw + x];
      if (cur < 0) continue;
      if (x < gw - 1) {
        int32_t r = grid[(size_t)y * gw + x + 1];
        if (r >= 0) total = total + 1.0f / (sim_idx(mat_lab, cur, r) + 1.0f);
      }
      if (y < gh - 1) {
        int32_t d = grid[(size_t)(y + 1) * gw + x];
        if (d >= 0) total = total + 1.0f / (sim_idx(mat_lab, cur, d) + 1.0f);
      }
    }
  return total;
}

/* adjacency.rs:133-214 */
float swap_delta(const float* mat_lab, const int32_t* grid, int gw, int gh, int x1, int y1, int x2, int y2) {
  if (gh == 0) return 0.0f;
  int32_t p1 = grid[(size_t)y1 * gw + x1];
  int32_t p2 = grid[(size_t)y2 * gw + x2];
  if (p1 < 0 || p2 < 0) return 0.0f;
  if (p1 == p2) return 0.0f;
  float old_cost = 0.0f, new_cost = 0.0f;
  int ax[4], ay[4];
  int c1 = adjacent(x1, y1, gw, gh, ax, ay);
  bool adjacent12 = false;
  for (int i = 0; i < c1; ++i) {
    if (ax[i] == x2 && ay[i] == y2) { adjacent12 = true; continue; }
    int32_t nb = grid[(size_t)ay[i] * gw + ax[i]];
    if (nb >= 0) {
      old_cost = old_cost + 1.0f / (sim_idx(mat_lab, p1, nb) + 1.0f);
      new_cost = new_cost + 1.0f / (sim_idx(mat_lab, p2, nb) + 1.0f);
    }
  }
  int bx[4], by[4];
  int c2 = adjacent(x2, y2, gw, gh, bx, by);
  for (int i = 0; i < c2; ++i) {
    if (bx[i] == x1 && by[i] == y1) continue;
    int32_t nb = grid[(size_t)by[i] * gw + bx[i]];
    if (nb >= 0) {
      old_cost = old_cost + 1.0f / (sim_idx(mat_lab, p2, nb) + 1.0f);
      new_cost = new_cost + 1.0f / (sim_idx(mat_lab, p1, nb) + 1.0f);
    }
  }
  if (adjacent12) {
    float cost = 1.0f / (sim_idx(mat_lab, p1, p2) + 1.0f);
    old_cost = old_cost + cost;
    new_cost = new_cost + cost;
  }
  return new_cost - old_cost;
}

/* ------------------------------------------------------------------ ChaCha12 / StdRng [3P] */
inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
#define ORC_QR(a, b, c, d)            \
  a += b; d ^= a; d = rotl32(d, 16);  \
  c += d; b ^= c; b = rotl32(b, 12);  \
  a += b; d ^= a; d = rotl32(d, 8);   \
  c += d; b ^= c; b = rotl32(b, 7);

void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                    key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u}; /* 64-bit counter, stream 0 */
  uint32_t x[16];
  std::memcpy(x, s, sizeof x);
  for (int r = 0; r < rounds; r += 2) {
    ORC_QR(x[0], x[4], x[8], x[12]) ORC_QR(x[1], x[5], x[9], x[13])
    ORC_QR(x[2], x[6], x[10], x[14]) ORC_QR(x[3], x[7], x[11], x[15])
    ORC_QR(x[0], x[5], x[10], x[15]) ORC_QR(x[1], x[6], x[11], x[12])
    ORC_QR(x[2], x[7], x[8], x[13]) ORC_QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

}  // namespace

struct orc_rng {
  uint32_t key[8];
  uint64_t counter;   /* next block counter */
  uint32_t buf[64];   /* rand_chacha: 4 blocks per refill */
  int index;          /* rand_core BlockRng index; 64 = empty */
  void refill() {
    for (int b = 0; b < 4; ++b) chacha_block(key, counter + b, 12, buf + 16 * b);
    counter += 4;
  }
  void generate_and_set(int idx) { refill(); index = idx; }
  uint32_t next_u32() {
    if (index >= 64) generate_and_set(0);
    return buf[index++];
  }
  uint64_t next_u64() { /* rand_core 0.6.4 BlockRng::next_u64 */
    if (index < 63) {
      uint64_t v = ((uint64_t)buf[index + 1] << 32) | buf[index];
      index += 2;
      return v;
    } else if (index >= 64) {
      generate_and_set(2);
      return ((uint64_t)buf[1] << 32) | buf[0];
    } else {
      uint64_t x = buf[63];
      generate_and_set(1);
      uint64_t y = buf[0];
      return (y << 32) | x;
    }
  }
  uint64_t gen_range(uint64_t n) { /* rand 0.8.5 UniformInt<usize>::sample_single, 64-bit */
    if (n == 0) return next_u64();
    int lz = __builtin_clzll(n);
    uint64_t zone = (n << lz) - 1;
    for (;;) {
      uint64_t v = next_u64();
      unsigned __int128 m = (unsigned __int128)v * n;
      uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
      if (lo <= zone) return hi;
    }
  }
  float gen_f32() { /* rand 0.8.5 Standard for f32: 24 high bits * 2^-24 */
    uint32_t v = next_u32();
    return (float)(v >> 8) * (1.0f / 16777216.0f);
  }
};

extern "C" {

void orc_chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
  chacha_block(key, counter, rounds, out);
}

orc_rng* orc_rng_new(uint64_t seed) {
  orc_rng* r = new orc_rng();
  /* rand_core 0.6.4 SeedableRng::seed_from_u64: PCG32 expansion into the 32-byte key */
  uint64_t state = seed;
  for (int i = 0; i < 8; ++i) {
    state = state * 6364136223846793005ULL + 11634580027462260723ULL;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    r->key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
  r->counter = 0;
  r->index = 64;
  return r;
}
void orc_rng_free(orc_rng* r) { delete r; }
uint32_t orc_rng_next_u32(orc_rng* r) { return r->next_u32(); }
uint64_t orc_rng_next_u64(orc_rng* r) { return r->next_u64(); }
uint64_t orc_rng_gen_range(orc_rng* r, uint64_t n) { return r->gen_range(n); }
float orc_rng_gen_f32(orc_rng* r) { return r->gen_f32(); }

/* f32 exp for optimizer.rs:82 `(-delta / temperature).exp()`.  Rust's f32::exp is the
 * platform libm expf; to make CPU and GPU bit-identical BY CONSTRUCTION the oracle defines
 * exp as: argument reduction + degree-13 Taylor/Horner in IEEE double (unfused, this exact
 * operation order), one final rounding to f32.  The double result is within ~2^-52 relative of
 * exp(x), so the f32 result is the correctly rounded one except when exp(x) lies within
 * ~2^-28 ulp of a rounding boundary; tests/ check it against glibc expf. */
float orc_expf(float xf) {
  if (xf != xf) return xf;
  if (xf > 88.8f) return INFINITY;
  if (xf < -104.0f) return 0.0f;
  double x = (double)xf;
  double kd = floor(x * 1.4426950408889634 + 0.5);
  double r = x - kd * 0.693147180369123816490e+00; /* ln2 hi */
  r = r - kd * 1.90821492927058770002e-10;         /* ln2 lo */
  static const double c[14] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
                               1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
                               1.0 / 479001600, 1.0 / 6227020800.0};
  double p = c[13];
  for (int i = 12; i >= 0; --i) p = p * r + c[i];
  int k = (int)kd;
  uint64_t bits = (uint64_t)(k + 1023) << 52; /* 2^k, k in [-151, 129] -> normal double */
  double scale;
  std::memcpy(&scale, &bits, 8);
  return (float)(p * scale);
}

void orc_rgb8_to_lab(uint8_t r, uint8_t g, uint8_t b, float out[3]) { rgb8_to_lab(r, g, b, out[0], out[1], out[2]); }

void orc_average_lab_region(const uint8_t* rgb, int img_w, int x0, int y0, int rw, int rh, float out[3]) {
  average_lab_region(rgb, img_w, x0, y0, rw, rh, out);
}

void orc_material_lab(const uint8_t* rgb, int n, int w, int h, float* out, int threads) {
  parallel_for(n, threads, [=](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
      average_lab_region(rgb + (size_t)i * w * h * 3, w, 0, 0, w, h, out + 3 * i);
  });
}

void orc_cell_lab(const uint8_t* rgb, int img_w, int img_h, int gw, int gh, float* out, int threads) {
  int tw = img_w / gw, th = img_h / gh; /* main.rs:495-496 u32 floor */
  parallel_for((int64_t)gw * gh, threads, [=](int64_t b, int64_t e) {
    for (int64_t c = b; c < e; ++c) {
      int x = (int)(c % gw), y = (int)(c / gw);
      average_lab_region(rgb, img_w, x * tw, y * th, tw, th, out + 3 * c); /* main.rs:536-541 */
    }
  });
}

void orc_average_rgb_region(const uint8_t* rgb, int img_w, int x0, int y0, int rw, int rh, uint8_t out[3]) {
  uint64_t s[3] = {0, 0, 0};
  for (int y = 0; y < rh; ++y) {
    const uint8_t* row = rgb + ((size_t)(y0 + y) * img_w + x0) * 3;
    for (int x = 0; x < rw; ++x) {
      s[0] += row[3 * x];
      s[1] += row[3 * x + 1];
      s[2] += row[3 * x + 2];
    }
  }
  uint64_t total = (uint64_t)rw * rh;
  for (int c = 0; c < 3; ++c) out[c] = (uint8_t)(s[c] / total);
}

void orc_topk(const float* mat_lab, int n, const float* cell_lab, int c, int k,
              uint32_t* out_idx, float* out_d2, int threads) {
  if (k > n) k = n;
  parallel_for(c, threads, [=](int64_t b, int64_t e) {
    std::vector<uint64_t> keys(n);
    for (int64_t q = b; q < e; ++q) {
      const float* ql = cell_lab + 3 * q;
      for (int i = 0; i < n; ++i) {
        float d2 = sq_dist(ql, mat_lab + 3 * (size_t)i); /* query - point order as kiddo: (a-b) with a = query */
        uint32_t bits;
        std::memcpy(&bits, &d2, 4);
        keys[i] = ((uint64_t)bits << 32) | (uint32_t)i; /* d2 >= 0 -> bit pattern is order-preserving */
      }
      std::partial_sort(keys.begin(), keys.begin() + k, keys.end());
      for (int j = 0; j < k; ++j) {
        uint32_t bits = (uint32_t)(keys[j] >> 32);
        out_idx[(size_t)q * k + j] = (uint32_t)keys[j];
        std::memcpy(&out_d2[(size_t)q * k + j], &bits, 4);
      }
    }
  });
}

// ---- the same k nearest by a bucketed k-d tree (what the reference's CPU path costs: kiddo, main.rs:90, 145-151, 348-352) --
// Same results as orc_topk — the k smallest (d2, index) pairs, d2 by the same unfused arithmetic — found in O(log n + k)
// per query instead of O(n): leaves of <= 32 points, median split on the axis of largest spread, near child first, far
// child only if the split plane is not farther than the current k-th distance.  The plane test is an exact bound:
// rounding is monotone, so fl(fl(|q - split|)^2) <= fl(dl^2) <= d2 for every point behind the plane; on equality the
// child is visited (it may hold an equal distance with a smaller index).
namespace {
struct KdNode {
  int axis;     // -1: leaf
  float split;
  int left, right;
  int begin, end;  // leaf: range in `order`
};
struct KdTree {
  std::vector<KdNode> nodes;
  std::vector<int> order;
  const float* pts;
  int build(int b, int e) {
    KdNode nd{-1, 0.f, -1, -1, b, e};
    const int id = (int)nodes.size();
    nodes.push_back(nd);
    if (e - b <= 32) return id;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = b; i < e; ++i)
      for (int a = 0; a < 3; ++a) {
        const float v = pts[3 * (size_t)order[i] + a];
        lo[a] = std::min(lo[a], v);
        hi[a] = std::max(hi[a], v);
      }
    int axis = 0;
    for (int a = 1; a < 3; ++a)
      if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
    if (!(hi[axis] > lo[axis])) return id;  // all points equal: one (large) leaf
    const int mid = (b + e) / 2;
    std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e,
                     [&](int x, int y) { return pts[3 * (size_t)x + axis] < pts[3 * (size_t)y + axis]; });
    const float split = pts[3 * (size_t)order[mid] + axis];  // left: values <= split, right: values >= split
    const int l = build(b, mid), r = build(mid, e);
    nodes[id].axis = axis;
    nodes[id].split = split;
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }
};
}  // namespace

void orc_topk_kdtree(const float* mat_lab, int n, const float* cell_lab, int c, int k, uint32_t* out_idx, float* out_d2,
                     int threads) {
  if (k > n) k = n;
  if (k <= 0) return;
  KdTree tree;
  tree.pts = mat_lab;
  tree.order.resize(n);
  for (int i = 0; i < n; ++i) tree.order[i] = i;
  tree.nodes.reserve((size_t)n / 8 + 16);
  tree.build(0, n);
  const KdTree* T = &tree;
  parallel_for(c, threads, [=](int64_t b, int64_t e) {
    std::vector<uint64_t> heap;  // max-heap of (d2 bits << 32 | index): the top is the current k-th
    heap.reserve(k + 1);
    struct Entry {
      int node;
      uint32_t bound;  // bit pattern of a lower bound of every d2 in the subtree (0: none known)
    };
    std::vector<Entry> stack;
    for (int64_t q = b; q < e; ++q) {
      const float* ql = cell_lab + 3 * q;
      heap.clear();
      stack.clear();
      stack.push_back(Entry{0, 0u});
      while (!stack.empty()) {
        const Entry en = stack.back();
        stack.pop_back();
        // the heap may have tightened since the entry was pushed; equality still visits (equal d2, smaller index)
        if ((int)heap.size() == k && en.bound > (uint32_t)(heap.front() >> 32)) continue;
        const KdNode& nd = T->nodes[en.node];
        if (nd.axis < 0) {
          for (int i = nd.begin; i < nd.end; ++i) {
            const int id = T->order[i];
            const float d2 = sq_dist(ql, mat_lab + 3 * (size_t)id);
            uint32_t bits;
            std::memcpy(&bits, &d2, 4);
            const uint64_t key = ((uint64_t)bits << 32) | (uint32_t)id;
            if ((int)heap.size() < k) {
              heap.push_back(key);
              std::push_heap(heap.begin(), heap.end());
            } else if (key < heap.front()) {
              std::pop_heap(heap.begin(), heap.end());
              heap.back() = key;
              std::push_heap(heap.begin(), heap.end());
            }
          }
          continue;
        }
        const float diff = ql[nd.axis] - nd.split;
        const float plane = diff * diff;  // every point p on the far side has |q - p| >= |q - split| on this axis
        uint32_t pbits;
        std::memcpy(&pbits, &plane, 4);
        const bool left_near = diff <= 0.f;
        stack.push_back(Entry{left_near ? nd.right : nd.left, std::max(pbits, en.bound)});  // far: popped after ...
        stack.push_back(Entry{left_near ? nd.left : nd.right, en.bound});                   // ... the near child
      }
      std::sort_heap(heap.begin(), heap.end());
      for (int j = 0; j < k; ++j) {
        const uint32_t bits = (uint32_t)(heap[j] >> 32);
        out_idx[(size_t)q * k + j] = (uint32_t)heap[j];
        std::memcpy(&out_d2[(size_t)q * k + j], &bits, 4);
      }
    }
  });
}

float orc_lab_distance(const float a[3], const float b[3]) { return lab_distance(a, b); }

uint64_t orc_tri_index(uint64_t i, uint64_t j, uint64_t n) { return i * n - i * (i + 1) / 2 + j - i - 1; }

void orc_similarity_tri(const float* lab, int n, float* out, int threads) {
  parallel_for(n, threads, [=](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      uint64_t base = orc_tri_index(i, i + 1, n);
      for (int64_t j = i + 1; j < n; ++j) out[base + (j - i - 1)] = lab_distance(lab + 3 * i, lab + 3 * j);
    }
  });
}

float orc_penalty(const float* mat_lab, const int32_t* grid, int gw, int gh, int x, int y,
                  int32_t candidate, float weight) {
  return penalty(mat_lab, grid, gw, gh, x, y, candidate, weight);
}
float orc_total_cost(const float* mat_lab, const int32_t* grid, int gw, int gh) {
  return total_cost(mat_lab, grid, gw, gh);
}
float orc_swap_delta(const float* mat_lab, const int32_t* grid, int gw, int gh, int x1, int y1, int x2, int y2) {
  return swap_delta(mat_lab, grid, gw, gh, x1, y1, x2, y2);
}

uint32_t orc_place(const uint32_t* cand_idx, const float* cand_d2, int k,
                   const float* mat_lab, int n, int gw, int gh, int max_usage, float weight,
                   int32_t* placed, uint32_t* usage_out, uint32_t* fallback_cells, uint32_t fallback_cap) {
  std::vector<uint32_t> usage(n, 0u);
  size_t cells = (size_t)gw * gh;
  for (size_t i = 0; i < cells; ++i) placed[i] = -1; /* main.rs:303-307 */
  uint32_t n_fallbacks = 0;
  for (int y = 0; y < gh; ++y) {
    for (int x = 0; x < gw; ++x) { /* main.rs:529-530 raster order */
      size_t cell = (size_t)y * gw + x;
      const uint32_t* ci = cand_idx + cell * k;
      const float* cd = cand_d2 + cell * k;
      /* main.rs:335-418 */
      bool have = false;
      float best_score = 0.0f;
      int32_t best = -1;
      for (int j = 0; j < k; ++j) {
        int32_t t = (int32_t)ci[j];
        if (!(usage[t] < (uint32_t)max_usage)) continue;        /* lib.rs:71-74, main.rs:375 */
        if (!can_place(placed, gw, gh, x, y, t)) continue;      /* main.rs:380 */
        float pen = 0.0f;
        if (weight > 0.0f) pen = penalty(mat_lab, placed, gw, gh, x, y, t, weight); /* main.rs:355-362, 386-396 */
        float total = cd[j] + pen;                              /* main.rs:398 */
        if (!have || total < best_score) {                      /* main.rs:401-407 strict < */
          have = true;
          best_score = total;
          best = t;
        }
      }
      if (have) {
        usage[best] += 1; /* main.rs:410-414 */
        placed[cell] = best;
        continue;
      }
      /* main.rs:420-475 fallback_tile_selection */
      if (fallback_cells && n_fallbacks < fallback_cap) fallback_cells[n_fallbacks] = (uint32_t)cell;
      ++n_fallbacks;
      std::fill(usage.begin(), usage.end(), 0u); /* lib.rs:86-88 reset clears ALL counts */
      int32_t chosen = -1;
      for (int j = 0; j < k; ++j) {
        int32_t t = (int32_t)ci[j];
        if (can_place(placed, gw, gh, x, y, t)) { chosen = t; break; } /* main.rs:441-453 */
      }
      if (chosen < 0) chosen = (int32_t)ci[0]; /* main.rs:456-474 nearest_one */
      usage[chosen] += 1;
      placed[cell] = chosen;
    }
  }
  if (usage_out) std::memcpy(usage_out, usage.data(), sizeof(uint32_t) * n);
  return n_fallbacks;
}

void orc_anneal(const float* mat_lab, int32_t* grid, int gw, int gh, uint64_t iterations,
                float t0, float decay, uint64_t seed, orc_opt_result* res) {
  orc_opt_result out;
  std::memset(&out, 0, sizeof out);
  if (gh == 0 || gw == 0) { /* optimizer.rs:42-45 */
    if (res) *res = out;
    return;
  }
  orc_rng* rng = orc_rng_new(seed); /* optimizer.rs:48 thread_rng() -> defined as StdRng::seed_from_u64 */
  float current = total_cost(mat_lab, grid, gw, gh);
  float initial = current, best = current;
  uint64_t improved = 0, accepted = 0;
  float temperature = t0;
  for (uint64_t it = 0; it < iterations; ++it) {
    int x1 = (int)rng->gen_range((uint64_t)gw); /* optimizer.rs:60-63: x before y */
    int y1 = (int)rng->gen_range((uint64_t)gh);
    int x2 = (int)rng->gen_range((uint64_t)gw);
    int y2 = (int)rng->gen_range((uint64_t)gh);
    if (x1 == x2 && y1 == y2) continue;                              /* optimizer.rs:66-68 */
    size_t i1 = (size_t)y1 * gw + x1, i2 = (size_t)y2 * gw + x2;
    if (grid[i1] < 0 || grid[i2] < 0) continue;                      /* optimizer.rs:71-73 */
    float delta = swap_delta(mat_lab, grid, gw, gh, x1, y1, x2, y2); /* optimizer.rs:76 */
    bool accept;
    if (delta < 0.0f) {
      accept = true;
    } else {
      float probability = orc_expf(-delta / temperature);            /* optimizer.rs:82 */
      accept = rng->gen_f32() < probability;                         /* optimizer.rs:83 */
    }
    if (accept) {
      std::swap(grid[i1], grid[i2]);
      current = current + delta;
      accepted += 1;
      if (current < best) {
        best = current;
        improved += 1;
      }
    }
    temperature = temperature * decay; /* optimizer.rs:102 */
  }
  orc_rng_free(rng);
  out.initial_cost = initial;
  out.final_cost = current;
  out.best_cost = best;
  out.improved_count = improved;
  out.accepted_count = accepted;
  out.iterations = iterations; /* optimizer.rs:127 echoes max_iterations */
  if (res) *res = out;
}

/* MosaicOptimizer::optimize_greedy (optimizer.rs:132-198): only improvements are accepted; no acceptance sample */
void orc_anneal_greedy(const float* mat_lab, int32_t* grid, int gw, int gh, uint64_t iterations, uint64_t seed,
                       orc_opt_result* res) {
  orc_opt_result out;
  std::memset(&out, 0, sizeof out);
  if (gh == 0 || gw == 0) { /* optimizer.rs:138-141 */
    if (res) *res = out;
    return;
  }
  orc_rng* rng = orc_rng_new(seed); /* optimizer.rs:144 thread_rng() -> defined as StdRng::seed_from_u64 */
  float current = total_cost(mat_lab, grid, gw, gh); /* optimizer.rs:145 */
  float initial = current;
  uint64_t improved = 0;
  for (uint64_t it = 0; it < iterations; ++it) {
    int x1 = (int)rng->gen_range((uint64_t)gw); /* optimizer.rs:152-155: x before y */
    int y1 = (int)rng->gen_range((uint64_t)gh);
    int x2 = (int)rng->gen_range((uint64_t)gw);
    int y2 = (int)rng->gen_range((uint64_t)gh);
    if (x1 == x2 && y1 == y2) continue;                              /* optimizer.rs:157-159 */
    size_t i1 = (size_t)y1 * gw + x1, i2 = (size_t)y2 * gw + x2;
    if (grid[i1] < 0 || grid[i2] < 0) continue;                      /* optimizer.rs:161-163 */
    float delta = swap_delta(mat_lab, grid, gw, gh, x1, y1, x2, y2); /* optimizer.rs:165 */
    if (delta < 0.0f) {                                              /* optimizer.rs:167-175 */
      std::swap(grid[i1], grid[i2]);
      current = current + delta;
      improved += 1;
    }
  }
  orc_rng_free(rng);
  out.initial_cost = initial; /* optimizer.rs:190-197 */
  out.final_cost = current;
  out.best_cost = current;
  out.improved_count = improved;
  out.accepted_count = improved;
  out.iterations = iterations;
  if (res) *res = out;
}

/* ------------------------------------------------------------------ resize [3P]
 * fast_image_resize 5.1.4, ResizeOptions::new() = Convolution(Lanczos3), PixelType::U8x3:
 * horizontal pass into an out_w x in_h u8 image, then vertical pass; an axis whose size does not
 * change is skipped.  Coefficients: f64, normalised to sum 1, then i16 fixed point with adaptive
 * precision (Normalizer16); accumulate in i32 starting from 1<<(precision-1), >> precision, clamp. */
}  // extern "C"

namespace {
double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return std::sin(x) / x;
}
double lanczos3(double x) { return (x >= -3.0 && x < 3.0) ? sinc_filter(x) * sinc_filter(x / 3.0) : 0.0; }

struct AxisCoeffs {
  int window = 0, precision = 0;
  std::vector<int32_t> start, size;
  std::vector<int16_t> w; /* out_size * window */
};

int axis_window(int in_size, int out_size) {
  double scale = (double)in_size / (double)out_size;
  double filter_scale = std::max(scale, 1.0);
  double radius = 3.0 * filter_scale;
  return (int)std::ceil(radius) * 2 + 1;
}

AxisCoeffs axis_coeffs(int in_size, int out_size) {
  AxisCoeffs c;
  double scale = (double)in_size / (double)out_size;
  double filter_scale = std::max(scale, 1.0);
  double radius = 3.0 * filter_scale;
  c.window = (int)std::ceil(radius) * 2 + 1;
  double recip = 1.0 / filter_scale;
  std::vector<double> vals((size_t)out_size * c.window, 0.0);
  c.start.resize(out_size);
  c.size.resize(out_size);
  for (int o = 0; o < out_size; ++o) {
    double in_center = ((double)o + 0.5) * scale;
    double lo = std::floor(in_center - radius);
    int x_min = lo < 0.0 ? 0 : (int)lo;
    int x_max = std::min((int)std::ceil(in_center + radius), in_size);
    double center = in_center - 0.5;
    double ww = 0.0;
    double* k = &vals[(size_t)o * c.window];
    for (int x = x_min; x < x_max; ++x) {
      double wv = lanczos3(((double)x - center) * recip);
      k[x - x_min] = wv;
      ww += wv;
    }
    if (ww != 0.0)
      for (int x = 0; x < x_max - x_min; ++x) k[x] /= ww;
    c.start[o] = x_min;
    c.size[o] = x_max - x_min;
  }
  /* Normalizer16: precision from the largest coefficient; PRECISION_BITS = 22, MAX_COEFS_PRECISION = 15 */
  double max_weight = 0.0;
  for (double v : vals) max_weight = std::max(max_weight, v);
  int precision = 0;
  for (int cur = 0; cur < 22; ++cur) {
    precision = cur;
    int32_t next_value = (int32_t)std::round(max_weight * (double)(1 << (precision + 1)));
    if (next_value >= (1 << 15)) break;
  }
  c.precision = precision;
  double s = (double)(1 << precision);
  c.w.resize(vals.size());
  for (size_t i = 0; i < vals.size(); ++i) c.w[i] = (int16_t)std::round(vals[i] * s);
  return c;
}

inline uint8_t clip8(int32_t v, int precision) {
  v = v >> precision;
  return (uint8_t)std::min(255, std::max(0, v));
}

void resize_rgb8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  std::vector<uint8_t> tmp;
  const uint8_t* hsrc = src;
  if (sw != dw) { /* horizontal pass */
    AxisCoeffs c = axis_coeffs(sw, dw);
    tmp.resize((size_t)dw * sh * 3);
    for (int y = 0; y < sh; ++y)
      for (int o = 0; o < dw; ++o) {
        const int16_t* k = &c.w[(size_t)o * c.window];
        int32_t ss[3] = {1 << (c.precision - 1), 1 << (c.precision - 1), 1 << (c.precision - 1)};
        for (int t = 0; t < c.size[o]; ++t) {
          const uint8_t* p = src + ((size_t)y * sw + c.start[o] + t) * 3;
          ss[0] += (int32_t)p[0] * k[t];
          ss[1] += (int32_t)p[1] * k[t];
          ss[2] += (int32_t)p[2] * k[t];
        }
        uint8_t* q = &tmp[((size_t)y * dw + o) * 3];
        q[0] = clip8(ss[0], c.precision);
        q[1] = clip8(ss[1], c.precision);
        q[2] = clip8(ss[2], c.precision);
      }
    hsrc = tmp.data();
  }
  if (sh != dh) { /* vertical pass */
    AxisCoeffs c = axis_coeffs(sh, dh);
    for (int o = 0; o < dh; ++o) {
      const int16_t* k = &c.w[(size_t)o * c.window];
      for (int x = 0; x < dw * 3; ++x) {
        int32_t ss = 1 << (c.precision - 1);
        for (int t = 0; t < c.size[o]; ++t) ss += (int32_t)hsrc[((size_t)(c.start[o] + t) * dw) * 3 + x] * k[t];
        dst[(size_t)o * dw * 3 + x] = clip8(ss, c.precision);
      }
    }
  } else {
    std::memcpy(dst, hsrc, (size_t)dw * dh * 3);
  }
}

/* ------------------------------------------------------------------ colour adjustment */
inline float clampf(float v, float lo, float hi) { /* Rust f32::clamp */
  if (v < lo) return lo;
  if (v > hi) return hi;
  return v;
}

/* [3P] palette 0.7.6 Hsv::from_color_unclamped(Rgb): hexcone; hue in raw (unnormalised) degrees */
void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
  float max, min, sep, coeff;
  if (r > g) { max = r; min = g; sep = g - b; coeff = 0.0f; }
  else       { max = g; min = r; sep = b - r; coeff = 2.0f; }
  if (b > max) { max = b; sep = r - g; coeff = 4.0f; }
  else if (b < min) { min = b; }
  h = 0.0f;
  s = 0.0f;
  v = max;
  if (max != min) {
    float d = max - min;
    s = d / max;
    h = ((sep / d) + coeff) * 60.0f;
  }
}

/* [3P] palette 0.7.6 Rgb::from_color_unclamped(Hsv) */
void hsv_to_rgb(float hue, float s, float v, float& r, float& g, float& b) {
  float c = v * s;
  float deg = hue - (floorf(hue / 360.0f) * 360.0f); /* RgbHue::into_positive_degrees */
  float h = deg / 60.0f;
  float h_mod_two = h - floorf(h * 0.5f) * 2.0f;
  float x = c * (1.0f - fabsf(h_mod_two - 1.0f));
  float m = v - c;
  bool z0 = h >= 0.0f && h < 1.0f, z1 = h >= 1.0f && h < 2.0f, z2 = h >= 2.0f && h < 3.0f,
       z3 = h >= 3.0f && h < 4.0f, z4 = h >= 4.0f && h < 5.0f;
  float red = (z1 || z4) ? x : ((z2 || z3) ? 0.0f : c);
  float green = (z0 || z3) ? x : ((z1 || z2) ? c : 0.0f);
  float blue = (z0 || z1) ? 0.0f : ((z3 || z4) ? c : x);
  r = red + m;
  g = green + m;
  b = blue + m;
}

/* color_adjustment.rs:91-96 */
inline float brightness_contrast(float value, float brightness, float contrast) {
  float contrasted = clampf((value - 0.5f) * contrast + 0.5f, 0.0f, 1.0f);
  return clampf(contrasted + brightness, 0.0f, 1.0f);
}

inline float rem_euclid(float a, float b) { /* Rust f32::rem_euclid */
  float r = fmodf(a, b);
  return r < 0.0f ? r + fabsf(b) : r;
}

inline uint8_t to_u8(float v) { /* `(v * 255.0).clamp(0.0, 255.0) as u8` : saturating truncation, NaN -> 0 */
  float t = v * 255.0f;
  if (t != t) return 0;
  t = clampf(t, 0.0f, 255.0f);
  return (uint8_t)t;
}

struct Adj { float brightness, contrast, hue_shift, saturation; };

/* color_adjustment.rs:53-87 */
void adjust_pixel(const Adj& a, const uint8_t in[3], uint8_t out[3]) {
  float r = (float)in[0] / 255.0f, g = (float)in[1] / 255.0f, b = (float)in[2] / 255.0f;
  float br = brightness_contrast(r, a.brightness, a.contrast);
  float bg = brightness_contrast(g, a.brightness, a.contrast);
  float bb = brightness_contrast(b, a.brightness, a.contrast);
  float fr = br, fg = bg, fb = bb;
  if (a.hue_shift != 0.0f || a.saturation != 1.0f) {
    float h, s, v;
    rgb_to_hsv(br, bg, bb, h, s, v);
    float nh = rem_euclid(h + a.hue_shift, 360.0f);
    float ns = clampf(s * a.saturation, 0.0f, 1.0f);
    hsv_to_rgb(nh, ns, v, fr, fg, fb);
  }
  out[0] = to_u8(fr);
  out[1] = to_u8(fg);
  out[2] = to_u8(fb);
}

/* color_adjustment.rs:28-37 */
Adj adj_new(float brightness, float contrast, float hue_shift, float saturation) {
  return Adj{clampf(brightness, -1.0f, 1.0f), clampf(contrast, 0.0f, 2.0f), clampf(hue_shift, -180.0f, 180.0f),
             clampf(saturation, 0.0f, 2.0f)};
}

/* color_adjustment.rs:99-154 */
Adj optimal_adjustment(const uint8_t tile[3], const uint8_t target[3], float strength) {
  float tr = (float)tile[0] / 255.0f, tg = (float)tile[1] / 255.0f, tb = (float)tile[2] / 255.0f;
  float gr = (float)target[0] / 255.0f, gg = (float)target[1] / 255.0f, gb = (float)target[2] / 255.0f;
  float tile_luma = 0.299f * tr + 0.587f * tg + 0.114f * tb;
  float target_luma = 0.299f * gr + 0.587f * gg + 0.114f * gb;
  float brightness_diff = (target_luma - tile_luma) * strength;
  float th, ts, tv, gh, gs, gv;
  rgb_to_hsv(tr, tg, tb, th, ts, tv);
  rgb_to_hsv(gr, gg, gb, gh, gs, gv);
  float hue_diff = 0.0f;
  if (ts > 0.1f && gs > 0.1f) {
    float diff = gh - th;
    float wrapped = diff > 180.0f ? diff - 360.0f : (diff < -180.0f ? diff + 360.0f : diff);
    hue_diff = wrapped * strength * 0.5f;
  }
  float sat_ratio = 1.0f;
  if (ts > 0.01f) {
    float ratio = gs / ts;
    sat_ratio = 1.0f + (ratio - 1.0f) * strength * 0.7f;
  }
  return adj_new(brightness_diff, 1.0f, hue_diff, sat_ratio);
}
}  // namespace

extern "C" {

int orc_resize_window(int in_size, int out_size) { return axis_window(in_size, out_size); }

int orc_resize_coeffs(int in_size, int out_size, int32_t* bounds_start, int32_t* bounds_size,
                      int16_t* coeffs, int* precision) {
  AxisCoeffs c = axis_coeffs(in_size, out_size);
  std::memcpy(bounds_start, c.start.data(), sizeof(int32_t) * out_size);
  std::memcpy(bounds_size, c.size.data(), sizeof(int32_t) * out_size);
  std::memcpy(coeffs, c.w.data(), sizeof(int16_t) * c.w.size());
  *precision = c.precision;
  return c.window;
}

void orc_resize_rgb8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  resize_rgb8(src, sw, sh, dst, dw, dh);
}

void orc_color_adjustment_new(float brightness, float contrast, float hue_shift, float saturation, float out[4]) {
  Adj a = adj_new(brightness, contrast, hue_shift, saturation);
  out[0] = a.brightness; out[1] = a.contrast; out[2] = a.hue_shift; out[3] = a.saturation;
}
void orc_optimal_adjustment(const uint8_t tile_avg[3], const uint8_t target_avg[3], float strength, float out[4]) {
  Adj a = optimal_adjustment(tile_avg, target_avg, strength);
  out[0] = a.brightness; out[1] = a.contrast; out[2] = a.hue_shift; out[3] = a.saturation;
}
void orc_adjust_pixel(const float adj[4], const uint8_t in[3], uint8_t out[3]) {
  Adj a{adj[0], adj[1], adj[2], adj[3]};
  adjust_pixel(a, in, out);
}
void orc_rgb_to_hsv(float r, float g, float b, float out[3]) { rgb_to_hsv(r, g, b, out[0], out[1], out[2]); }
void orc_hsv_to_rgb(float h, float s, float v, float out[3]) { hsv_to_rgb(h, s, v, out[0], out[1], out[2]); }

void orc_render(const uint8_t* target, int img_w, int img_h, int gw, int gh,
                const uint8_t* materials, int n, int mw, int mh,
                const int32_t* placed, float strength, uint8_t* out, int threads) {
  (void)n;
  int tw = img_w / gw, th = img_h / gh;
  int ow = gw * tw;
  strength = clampf(strength, 0.0f, 1.0f); /* main.rs:162 */
  std::memset(out, 0, (size_t)ow * (size_t)(gh * th) * 3); /* ImageBuffer::new is zero-filled (main.rs:518) */
  parallel_for((int64_t)gw * gh, threads, [=](int64_t b, int64_t e) {
    std::vector<uint8_t> tile((size_t)tw * th * 3);
    for (int64_t c = b; c < e; ++c) {
      int x = (int)(c % gw), y = (int)(c / gw);
      int32_t m = placed[c];
      if (m < 0) continue; /* main.rs:628 `if let Some(tile_path)` ; output stays zero */
      resize_rgb8(materials + (size_t)m * mw * mh * 3, mw, mh, tile.data(), tw, th); /* main.rs:549 */
      if (strength > 0.0f) { /* main.rs:552-565 */
        uint8_t target_avg[3], tile_avg[3];
        orc_average_rgb_region(target, img_w, x * tw, y * th, tw, th, target_avg);
        orc_average_rgb_region(tile.data(), tw, 0, 0, tw, th, tile_avg);
        Adj a = optimal_adjustment(tile_avg, target_avg, strength);
        for (int p = 0; p < tw * th; ++p) adjust_pixel(a, &tile[3 * (size_t)p], &tile[3 * (size_t)p]);
      }
      for (int dy = 0; dy < th; ++dy) /* main.rs:583-592 composite */
        std::memcpy(out + (((size_t)y * th + dy) * ow + (size_t)x * tw) * 3, &tile[(size_t)dy * tw * 3], (size_t)tw * 3);
    }
  });
}

}  // extern "C"
