// lab_mean.cu — K1: mean CIE L*a*b* (D65) of RGB8 images / target-grid cells.
//
// Replaces MosaicGeneratorImpl::calculate_average_lab (reference src/lib.rs:94-119) at its call
// sites main.rs:280/294 (materials) and main.rs:540-541 (crop of every grid cell).
//
// Per pixel: 3 sRGB->linear lookups (256-entry table, replicated x8 in shared memory so that the
// 32 lanes of a warp rarely collide on a bank), a 3x3 linear-RGB -> XYZ/white matrix (9 FMA), three
// f(t) = cbrt(t) | kappa*t+delta evaluations (cbrt = MUFU lg2/ex2 seed + one Newton step, ~1 ulp),
// and three running sums  sum(fy), sum(fx-fy), sum(fy-fz)  — L, a, b are affine in those, so the
// affine map is applied once per image.  One warp owns one image / one cell: per-lane partial sums
// in a fixed pixel order + a fixed shuffle tree => results are run-to-run deterministic.
// The reference folds the pixels sequentially in f32; a tree differs from that in the last bits,
// which is what the <= 1e-4 relative tolerance of the north star covers.
#include "common.cuh"

#include <cmath>

namespace mosaic {

__constant__ float c_srgb_lin[256];

constexpr int kLutRep = 8;

#ifndef MOSAIC_LABF_VARIANT
#define MOSAIC_LABF_VARIANT 2
#endif
__device__ __forceinline__ float lab_f(float t) {
  // palette lab.rs: epsilon = (6/29)^3, kappa = 841/108, delta = 4/29
  const float eps = 0.008856452f, kappa = 7.787037f, delta = 0.13793103f;
  float lg, z;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(t));
#if MOSAIC_LABF_VARIANT == 2
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(z) : "f"(lg * 0.33333334f));  // MUFU only, ~2^-21 relative
  const float y = z;
#else
  // z ~ t^(-1/3) from two MUFUs (~2^-21 relative), one division-free Newton step on z^-3 = t, then cbrt = t z^2
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(z) : "f"(lg * -0.33333334f));
  const float z2 = z * z;
  const float e = fmaf(-(t * z2), z, 1.0f);
  z = fmaf(z * e, 0.33333334f, z);
  const float y = t * z * z;
#endif
  const float lin = fmaf(kappa, t, delta);
  return t > eps ? y : lin;  // (t == 0: the MUFU path produces inf/NaN, which the select discards)
}

struct LabAcc {
  float fy = 0.f, dxy = 0.f, dyz = 0.f;
  unsigned r = 0, g = 0, b = 0;  // integer channel sums (calculate_average_rgb, main.rs:687-708)
};

__device__ __forceinline__ void accumulate_pixel(const float* __restrict__ lut, unsigned r8, unsigned g8,
                                                 unsigned b8, LabAcc& acc) {
  float r = lut[r8 * kLutRep], g = lut[g8 * kLutRep], b = lut[b8 * kLutRep];
  // sRGB/D65 matrix (Lindbloom constants used by palette) with the division by the white point
  // (0.95047, 1, 1.08883) folded into the rows.
  float x = fmaf(0.18984029f, b, fmaf(0.37620977f, g, 0.433949941f * r));
  float y = fmaf(0.0721750f, b, fmaf(0.7151522f, g, 0.2126729f * r));
  float z = fmaf(0.872775456f, b, fmaf(0.109467961f, g, 0.0177565828f * r));
  float fx = lab_f(x), fy = lab_f(y), fz = lab_f(z);
  acc.fy += fy;
  acc.dxy += fx - fy;
  acc.dyz += fy - fz;
}

__device__ __forceinline__ void load_lut(float* lut) {
  for (int i = threadIdx.x; i < 256 * kLutRep; i += blockDim.x) lut[i] = c_srgb_lin[i / kLutRep];
  __syncthreads();
}

__device__ __forceinline__ void finish(LabAcc acc, float inv_count_num, float* out) {
  float sfy = warp_sum(acc.fy), sxy = warp_sum(acc.dxy), syz = warp_sum(acc.dyz);
  if ((threadIdx.x & 31) == 0) {
    float mfy = sfy / inv_count_num, mxy = sxy / inv_count_num, myz = syz / inv_count_num;
    out[0] = fmaf(116.0f, mfy, -16.0f);
    out[1] = 500.0f * mxy;
    out[2] = 200.0f * myz;
  }
}

// One warp per image; images are contiguous RGB8 of `pixels` pixels, 16-byte aligned when
// (pixels*3) % 16 == 0 (the vector path: each lane streams 48-byte groups = 16 pixels as 3 x LDG.128).
__global__ void __launch_bounds__(256) lab_mean_images_kernel(const uint8_t* __restrict__ imgs, int n, int pixels,
                                                              int vec_ok, float* __restrict__ lab_out) {
  __shared__ float lut_s[256 * kLutRep];
  load_lut(lut_s);
  const int lane = threadIdx.x & 31;
  const float* lut = lut_s + (lane & (kLutRep - 1));
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int img = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; img < n; img += warps_per_grid) {
    const uint8_t* base = imgs + (size_t)img * pixels * 3;
    LabAcc acc;
    int done = 0;
    if (vec_ok) {
      const uint4* p = reinterpret_cast<const uint4*>(base);
      const int groups = pixels / 16;
      for (int gidx = lane; gidx < groups; gidx += 32) {
        uint4 v0 = __ldg(p + 3 * gidx), v1 = __ldg(p + 3 * gidx + 1), v2 = __ldg(p + 3 * gidx + 2);
        unsigned w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // 12 bytes = 4 pixels per 3 words
          unsigned a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
          accumulate_pixel(lut, a & 255u, (a >> 8) & 255u, (a >> 16) & 255u, acc);
          accumulate_pixel(lut, a >> 24, b & 255u, (b >> 8) & 255u, acc);
          accumulate_pixel(lut, (b >> 16) & 255u, b >> 24, c & 255u, acc);
          accumulate_pixel(lut, (c >> 8) & 255u, (c >> 16) & 255u, c >> 24, acc);
        }
      }
      done = groups * 16;
    }
    for (int px = done + lane; px < pixels; px += 32) {
      const uint8_t* q = base + 3 * (size_t)px;
      accumulate_pixel(lut, q[0], q[1], q[2], acc);
    }
    finish(acc, (float)pixels, lab_out + 3 * (size_t)img);
  }
}

// Materials of individual sizes (the reference opens files of any size, main.rs:263-301): one CTA per image — sizes
// differ by orders of magnitude, so a warp per image would leave the GPU waiting for the largest ones.  Threads take
// groups of 4 pixels (12 bytes) strided by the CTA; sums are kept in two levels (64 pixels per thread, then folded)
// so that a multi-megapixel image does not lose the small a/b terms in a long single-precision accumulation, and the
// per-warp totals are combined in warp order: run-to-run deterministic.
__global__ void __launch_bounds__(256) lab_mean_images_ragged_kernel(const uint8_t* __restrict__ imgs,
                                                                     const unsigned long long* __restrict__ offs,
                                                                     const int32_t* __restrict__ widths,
                                                                     const int32_t* __restrict__ heights, int mat_begin,
                                                                     int mat_end, float* __restrict__ lab_out) {
  __shared__ float lut_s[256 * kLutRep];
  __shared__ float s_part[8][3];
  load_lut(lut_s);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lut = lut_s + (lane & (kLutRep - 1));
  for (int img = mat_begin + blockIdx.x; img < mat_end; img += gridDim.x) {
    const uint8_t* base = imgs + offs[img];
    const int pixels = widths[img] * heights[img];
    LabAcc total;
    // the image starts at an arbitrary byte: a thread reads the four ALIGNED words that cover its 12 bytes and
    // funnel-shifts them by the image's byte phase (the buffer is allocated with slack for the last word)
    const unsigned phase = (unsigned)((uintptr_t)base & 3u) * 8u;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(base - ((uintptr_t)base & 3u));
    const int groups = pixels / 4;
    for (int g0 = 0; g0 < groups; g0 += 256 * 16) {
      LabAcc acc;
      const int g1 = min(groups, g0 + 256 * 16);
      for (int g = g0 + tid; g < g1; g += 256) {
        const uint32_t* w = words + 3 * (size_t)g;
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
        const unsigned a = __funnelshift_r(w0, w1, phase), b = __funnelshift_r(w1, w2, phase), c = __funnelshift_r(w2, w3, phase);
        accumulate_pixel(lut, a & 255u, (a >> 8) & 255u, (a >> 16) & 255u, acc);
        accumulate_pixel(lut, a >> 24, b & 255u, (b >> 8) & 255u, acc);
        accumulate_pixel(lut, (b >> 16) & 255u, b >> 24, c & 255u, acc);
        accumulate_pixel(lut, (c >> 8) & 255u, (c >> 16) & 255u, c >> 24, acc);
      }
      total.fy += acc.fy;
      total.dxy += acc.dxy;
      total.dyz += acc.dyz;
    }
    for (int px = groups * 4 + tid; px < pixels; px += 256) {
      const uint8_t* q = base + 3 * (size_t)px;
      accumulate_pixel(lut, q[0], q[1], q[2], total);
    }
    const float sfy = warp_sum(total.fy), sxy = warp_sum(total.dxy), syz = warp_sum(total.dyz);
    if (lane == 0) {
      s_part[warp][0] = sfy;
      s_part[warp][1] = sxy;
      s_part[warp][2] = syz;
    }
    __syncthreads();
    if (tid == 0) {
      float fy = 0.f, xy = 0.f, yz = 0.f;
      for (int w = 0; w < 8; ++w) {
        fy += s_part[w][0];
        xy += s_part[w][1];
        yz += s_part[w][2];
      }
      const float cnt = (float)pixels;
      lab_out[3 * (size_t)img] = fmaf(116.0f, fy / cnt, -16.0f);
      lab_out[3 * (size_t)img + 1] = 500.0f * (xy / cnt);
      lab_out[3 * (size_t)img + 2] = 200.0f * (yz / cnt);
    }
    __syncthreads();  // s_part is reused by the CTA's next image
  }
}

// One warp per grid cell: the tile_w x tile_h crop at (x*tile_w, y*tile_h) of the target
// (main.rs:536-541).  Pixels are visited in the crop's row-major order, lane-strided.  Also emits
// the integer channel sums of the crop (target_avg_rgb of main.rs:554, used by the render kernel).
__global__ void __launch_bounds__(256) lab_mean_cells_kernel(const uint8_t* __restrict__ img, int img_w, int gw,
                                                             int cell_begin, int cell_end, int tile_w, int tile_h,
                                                             float* __restrict__ lab_out,
                                                             unsigned long long* __restrict__ rgbsum_out) {
  __shared__ float lut_s[256 * kLutRep];
  load_lut(lut_s);
  const int lane = threadIdx.x & 31;
  const float* lut = lut_s + (lane & (kLutRep - 1));
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const int pixels = tile_w * tile_h;
  for (int cell = cell_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); cell < cell_end;
       cell += warps_per_grid) {
    const int cx = cell % gw, cy = cell / gw;
    const uint8_t* base = img + ((size_t)cy * tile_h * img_w + (size_t)cx * tile_w) * 3;
    LabAcc acc;
    int py = lane / tile_w, px = lane - py * tile_w;  // lane < 32 <= any sensible tile_w is NOT assumed
    const int step_y = 32 / tile_w, step_x = 32 - step_y * tile_w;
    for (int p = lane; p < pixels; p += 32) {
      const uint8_t* q = base + ((size_t)py * img_w + px) * 3;
      unsigned r8 = q[0], g8 = q[1], b8 = q[2];
      accumulate_pixel(lut, r8, g8, b8, acc);
      acc.r += r8;
      acc.g += g8;
      acc.b += b8;
      px += step_x;
      py += step_y;
      if (px >= tile_w) {
        px -= tile_w;
        py += 1;
      }
    }
    finish(acc, (float)pixels, lab_out + 3 * (size_t)cell);
    unsigned long long sr = acc.r, sg = acc.g, sb = acc.b;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane == 0) {
      rgbsum_out[3 * (size_t)cell] = sr;
      rgbsum_out[3 * (size_t)cell + 1] = sg;
      rgbsum_out[3 * (size_t)cell + 2] = sb;
    }
  }
}

// TMA-staged variant of the cell kernel.  One warp per cell; the tile is brought into shared memory by 1-D bulk
// copies — one per tile row, issued in parallel by the lanes (lane r copies the 16-byte-aligned span that covers
// row r of the crop), all completing on one mbarrier — and the NEXT cell of the warp is already in flight while
// the current one is reduced (two buffers per warp).  Pixels are then read from shared memory in the crop's
// row-major order, lane-strided, exactly like the direct kernel, so both produce identical sums.
__global__ void __launch_bounds__(256) lab_mean_cells_tma_kernel(const uint8_t* __restrict__ img, int img_w, int gw,
                                                                 int cell_begin, int cell_end, int tile_w, int tile_h,
                                                                 int row_pitch /* bytes per staged row, multiple of 16 */,
                                                                 float* __restrict__ lab_out,
                                                                 unsigned long long* __restrict__ rgbsum_out) {
  extern __shared__ __align__(128) unsigned char stage_raw[];
  __shared__ float lut_s[256 * kLutRep];
  __shared__ unsigned long long bars[8][2];
  load_lut(lut_s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* lut = lut_s + (lane & (kLutRep - 1));
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const int pixels = tile_w * tile_h;
  const size_t buf_bytes = (size_t)row_pitch * tile_h;
  unsigned char* buf0 = stage_raw + (size_t)warp * 2 * buf_bytes;
  if (lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const size_t img_row_bytes = (size_t)img_w * 3;
  auto issue = [&](int cell, int b) {  // stage the crop of `cell` into buffer b
    const int cx = cell % gw, cy = cell / gw;
    const size_t first = ((size_t)cy * tile_h) * img_row_bytes + (size_t)cx * tile_w * 3;  // byte offset of the crop
    unsigned total = 0;  // the barrier must expect every byte BEFORE its single arrival
    for (int r = lane; r < tile_h; r += 32) {
      const size_t start = first + (size_t)r * img_row_bytes;
      total += (unsigned)(((start + (size_t)tile_w * 3 + 15) & ~(size_t)15) - (start & ~(size_t)15));
    }
    total = warp_sum_u32(total);
    if (lane == 0) mbar_expect_tx(&bars[warp][b], total);
    __syncwarp();
    for (int r = lane; r < tile_h; r += 32) {
      const size_t start = first + (size_t)r * img_row_bytes;
      const size_t a0 = start & ~(size_t)15, a1 = (start + (size_t)tile_w * 3 + 15) & ~(size_t)15;
      tma_load_1d(buf0 + (size_t)b * buf_bytes + (size_t)r * row_pitch, img + a0, (unsigned)(a1 - a0), &bars[warp][b]);
    }
  };
  int cell = cell_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int it = 0;
  if (cell < cell_end) issue(cell, 0);
  for (; cell < cell_end; cell += warps_per_grid, ++it) {
    const int b = it & 1;
    const int next = cell + warps_per_grid;
    if (next < cell_end) issue(next, b ^ 1);
    mbar_wait(&bars[warp][b], (unsigned)((it >> 1) & 1));
    const int cx = cell % gw, cy = cell / gw;
    const size_t first = ((size_t)cy * tile_h) * img_row_bytes + (size_t)cx * tile_w * 3;
    const unsigned char* base = buf0 + (size_t)b * buf_bytes;
    LabAcc acc;
    int py = lane / tile_w, px = lane - py * tile_w;
    const int step_y = 32 / tile_w, step_x = 32 - step_y * tile_w;
    for (int p = lane; p < pixels; p += 32) {
      const unsigned off = (unsigned)((first + (size_t)py * img_row_bytes) & 15);  // the row's offset inside its aligned span
      const unsigned char* q = base + (size_t)py * row_pitch + off + 3 * px;
      const unsigned r8 = q[0], g8 = q[1], b8 = q[2];
      accumulate_pixel(lut, r8, g8, b8, acc);
      acc.r += r8;
      acc.g += g8;
      acc.b += b8;
      px += step_x;
      py += step_y;
      if (px >= tile_w) {
        px -= tile_w;
        py += 1;
      }
    }
    finish(acc, (float)pixels, lab_out + 3 * (size_t)cell);
    unsigned long long sr = acc.r, sg = acc.g, sb = acc.b;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane == 0) {
      rgbsum_out[3 * (size_t)cell] = sr;
      rgbsum_out[3 * (size_t)cell + 1] = sg;
      rgbsum_out[3 * (size_t)cell + 2] = sb;
    }
    __syncwarp();  // every lane is done with buffer b before it is refilled two iterations later
  }
}

// n*3 packed Lab -> float4 gather table + SoA (l[], a[], b[]; padded with +inf so that padding can
// never rank among the nearest).
__global__ void pack_lab_kernel(const float* __restrict__ lab, int n, int n_pad, float4* __restrict__ lab4,
                                float* __restrict__ soa) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float l = __int_as_float(0x7f800000), a = l, b = l;
  if (i < n) {
    l = lab[3 * i];
    a = lab[3 * i + 1];
    b = lab[3 * i + 2];
    lab4[i] = make_float4(l, a, b, 0.f);
  }
  soa[i] = l;
  soa[n_pad + i] = a;
  soa[2 * n_pad + i] = b;
}

static int upload_lut(mosaic_ctx* ctx) {
  static bool done[64] = {};
  if (ctx->device < 64 && done[ctx->device]) return MOSAIC_OK;
  float lut[256];
  for (int i = 0; i < 256; ++i) {
    // palette 0.7.6 Srgb::into_linear on `u8 as f32 / 255.0` (lib.rs:103): evaluated once per code
    // value on the host with the same f32 formula the reference applies per pixel.
    float x = (float)i / 255.0f;
    lut[i] = x <= 0.04045f ? (float)(1.0 / 12.92) * x
                           : powf(fmaf(x, (float)(1.0 / 1.055), (float)(0.055 / 1.055)), 2.4f);
  }
  MOSAIC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_srgb_lin, lut, sizeof lut, 0, cudaMemcpyHostToDevice, ctx->stream));
  MOSAIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->device < 64) done[ctx->device] = true;
  return MOSAIC_OK;
}

int launch_material_lab(mosaic_ctx* ctx) {
  MOSAIC_TRY(upload_lut(ctx));
  const int pixels = ctx->mw * ctx->mh;
  const int vec_ok = ((size_t)pixels * 3) % 16 == 0 && ((uintptr_t)ctx->d_materials % 16) == 0;
  // one warp per image, 8 warps per CTA; at most 8 resident CTAs per SM worth of grid
  int blocks = (ctx->n + 7) / 8;
  int max_blocks = ctx->sm_count * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  {
    StageTimer t(ctx, MOSAIC_STAGE_MATERIAL_LAB);
    lab_mean_images_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_materials, ctx->n, pixels, vec_ok, ctx->d_mat_lab);
    ctx->launches++;
  }
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

int launch_material_lab_ragged(mosaic_ctx* ctx, int mat_begin, int mat_end) {
  MOSAIC_TRY(upload_lut(ctx));
  int blocks = mat_end - mat_begin;
  const int max_blocks = ctx->sm_count * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  {
    StageTimer t(ctx, MOSAIC_STAGE_MATERIAL_LAB);
    lab_mean_images_ragged_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_materials, ctx->d_mat_off, ctx->d_mat_w, ctx->d_mat_h,
                                                                   mat_begin, mat_end, ctx->d_mat_lab);
    ctx->launches++;
  }
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

int launch_pack_lab(mosaic_ctx* ctx) {
  int threads = 256, blocks = (ctx->n_pad + threads - 1) / threads;
  pack_lab_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->d_mat_lab, ctx->n, ctx->n_pad, ctx->d_mat_lab4,
                                                       ctx->d_mat_soa);
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

int launch_cell_lab_range(mosaic_ctx* ctx, int cell_begin, int cell_end) {
  MOSAIC_TRY(upload_lut(ctx));
  int cells = cell_end - cell_begin;
  if (cells <= 0) return MOSAIC_OK;
  int blocks = (cells + 7) / 8;
  int max_blocks = ctx->sm_count * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  // TMA-staged path: the aligned spans must stay inside the allocation (owned buffers are cudaMalloc'ed, i.e. padded
  // to 256 B; a caller-owned image qualifies when its size is a multiple of 16) and two tiles per warp must fit
  const size_t row_pitch = (((size_t)ctx->tile_w * 3 + 15 + 15) & ~(size_t)15);
  const size_t stage = 8 * 2 * row_pitch * (size_t)ctx->tile_h;
  const bool owned = ctx->d_target == ctx->d_target_owned;
  const bool tma_ok = ((uintptr_t)ctx->d_target % 16 == 0) && (owned || ((size_t)ctx->tw_img * ctx->th_img * 3) % 16 == 0) &&
                      stage <= 96 * 1024;
  {
    StageTimer t(ctx, MOSAIC_STAGE_CELL_LAB);
    if (tma_ok) {
      MOSAIC_CUDA(ctx, cudaFuncSetAttribute(lab_mean_cells_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
      int tblocks = (cells + 7) / 8;
      if (tblocks > ctx->sm_count * 2) tblocks = ctx->sm_count * 2;  // two resident CTAs per SM, several cells per warp
      lab_mean_cells_tma_kernel<<<tblocks, 256, stage, ctx->stream>>>(ctx->d_target, ctx->tw_img, ctx->gw, cell_begin,
                                                                       cell_end, ctx->tile_w, ctx->tile_h, (int)row_pitch,
                                                                       ctx->d_cell_lab, ctx->d_cell_rgbsum);
    } else {
      lab_mean_cells_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_target, ctx->tw_img, ctx->gw, cell_begin, cell_end,
                                                             ctx->tile_w, ctx->tile_h, ctx->d_cell_lab,
                                                             ctx->d_cell_rgbsum);
    }
    ctx->launches++;
  }
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

}  // namespace mosaic
