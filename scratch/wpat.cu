// store-pattern microbenchmark for the packed triangular matrix (K3): which write pattern reaches the HBM write roof?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 row_off(u64 i, u64 n) { return i * n - i * (i + 1) / 2; }
// MODE 0: 8 rows per band, thread writes STG.32 at columns j and j+256 of each row (the kernel's pattern)
// MODE 1: same, but every row base rounded down to 16 B (alignment effect only; overlaps neighbours, timing only)
// MODE 2: 8 rows per band, thread writes one STG.128 (4 consecutive floats), bases 16-B aligned
// MODE 3: 1 row per band, STG.128, aligned
// MODE 4: 8 rows per band, STG.32, but rows visited one after the other (row loop outside the column loop)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, unsigned n, unsigned* ctr) {
  __shared__ unsigned sb;
  const unsigned tid = threadIdx.x;
  constexpr unsigned ROWS = MODE == 3 ? 1 : 8;
  const unsigned nb = (n - 1 + ROWS - 1) / ROWS;
  for (;;) {
    __syncthreads();
    if (tid == 0) sb = atomicAdd(ctr, 1u);
    __syncthreads();
    const unsigned band = sb;
    if (band >= nb) break;
    const unsigned i0 = band * ROWS;
    long long base[ROWS];
#pragma unroll
    for (unsigned r = 0; r < ROWS; ++r) {
      const unsigned i = min(i0 + r, n - 2);
      long long b = (long long)row_off(i, n) - (long long)i - 1ll;
      if (MODE == 1 || MODE == 2 || MODE == 3) b &= ~3ll;
      base[r] = b;
    }
    const float v = (float)band;
    if (MODE == 0 || MODE == 1) {
      for (unsigned j = i0 + ROWS + tid; j + 256 < n; j += 512)
#pragma unroll
        for (unsigned r = 0; r < ROWS; ++r) {
          out[base[r] + j] = v;
          out[base[r] + j + 256] = v;
        }
    } else if (MODE == 2 || MODE == 3) {
      for (unsigned j = ((i0 + ROWS + 3) & ~3u) + 4 * tid; j + 3 < n; j += 1024)
#pragma unroll
        for (unsigned r = 0; r < ROWS; ++r) *reinterpret_cast<float4*>(out + base[r] + j) = make_float4(v, v, v, v);
    } else {
#pragma unroll 1
      for (unsigned r = 0; r < ROWS; ++r)
        for (unsigned j = i0 + ROWS + tid; j + 256 < n; j += 512) {
          out[base[r] + j] = v;
          out[base[r] + j + 256] = v;
        }
    }
  }
}
template <int MODE>
void run(float* out, unsigned n, unsigned* ctr, int blocks) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaMemset(ctr, 0, 4);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(out, n, ctr);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double bytes = (double)n * (n - 1) / 2 * 4;
  printf("{\"mode\": %d, \"blocks\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n", MODE, blocks, best, bytes / best / 1e6);
}
int main(int argc, char** argv) {
  const unsigned n = argc > 1 ? atoi(argv[1]) : 50000;
  float* out;
  unsigned* ctr;
  cudaMalloc(&out, (size_t)n * (n - 1) / 2 * 4 + 4096);
  cudaMalloc(&ctr, 4);
  for (int blocks : {296, 444, 592, 1184}) {
    run<0>(out, n, ctr, blocks);
    run<1>(out, n, ctr, blocks);
    run<2>(out, n, ctr, blocks);
    run<3>(out, n, ctr, blocks);
    run<4>(out, n, ctr, blocks);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
