// micro-benchmark: dependent-chain latency of warp-level ops on one warp (cycles per op)
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(unsigned* out, unsigned seed, long long* cyc) {
  __shared__ unsigned sm[64];
  const int lane = threadIdx.x;
  sm[lane] = lane; sm[lane + 32] = lane;
  __syncwarp();
  unsigned v = seed + lane, acc = 0;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) { v = __ballot_sync(0xffffffffu, (v + lane) & 1) + i; }
    else if (OP == 1) { v = __reduce_min_sync(0xffffffffu, v ^ lane) + i; }
    else if (OP == 2) { v = __shfl_sync(0xffffffffu, v, (v + 1) & 31) + i; }
    else if (OP == 3) { v = __ffs(v | 0x80000000u) + i * 77; }
    else if (OP == 4) { v = sm[(v & 31)] + i; }
    else if (OP == 5) { v = *((volatile unsigned*)&sm[(v & 31)]) + i; }
    else if (OP == 6) { v = __popc(v) + i * 77; }
    else if (OP == 7) { v = __any_sync(0xffffffffu, (v + lane) & 1) + i; }
    else if (OP == 8) { v = (v * 3 + i) ^ lane; }
    else if (OP == 9) { sm[(v & 31)] = v; v = sm[(v & 31) + 0] + i; }
    else if (OP == 10) { v = __float_as_uint(__fsqrt_rn(__uint_as_float((v & 0x3fffff) | 0x3f800000))) + i; }
    else if (OP == 11) { v = __clz(v | 1) + i * 77; }
    else if (OP == 12) { unsigned m = __ballot_sync(0xffffffffu, (v + lane) & 1); v = __shfl_sync(0xffffffffu, v, (__ffs(m | 0x80000000u) - 1)) + i; }
    else if (OP == 13) { v = __match_any_sync(0xffffffffu, v & 3) + i; }
    else if (OP == 14) { v = __reduce_or_sync(0xffffffffu, v ^ lane) + i; }
    else if (OP == 15) { v = __reduce_add_sync(0xffffffffu, v ^ lane) + i; }
    acc ^= v;
  }
  long long t1 = clock64();
  out[lane] = acc;
  if (lane == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, unsigned* d, long long* dc) {
  k<OP><<<1, 32>>>(d, 1234, dc); cudaDeviceSynchronize();
  k<OP><<<1, 32>>>(d, 1234, dc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.1f cycles/iter\n", name, (double)c / N);
}
int main() {
  unsigned* d; long long* dc; cudaMalloc(&d, 256); cudaMalloc(&dc, 8);
  run<8>("alu (imad+lop) baseline", d, dc);
  run<0>("ballot (+iadd)", d, dc);
  run<7>("vote.any (+iadd)", d, dc);
  run<1>("redux.min (+lop+iadd)", d, dc);
  run<14>("redux.or", d, dc);
  run<15>("redux.add", d, dc);
  run<2>("shfl.idx (+2 alu)", d, dc);
  run<3>("ffs (brev+flo)", d, dc);
  run<11>("clz (flo)", d, dc);
  run<6>("popc", d, dc);
  run<4>("lds", d, dc);
  run<5>("lds volatile", d, dc);
  run<9>("sts+lds same addr", d, dc);
  run<10>("fsqrt_rn", d, dc);
  run<12>("ballot+ffs+shfl", d, dc);
  run<13>("match.any", d, dc);
  return 0;
}
