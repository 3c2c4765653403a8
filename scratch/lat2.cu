// micro-benchmark of the placement chain: cycles per iteration of the loop-carried dependency
#include <cstdio>
#include <cuda_runtime.h>
#define N 8192
__device__ __forceinline__ unsigned redux_min_ptx(unsigned v) {
  unsigned r;
  asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
template <int OP>
__global__ void k(unsigned* out, const unsigned* in, long long* cyc) {
  __shared__ unsigned sm[256];
  __shared__ unsigned char us[256];
  const unsigned lane = threadIdx.x;
  for (int i = lane; i < 256; i += 32) { sm[i] = in[i]; us[i] = 0; }
  __syncwarp();
  unsigned prev = in[300], acc = 0;
  unsigned id0 = (in[lane] & 0xffff) | (lane << 26), id1 = (in[32 + lane] & 0xffff) | (lane << 26);
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
    const unsigned word = (i & 1) ? id1 : id0;
    const unsigned id = word & 0x3ffffffu;
    const unsigned k0 = word + (i & 2), k1 = word | 0x80000000u;  // off-chain
    if (OP == 0) {  // compare -> select -> redux.min -> mask
      const unsigned key = id == prev ? k1 : k0;
      prev = __reduce_min_sync(0xffffffffu, key) & 0x3ffffffu;
    } else if (OP == 1) {  // + vote.any + never-taken branch
      const unsigned key = id == prev ? k1 : k0;
      const unsigned r = __reduce_min_sync(0xffffffffu, key);
      const unsigned j0 = r >> 26;
      if (__any_sync(0xffffffffu, (j0 - (lane + 40)) < 2u)) { acc += sm[r & 255]; __nanosleep(100); }
      prev = r & 0x3ffffffu;
    } else if (OP == 2) {  // ballot -> ffs -> shfl
      const bool valid = id != prev && (k0 & 4) == 0;
      const unsigned m = __ballot_sync(0xffffffffu, valid);
      const int j0 = (__ffs(m) - 1) & 31;
      prev = __shfl_sync(0xffffffffu, id, j0);
    } else if (OP == 3) {  // ballot -> lanemask compare -> redux on single contributor (or)
      const bool valid = id != prev && (k0 & 4) == 0;
      const unsigned m = __ballot_sync(0xffffffffu, valid);
      unsigned lt;
      asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
      const bool first = valid && (m & lt) == 0;
      prev = __reduce_or_sync(0xffffffffu, first ? id : 0u);
    } else if (OP == 4) {  // PTX redux
      const unsigned key = id == prev ? k1 : k0;
      prev = redux_min_ptx(key) & 0x3ffffffu;
    } else if (OP == 5) {  // full step shape: redux + vote + predicated sts + syncwarp + lds for next
      const unsigned key = id == prev ? k1 : k0;
      const unsigned r = __reduce_min_sync(0xffffffffu, key);
      const unsigned u = us[(word >> 3) & 255];
      const unsigned j0 = r >> 26;
      if (__any_sync(0xffffffffu, (j0 - (lane + 40)) < 2u)) { acc += sm[r & 255]; __nanosleep(100); }
      if (lane == j0) us[id & 255] = (unsigned char)(u + 1);
      if (lane == 0) { ((volatile unsigned*)sm)[i & 63] = r; }
      __syncwarp();
      acc += u;
      prev = r & 0x3ffffffu;
    } else if (OP == 6) {  // OP 5 without syncwarp
      const unsigned key = id == prev ? k1 : k0;
      const unsigned r = __reduce_min_sync(0xffffffffu, key);
      const unsigned u = us[(word >> 3) & 255];
      const unsigned j0 = r >> 26;
      if (__any_sync(0xffffffffu, (j0 - (lane + 40)) < 2u)) { acc += sm[r & 255]; __nanosleep(100); }
      if (lane == j0) us[id & 255] = (unsigned char)(u + 1);
      if (lane == 0) { ((volatile unsigned*)sm)[i & 63] = r; }
      acc += u;
      prev = r & 0x3ffffffu;
    } else if (OP == 7) {  // fminf-style: key via shfl-free butterfly? (5 shfl.bfly + min) for reference
      unsigned key = id == prev ? k1 : k0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
      prev = key & 0x3ffffffu;
    } else if (OP == 8) {  // match/elect style: ballot + brev/flo only (no broadcast) -> j0 chain
      const bool valid = (lane + prev) & 1;
      const unsigned m = __ballot_sync(0xffffffffu, valid);
      prev = __ffs(m) + i;
    } else if (OP == 9) {  // ballot + (m & -m) isolate + popc-free: chain on mask only
      const bool valid = ((prev >> lane) & 1) ^ (i & 1);
      const unsigned m = __ballot_sync(0xffffffffu, valid);
      prev = (m & (0u - m)) + i;
    }
    acc ^= prev;
  }
  long long t1 = clock64();
  out[lane] = acc;
  if (lane == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, unsigned* d, unsigned* in, long long* dc) {
  k<OP><<<1, 32>>>(d, in, dc); cudaDeviceSynchronize();
  k<OP><<<1, 32>>>(d, in, dc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-60s %7.1f cycles/iter  (%s)\n", name, (double)c / N, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  unsigned *d, *in; long long* dc; cudaMalloc(&d, 256); cudaMalloc(&dc, 8); cudaMalloc(&in, 4096);
  unsigned h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1000 + i * 7919u;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  run<0>("cmp+sel+reduce_min+mask", d, in, dc);
  run<4>("cmp+sel+redux.sync.min(ptx)+mask", d, in, dc);
  run<1>("cmp+sel+reduce_min+vote.any+branch", d, in, dc);
  run<2>("cmp+ballot+ffs+shfl", d, in, dc);
  run<3>("cmp+ballot+lanemask+reduce_or", d, in, dc);
  run<5>("full step (redux, vote, sts.u8, sts, syncwarp, lds)", d, in, dc);
  run<6>("full step without syncwarp", d, in, dc);
  run<7>("cmp+sel+5x(shfl.bfly+min)", d, in, dc);
  run<8>("ballot+ffs chain", d, in, dc);
  run<9>("ballot+isolate-lowest chain", d, in, dc);
  return 0;
}
