// selftest.cu — device-side self check of the branch-free IEEE helpers (common.cuh) used on the serial
// chains: sqrt_rn_fast / rcp_rn_fast must equal the correctly rounded builtins on their whole domain.
// Runs over ALL 2^32 float bit patterns (a few milliseconds on a B200).
#include "common.cuh"

namespace mosaic {

namespace {
__global__ void selftest_kernel(unsigned long long* mismatches) {
  unsigned long long bad_sqrt = 0, bad_rcp = 0;
  const unsigned long long total = 1ull << 32;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((unsigned)i);
    bool oor = false;
    const float s = sqrt_rn_fast(x, oor);
    if (!oor && __float_as_uint(s) != __float_as_uint(__fsqrt_rn(x))) ++bad_sqrt;
    // every pattern the fast path does not claim must be flagged (so that callers fall back)
    if (oor && (x == 0.0f)) ++bad_sqrt;
    if (x >= 1.0f && x <= 4.2535296e37f) {  // [1, 2^125]
      const float r = rcp_rn_fast(x);
      if (__float_as_uint(r) != __float_as_uint(__frcp_rn(x)) || __float_as_uint(r) != __float_as_uint(__fdiv_rn(1.0f, x)))
        ++bad_rcp;
    }
  }
  if (bad_sqrt) atomicAdd(&mismatches[0], bad_sqrt);
  if (bad_rcp) atomicAdd(&mismatches[1], bad_rcp);
}
}  // namespace

int launch_selftest(mosaic_ctx* ctx, unsigned long long* out) {
  unsigned long long* d = nullptr;
  MOSAIC_CUDA(ctx, cudaMalloc((void**)&d, 2 * sizeof(unsigned long long)));
  MOSAIC_CUDA(ctx, cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
  selftest_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "selftest");
  return MOSAIC_OK;
}

}  // namespace mosaic
