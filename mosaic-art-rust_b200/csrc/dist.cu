// dist.cu — NCCL plumbing for sharded contexts (one process per GPU, 1/2/4/8 GPUs of one host).
//
// The reference is single-process (no collective call sites).  The exchange steps
// the sharded path needs are all "every rank ends up with every rank's contiguous slice":
// material pixels and Lab means (index-sharded), candidate lists (row-sharded) and similarity row
// blocks (range-sharded), plus a broadcast of the placement from the rank that runs the serial
// chain.  Slices are unequal in general, so the allgather is a group of ncclBroadcast calls (one
// per source rank) — NCCL fuses the group into one launch over NVLink 5 / NVSwitch.
//
// libnccl is loaded with dlopen at first use so that single-GPU users do not need it installed and
// so that the process shares whatever libnccl.so.2 is already loaded (e.g. the one PyTorch ships).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace mosaic {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
constexpr int ncclInt8 = 0;  // ncclDataType_t: ncclInt8 = ncclChar = 0
}  // namespace

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* load_nccl(std::string* err) {
  static NcclApi api;
  static bool tried = false, ok = false;
  static std::string load_err;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      load_err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    } else {
      auto sym = [&](const char* s) { return dlsym(api.handle, s); };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
      api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
      api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather && api.AllReduce && api.GroupStart &&
           api.GroupEnd && api.GetErrorString;
      if (!ok) load_err = "libnccl is missing required symbols";
    }
  }
  if (!ok) {
    if (err) *err = load_err;
    return nullptr;
  }
  return &api;
}

static int nccl_fail(mosaic_ctx* ctx, NcclApi* api, ncclResult_t r, const char* what) {
  return fail(ctx, MOSAIC_ERR_NCCL, std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(r) : "?"));
}

#define MOSAIC_NCCL(ctx, api, expr)                                  \
  do {                                                               \
    ncclResult_t _r = (expr);                                        \
    if (_r != 0) return nccl_fail((ctx), (api), _r, #expr);          \
  } while (0)

int dist_unique_id(void* out, size_t bytes) {
  if (!out || bytes < sizeof(ncclUniqueId)) return fail(nullptr, MOSAIC_ERR_INVALID, "unique id buffer must be >= 128 bytes");
  std::string err;
  NcclApi* api = load_nccl(&err);
  if (!api) return fail(nullptr, MOSAIC_ERR_NCCL, err);
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != 0) return nccl_fail(nullptr, api, r, "ncclGetUniqueId");
  std::memcpy(out, &id, sizeof id);
  return MOSAIC_OK;
}

int dist_init(mosaic_ctx* ctx, const void* id_bytes, size_t n) {
  if (!id_bytes || n < sizeof(ncclUniqueId)) return fail(ctx, MOSAIC_ERR_INVALID, "nccl unique id must be 128 bytes");
  std::string err;
  NcclApi* api = load_nccl(&err);
  if (!api) return fail(ctx, MOSAIC_ERR_NCCL, err);
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof id);
  ncclComm_t comm = nullptr;
  MOSAIC_NCCL(ctx, api, api->CommInitRank(&comm, ctx->world, id, ctx->rank));
  ctx->nccl = api;
  ctx->nccl_comm = comm;
  return MOSAIC_OK;
}

void dist_destroy(mosaic_ctx* ctx) {
  dist_peer_close(ctx);
  if (ctx->nccl && ctx->nccl_comm) ctx->nccl->CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

int dist_allgather(mosaic_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (ctx->world == 1) return MOSAIC_OK;
  NcclApi* api = ctx->nccl;
  CollTimer t(ctx, ctx->stream);
  MOSAIC_NCCL(ctx, api, api->AllGather(send, recv, bytes_per_rank, ncclInt8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return MOSAIC_OK;
}

int dist_allgatherv_inplace(mosaic_ctx* ctx, void* buf, const size_t* offsets, const size_t* sizes) {
  if (ctx->world == 1) return MOSAIC_OK;
  NcclApi* api = ctx->nccl;
  CollTimer t(ctx, ctx->stream);
  MOSAIC_NCCL(ctx, api, api->GroupStart());
  for (int r = 0; r < ctx->world; ++r) {
    if (sizes[r] == 0) continue;
    char* p = (char*)buf + offsets[r];
    ncclResult_t rc = api->Broadcast(p, p, sizes[r], ncclInt8, r, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != 0) {
      api->GroupEnd();
      return nccl_fail(ctx, api, rc, "ncclBroadcast");
    }
  }
  MOSAIC_NCCL(ctx, api, api->GroupEnd());
  return MOSAIC_OK;
}

int dist_broadcast(mosaic_ctx* ctx, void* buf, size_t bytes, int root) {
  if (ctx->world == 1) return MOSAIC_OK;
  NcclApi* api = ctx->nccl;
  CollTimer t(ctx, ctx->stream);
  MOSAIC_NCCL(ctx, api, api->Broadcast(buf, buf, bytes, ncclInt8, root, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return MOSAIC_OK;
}

// ---- peer windows ------------------------------------------------------------------------------------------------
// Every rank maps the other ranks' exchange buffers (CUDA IPC; NVLink 5 / NVSwitch underneath).  The data path of
// mosaic_generate then needs no collective: a producer kernel stores into the consumer rank's buffer and a one-warp
// kernel behind it publishes a counter with release semantics at system scope; the consumer's stream holds a one-warp
// kernel that acquires the counter.  NCCL is used once per buffer generation, to hand the IPC handles around.
namespace {

struct WindowInfo {
  cudaIpcMemHandle_t handle;
  unsigned long long offset;  // of the window inside the allocation the handle names
  unsigned long long valid;
};

typedef int (*cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);
cuMemGetAddressRange_t address_range_fn() {
  static cuMemGetAddressRange_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = (cuMemGetAddressRange_t)dlsym(h, "cuMemGetAddressRange_v2");
  }
  return fn;
}

struct PushDst {
  void* p[mosaic_ctx::kMaxWorld];
};
__global__ void peer_push_kernel_v(PushDst dst, int ndst, const uint32_t* __restrict__ src, size_t words) {
  for (int d = 0; d < ndst; ++d) {
    uint32_t* out = reinterpret_cast<uint32_t*>(dst.p[d]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) out[i] = src[i];
  }
}
struct FlagDst {
  unsigned* p[mosaic_ctx::kMaxWorld];
};
// everything this stream wrote before (also into peer memory) is visible to whoever acquires the counter
__global__ void peer_signal_kernel(FlagDst dst, int ndst, int slot, unsigned value) {
  if ((int)threadIdx.x < ndst) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst.p[threadIdx.x] + slot), "r"(value) : "memory");
  }
}
__global__ void peer_wait_kernel(const unsigned* flags, int first_slot, int count, unsigned value, unsigned* err) {
  if ((int)threadIdx.x < count) {
    const unsigned* f = flags + first_slot + threadIdx.x;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - value) >= 0) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {  // 20 s: a peer died — do not hang the GPU
        atomicExch(err, 1u);
        break;
      }
      __nanosleep(100);
    }
  }
  __syncwarp();
  __threadfence_system();
}

}  // namespace

void dist_peer_close(mosaic_ctx* ctx) {
  // small buffers can share one allocation (one IPC handle): every mapped base is closed once
  std::vector<void*> closed;
  for (int w = 0; w < mosaic_ctx::PW_COUNT; ++w)
    for (int r = 0; r < mosaic_ctx::kMaxWorld; ++r) {
      void* base = ctx->pw_open_base[w][r];
      if (ctx->pw_open[w][r] && base && std::find(closed.begin(), closed.end(), base) == closed.end()) {
        cudaIpcCloseMemHandle(base);
        closed.push_back(base);
      }
      ctx->pw_open[w][r] = false;
      ctx->pw_peer[w][r] = nullptr;
      ctx->pw_open_base[w][r] = nullptr;
    }
  (void)cudaGetLastError();
  ctx->peer_ok = false;
}

int dist_peer_setup(mosaic_ctx* ctx) {
  if (ctx->world == 1 || ctx->peer_failed) return MOSAIC_OK;
  NcclApi* api = ctx->nccl;
  const int W = mosaic_ctx::PW_COUNT, world = ctx->world;
  if (world > mosaic_ctx::kMaxWorld) {
    ctx->peer_failed = true;
    return MOSAIC_OK;
  }
  if (!ctx->d_flags) {
    MOSAIC_CUDA(ctx, cudaMalloc((void**)&ctx->d_flags, 256 * sizeof(unsigned)));
    MOSAIC_CUDA(ctx, cudaMemset(ctx->d_flags, 0, 256 * sizeof(unsigned)));
    MOSAIC_CUDA(ctx, cudaMalloc((void**)&ctx->d_peer_err, sizeof(unsigned)));
    MOSAIC_CUDA(ctx, cudaMemset(ctx->d_peer_err, 0, sizeof(unsigned)));
  }
  void* cur[mosaic_ctx::PW_COUNT] = {ctx->d_flags,    ctx->d_mat_lab, ctx->d_cand_idx,
                                     ctx->d_cand_d2,  ctx->d_result,  ctx->d_materials_owned};
  // Every rank takes the same decision here: buffers move only when a grow-only allocation grows, and every rank
  // allocates the same sizes (whole-problem sizes, also for sliced inputs) in the same order of API calls.
  bool same = ctx->peer_ok;
  for (int w = 0; w < W; ++w) same = same && cur[w] == ctx->pw_local[w];
  if (same) return MOSAIC_OK;
  MOSAIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dist_peer_close(ctx);
  std::vector<WindowInfo> mine(W), all((size_t)W * world);
  int ok = 1;
  cuMemGetAddressRange_t range = address_range_fn();
  for (int w = 0; w < W; ++w) {
    std::memset(&mine[w], 0, sizeof(WindowInfo));
    if (!cur[w]) continue;
    unsigned long long base = (unsigned long long)(uintptr_t)cur[w];
    size_t size = 0;
    if (!range || range(&base, &size, (unsigned long long)(uintptr_t)cur[w]) != 0) base = (unsigned long long)(uintptr_t)cur[w];
    if (cudaIpcGetMemHandle(&mine[w].handle, (void*)(uintptr_t)base) != cudaSuccess) {
      ctx->peer_note = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
      ok = 0;
      continue;
    }
    mine[w].offset = (unsigned long long)(uintptr_t)cur[w] - base;
    mine[w].valid = 1;
  }
  // hand the handles around (device staging: NCCL moves device memory)
  WindowInfo* d_stage = nullptr;
  MOSAIC_CUDA(ctx, cudaMalloc((void**)&d_stage, sizeof(WindowInfo) * (size_t)W * (world + 1)));
  MOSAIC_CUDA(ctx, cudaMemcpyAsync(d_stage, mine.data(), sizeof(WindowInfo) * W, cudaMemcpyHostToDevice, ctx->stream));
  {
    ncclResult_t r = api->AllGather(d_stage, d_stage + W, sizeof(WindowInfo) * W, ncclInt8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) {
      cudaFree(d_stage);
      return nccl_fail(ctx, api, r, "ncclAllGather (peer windows)");
    }
  }
  MOSAIC_CUDA(ctx, cudaMemcpyAsync(all.data(), d_stage + W, sizeof(WindowInfo) * (size_t)W * world, cudaMemcpyDeviceToHost, ctx->stream));
  MOSAIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int w = 0; w < W && ok; ++w)
    for (int r = 0; r < world; ++r) {
      const WindowInfo& wi = all[(size_t)r * W + w];
      if (!cur[w]) continue;
      if (r == ctx->rank) {
        ctx->pw_peer[w][r] = cur[w];
        continue;
      }
      if (!wi.valid) {
        ok = 0;
        break;
      }
      void* p = nullptr;
      for (int w2 = 0; w2 < w && !p; ++w2)  // the same allocation (handle) is mapped once per process
        if (ctx->pw_open[w2][r] && std::memcmp(&all[(size_t)r * W + w2].handle, &wi.handle, sizeof wi.handle) == 0)
          p = ctx->pw_open_base[w2][r];
      if (!p && cudaIpcOpenMemHandle(&p, wi.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ctx->peer_note = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(cudaGetLastError());
        ok = 0;
        break;
      }
      ctx->pw_open[w][r] = true;
      ctx->pw_open_base[w][r] = p;
      ctx->pw_peer[w][r] = (char*)p + wi.offset;
    }
  // all ranks use the windows, or none does
  int* d_ok = reinterpret_cast<int*>(d_stage);
  MOSAIC_CUDA(ctx, cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  {
    ncclResult_t r = api->AllReduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) {
      cudaFree(d_stage);
      return nccl_fail(ctx, api, r, "ncclAllReduce (peer windows)");
    }
  }
  int all_ok = 0;
  MOSAIC_CUDA(ctx, cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MOSAIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d_stage);
  if (!all_ok) {
    dist_peer_close(ctx);
    ctx->peer_failed = true;  // the NCCL exchange path stays in use
    return MOSAIC_OK;
  }
  for (int w = 0; w < W; ++w) ctx->pw_local[w] = cur[w];
  ctx->peer_ok = true;
  return MOSAIC_OK;
}

int dist_peer_push(mosaic_ctx* ctx, int window, size_t offset, size_t bytes, int only_rank, cudaStream_t s) {
  if (bytes == 0) return MOSAIC_OK;
  PushDst d;
  int nd = 0;
  for (int r = 0; r < ctx->world; ++r)
    if (r != ctx->rank && (only_rank < 0 || r == only_rank)) d.p[nd++] = (char*)ctx->pw_peer[window][r] + offset;
  if (!nd) return MOSAIC_OK;
  const size_t words = bytes / 4;  // every exchanged array is made of 4-byte elements
  int blocks = (int)std::min<size_t>((words + 255) / 256, (size_t)ctx->sm_count * 2);
  peer_push_kernel_v<<<blocks, 256, 0, s>>>(d, nd, reinterpret_cast<const uint32_t*>((const char*)ctx->pw_local[window] + offset), words);
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

int dist_peer_signal(mosaic_ctx* ctx, int slot, unsigned value, int only_rank, cudaStream_t s) {
  FlagDst d;
  int nd = 0;
  for (int r = 0; r < ctx->world; ++r)
    if (only_rank < 0 || r == only_rank) d.p[nd++] = reinterpret_cast<unsigned*>(ctx->pw_peer[mosaic_ctx::PW_FLAGS][r]);
  peer_signal_kernel<<<1, 32, 0, s>>>(d, nd, slot, value);
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

int dist_peer_wait(mosaic_ctx* ctx, int first_slot, int count, unsigned value, cudaStream_t s) {
  peer_wait_kernel<<<1, 32, 0, s>>>(ctx->d_flags, first_slot, count, value, ctx->d_peer_err);
  ctx->launches++;
  MOSAIC_CUDA(ctx, cudaGetLastError());
  return MOSAIC_OK;
}

// every rank has passed this point of its stream (and everything it read from peer windows before it is done)
int dist_peer_barrier(mosaic_ctx* ctx, cudaStream_t s) {
  const unsigned e = ++ctx->ep_bar;
  MOSAIC_TRY(dist_peer_signal(ctx, FLAG_BAR + ctx->rank, e, -1, s));
  return dist_peer_wait(ctx, FLAG_BAR, ctx->world, e, s);
}

int dist_peer_check(mosaic_ctx* ctx) {
  if (!ctx->peer_ok || !ctx->d_peer_err) return MOSAIC_OK;
  unsigned e = 0;
  MOSAIC_CUDA(ctx, cudaMemcpy(&e, ctx->d_peer_err, sizeof e, cudaMemcpyDeviceToHost));
  if (e) {
    cudaMemset(ctx->d_peer_err, 0, sizeof e);
    return fail(ctx, MOSAIC_ERR_NCCL, "a peer hand-off timed out (another rank died or the call sequences of the ranks differ)");
  }
  return MOSAIC_OK;
}

}  // namespace mosaic
