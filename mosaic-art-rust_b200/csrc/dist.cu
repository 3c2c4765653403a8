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
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather && api.GroupStart &&
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
  if (ctx->nccl && ctx->nccl_comm) ctx->nccl->CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

int dist_allgather(mosaic_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (ctx->world == 1) return MOSAIC_OK;
  NcclApi* api = ctx->nccl;
  MOSAIC_NCCL(ctx, api, api->AllGather(send, recv, bytes_per_rank, ncclInt8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return MOSAIC_OK;
}

int dist_allgatherv_inplace(mosaic_ctx* ctx, void* buf, const size_t* offsets, const size_t* sizes) {
  if (ctx->world == 1) return MOSAIC_OK;
  NcclApi* api = ctx->nccl;
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
  MOSAIC_NCCL(ctx, api, api->Broadcast(buf, buf, bytes, ncclInt8, root, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return MOSAIC_OK;
}

}  // namespace mosaic
