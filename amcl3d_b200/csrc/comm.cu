// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The measurement update shards by particle with the grid replicated, so the only data that ever crosses
// GPUs is (a) ten fp64 partial sums per update and (b) the particle set during a global resample.
// NCCL is bound at run time with dlopen so that single-GPU users have no NCCL dependency; if a copy is
// already loaded in the process (e.g. torch's bundled libnccl.so.2) that one is reused.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace
{
// Minimal slice of the public NCCL API (nccl.h), declared locally so the build needs no NCCL headers.
typedef struct ncclComm* ncclComm_t;
typedef struct
{
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
enum
{
  kNcclUint8 = 1,
  kNcclFloat64 = 8
};
enum
{
  kNcclSum = 0
};

struct Nccl
{
  void* handle{ nullptr };
  ncclResult_t (*GetUniqueId)(ncclUniqueId*){ nullptr };
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int){ nullptr };
  ncclResult_t (*CommDestroy)(ncclComm_t){ nullptr };
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t){ nullptr };
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t){ nullptr };
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t){ nullptr };
  const char* (*GetErrorString)(ncclResult_t){ nullptr };
  bool ok{ false };
};

Nccl& nccl()
{
  static Nccl n;
  static bool tried = false;
  if (tried)
    return n;
  tried = true;
  const char* names[] = { "libnccl.so.2", "libnccl.so" };
  for (const char* nm : names)
  {
    n.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process?
    if (n.handle)
      break;
  }
  if (!n.handle)
    for (const char* nm : names)
    {
      n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.handle)
        break;
    }
  if (!n.handle)
    return n;
  n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(n.handle, "ncclGetUniqueId"));
  n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(n.handle, "ncclCommInitRank"));
  n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.handle, "ncclCommDestroy"));
  n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.handle, "ncclAllReduce"));
  n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(n.handle, "ncclAllGather"));
  n.Broadcast = reinterpret_cast<decltype(n.Broadcast)>(dlsym(n.handle, "ncclBroadcast"));
  n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.handle, "ncclGetErrorString"));
  n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.AllGather && n.Broadcast;
  return n;
}

int nccl_fail(const char* what, ncclResult_t r)
{
  Nccl& n = nccl();
  std::string msg = std::string(what) + ": NCCL error " + std::to_string(r);
  if (n.GetErrorString)
    msg += std::string(" (") + n.GetErrorString(r) + ")";
  return amcl3d_b200::fail(AMCL3D_CUDA_ERR_NCCL, msg);
}
}  // namespace

namespace amcl3d_b200
{
int comm_all_reduce_f64(amcl3d_cuda_ctx* ctx, double* d_buf, size_t count)
{
  if (ctx->n_ranks <= 1)
    return 0;
  Nccl& n = nccl();
  if (!n.ok || !ctx->nccl_comm)
    return fail(AMCL3D_CUDA_ERR_NCCL, "all_reduce: no communicator");
  ncclResult_t r =
      n.AllReduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream);
  if (r != 0)
    return nccl_fail("ncclAllReduce", r);
  return 0;
}

int comm_all_gather(amcl3d_cuda_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank)
{
  if (ctx->n_ranks <= 1)
  {
    cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream);
    return 0;
  }
  Nccl& n = nccl();
  if (!n.ok || !ctx->nccl_comm)
    return fail(AMCL3D_CUDA_ERR_NCCL, "all_gather: no communicator");
  ncclResult_t r =
      n.AllGather(d_send, d_recv, bytes_per_rank, kNcclUint8, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream);
  if (r != 0)
    return nccl_fail("ncclAllGather", r);
  return 0;
}

int comm_broadcast(amcl3d_cuda_ctx* ctx, void* d_buf, size_t bytes, int root)
{
  if (ctx->n_ranks <= 1)
    return 0;
  Nccl& n = nccl();
  if (!n.ok || !ctx->nccl_comm)
    return fail(AMCL3D_CUDA_ERR_NCCL, "broadcast: no communicator");
  ncclResult_t r = n.Broadcast(d_buf, d_buf, bytes, kNcclUint8, root, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream);
  if (r != 0)
    return nccl_fail("ncclBroadcast", r);
  return 0;
}
}  // namespace amcl3d_b200

using namespace amcl3d_b200;

extern "C" {

int amcl3d_cuda_comm_unique_id(uint8_t id_out[128])
{
  if (!id_out)
    return fail(AMCL3D_CUDA_ERR_INVALID, "comm_unique_id: NULL argument");
  Nccl& n = nccl();
  if (!n.ok)
    return fail(AMCL3D_CUDA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = n.GetUniqueId(&id);
  if (r != 0)
    return nccl_fail("ncclGetUniqueId", r);
  std::memcpy(id_out, id.internal, 128);
  return 0;
}

int amcl3d_cuda_comm_init(amcl3d_cuda_ctx* ctx, const uint8_t id[128], int rank, int n_ranks)
{
  if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return fail(AMCL3D_CUDA_ERR_INVALID, "comm_init: bad argument");
  if (ctx->nccl_comm)
    return fail(AMCL3D_CUDA_ERR_INVALID, "comm_init: communicator already attached");
  Nccl& n = nccl();
  if (!n.ok)
    return fail(AMCL3D_CUDA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  std::memcpy(uid.internal, id, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = n.CommInitRank(&comm, n_ranks, uid, rank);
  if (r != 0)
    return nccl_fail("ncclCommInitRank", r);
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  return 0;
}

int amcl3d_cuda_comm_destroy(amcl3d_cuda_ctx* ctx)
{
  if (!ctx || !ctx->nccl_comm)
    return 0;
  Nccl& n = nccl();
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (n.ok)
    n.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  ctx->rank = 0;
  ctx->n_ranks = 1;
  return 0;
}

int amcl3d_cuda_comm_rank(const amcl3d_cuda_ctx* ctx, int* rank, int* n_ranks)
{
  if (!ctx || !rank || !n_ranks)
    return fail(AMCL3D_CUDA_ERR_INVALID, "comm_rank: NULL argument");
  *rank = ctx->rank;
  *n_ranks = ctx->n_ranks;
  return 0;
}

}  // extern "C"
