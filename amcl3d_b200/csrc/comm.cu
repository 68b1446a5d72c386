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
  kNcclUint32 = 3,
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

// In-place sum of uint32 words.  Used for arrays in which every element is non-zero on at most ONE rank (a float's bit
// pattern + 0 = the same bit pattern): an exact exchange, whatever the element type.
int comm_all_reduce_u32(amcl3d_cuda_ctx* ctx, uint32_t* d_buf, size_t count)
{
  if (ctx->n_ranks <= 1)
    return 0;
  Nccl& n = nccl();
  if (!n.ok || !ctx->nccl_comm)
    return fail(AMCL3D_CUDA_ERR_NCCL, "all_reduce: no communicator");
  ncclResult_t r =
      n.AllReduce(d_buf, d_buf, count, kNcclUint32, kNcclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream);
  if (r != 0)
    return nccl_fail("ncclAllReduce (u32)", r);
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

// Maps every rank's PeerBox into this process (CUDA IPC; the handles travel through one ncclAllGather).  Any failure
// -- more than kMaxPeers ranks, IPC not permitted in this container, no peer access -- leaves peer_ok false and the
// update falls back to ncclAllReduce for the ten partial sums (still a GPU path; nothing runs on the CPU).
static void peer_teardown(amcl3d_cuda_ctx* ctx)
{
  for (int r = 0; r < kMaxPeers; ++r)
  {
    if (ctx->peer_box[r])
    {
      if (r == ctx->rank)
        cudaFree(ctx->peer_box[r]);
      else
        cudaIpcCloseMemHandle(ctx->peer_box[r]);
    }
    ctx->peer_box[r] = nullptr;
  }
  ctx->peer_ok = false;
  cudaGetLastError();
}

static void peer_setup(amcl3d_cuda_ctx* ctx)
{
  ctx->peer_ok = false;
  ctx->peer_seq = 0;
  Nccl& n = nccl();
  const int world = ctx->n_ranks;
  // every rank must reach the all-gather below (it is collective) and agree on the outcome (a second all-gather of
  // "I mapped everything" votes), so failures are recorded, not returned early
  bool mine_ok = world >= 2 && world <= kMaxPeers;
  void* local = nullptr;
  cudaIpcMemHandle_t handle;
  std::memset(&handle, 0, sizeof(handle));
  if (mine_ok && cudaMalloc(&local, sizeof(PeerBox)) != cudaSuccess)
    mine_ok = false;
  if (mine_ok && cudaMemset(local, 0, sizeof(PeerBox)) != cudaSuccess)
    mine_ok = false;
  if (mine_ok && cudaIpcGetMemHandle(&handle, local) != cudaSuccess)
    mine_ok = false;
  cudaGetLastError();
  if (world > kMaxPeers)
    return;  // known on every rank: no collective needed
  struct Msg
  {
    cudaIpcMemHandle_t h;
    int ok;
    int pad[15];
  };
  static_assert(sizeof(Msg) == 128, "IPC message size");
  Msg* d_msgs = nullptr;
  std::vector<Msg> msgs(static_cast<size_t>(world));
  bool coll_ok = cudaMalloc(&d_msgs, sizeof(Msg) * world) == cudaSuccess;
  auto all_gather_msgs = [&](const Msg& m) -> bool {
    if (!coll_ok)
      return false;
    if (cudaMemcpyAsync(d_msgs + ctx->rank, &m, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      return false;
    if (n.AllGather(d_msgs + ctx->rank, d_msgs, sizeof(Msg), kNcclUint8, static_cast<ncclComm_t>(ctx->nccl_comm),
                    ctx->stream) != 0)
      return false;
    if (cudaMemcpyAsync(msgs.data(), d_msgs, sizeof(Msg) * world, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      return false;
    return cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  };
  Msg m;
  std::memset(&m, 0, sizeof(m));
  m.h = handle;
  m.ok = mine_ok ? 1 : 0;
  bool all_ok = all_gather_msgs(m);
  for (int r = 0; all_ok && r < world; ++r)
    all_ok = msgs[r].ok == 1;
  if (all_ok)
  {
    ctx->peer_box[ctx->rank] = local;
    local = nullptr;
    for (int r = 0; r < world && all_ok; ++r)
    {
      if (r == ctx->rank)
        continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, msgs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        all_ok = false;
      else
        ctx->peer_box[r] = p;
    }
  }
  cudaGetLastError();
  // second round: did EVERY rank map every box?
  std::memset(&m, 0, sizeof(m));
  m.ok = all_ok ? 1 : 0;
  bool agreed = all_gather_msgs(m);
  for (int r = 0; agreed && r < world; ++r)
    agreed = msgs[r].ok == 1;
  if (d_msgs)
    cudaFree(d_msgs);
  if (local)
    cudaFree(local);
  if (agreed)
    ctx->peer_ok = true;
  else
    peer_teardown(ctx);
  cudaGetLastError();
}

int comm_peer_view(amcl3d_cuda_ctx* ctx, PeerView* pv)
{
  std::memset(pv, 0, sizeof(*pv));
  pv->n_ranks = 1;
  if (ctx->n_ranks <= 1 || !ctx->peer_ok || ctx->opt_peer_reduce == 1)
    return 0;
  for (int r = 0; r < ctx->n_ranks; ++r)
    pv->box[r] = static_cast<PeerBox*>(ctx->peer_box[r]);
  pv->n_ranks = ctx->n_ranks;
  pv->rank = ctx->rank;
  pv->seq = ++ctx->peer_seq;
  pv->timeout_clocks = ctx->opt_peer_timeout_ms * ctx->clock_khz;
  return 1;
}

void comm_release_shards(amcl3d_cuda_ctx* ctx, ShardView* sv)
{
  if (!sv)
    return;
  for (int r = 0; r < kMaxPeers; ++r)
  {
    if (sv->block[r] && sv->n_ranks > 1 && r != sv->rank)
      cudaIpcCloseMemHandle(sv->block[r]);
    sv->block[r] = nullptr;
  }
  sv->n_ranks = 0;
  cudaGetLastError();
  (void)ctx;
}

// Collective over the communicator.  Every rank contributes (particle count, capacity, IPC handle of its particle
// block); afterwards every rank knows all counts / first global indices and has every block mapped.  Unequal counts
// (including empty shards) are fine: all offsets come from the exchanged table.
int comm_exchange_shards(amcl3d_cuda_ctx* ctx, float* local_block, uint64_t cap, uint64_t n, ShardView* out)
{
  comm_release_shards(ctx, out);
  std::memset(out, 0, sizeof(*out));
  out->rank = ctx->rank;
  out->n_ranks = ctx->n_ranks;
  if (ctx->n_ranks <= 1)
  {
    out->n_ranks = 1;
    out->rank = 0;
    out->n[0] = n;
    out->cap[0] = cap;
    out->block[0] = local_block;
    out->n_total = n;
    return 0;
  }
  Nccl& nc = nccl();
  if (!nc.ok || !ctx->nccl_comm)
    return fail(AMCL3D_CUDA_ERR_NCCL, "exchange_shards: no communicator");
  if (!ctx->peer_ok)
    return fail(AMCL3D_CUDA_ERR_NCCL,
                "exchange_shards: peer memory (CUDA IPC over NVLink) is not available between the ranks; the sharded "
                "particle filter needs it");
  const int world = ctx->n_ranks;
  struct Msg
  {
    cudaIpcMemHandle_t h;
    uint64_t n, cap;
    int ok;
    int pad[11];
  };
  static_assert(sizeof(Msg) == 128, "shard message size");
  Msg mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.n = n;
  mine.cap = cap;
  mine.ok = 1;
  if (local_block && cudaIpcGetMemHandle(&mine.h, local_block) != cudaSuccess)
    mine.ok = 0;
  if (!local_block)
    mine.ok = cap == 0 ? 1 : 0;
  cudaGetLastError();
  Msg* d_msgs = nullptr;
  std::vector<Msg> msgs(static_cast<size_t>(world));
  A3D_CUDA_TRY(cudaMalloc(&d_msgs, sizeof(Msg) * world));
  auto all_gather = [&](const Msg& m) -> int {
    A3D_CUDA_TRY(cudaMemcpyAsync(d_msgs + ctx->rank, &m, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = nc.AllGather(d_msgs + ctx->rank, d_msgs, sizeof(Msg), kNcclUint8,
                                  static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream);
    if (r != 0)
      return nccl_fail("ncclAllGather (shard table)", r);
    A3D_CUDA_TRY(cudaMemcpyAsync(msgs.data(), d_msgs, sizeof(Msg) * world, cudaMemcpyDeviceToHost, ctx->stream));
    A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
  };
  int rc = all_gather(mine);
  bool all_ok = rc == 0;
  for (int r = 0; all_ok && r < world; ++r)
    all_ok = msgs[r].ok == 1;
  uint64_t first = 0;
  if (all_ok)
  {
    for (int r = 0; r < world; ++r)
    {
      out->n[r] = msgs[r].n;
      out->cap[r] = msgs[r].cap;
      out->first[r] = first;
      first += msgs[r].n;
      if (r == ctx->rank)
        out->block[r] = local_block;
      else if (msgs[r].cap > 0)
      {
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, msgs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
          all_ok = false;
        else
          out->block[r] = static_cast<float*>(p);
      }
    }
    out->n_total = first;
  }
  cudaGetLastError();
  // second round: did EVERY rank map everything?  (all ranks must agree before anybody dereferences a peer pointer)
  Msg vote;
  std::memset(&vote, 0, sizeof(vote));
  vote.ok = all_ok ? 1 : 0;
  if (rc == 0)
    rc = all_gather(vote);
  bool agreed = rc == 0;
  for (int r = 0; agreed && r < world; ++r)
    agreed = msgs[r].ok == 1;
  cudaFree(d_msgs);
  if (!agreed)
  {
    comm_release_shards(ctx, out);
    if (rc != 0)
      return rc;
    return fail(AMCL3D_CUDA_ERR_NCCL, "exchange_shards: a rank could not map a peer's particle block (CUDA IPC)");
  }
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
  peer_setup(ctx);
  return 0;
}

int amcl3d_cuda_comm_destroy(amcl3d_cuda_ctx* ctx)
{
  if (!ctx || !ctx->nccl_comm)
    return 0;
  Nccl& n = nccl();
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  peer_teardown(ctx);
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

int amcl3d_cuda_comm_peer_active(const amcl3d_cuda_ctx* ctx, int* active)
{
  if (!ctx || !active)
    return fail(AMCL3D_CUDA_ERR_INVALID, "comm_peer_active: NULL argument");
  *active = (ctx->n_ranks > 1 && ctx->peer_ok && ctx->opt_peer_reduce != 1) ? 1 : 0;
  return 0;
}

}  // extern "C"
