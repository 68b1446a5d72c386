// api.cu -- context, options and grid storage behind include/amcl3d_cuda.h.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace amcl3d_b200
{
static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg)
{
  g_last_error = msg;
  return code;
}

int ensure_pinned(amcl3d_cuda_ctx* ctx, size_t bytes)
{
  if (ctx->pinned_bytes >= bytes)
    return 0;
  if (ctx->pinned)
    cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_bytes = 0;
  size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
  A3D_CUDA_TRY(cudaMallocHost(&ctx->pinned, want));
  ctx->pinned_bytes = want;
  return 0;
}

int ensure_scratch(amcl3d_cuda_ctx* ctx, size_t bytes)
{
  if (ctx->scratch_bytes >= bytes)
    return 0;
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->scratch)
    cudaFree(ctx->scratch);
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  size_t want = bytes < (4u << 20) ? (4u << 20) : (bytes + bytes / 4);
  want = (want + 4095) / 4096 * 4096;
  A3D_CUDA_TRY(cudaMalloc(&ctx->scratch, want));
  ctx->scratch_bytes = want;
  return 0;
}

// AoS (dist, prob) cells <-> SoA planes
// `cells` holds the n cells whose LOGICAL (reference) linear indices start at `first`; the planes are addressed
// through the grid's physical layout.
__global__ void split_cells_kernel(const GridView g, const float2* __restrict__ cells, float* __restrict__ dist,
                                   float* __restrict__ prob, uint64_t first, uint64_t n)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float2 c = cells[i];
    const uint32_t p = logical_to_phys(g, static_cast<uint32_t>(first + i));
    if (dist)
      dist[p] = c.x;
    prob[p] = c.y;
  }
}

__global__ void merge_cells_kernel(const GridView g, float2* __restrict__ cells, const float* __restrict__ dist,
                                   const float* __restrict__ prob, uint64_t first, uint64_t n)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const uint32_t p = logical_to_phys(g, static_cast<uint32_t>(first + i));
    cells[i] = make_float2(dist ? dist[p] : -1.f, prob[p]);
  }
}

// n probabilities whose logical indices start at `first`, gathered into a linear buffer
__global__ void gather_prob_kernel(const GridView g, float* __restrict__ out, uint64_t first, uint64_t n)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = g.prob[logical_to_phys(g, static_cast<uint32_t>(first + i))];
}
// probabilities at arbitrary logical indices (0xFFFFFFFF / out of range -> 0)
__global__ void gather_prob_idx_kernel(const GridView g, const uint32_t* __restrict__ idx, float* __restrict__ out, uint64_t n)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const uint32_t gi = idx[i];
    out[i] = (static_cast<uint64_t>(gi) < g.n_cells) ? g.prob[logical_to_phys(g, gi)] : 0.f;
  }
}
}  // namespace amcl3d_b200

using namespace amcl3d_b200;

amcl3d_b200::GridView amcl3d_cuda_grid::view() const
{
  GridView v;
  v.prob = d_prob;
  v.size_x = dims[0];
  v.size_y = dims[1];
  v.size_z = dims[2];
  v.step_y = dims[0];
  v.step_z = dims[0] * dims[1];
  v.n_cells = n_cells;
  v.zero_index = static_cast<uint32_t>(n_phys);
  v.min_x = bounds[0];
  v.min_y = bounds[1];
  v.min_z = bounds[2];
  v.max_x = bounds[3];
  v.max_y = bounds[4];
  v.max_z = bounds[5];
  v.ext_x = bounds[3] - bounds[0];
  v.ext_y = bounds[4] - bounds[1];
  v.ext_z = bounds[5] - bounds[2];
  v.res = bounds[6];
  v.inv_res_f = static_cast<float>(1.0 / bounds[6]);
  v.inv_res_lo = static_cast<float>(1.0 / bounds[6] - static_cast<double>(v.inv_res_f));
  auto up = [](double e) {
    float f = static_cast<float>(e);
    if (static_cast<double>(f) < e)
      f = std::nextafterf(f, INFINITY);
    return f;
  };
  v.ext_up_x = up(v.ext_x);
  v.ext_up_y = up(v.ext_y);
  v.ext_up_z = up(v.ext_z);
  v.brick_shift = brick_shift;
  v.nbx = nb[0];
  v.nby = nb[1];
  return v;
}

extern "C" {

int amcl3d_cuda_abi_version(void) { return AMCL3D_CUDA_ABI_VERSION; }
const char* amcl3d_cuda_last_error(void) { return g_last_error.c_str(); }

int amcl3d_cuda_ctx_create(int device, void* stream, amcl3d_cuda_ctx** out)
{
  if (!out)
    return fail(AMCL3D_CUDA_ERR_INVALID, "ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(AMCL3D_CUDA_ERR_NO_DEVICE,
                std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return fail(AMCL3D_CUDA_ERR_INVALID, "ctx_create: device index out of range");
  A3D_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  A3D_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  amcl3d_cuda_ctx* c = new amcl3d_cuda_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->l2_bytes = prop.l2CacheSize;
  c->l2_persist_max = prop.persistingL2CacheMaxSize;
  c->cc = prop.major * 10 + prop.minor;
  {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    c->clock_khz = khz > 0 ? khz : 2000000;
  }
  if (stream)
  {
    c->stream = static_cast<cudaStream_t>(stream);
    c->own_stream = false;
  }
  else
  {
    cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess)
    {
      delete c;
      return fail(AMCL3D_CUDA_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(se));
    }
    c->own_stream = true;
  }
  cudaEventCreate(&c->ev_k0);
  cudaEventCreate(&c->ev_k1);
  cudaEventCreate(&c->ev_k2);
  cudaEventCreate(&c->ev_k3);
  *out = c;
  return 0;
}

int amcl3d_cuda_ctx_destroy(amcl3d_cuda_ctx* ctx)
{
  if (!ctx)
    return 0;
  cudaSetDevice(ctx->device);
  amcl3d_cuda_comm_destroy(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream)
    cudaStreamDestroy(ctx->stream);
  if (ctx->ev_k0)
    cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1)
    cudaEventDestroy(ctx->ev_k1);
  if (ctx->ev_k2)
    cudaEventDestroy(ctx->ev_k2);
  if (ctx->ev_k3)
    cudaEventDestroy(ctx->ev_k3);
  if (ctx->scratch)
    cudaFree(ctx->scratch);
  if (ctx->pinned)
    cudaFreeHost(ctx->pinned);
  delete ctx;
  return 0;
}

int amcl3d_cuda_ctx_set_stream(amcl3d_cuda_ctx* ctx, void* stream)
{
  if (!ctx)
    return fail(AMCL3D_CUDA_ERR_INVALID, "set_stream: ctx is NULL");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream)
    cudaStreamDestroy(ctx->stream);
  if (stream)
  {
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->own_stream = false;
  }
  else
  {
    A3D_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return 0;
}

int amcl3d_cuda_ctx_synchronize(amcl3d_cuda_ctx* ctx)
{
  if (!ctx)
    return fail(AMCL3D_CUDA_ERR_INVALID, "synchronize: ctx is NULL");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int amcl3d_cuda_ctx_device_info(amcl3d_cuda_ctx* ctx, int64_t info[4])
{
  if (!ctx || !info)
    return fail(AMCL3D_CUDA_ERR_INVALID, "device_info: NULL argument");
  info[0] = ctx->sm_count;
  info[1] = ctx->l2_bytes;
  info[2] = ctx->l2_persist_max;
  info[3] = ctx->cc;
  return 0;
}

static int64_t* option_slot(amcl3d_cuda_ctx* ctx, const char* name)
{
  if (!std::strcmp(name, "weight_point_splits"))
    return &ctx->opt_point_splits;
  if (!std::strcmp(name, "sum_mode"))
    return &ctx->opt_sum_mode;
  if (!std::strcmp(name, "resample_mode"))
    return &ctx->opt_resample_mode;
  if (!std::strcmp(name, "kernel_timing"))
    return &ctx->opt_kernel_timing;
  if (!std::strcmp(name, "l2_persist"))
    return &ctx->opt_l2_persist;
  if (!std::strcmp(name, "max_cells"))
    return &ctx->opt_max_cells;
  if (!std::strcmp(name, "weight_block_threads"))
    return &ctx->opt_block_threads;
  if (!std::strcmp(name, "weight_variant"))
    return &ctx->opt_weight_variant;
  if (!std::strcmp(name, "l2_fetch_granularity"))
    return &ctx->opt_l2_fetch;
  if (!std::strcmp(name, "weight_chunk_points"))
    return &ctx->opt_chunk_points;
  if (!std::strcmp(name, "grid_layout"))
    return &ctx->opt_grid_layout;
  if (!std::strcmp(name, "cloud_order"))
    return &ctx->opt_cloud_order;
  if (!std::strcmp(name, "serial_chain"))
    return &ctx->opt_serial_chain;
  if (!std::strcmp(name, "particle_order"))
    return &ctx->opt_particle_order;
  if (!std::strcmp(name, "peer_reduce"))
    return &ctx->opt_peer_reduce;
  if (!std::strcmp(name, "reference_order"))
    return &ctx->opt_reference_order;
  if (!std::strcmp(name, "replay"))
    return &ctx->opt_replay;
  if (!std::strcmp(name, "replay_max_mb"))
    return &ctx->opt_replay_max_mb;
  if (!std::strcmp(name, "ordered_mode"))
    return &ctx->opt_ordered;
  if (!std::strcmp(name, "global_schedule"))
    return &ctx->opt_global_schedule;
  if (!std::strcmp(name, "global_schedule_chunk"))
    return &ctx->opt_deal_chunk;
  if (!std::strcmp(name, "order_clip_sigma_x10"))
    return &ctx->opt_order_clip;
  if (!std::strcmp(name, "order_key_bits"))
    return &ctx->opt_order_bits;
  if (!std::strcmp(name, "order_weight_x"))
    return &ctx->opt_order_w[0];
  if (!std::strcmp(name, "order_weight_y"))
    return &ctx->opt_order_w[1];
  if (!std::strcmp(name, "order_weight_z"))
    return &ctx->opt_order_w[2];
  if (!std::strcmp(name, "order_weight_yaw"))
    return &ctx->opt_order_w[3];
  if (!std::strcmp(name, "peer_timeout_ms"))
    return &ctx->opt_peer_timeout_ms;
  return nullptr;
}

int amcl3d_cuda_ctx_set_option(amcl3d_cuda_ctx* ctx, const char* name, int64_t value)
{
  if (!ctx || !name)
    return fail(AMCL3D_CUDA_ERR_INVALID, "set_option: NULL argument");
  int64_t* s = option_slot(ctx, name);
  if (!s)
    return fail(AMCL3D_CUDA_ERR_INVALID, std::string("set_option: unknown option ") + name);
  *s = value;
  if (s == &ctx->opt_l2_fetch && (value == 32 || value == 64 || value == 128))
  {
    // device-wide: how many bytes L2 pulls from HBM per miss.  Random 4-byte gathers want 32 (one sector), the
    // default of 64 doubles the DRAM traffic of every miss.
    A3D_CUDA_TRY(cudaSetDevice(ctx->device));
    A3D_CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, static_cast<size_t>(value)));
  }
  return 0;
}

int amcl3d_cuda_ctx_get_option(amcl3d_cuda_ctx* ctx, const char* name, int64_t* value)
{
  if (!ctx || !name || !value)
    return fail(AMCL3D_CUDA_ERR_INVALID, "get_option: NULL argument");
  int64_t* s = option_slot(ctx, name);
  if (!s)
    return fail(AMCL3D_CUDA_ERR_INVALID, std::string("get_option: unknown option ") + name);
  *value = *s;
  return 0;
}

int amcl3d_cuda_ctx_last_kernel_ms(amcl3d_cuda_ctx* ctx, float* ms)
{
  if (!ctx || !ms)
    return fail(AMCL3D_CUDA_ERR_INVALID, "last_kernel_ms: NULL argument");
  if (!ctx->ev_valid)
    return fail(AMCL3D_CUDA_ERR_INVALID, "last_kernel_ms: no timed kernel yet (set option kernel_timing = 1)");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  A3D_CUDA_TRY(cudaEventSynchronize(ctx->ev_k1));
  A3D_CUDA_TRY(cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
  return 0;
}

int amcl3d_cuda_ctx_last_update_phases_ms(amcl3d_cuda_ctx* ctx, float ms3[3])
{
  if (!ctx || !ms3)
    return fail(AMCL3D_CUDA_ERR_INVALID, "last_update_phases_ms: NULL argument");
  if (!ctx->ev_valid || !ctx->ev_phases_valid)
    return fail(AMCL3D_CUDA_ERR_INVALID, "last_update_phases_ms: no timed update yet (set option kernel_timing = 1)");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  A3D_CUDA_TRY(cudaEventSynchronize(ctx->ev_k3));
  A3D_CUDA_TRY(cudaEventElapsedTime(&ms3[0], ctx->ev_k0, ctx->ev_k1));
  A3D_CUDA_TRY(cudaEventElapsedTime(&ms3[1], ctx->ev_k1, ctx->ev_k2));
  A3D_CUDA_TRY(cudaEventElapsedTime(&ms3[2], ctx->ev_k2, ctx->ev_k3));
  return 0;
}

int amcl3d_cuda_ctx_launch_count(amcl3d_cuda_ctx* ctx, uint64_t* count)
{
  if (!ctx || !count)
    return fail(AMCL3D_CUDA_ERR_INVALID, "launch_count: NULL argument");
  *count = ctx->launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------ grid

int amcl3d_cuda_grid_create(amcl3d_cuda_ctx* ctx, const double bounds7[7], amcl3d_cuda_grid** out)
{
  if (!ctx || !bounds7 || !out)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_create: NULL argument");
  *out = nullptr;
  if (!(bounds7[6] > 0.0))
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_create: resolution must be positive");
  uint32_t dims[3];
  for (int a = 0; a < 3; ++a)
  {
    // PointCloudTools.cpp:93-98 -- (uint32) ceil((max - min) / resolution) evaluated in double
    const double extent = bounds7[3 + a] - bounds7[a];
    const double cells = std::ceil(extent / bounds7[6]);
    if (!(cells >= 1.0) || cells > 2097152.0)
      return fail(AMCL3D_CUDA_ERR_INVALID, "grid_create: degenerate or oversized axis");
    dims[a] = static_cast<uint32_t>(cells);
  }
  const uint64_t total = static_cast<uint64_t>(dims[0]) * dims[1] * dims[2];
  if (ctx->opt_max_cells > 0 && total > static_cast<uint64_t>(ctx->opt_max_cells))
    return fail(AMCL3D_CUDA_ERR_TOO_BIG, "Octomap size is too big. Grid size over the configured cell cap.");
  if (total >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_TOO_BIG, "grid_create: linear voxel indices must fit 32 bits");
  amcl3d_cuda_grid* g = new amcl3d_cuda_grid();
  g->ctx = ctx;
  std::memcpy(g->bounds, bounds7, sizeof(g->bounds));
  std::memcpy(g->dims, dims, sizeof(dims));
  g->n_cells = total;
  // Physical layout: option "grid_layout" 1 = linear, 2 = bricked (32^3 voxels = 128 KB per brick), 0 = auto:
  // bricked once the probability plane no longer fits L2 (then page/TLB locality of the gather starts to matter).
  int layout = static_cast<int>(ctx->opt_grid_layout);
  if (layout == 0)
    layout = (total * sizeof(float) > static_cast<uint64_t>(ctx->l2_bytes > 0 ? ctx->l2_bytes : (96 << 20))) ? 2 : 1;
  g->brick_shift = 0;
  g->n_phys = total;
  if (layout == 2)
  {
    const uint32_t b = kBrickShift;
    uint64_t padded = 1;
    for (int a = 0; a < 3; ++a)
    {
      g->nb[a] = (dims[a] + (1u << b) - 1u) >> b;
      padded *= static_cast<uint64_t>(g->nb[a]) << b;
    }
    if (padded < 0xFFFFFFFFull)  // physical addresses are 32-bit too; otherwise stay linear
    {
      g->brick_shift = b;
      g->n_phys = padded;
    }
  }
  *out = g;
  return 0;
}

int amcl3d_cuda_grid_destroy(amcl3d_cuda_grid* grid)
{
  if (!grid)
    return 0;
  cudaSetDevice(grid->ctx->device);
  cudaStreamSynchronize(grid->ctx->stream);
  if (grid->d_prob)
    cudaFree(grid->d_prob);
  if (grid->d_dist)
    cudaFree(grid->d_dist);
  delete grid;
  return 0;
}

int amcl3d_cuda_grid_dims(const amcl3d_cuda_grid* grid, uint32_t dims3[3])
{
  if (!grid || !dims3)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_dims: NULL argument");
  std::memcpy(dims3, grid->dims, sizeof(grid->dims));
  return 0;
}

int amcl3d_cuda_grid_bounds(const amcl3d_cuda_grid* grid, double bounds7[7])
{
  if (!grid || !bounds7)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_bounds: NULL argument");
  std::memcpy(bounds7, grid->bounds, sizeof(grid->bounds));
  return 0;
}

int amcl3d_cuda_grid_upload_cells(amcl3d_cuda_grid* grid, const float* cells, double sensor_dev)
{
  if (!grid || !cells)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_upload_cells: NULL argument");
  amcl3d_cuda_ctx* ctx = grid->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  const uint64_t n = grid->n_cells;
  if (!grid->d_prob)
  {
    // + kZeroCellPad floats: GridView::zero_index points at a cell that always reads 0 (skipped points gather it)
    A3D_CUDA_TRY(cudaMalloc(&grid->d_prob, (grid->n_phys + kZeroCellPad) * sizeof(float)));
    A3D_CUDA_TRY(cudaMemsetAsync(grid->d_prob, 0, (grid->n_phys + kZeroCellPad) * sizeof(float), ctx->stream));
  }
  if (!grid->d_dist)
  {
    A3D_CUDA_TRY(cudaMalloc(&grid->d_dist, grid->n_phys * sizeof(float)));
    A3D_CUDA_TRY(cudaMemsetAsync(grid->d_dist, 0, grid->n_phys * sizeof(float), ctx->stream));
  }
  const GridView gv = grid->view();
  // stream the AoS cells through a bounded device staging buffer
  const uint64_t chunk = 1ull << 24;  // 16 M cells = 128 MB
  float2* d_stage = nullptr;
  A3D_CUDA_TRY(cudaMalloc(&d_stage, (n < chunk ? n : chunk) * sizeof(float2)));
  for (uint64_t off = 0; off < n; off += chunk)
  {
    const uint64_t m = (n - off < chunk) ? n - off : chunk;
    cudaError_t e = cudaMemcpyAsync(d_stage, cells + 2 * off, m * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
    {
      split_cells_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(gv, d_stage, grid->d_dist, grid->d_prob, off, m);
      ctx->launches++;
      e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess)
    {
      cudaFree(d_stage);
      return fail(AMCL3D_CUDA_ERR_CUDA, std::string("grid_upload_cells: ") + cudaGetErrorString(e));
    }
  }
  cudaFree(d_stage);
  grid->sensor_dev = sensor_dev;
  grid->has_cells = true;
  return 0;
}

int amcl3d_cuda_grid_download_cells(const amcl3d_cuda_grid* grid, float* cells)
{
  if (!grid || !cells)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_download_cells: NULL argument");
  if (!grid->has_cells)
    return fail(AMCL3D_CUDA_ERR_NOT_OPEN, "grid_download_cells: grid has no cells");
  amcl3d_cuda_ctx* ctx = grid->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  const uint64_t n = grid->n_cells;
  const uint64_t chunk = 1ull << 24;
  float2* d_stage = nullptr;
  A3D_CUDA_TRY(cudaMalloc(&d_stage, (n < chunk ? n : chunk) * sizeof(float2)));
  for (uint64_t off = 0; off < n; off += chunk)
  {
    const uint64_t m = (n - off < chunk) ? n - off : chunk;
    merge_cells_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(grid->view(), d_stage, grid->d_dist, grid->d_prob, off,
                                                                   m);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(cells + 2 * off, d_stage, m * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess)
    {
      cudaFree(d_stage);
      return fail(AMCL3D_CUDA_ERR_CUDA, std::string("grid_download_cells: ") + cudaGetErrorString(e));
    }
  }
  cudaFree(d_stage);
  return 0;
}

int amcl3d_cuda_grid_download_prob(const amcl3d_cuda_grid* grid, float* prob)
{
  if (!grid || !prob)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_download_prob: NULL argument");
  if (!grid->has_cells)
    return fail(AMCL3D_CUDA_ERR_NOT_OPEN, "grid_download_prob: grid has no cells");
  return amcl3d_cuda_grid_download_prob_range(grid, 0, grid->n_cells, prob);
}

int amcl3d_cuda_grid_download_prob_range(const amcl3d_cuda_grid* grid, uint64_t first, uint64_t count, float* prob)
{
  if (!grid || (count && !prob))
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_download_prob_range: NULL argument");
  if (!grid->has_cells)
    return fail(AMCL3D_CUDA_ERR_NOT_OPEN, "grid_download_prob_range: grid has no cells");
  A3D_CUDA_TRY(cudaSetDevice(grid->ctx->device));
  const uint64_t avail = first < grid->n_cells ? grid->n_cells - first : 0;
  const uint64_t m = count < avail ? count : avail;
  amcl3d_cuda_ctx* ctx = grid->ctx;
  if (m && grid->brick_shift == 0)
  {
    A3D_CUDA_TRY(cudaMemcpyAsync(prob, grid->d_prob + first, m * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  else if (m)
  {
    // bricked storage: gather the logical range into a linear staging buffer, chunk by chunk
    const uint64_t chunk = 1ull << 25;
    float* d_stage = nullptr;
    A3D_CUDA_TRY(cudaMalloc(&d_stage, (m < chunk ? m : chunk) * sizeof(float)));
    const GridView gv = grid->view();
    for (uint64_t off = 0; off < m; off += chunk)
    {
      const uint64_t k = (m - off < chunk) ? m - off : chunk;
      gather_prob_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(gv, d_stage, first + off, k);
      ctx->launches++;
      cudaError_t e = cudaMemcpyAsync(prob + off, d_stage, k * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess)
      {
        cudaFree(d_stage);
        return fail(AMCL3D_CUDA_ERR_CUDA, std::string("grid_download_prob_range: ") + cudaGetErrorString(e));
      }
    }
    cudaFree(d_stage);
  }
  for (uint64_t i = m; i < count; ++i)
    prob[i] = 0.f;
  return 0;
}

int amcl3d_cuda_grid_gather_prob(const amcl3d_cuda_grid* grid, const uint32_t* idx, uint64_t n, float* prob_out)
{
  if (!grid || (n && (!idx || !prob_out)))
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_gather_prob: NULL argument");
  if (!grid->has_cells)
    return fail(AMCL3D_CUDA_ERR_NOT_OPEN, "grid_gather_prob: grid has no cells");
  if (n == 0)
    return 0;
  amcl3d_cuda_ctx* ctx = grid->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t half = (n * 4 + 255) / 256 * 256;
  A3D_TRY(ensure_scratch(ctx, 2 * half));
  uint32_t* d_idx = static_cast<uint32_t*>(ctx->scratch);
  float* d_out = reinterpret_cast<float*>(static_cast<char*>(ctx->scratch) + half);
  A3D_CUDA_TRY(cudaMemcpyAsync(d_idx, idx, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  const uint64_t blocks = (n + 255) / 256;
  gather_prob_idx_kernel<<<static_cast<unsigned>(blocks > 65535 ? 65535 : blocks), 256, 0, ctx->stream>>>(grid->view(), d_idx,
                                                                                                         d_out, n);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  A3D_CUDA_TRY(cudaMemcpyAsync(prob_out, d_out, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int amcl3d_cuda_grid_has_cells(const amcl3d_cuda_grid* grid, int* has_cells)
{
  if (!grid || !has_cells)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_has_cells: NULL argument");
  *has_cells = grid->has_cells ? 1 : 0;
  return 0;
}

int amcl3d_cuda_is_into_map(const amcl3d_cuda_grid* grid, float x, float y, float z, int* inside)
{
  if (!grid || !inside)
    return fail(AMCL3D_CUDA_ERR_INVALID, "is_into_map: NULL argument");
  // Grid3d.cpp:206-207 (pure host arithmetic on the bounds the grid was created with)
  const double* b = grid->bounds;
  *inside = (x >= b[0] && x < b[3] && y >= b[1] && y < b[4] && z >= b[2] && z < b[5]) ? 1 : 0;
  return 0;
}

}  // extern "C"
