// distance_field.cu -- Grid3d::open's map preprocessing: computeGrid (PointCloudTools.cpp:84-149).
//
// The reference asks a kd-tree for the exact nearest map point of every voxel corner, one voxel at a time.
// Here the map points are bucketed into coarse blocks of 8x8x8 voxels (counting sort on the device) and ONE
// CTA OWNS ONE 8x8x8 VOXEL TILE (one thread per voxel).  The CTA visits the surrounding blocks in growing
// Chebyshev rings; the points of all non-empty candidate blocks of a ring are streamed through shared memory
// and every thread keeps the minimum of the FLOAT squared distance ((dx*dx + dy*dy) + dz*dz, no FMA --
// FLANN's L2_Simple<float>, which is what the reference's kd-tree returns).  A ring is not started once
// every voxel of the tile already has a neighbour closer than anything that ring could contain, and blocks
// farther than the tile's current worst distance are skipped, so the result is the exact minimum over ALL
// points -- the same number the kd-tree returns -- for every voxel, near or far.
//
// Multi-GPU: tiles are split by z-slab across ranks (amcl3d_cuda_comm_*); every rank holds all points, no halo
// exchange is needed because a CTA can reach any block; slabs are exchanged with one broadcast per rank.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace amcl3d_b200
{
constexpr int kBlk = 8;             // voxels per block edge
constexpr int kTileThreads = 512;   // kBlk^3
#define kInf (__int_as_float(0x7f800000))

struct BlockGrid
{
  int nb[3];        // blocks per axis (including padding)
  int pad[3];       // blocks of padding in front of the map's block (0,0,0)
  double origin[3]; // metric origin of block (0,0,0) = min - pad*kBlk*res
  double bsize;     // block edge in metres
  double res;
  double min[3];
  uint32_t dims[3]; // voxel grid size
};

__device__ __forceinline__ int point_block_axis(const BlockGrid& bg, int a, float c)
{
  int k = static_cast<int>(floor((static_cast<double>(c) - bg.origin[a]) / bg.bsize));
  k = k < 0 ? 0 : k;
  k = k >= bg.nb[a] ? bg.nb[a] - 1 : k;
  return k;
}

__device__ __forceinline__ uint32_t point_block(const BlockGrid& bg, const float4 p)
{
  const int kx = point_block_axis(bg, 0, p.x), ky = point_block_axis(bg, 1, p.y), kz = point_block_axis(bg, 2, p.z);
  return (static_cast<uint32_t>(kz) * bg.nb[1] + ky) * bg.nb[0] + kx;
}

__global__ void df_count_kernel(const BlockGrid bg, const float4* __restrict__ pts, const uint64_t n,
                                uint32_t* __restrict__ counts)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&counts[point_block(bg, pts[i])], 1u);
}

__global__ void df_scatter_kernel(const BlockGrid bg, const float4* __restrict__ pts, const uint64_t n,
                                  const uint32_t* __restrict__ start, uint32_t* __restrict__ fill,
                                  float4* __restrict__ sorted)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float4 p = pts[i];
    const uint32_t b = point_block(bg, p);
    const uint32_t slot = start[b] + atomicAdd(&fill[b], 1u);
    sorted[slot] = p;
  }
}

// ---- exclusive uint32 scan (three passes) -------------------------------------------------------------------
constexpr int kUScanBlock = 256;
constexpr int kUScanItems = 8;
__global__ void __launch_bounds__(kUScanBlock) uscan_sums_kernel(const uint32_t* __restrict__ in, const uint64_t n,
                                                               uint32_t* __restrict__ sums)
{
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kUScanBlock * kUScanItems;
  uint32_t s = 0;
  for (int k = 0; k < kUScanItems; ++k)
  {
    const uint64_t i = base + static_cast<uint64_t>(k) * kUScanBlock + threadIdx.x;
    if (i < n)
      s += in[i];
  }
  __shared__ uint32_t red[kUScanBlock / 32];
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint32_t t = 0;
    for (int k = 0; k < kUScanBlock / 32; ++k)
      t += red[k];
    sums[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) uscan_offsets_kernel(uint32_t* __restrict__ sums, const uint32_t n_blocks)
{
  // one block; chunked exclusive scan of the per-block totals
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0)
    carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_blocks; base += 1024)
  {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_blocks ? sums[i] : 0u;
    uint32_t incl = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o)
        incl += t;
    }
    if (lane == 31)
      warp_tot[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int k = 0; k < warp; ++k)
      woff += warp_tot[k];
    const uint32_t c = carry;
    if (i < n_blocks)
      sums[i] = c + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023)
      carry = c + woff + incl;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kUScanBlock) uscan_final_kernel(const uint32_t* __restrict__ in, const uint64_t n,
                                                                const uint32_t* __restrict__ offsets,
                                                                uint32_t* __restrict__ out)
{
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kUScanBlock * kUScanItems + static_cast<uint64_t>(threadIdx.x) * kUScanItems;
  uint32_t v[kUScanItems];
  uint32_t s = 0;
  for (int k = 0; k < kUScanItems; ++k)
  {
    const uint64_t i = base + k;
    v[k] = s;  // exclusive within the thread
    s += (i < n) ? in[i] : 0u;
  }
  __shared__ uint32_t warp_tot[kUScanBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = s;
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o)
      incl += t;
  }
  if (lane == 31)
    warp_tot[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int k = 0; k < warp; ++k)
    woff += warp_tot[k];
  const uint32_t off = offsets[blockIdx.x] + woff + (incl - s);
  for (int k = 0; k < kUScanItems; ++k)
  {
    const uint64_t i = base + k;
    if (i < n)
      out[i] = off + v[k];
  }
}

// ---- the tile kernel ----------------------------------------------------------------------------------------
struct DfParams
{
  BlockGrid bg;
  int tiles[3];      // voxel tiles per axis
  int tz0, tz1;      // z-tile range handled by this launch (z-slab sharding)
  float g1, g2;      // PointCloudTools.cpp:114-115
  // Probability-only builds (no distance plane): a squared distance beyond which prob = g1 * expf(-d2*d2*g2) is
  // exactly +0 in float (argument below -110; expf underflows to 0 below -104), so the search may stop there.
  // +inf when the distance plane is kept (far distances are then part of the result).
  float d2_cut;
  uint64_t n_points;
  uint32_t brick_shift, nbx, nby;  // physical layout of the output planes (GridView::brick_shift)
};

__global__ void __launch_bounds__(kTileThreads)
    df_tile_kernel(const DfParams P, const uint32_t* __restrict__ start, const float4* __restrict__ pts,
                   float* __restrict__ dist_out, float* __restrict__ prob_out)
{
  __shared__ float4 s_pts[kTileThreads];
  __shared__ uint32_t s_cstart[kTileThreads];
  __shared__ uint32_t s_prefix[kTileThreads + 1];
  __shared__ uint32_t s_warp[kTileThreads / 32];
  __shared__ uint32_t s_nc;
  __shared__ float s_wmax[kTileThreads / 32];
  __shared__ float s_tile_worst;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // tile coordinates
  const uint32_t tiles_xy = static_cast<uint32_t>(P.tiles[0]) * P.tiles[1];
  const uint32_t tlin = blockIdx.x;
  const int tz = P.tz0 + static_cast<int>(tlin / tiles_xy);
  const int ty = static_cast<int>((tlin % tiles_xy) / P.tiles[0]);
  const int tx = static_cast<int>(tlin % P.tiles[0]);
  const uint32_t ix = tx * kBlk + (tid & 7), iy = ty * kBlk + ((tid >> 3) & 7), iz = tz * kBlk + (tid >> 6);
  const bool valid = ix < P.bg.dims[0] && iy < P.bg.dims[1] && iz < P.bg.dims[2];
  // PointCloudTools.cpp:127-129: search point = min + i*res in double, stored to float
  const float qx = static_cast<float>(__dadd_rn(P.bg.min[0], __dmul_rn(static_cast<double>(ix), P.bg.res)));
  const float qy = static_cast<float>(__dadd_rn(P.bg.min[1], __dmul_rn(static_cast<double>(iy), P.bg.res)));
  const float qz = static_cast<float>(__dadd_rn(P.bg.min[2], __dmul_rn(static_cast<double>(iz), P.bg.res)));
  // this tile's block coordinates in the padded block grid
  const int bx = tx + P.bg.pad[0], by = ty + P.bg.pad[1], bz = tz + P.bg.pad[2];

  float best = valid ? kInf : 0.f;
  if (tid == 0)
    s_tile_worst = kInf;
  __syncthreads();

  const float bsz = static_cast<float>(P.bg.bsize);
  const float res = static_cast<float>(P.bg.res);
  int max_ring = 0;
  max_ring = max(max_ring, max(bx, P.bg.nb[0] - 1 - bx));
  max_ring = max(max_ring, max(by, P.bg.nb[1] - 1 - by));
  max_ring = max(max_ring, max(bz, P.bg.nb[2] - 1 - bz));
  if (P.n_points == 0)
    max_ring = -1;  // nothing to search: every voxel takes the "no neighbour" branch

  for (int r = 0; r <= max_ring; ++r)
  {
    const float worst = fminf(s_tile_worst, P.d2_cut);
    if (r >= 2)
    {
      // every point of ring r is farther than (r-1)*block from every corner of this tile
      const float reach = static_cast<float>(r - 1) * bsz;
      if (reach * reach > worst * 1.0001f)
        break;
    }
    const int side = 2 * r + 1;
    const int n_pos = side * side * side;
    for (int pbase = 0; pbase < n_pos; pbase += kTileThreads)
    {
      // ---- each thread inspects one position of the (2r+1)^3 cube; only the shell (ring r) counts
      uint32_t my_start = 0, my_cnt = 0;
      const int pos = pbase + tid;
      if (pos < n_pos)
      {
        const int dz = pos / (side * side) - r, dy = (pos / side) % side - r, dx = pos % side - r;
        const int m = max(abs(dx), max(abs(dy), abs(dz)));
        const int cx = bx + dx, cy = by + dy, cz = bz + dz;
        if (m == r && cx >= 0 && cx < P.bg.nb[0] && cy >= 0 && cy < P.bg.nb[1] && cz >= 0 && cz < P.bg.nb[2])
        {
          const uint32_t b = (static_cast<uint32_t>(cz) * P.bg.nb[1] + cy) * P.bg.nb[0] + cx;
          const uint32_t s = start[b], e = start[b + 1];
          if (e > s)
          {
            // smallest possible distance between a corner of this tile and a point of that block
            auto gap = [&](int k) -> float {
              if (k >= 1)
                return static_cast<float>((k - 1) * kBlk + 1) * res;
              if (k <= -1)
                return static_cast<float>((-k - 1) * kBlk) * res;
              return 0.f;
            };
            const float gx = gap(dx), gy = gap(dy), gz = gap(dz);
            const float lb = gx * gx + gy * gy + gz * gz;
            if (!(lb * 0.9999f > worst))
            {
              my_start = s;
              my_cnt = e - s;
            }
          }
        }
      }
      // ---- compact the candidates and build the prefix of their point counts
      const unsigned ballot = __ballot_sync(0xffffffffu, my_cnt > 0);
      if (lane == 0)
        s_warp[warp] = __popc(ballot);
      __syncthreads();
      uint32_t woff = 0, nc = 0;
      for (int k = 0; k < kTileThreads / 32; ++k)
      {
        const uint32_t c = s_warp[k];
        if (k < warp)
          woff += c;
        nc += c;
      }
      if (my_cnt > 0)
      {
        const uint32_t slot = woff + __popc(ballot & ((1u << lane) - 1u));
        s_cstart[slot] = my_start;
        s_prefix[slot + 1] = my_cnt;  // turned into an inclusive prefix below
      }
      if (tid == 0)
      {
        s_prefix[0] = 0;
        s_nc = nc;
      }
      __syncthreads();
      if (nc == 0)
        continue;  // uniform across the CTA
      // inclusive scan of s_prefix[1..nc] (nc <= 512): one element per thread
      {
        uint32_t v = (static_cast<uint32_t>(tid) < nc) ? s_prefix[tid + 1] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1)
        {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o)
            incl += t;
        }
        __syncthreads();
        if (lane == 31)
          s_warp[warp] = incl;
        __syncthreads();
        uint32_t wo = 0;
        for (int k = 0; k < warp; ++k)
          wo += s_warp[k];
        if (static_cast<uint32_t>(tid) < nc)
          s_prefix[tid + 1] = wo + incl;
        __syncthreads();
      }
      const uint32_t total = s_prefix[nc];
      // ---- stream the candidate points through shared memory
      for (uint32_t base = 0; base < total; base += kTileThreads)
      {
        const uint32_t idx = base + tid;
        if (idx < total)
        {
          // candidate c with s_prefix[c] <= idx < s_prefix[c+1]
          uint32_t lo = 0, hi = nc;
          while (hi - lo > 1)
          {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_prefix[mid] <= idx)
              lo = mid;
            else
              hi = mid;
          }
          s_pts[tid] = pts[s_cstart[lo] + (idx - s_prefix[lo])];
        }
        __syncthreads();
        const int len = static_cast<int>(min(static_cast<uint32_t>(kTileThreads), total - base));
        if (valid)
        {
#pragma unroll 4
          for (int j = 0; j < len; ++j)
          {
            const float4 p = s_pts[j];
            // FLANN L2_Simple<float>: diff = query - point; result += diff*diff, in float, no FMA
            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            best = fminf(best, d);
          }
        }
        __syncthreads();
      }
      // ---- refresh the tile's worst distance (prunes the rest of this ring and later rings)
      {
        float w = best;
        for (int o = 16; o > 0; o >>= 1)
          w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (lane == 0)
          s_wmax[warp] = w;
        __syncthreads();
        if (tid == 0)
        {
          float m = 0.f;
          for (int k = 0; k < kTileThreads / 32; ++k)
            m = fmaxf(m, s_wmax[k]);
          s_tile_worst = m;
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }

  if (valid)
  {
    uint64_t index;
    if (P.brick_shift == 0)
      index = static_cast<uint64_t>(ix) + static_cast<uint64_t>(iy) * P.bg.dims[0] +
              static_cast<uint64_t>(iz) * P.bg.dims[0] * P.bg.dims[1];
    else
    {
      const uint32_t b = P.brick_shift, m = (1u << b) - 1u;
      const uint32_t brick = ((iz >> b) * P.nby + (iy >> b)) * P.nbx + (ix >> b);
      index = (brick << (3 * b)) | ((iz & m) << (2 * b)) | ((iy & m) << b) | (ix & m);
    }
    if (P.n_points == 0 || best == kInf)
    {
      // PointCloudTools.cpp:139-143: no neighbour
      if (dist_out)
        dist_out[index] = -1.0f;
      prob_out[index] = 0.0f;
    }
    else
    {
      if (dist_out)
        dist_out[index] = best;  // :136 the SQUARED distance
      // :137 prob = gauss_const1 * expf(-dist * dist * gauss_const2), all float
      prob_out[index] = __fmul_rn(P.g1, expf(__fmul_rn(__fmul_rn(-best, best), P.g2)));
    }
  }
}

int scan_u32(amcl3d_cuda_ctx* ctx, const uint32_t* d_in, uint64_t n, uint32_t* d_out)
{
  const uint32_t n_blocks = static_cast<uint32_t>((n + kUScanBlock * kUScanItems - 1) / (kUScanBlock * kUScanItems));
  uint32_t* d_sums = nullptr;
  A3D_CUDA_TRY(cudaMalloc(&d_sums, static_cast<size_t>(n_blocks) * 4));
  uscan_sums_kernel<<<n_blocks, kUScanBlock, 0, ctx->stream>>>(d_in, n, d_sums);
  uscan_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(d_sums, n_blocks);
  uscan_final_kernel<<<n_blocks, kUScanBlock, 0, ctx->stream>>>(d_in, n, d_sums, d_out);
  ctx->launches += 3;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_sums);
  if (e != cudaSuccess)
    return fail(AMCL3D_CUDA_ERR_CUDA, std::string("scan_u32: ") + cudaGetErrorString(e));
  return 0;
}

int comm_broadcast(amcl3d_cuda_ctx* ctx, void* d_buf, size_t bytes, int root);  // comm.cu

}  // namespace amcl3d_b200

using namespace amcl3d_b200;

extern "C" int amcl3d_cuda_grid_compute(amcl3d_cuda_grid* grid, const float* points_xyzw, uint64_t n_points,
                                        double sensor_dev, int keep_dist)
{
  if (!grid || (n_points && !points_xyzw))
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_compute: NULL argument");
  if (!(sensor_dev > 0.0))
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_compute: sensor_dev must be positive");
  if (n_points >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "grid_compute: too many map points");
  amcl3d_cuda_ctx* ctx = grid->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  if (!grid->d_prob)
  {
    // + kZeroCellPad floats: GridView::zero_index points at a cell that always reads 0 (skipped points gather it)
    A3D_CUDA_TRY(cudaMalloc(&grid->d_prob, (grid->n_phys + kZeroCellPad) * sizeof(float)));
    if (grid->brick_shift)  // padding voxels of partial bricks are never read, but keep them defined
      A3D_CUDA_TRY(cudaMemsetAsync(grid->d_prob, 0, (grid->n_phys + kZeroCellPad) * sizeof(float), ctx->stream));
    else
      A3D_CUDA_TRY(cudaMemsetAsync(grid->d_prob + grid->n_phys, 0, kZeroCellPad * sizeof(float), ctx->stream));
  }
  if (keep_dist && !grid->d_dist)
  {
    A3D_CUDA_TRY(cudaMalloc(&grid->d_dist, grid->n_phys * sizeof(float)));
    if (grid->brick_shift)
      A3D_CUDA_TRY(cudaMemsetAsync(grid->d_dist, 0, grid->n_phys * sizeof(float), ctx->stream));
  }
  if (!keep_dist && grid->d_dist)
  {
    cudaFree(grid->d_dist);
    grid->d_dist = nullptr;
  }

  // block grid covering the map bounds and every point
  DfParams P;
  std::memset(&P, 0, sizeof(P));
  BlockGrid& bg = P.bg;
  bg.res = grid->bounds[6];
  bg.bsize = bg.res * kBlk;
  double lo[3], hi[3];
  for (int a = 0; a < 3; ++a)
  {
    bg.min[a] = grid->bounds[a];
    bg.dims[a] = grid->dims[a];
    P.tiles[a] = static_cast<int>((grid->dims[a] + kBlk - 1) / kBlk);
    lo[a] = grid->bounds[a];
    hi[a] = grid->bounds[a] + P.tiles[a] * bg.bsize;
  }
  for (uint64_t i = 0; i < n_points; ++i)
    for (int a = 0; a < 3; ++a)
    {
      const double c = points_xyzw[4 * i + a];
      if (c < lo[a])
        lo[a] = c;
      if (c > hi[a])
        hi[a] = c;
    }
  uint64_t n_blocks = 1;
  for (int a = 0; a < 3; ++a)
  {
    bg.pad[a] = static_cast<int>(std::ceil((grid->bounds[a] - lo[a]) / bg.bsize));
    bg.origin[a] = grid->bounds[a] - bg.pad[a] * bg.bsize;
    bg.nb[a] = static_cast<int>(std::floor((hi[a] - bg.origin[a]) / bg.bsize)) + 1;
    if (bg.nb[a] < P.tiles[a] + bg.pad[a])
      bg.nb[a] = P.tiles[a] + bg.pad[a];
    n_blocks *= static_cast<uint64_t>(bg.nb[a]);
  }
  if (n_blocks >= 0x7FFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_TOO_BIG, "grid_compute: map points lie too far outside the map bounds");
  // PointCloudTools.cpp:114-115
  P.g1 = static_cast<float>(1. / (sensor_dev * std::sqrt(2 * M_PI)));
  P.g2 = static_cast<float>(1. / (2. * sensor_dev * sensor_dev));
  P.n_points = n_points;
  P.d2_cut = keep_dist ? INFINITY : std::sqrt(110.f / P.g2);

  float4 *d_pts = nullptr, *d_sorted = nullptr;
  uint32_t *d_counts = nullptr, *d_start = nullptr;
  auto cleanup = [&]() {
    if (d_pts)
      cudaFree(d_pts);
    if (d_sorted)
      cudaFree(d_sorted);
    if (d_counts)
      cudaFree(d_counts);
    if (d_start)
      cudaFree(d_start);
  };
#define DF_TRY(expr)                                                                                                   \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e_ = (expr);                                                                                           \
    if (e_ != cudaSuccess)                                                                                             \
    {                                                                                                                  \
      cleanup();                                                                                                       \
      return fail(AMCL3D_CUDA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                           \
    }                                                                                                                  \
  } while (0)
  DF_TRY(cudaMalloc(&d_pts, (n_points ? n_points : 1) * sizeof(float4)));
  DF_TRY(cudaMalloc(&d_sorted, (n_points ? n_points : 1) * sizeof(float4)));
  DF_TRY(cudaMalloc(&d_counts, (n_blocks + 1) * sizeof(uint32_t)));
  DF_TRY(cudaMalloc(&d_start, (n_blocks + 1) * sizeof(uint32_t)));
  DF_TRY(cudaMemsetAsync(d_counts, 0, (n_blocks + 1) * sizeof(uint32_t), ctx->stream));
  if (n_points)
  {
    DF_TRY(cudaMemcpyAsync(d_pts, points_xyzw, n_points * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    df_count_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(bg, d_pts, n_points, d_counts);
    ctx->launches++;
  }
  {
    int rc = scan_u32(ctx, d_counts, n_blocks + 1, d_start);
    if (rc != 0)
    {
      cleanup();
      return rc;
    }
  }
  if (n_points)
  {
    DF_TRY(cudaMemsetAsync(d_counts, 0, (n_blocks + 1) * sizeof(uint32_t), ctx->stream));
    df_scatter_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(bg, d_pts, n_points, d_start, d_counts, d_sorted);
    ctx->launches++;
  }
  // z-slab of tiles owned by this rank
  P.brick_shift = grid->brick_shift;
  P.nbx = grid->nb[0];
  P.nby = grid->nb[1];
  const int tz_total = P.tiles[2];
  // Slab boundaries (in z-tiles; whole brick rows when bricked so that each rank's output is one contiguous address
  // range).  Equal heights are badly balanced on real maps (a warehouse is occupied in its lower third), so the rows
  // are weighted: a tile layer costs ~20 point-visits per tile for the empty-neighbourhood search plus one per map
  // point in its own and the two adjacent block layers (calibrated on map L: empty layers 3 ms, occupied 13 ms).
  // Every rank holds all points, hence the same counts, hence the same boundaries.
  std::vector<int> slab(static_cast<size_t>(ctx->n_ranks) + 1, tz_total);
  slab[0] = 0;
  {
    const int row_tiles = grid->brick_shift ? (1 << grid->brick_shift) / kBlk : 1;
    const int n_rows = (tz_total + row_tiles - 1) / row_tiles;
    std::vector<double> cost(static_cast<size_t>(n_rows), 0.0);
    std::vector<uint32_t> layer_start(static_cast<size_t>(bg.nb[2]) + 1, 0u);
    if (n_points && ctx->n_ranks > 1)
    {
      const size_t per_layer = static_cast<size_t>(bg.nb[0]) * bg.nb[1];
      DF_TRY(cudaMemcpy2DAsync(layer_start.data(), sizeof(uint32_t), d_start, per_layer * sizeof(uint32_t), sizeof(uint32_t),
                               static_cast<size_t>(bg.nb[2]) + 1, cudaMemcpyDeviceToHost, ctx->stream));
      DF_TRY(cudaStreamSynchronize(ctx->stream));
    }
    const double tiles_xy = static_cast<double>(P.tiles[0]) * P.tiles[1];
    for (int tz = 0; tz < tz_total; ++tz)
    {
      double pts = 0.0;
      for (int dz = -1; dz <= 1; ++dz)
      {
        const int bz = tz + bg.pad[2] + dz;
        if (bz >= 0 && bz < bg.nb[2])
          pts += static_cast<double>(layer_start[bz + 1] - layer_start[bz]);
      }
      cost[tz / row_tiles] += 20.0 * tiles_xy + pts;
    }
    double total = 0.0;
    for (double c : cost)
      total += c;
    double acc = 0.0;
    int r = 1;
    for (int row = 0; row < n_rows && r < ctx->n_ranks; ++row)
    {
      acc += cost[row];
      while (r < ctx->n_ranks && acc >= total * r / ctx->n_ranks)
        slab[r++] = std::min(tz_total, (row + 1) * row_tiles);
    }
  }
  P.tz0 = slab[ctx->rank];
  P.tz1 = slab[ctx->rank + 1];
  const uint64_t n_tiles = static_cast<uint64_t>(P.tiles[0]) * P.tiles[1] * (P.tz1 - P.tz0);
  if (n_tiles >= 0x7FFFFFFFull)
  {
    cleanup();
    return fail(AMCL3D_CUDA_ERR_TOO_BIG, "grid_compute: too many tiles for one launch");
  }
  if (n_tiles > 0)
  {
    if (ctx->opt_kernel_timing)
      cudaEventRecord(ctx->ev_k0, ctx->stream);
    df_tile_kernel<<<static_cast<unsigned>(n_tiles), kTileThreads, 0, ctx->stream>>>(P, d_start, d_sorted, grid->d_dist,
                                                                                    grid->d_prob);
    if (ctx->opt_kernel_timing)
    {
      cudaEventRecord(ctx->ev_k1, ctx->stream);
      ctx->ev_valid = true;
    }
    ctx->launches++;
  }
  DF_TRY(cudaGetLastError());
  if (ctx->n_ranks > 1)
  {
    // replicate: every rank broadcasts its slab of layers
    // linear storage: a slab is dims_x*dims_y floats per layer; bricked: whole brick rows, (nbx*nby << 3b) >> b per layer
    const uint64_t layer = grid->brick_shift ? (static_cast<uint64_t>(grid->nb[0]) * grid->nb[1]) << (2 * grid->brick_shift) :
                                               static_cast<uint64_t>(grid->dims[0]) * grid->dims[1];
    const uint64_t z_limit = grid->brick_shift ? (static_cast<uint64_t>(grid->nb[2]) << grid->brick_shift) : grid->dims[2];
    for (int r = 0; r < ctx->n_ranks; ++r)
    {
      const uint64_t z0 = std::min<uint64_t>(z_limit, static_cast<uint64_t>(slab[r]) * kBlk);
      uint64_t z1 = std::min<uint64_t>(z_limit, static_cast<uint64_t>(slab[r + 1]) * kBlk);
      if (grid->brick_shift && slab[r + 1] == tz_total && slab[r] < tz_total)
        z1 = z_limit;  // the last non-empty slab owns the padding layers of the last brick row
      if (z1 <= z0)
        continue;
      int rc = comm_broadcast(ctx, grid->d_prob + z0 * layer, (z1 - z0) * layer * sizeof(float), r);
      if (rc == 0 && grid->d_dist)
        rc = comm_broadcast(ctx, grid->d_dist + z0 * layer, (z1 - z0) * layer * sizeof(float), r);
      if (rc != 0)
      {
        cleanup();
        return rc;
      }
    }
  }
  DF_TRY(cudaStreamSynchronize(ctx->stream));
#undef DF_TRY
  cleanup();
  grid->sensor_dev = sensor_dev;
  grid->has_cells = true;
  return 0;
}
