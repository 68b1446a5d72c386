// voxel_grid.cu -- the step immediately before the hot path: down-sampling of the incoming sensor cloud
// (Node.cpp:131-137: pcl::VoxelGrid<pcl::PointXYZ>, setLeafSize(voxel_size x3), filter()).  SURVEY.md 8(f) rank 3.
//
// PCL is not vendored in the reference tree (package.xml:27 pins only `pcl_ros`); this restates the published
// algorithm of pcl/filters/impl/voxel_grid.hpp (PCL 1.7 / 1.8, PointXYZ, default settings: downsample_all_data,
// min_points_per_voxel 0, no filter limits):
//   1. bounding box of the finite points; inverse leaf = 1.f / leaf (float);
//      min_b = (int) floor(min * inv), max_b = (int) floor(max * inv), div_b = max_b - min_b + 1;
//      if dx * dy * dz overflows int32 the input is returned unchanged (PCL warns "leaf size is too small");
//   2. per point: ijk = (int)(floor(p * inv) - (float) min_b);  idx = ijk0 + ijk1 * div_b0 + ijk2 * div_b0 * div_b1;
//   3. points sorted by idx; one output point per distinct idx, in ascending idx order (x fastest, then y, z) --
//      this is the order of the cloud the weighting kernel walks (Grid3d.cpp:168-196);
//   4. output = (float sum of the cell's points) / (float) count.
// PCL sorts with std::sort, which leaves the order of the points INSIDE a cell unspecified; here it is the input
// order (the sort key is (idx, input index)), so centroids are deterministic and bit-equal to the CPU restatement the
// tests check against; against a real PCL build they may differ in the last bits of a centroid.
//
// Device path: bounding box (atomics on order-preserving keys) -> 64-bit keys -> bitonic sort (shared-memory
// passes for strides < 2048, one launch per larger stride) -> cell heads -> exclusive scan -> one thread per cell
// sums its run sequentially.
#include <cmath>
#include <cstring>
#include <limits>

#include "common.cuh"

namespace amcl3d_b200
{
namespace
{
__device__ __forceinline__ uint32_t vg_f2o(float f)
{
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float vg_o2f(uint32_t o)
{
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

__device__ __forceinline__ bool vg_finite(const float4 p)
{
  return fabsf(p.x) <= 3.4028234e38f && fabsf(p.y) <= 3.4028234e38f && fabsf(p.z) <= 3.4028234e38f;  // false for NaN
}

// box[0..2] = min, box[3..5] = max (order-preserving uint encoding), box[6] = number of finite points
__global__ void vg_bbox_kernel(const float4* __restrict__ pts, const uint32_t n, uint32_t* __restrict__ box)
{
  uint32_t lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0u, 0u, 0u }, cnt = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float4 p = pts[i];
    if (vg_finite(p))
    {
      const float c[3] = { p.x, p.y, p.z };
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        const uint32_t o = vg_f2o(c[a]);
        lo[a] = min(lo[a], o);
        hi[a] = max(hi[a], o);
      }
      ++cnt;
    }
  }
  for (int a = 0; a < 3; ++a)
  {
    for (int s = 16; s > 0; s >>= 1)
    {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], s));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], s));
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicMin(&box[a], lo[a]);
      atomicMax(&box[3 + a], hi[a]);
    }
  }
  for (int s = 16; s > 0; s >>= 1)
    cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
  if ((threadIdx.x & 31) == 0 && cnt)
    atomicAdd(&box[6], cnt);
}

struct VgParams
{
  float inv[3];
  float min_b[3];      // (float) min_b, as PCL subtracts it
  int mul1, mul2;      // divb_mul
};

// key = (idx << 32) | input index; non-finite points and the padding up to the power of two sort to the end
__global__ void vg_keys_kernel(const float4* __restrict__ pts, const uint32_t n, const uint32_t n_pad, const VgParams P,
                               unsigned long long* __restrict__ keys)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
  {
    unsigned long long k = 0xFFFFFFFFFFFFFFFFull;
    if (i < n)
    {
      const float4 p = pts[i];
      if (vg_finite(p))
      {
        const int i0 = static_cast<int>(__fsub_rn(floorf(__fmul_rn(p.x, P.inv[0])), P.min_b[0]));
        const int i1 = static_cast<int>(__fsub_rn(floorf(__fmul_rn(p.y, P.inv[1])), P.min_b[1]));
        const int i2 = static_cast<int>(__fsub_rn(floorf(__fmul_rn(p.z, P.inv[2])), P.min_b[2]));
        const int idx = i0 + i1 * P.mul1 + i2 * P.mul2;
        k = (static_cast<unsigned long long>(static_cast<uint32_t>(idx)) << 32) | i;
      }
    }
    keys[i] = k;
  }
}

// ---- bitonic sort of n_pad (power of two, >= 2048) 64-bit keys, ascending
constexpr uint32_t kSortTile = 2048;

__device__ __forceinline__ void cmp_swap(unsigned long long& a, unsigned long long& b, const bool ascending)
{
  if ((a > b) == ascending)
  {
    const unsigned long long t = a;
    a = b;
    b = t;
  }
}

// sorts every 2048-key tile completely (all stages k = 2 .. 2048); direction of the last stages by the global index
__global__ void __launch_bounds__(1024) vg_sort_tiles_kernel(unsigned long long* __restrict__ keys)
{
  __shared__ unsigned long long s[kSortTile];
  const uint32_t base = blockIdx.x * kSortTile;
  s[threadIdx.x] = keys[base + threadIdx.x];
  s[threadIdx.x + 1024] = keys[base + threadIdx.x + 1024];
  __syncthreads();
  for (uint32_t k = 2; k <= kSortTile; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1)
    {
      const uint32_t t = threadIdx.x;
      const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
      const bool asc = ((base + i) & k) == 0;
      cmp_swap(s[i], s[i | j], asc);
      __syncthreads();
    }
  keys[base + threadIdx.x] = s[threadIdx.x];
  keys[base + threadIdx.x + 1024] = s[threadIdx.x + 1024];
}

// one global stage (stride j >= 2048) of merge step k
__global__ void vg_sort_global_kernel(unsigned long long* __restrict__ keys, const uint32_t n_half, const uint32_t j,
                                      const uint32_t k)
{
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_half; t += gridDim.x * blockDim.x)
  {
    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    unsigned long long a = keys[i], b = keys[i | j];
    const bool asc = (i & k) == 0;
    if ((a > b) == asc)
    {
      keys[i] = b;
      keys[i | j] = a;
    }
  }
}

// the strides 1024 .. 1 of merge step k (k > 2048), inside shared memory
__global__ void __launch_bounds__(1024) vg_sort_merge_tail_kernel(unsigned long long* __restrict__ keys, const uint32_t k)
{
  __shared__ unsigned long long s[kSortTile];
  const uint32_t base = blockIdx.x * kSortTile;
  s[threadIdx.x] = keys[base + threadIdx.x];
  s[threadIdx.x + 1024] = keys[base + threadIdx.x + 1024];
  __syncthreads();
  const bool asc = (base & k) == 0;  // k > tile: one direction per tile
  for (uint32_t j = kSortTile >> 1; j > 0; j >>= 1)
  {
    const uint32_t t = threadIdx.x;
    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    cmp_swap(s[i], s[i | j], asc);
    __syncthreads();
  }
  keys[base + threadIdx.x] = s[threadIdx.x];
  keys[base + threadIdx.x + 1024] = s[threadIdx.x + 1024];
}

// head[i] = 1 where a new cell starts among the sorted valid keys
__global__ void vg_heads_kernel(const unsigned long long* __restrict__ keys, const uint32_t n, uint32_t* __restrict__ head)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const unsigned long long k = keys[i];
    const bool valid = k != 0xFFFFFFFFFFFFFFFFull;
    head[i] = (valid && (i == 0 || (keys[i - 1] >> 32) != (k >> 32))) ? 1u : 0u;
  }
}

// one thread per cell: float sums in sorted (= input) order, divided by the float count (voxel_grid.hpp centroid)
__global__ void vg_centroid_kernel(const float4* __restrict__ pts, const unsigned long long* __restrict__ keys,
                                   const uint32_t n, const uint32_t* __restrict__ head, const uint32_t* __restrict__ cell,
                                   float4* __restrict__ out)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    if (!head[i])
      continue;
    const unsigned long long idx = keys[i] >> 32;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t m = 0;
    for (uint32_t j = i; j < n && (keys[j] >> 32) == idx && keys[j] != 0xFFFFFFFFFFFFFFFFull; ++j, ++m)
    {
      const float4 p = pts[static_cast<uint32_t>(keys[j] & 0xFFFFFFFFull)];
      sx = __fadd_rn(sx, p.x);
      sy = __fadd_rn(sy, p.y);
      sz = __fadd_rn(sz, p.z);
    }
    const float d = static_cast<float>(m);
    out[cell[i]] = make_float4(__fdiv_rn(sx, d), __fdiv_rn(sy, d), __fdiv_rn(sz, d), 1.0f);  // PointXYZ: data[3] = 1
  }
}
}  // namespace
}  // namespace amcl3d_b200

using namespace amcl3d_b200;

extern "C" int amcl3d_cuda_voxel_grid(amcl3d_cuda_ctx* ctx, const float* cloud_xyzw, uint64_t n_cloud, float leaf_x,
                                      float leaf_y, float leaf_z, float* out_xyzw, uint64_t out_capacity,
                                      uint64_t* n_out)
{
  if (!ctx || !n_out || (n_cloud && (!cloud_xyzw || !out_xyzw)))
    return fail(AMCL3D_CUDA_ERR_INVALID, "voxel_grid: NULL argument");
  if (!(leaf_x > 0.f) || !(leaf_y > 0.f) || !(leaf_z > 0.f))
    return fail(AMCL3D_CUDA_ERR_INVALID, "voxel_grid: leaf sizes must be positive");
  *n_out = 0;
  if (n_cloud == 0)
    return 0;
  if (n_cloud >= 0x7FFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "voxel_grid: cloud too large");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  const uint32_t n = static_cast<uint32_t>(n_cloud);
  uint32_t n_pad = kSortTile;
  while (n_pad < n)
    n_pad <<= 1;

  // all device buffers are pieces of the context's persistent scratch arena (grown on demand, never freed per call)
  auto pad = [](size_t bytes) { return (bytes + 255) / 256 * 256; };
  const size_t b_pts = pad(static_cast<size_t>(n) * sizeof(float4)), b_keys = pad(static_cast<size_t>(n_pad) * sizeof(unsigned long long)),
               b_idx = pad((static_cast<size_t>(n) + 1) * sizeof(uint32_t));
  A3D_TRY(ensure_scratch(ctx, 2 * b_pts + b_keys + 2 * b_idx + 256));
  char* const base = static_cast<char*>(ctx->scratch);
  float4* const d_pts = reinterpret_cast<float4*>(base);
  float4* const d_out = reinterpret_cast<float4*>(base + b_pts);
  unsigned long long* const d_keys = reinterpret_cast<unsigned long long*>(base + 2 * b_pts);
  uint32_t* const d_head = reinterpret_cast<uint32_t*>(base + 2 * b_pts + b_keys);
  uint32_t* const d_cell = reinterpret_cast<uint32_t*>(base + 2 * b_pts + b_keys + b_idx);
  uint32_t* const d_box = reinterpret_cast<uint32_t*>(base + 2 * b_pts + b_keys + 2 * b_idx);
  auto cleanup = [&]() {};
#define VG_TRY(expr)                                                                                                  \
  do                                                                                                                  \
  {                                                                                                                   \
    cudaError_t vg_e_ = (expr);                                                                                       \
    if (vg_e_ != cudaSuccess)                                                                                         \
    {                                                                                                                 \
      cleanup();                                                                                                      \
      return fail(AMCL3D_CUDA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(vg_e_));                       \
    }                                                                                                                 \
  } while (0)
  VG_TRY(cudaMemcpyAsync(d_pts, cloud_xyzw, static_cast<size_t>(n) * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  VG_TRY(cudaMemsetAsync(d_box, 0xFF, 3 * sizeof(uint32_t), ctx->stream));
  VG_TRY(cudaMemsetAsync(d_box + 3, 0, 5 * sizeof(uint32_t), ctx->stream));
  const int blocks = static_cast<int>(std::min<uint32_t>((n + 255) / 256, static_cast<uint32_t>(ctx->sm_count) * 8));
  vg_bbox_kernel<<<blocks, 256, 0, ctx->stream>>>(d_pts, n, d_box);
  ctx->launches++;
  uint32_t box[8];
  VG_TRY(cudaMemcpyAsync(box, d_box, sizeof(box), cudaMemcpyDeviceToHost, ctx->stream));
  VG_TRY(cudaStreamSynchronize(ctx->stream));
  if (box[6] == 0)
  {
    cleanup();
    return 0;  // no finite point: empty output
  }
  // voxel_grid.hpp: inverse leaf size, overflow check, min_b / div_b / divb_mul -- all in float / int as PCL does
  const float leaf[3] = { leaf_x, leaf_y, leaf_z };
  VgParams P;
  int64_t d[3];
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; ++a)
  {
    const float inv = 1.0f / leaf[a];
    const float lo = vg_o2f(box[a]), hi = vg_o2f(box[3 + a]);
    d[a] = static_cast<int64_t>((hi - lo) * inv) + 1;
    min_b[a] = static_cast<int>(std::floor(lo * inv));
    const int max_b = static_cast<int>(std::floor(hi * inv));
    div_b[a] = max_b - min_b[a] + 1;
    P.inv[a] = inv;
    P.min_b[a] = static_cast<float>(min_b[a]);
  }
  if (d[0] * d[1] * d[2] > static_cast<int64_t>(std::numeric_limits<int32_t>::max()))
  {
    // "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = input
    cleanup();
    if (out_capacity < n_cloud)
      return fail(AMCL3D_CUDA_ERR_INVALID, "voxel_grid: output buffer too small for the pass-through case");
    std::memcpy(out_xyzw, cloud_xyzw, static_cast<size_t>(n) * sizeof(float4));
    *n_out = n_cloud;
    return 0;
  }
  P.mul1 = div_b[0];
  P.mul2 = div_b[0] * div_b[1];
  vg_keys_kernel<<<blocks, 256, 0, ctx->stream>>>(d_pts, n, n_pad, P, d_keys);
  vg_sort_tiles_kernel<<<n_pad / kSortTile, 1024, 0, ctx->stream>>>(d_keys);
  ctx->launches += 2;
  for (uint32_t k = 2 * kSortTile; k <= n_pad && k != 0; k <<= 1)
  {
    for (uint32_t j = k >> 1; j >= kSortTile; j >>= 1)
    {
      vg_sort_global_kernel<<<blocks, 256, 0, ctx->stream>>>(d_keys, n_pad / 2, j, k);
      ctx->launches++;
    }
    vg_sort_merge_tail_kernel<<<n_pad / kSortTile, 1024, 0, ctx->stream>>>(d_keys, k);
    ctx->launches++;
  }
  vg_heads_kernel<<<blocks, 256, 0, ctx->stream>>>(d_keys, n, d_head);
  ctx->launches++;
  VG_TRY(cudaMemsetAsync(d_head + n, 0, sizeof(uint32_t), ctx->stream));
  if (scan_u32(ctx, d_head, static_cast<uint64_t>(n) + 1, d_cell) != 0)
  {
    cleanup();
    return AMCL3D_CUDA_ERR_CUDA;
  }
  uint32_t n_cells = 0;
  VG_TRY(cudaMemcpyAsync(&n_cells, d_cell + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VG_TRY(cudaStreamSynchronize(ctx->stream));
  if (n_cells > out_capacity)
  {
    cleanup();
    return fail(AMCL3D_CUDA_ERR_INVALID, "voxel_grid: output buffer too small");
  }
  vg_centroid_kernel<<<blocks, 256, 0, ctx->stream>>>(d_pts, d_keys, n, d_head, d_cell, d_out);
  ctx->launches++;
  VG_TRY(cudaGetLastError());
  VG_TRY(cudaMemcpyAsync(out_xyzw, d_out, static_cast<size_t>(n_cells) * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  VG_TRY(cudaStreamSynchronize(ctx->stream));
#undef VG_TRY
  cleanup();
  *n_out = n_cells;
  return 0;
}
