// cloud.cu -- spatial re-ordering of the staged sensor cloud for the large-map regime.
//
// When the likelihood grid does not fit L2 (bricked layout, see GridView::brick_shift) the gather is only fast if the
// points that are processed at about the same time reach voxels that lie close together.  A voxel-filtered cloud
// arrives ordered by (z, y, x) voxel index, i.e. 512 consecutive points span a whole horizontal slab of the scene.
// This file reorders the cloud along a 3-D Morton curve with a counting sort (15-bit keys = 32^3 cells over the
// cloud's bounding box): consecutive points become spatial neighbours and a 512-point chunk touches a few MB of grid.
//
// The order of the points is the order of the reference's float accumulation (Grid3d.cpp:191), so a re-ordered cloud
// gives weights that differ from the reference's in the last bits (well inside the 1e-5 tolerance); it is therefore
// only applied where the caller did not ask for the bit-exact order (option "cloud_order").
#include "common.cuh"

namespace amcl3d_b200
{
constexpr uint32_t kMortonBits = 5;                       // per axis
constexpr uint32_t kMortonCells = 1u << (3 * kMortonBits);  // 32768 buckets

// monotone float <-> uint mapping so that atomicMin/atomicMax on the bit pattern order like the floats
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o)
{
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// box[0..2] = min x,y,z ; box[3..5] = max x,y,z   (ordered-uint encoding; initialise to 0xFFFFFFFF / 0)
__global__ void cloud_bbox_kernel(const float4* __restrict__ cloud, const uint32_t n, uint32_t* __restrict__ box)
{
  uint32_t lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0u, 0u, 0u };
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float4 p = cloud[i];
    const float c[3] = { p.x, p.y, p.z };
    for (int a = 0; a < 3; ++a)
    {
      if (c[a] == c[a] && fabsf(c[a]) < 1e30f)  // ignore NaN / infinite points
      {
        const uint32_t o = float_to_ordered(c[a]);
        lo[a] = min(lo[a], o);
        hi[a] = max(hi[a], o);
      }
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
}

__device__ __forceinline__ uint32_t spread3(uint32_t v)  // 5 bits -> every third bit
{
  uint32_t r = 0;
#pragma unroll
  for (uint32_t b = 0; b < kMortonBits; ++b)
    r |= ((v >> b) & 1u) << (3 * b);
  return r;
}

__device__ __forceinline__ uint32_t morton_cell(const float4 p, const uint32_t* box)
{
  uint32_t key = 0;
  const float c[3] = { p.x, p.y, p.z };
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    const float lo = ordered_to_float(box[a]), hi = ordered_to_float(box[3 + a]);
    const float span = hi - lo;
    float t = span > 0.f ? (c[a] - lo) / span : 0.f;
    t = (t == t) ? fminf(fmaxf(t, 0.f), 1.f) : 0.f;
    const uint32_t q = min(static_cast<uint32_t>(t * static_cast<float>(1u << kMortonBits)), (1u << kMortonBits) - 1u);
    key |= spread3(q) << a;
  }
  return key;
}

__global__ void cloud_hist_kernel(const float4* __restrict__ cloud, const uint32_t n, const uint32_t* __restrict__ box,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ hist)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const uint32_t k = morton_cell(cloud[i], box);
    keys[i] = k;
    atomicAdd(&hist[k], 1u);
  }
}

// One block: exclusive scan of the 32768 bucket counts (in place).
__global__ void __launch_bounds__(1024) cloud_scan_kernel(uint32_t* __restrict__ hist)
{
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0)
    carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < kMortonCells; base += 1024)
  {
    const uint32_t v = hist[base + threadIdx.x];
    uint32_t incl = v;
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
    hist[base + threadIdx.x] = c + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023)
      carry = c + woff + incl;
    __syncthreads();
  }
}

// Scatter with one atomic cursor per bucket; the arrival order inside a bucket is fixed up afterwards.
__global__ void cloud_scatter_kernel(const float4* __restrict__ cloud, const uint32_t n, const uint32_t* __restrict__ keys,
                                     uint32_t* __restrict__ cursor, float4* __restrict__ sorted)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const uint32_t slot = atomicAdd(&cursor[keys[i]], 1u);
    float4 p = cloud[i];
    p.w = __uint_as_float(i);  // remember the original index: buckets are put back into input order below
    sorted[slot] = p;
  }
}

// Restores input order inside every bucket (insertion sort on the stored original index; buckets hold a handful of
// points), which makes the permutation -- and therefore every weight -- independent of atomic scheduling.
__global__ void cloud_fix_order_kernel(float4* __restrict__ sorted, const uint32_t* __restrict__ start,
                                       const uint32_t* __restrict__ end_cursor)
{
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < kMortonCells; b += gridDim.x * blockDim.x)
  {
    const uint32_t s = start[b], e = end_cursor[b];
    for (uint32_t i = s + 1; i < e; ++i)
    {
      const float4 v = sorted[i];
      uint32_t j = i;
      while (j > s && __float_as_uint(sorted[j - 1].w) > __float_as_uint(v.w))
      {
        sorted[j] = sorted[j - 1];
        --j;
      }
      sorted[j] = v;
    }
  }
}

// d_cloud (n points) is replaced by its Morton-bucket ordering; d_tmp must hold n points; d_work 2*32768+8+n uint32.
int sort_cloud_morton(amcl3d_cuda_ctx* ctx, float4* d_cloud, float4* d_tmp, uint32_t* d_work, uint32_t n)
{
  if (n < 2)
    return 0;
  uint32_t* box = d_work;                       // 8 words (6 used)
  uint32_t* start = d_work + 8;                 // kMortonCells
  uint32_t* cursor = start + kMortonCells;      // kMortonCells
  uint32_t* keys = cursor + kMortonCells;       // n
  const int blocks = static_cast<int>(std::min<uint32_t>((n + 255) / 256, static_cast<uint32_t>(ctx->sm_count) * 4));
  A3D_CUDA_TRY(cudaMemsetAsync(box, 0xFF, 3 * sizeof(uint32_t), ctx->stream));
  A3D_CUDA_TRY(cudaMemsetAsync(box + 3, 0, 3 * sizeof(uint32_t), ctx->stream));
  A3D_CUDA_TRY(cudaMemsetAsync(start, 0, kMortonCells * sizeof(uint32_t), ctx->stream));
  cloud_bbox_kernel<<<blocks, 256, 0, ctx->stream>>>(d_cloud, n, box);
  cloud_hist_kernel<<<blocks, 256, 0, ctx->stream>>>(d_cloud, n, box, keys, start);
  cloud_scan_kernel<<<1, 1024, 0, ctx->stream>>>(start);
  A3D_CUDA_TRY(cudaMemcpyAsync(cursor, start, kMortonCells * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  cloud_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(d_cloud, n, keys, cursor, d_tmp);
  cloud_fix_order_kernel<<<kMortonCells / 256, 256, 0, ctx->stream>>>(d_tmp, start, cursor);
  A3D_CUDA_TRY(cudaMemcpyAsync(d_cloud, d_tmp, static_cast<size_t>(n) * sizeof(float4), cudaMemcpyDeviceToDevice,
                               ctx->stream));
  ctx->launches += 5;
  A3D_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace amcl3d_b200
