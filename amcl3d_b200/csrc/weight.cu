// weight.cu -- the measurement-update weighting kernel (Grid3d::computeCloudWeight, Grid3d.cpp:133-199,
// batched over particles as ParticleFilter::update's first loop does, ParticleFilter.cpp:129-153).
//
// Mapping: ONE LANE PER PARTICLE, cloud points broadcast from shared memory.
//   * every lane walks the cloud in cloud order, so its running float sum is the reference's own sequential
//     sum (bit-exact when the cloud is not split across blocks);
//   * the 32 lanes of a warp evaluate the SAME point for 32 neighbouring particles; in tracking mode those
//     poses differ by centimetres, so the 32 gathers fall into a handful of 32-byte sectors instead of 32;
//   * the point tile is read with conflict-free broadcast LDS.128.
// The cloud may additionally be split into `n_splits` contiguous chunks (gridDim.y) so that small particle
// counts still fill 148 SMs; chunk partials are combined in chunk order by the caller.
#include <cmath>
#include <cstring>

#include "chain.cuh"
#include "common.cuh"

namespace amcl3d_b200
{
constexpr int kTilePoints = 512;

template <int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK)
    weight_lane_per_particle_kernel(const GridView g, const float4* __restrict__ cloud, const uint32_t n_cloud,
                                    const uint32_t chunk_len, const float* __restrict__ px, const float* __restrict__ py,
                                    const float* __restrict__ pz, const float* __restrict__ pa, const uint32_t n_poses,
                                    const RollPitch rp, float* __restrict__ part_sum, uint32_t* __restrict__ part_cnt)
{
  __shared__ float4 tile[kTilePoints];
  const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t chunk = blockIdx.y;
  const uint32_t begin = chunk * chunk_len;
  const uint32_t end = min(begin + chunk_len, n_cloud);

  bool active = i < n_poses;
  Pose3x3 P = {};
  if (active)
  {
    const float tx = px[i], ty = py[i], tz = pz[i];
    active = is_into_map(g, tx, ty, tz);  // ParticleFilter.cpp:137
    if (active)
      P = make_pose(g, rp, tx, ty, tz, pa[i]);
  }

  float sum = 0.f;
  uint32_t cnt = 0;
  for (uint32_t base = begin; base < end; base += kTilePoints)
  {
    const int len = static_cast<int>(min(static_cast<uint32_t>(kTilePoints), end - base));
    for (int j = threadIdx.x; j < len; j += BLOCK)
      tile[j] = cloud[base + j];
    __syncthreads();
    if (active)
    {
      int j = 0;
      for (; j + UNROLL <= len; j += UNROLL)
      {
        uint32_t gi[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
          const float4 p = tile[j + u];
          const float nx = transform_axis(p.x, p.y, p.z, P.r00, P.r01, P.r02, P.off_x);
          const float ny = transform_axis(p.x, p.y, p.z, P.r10, P.r11, P.r12, P.off_y);
          const float nz = transform_axis(p.x, p.y, p.z, P.r20, P.r21, P.r22, P.off_z);
          gi[u] = voxel_index(nx, ny, nz, g);
        }
        float v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          v[u] = (gi[u] != 0xFFFFFFFFu) ? __ldg(g.prob + gi[u]) : 0.f;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
          // prob >= 0 and sum starts at +0, so adding +0 for a skipped point leaves the bits unchanged
          sum = __fadd_rn(sum, v[u]);
          cnt += (gi[u] != 0xFFFFFFFFu) ? 1u : 0u;
        }
      }
      for (; j < len; ++j)
      {
        const float4 p = tile[j];
        const float nx = transform_axis(p.x, p.y, p.z, P.r00, P.r01, P.r02, P.off_x);
        const float ny = transform_axis(p.x, p.y, p.z, P.r10, P.r11, P.r12, P.off_y);
        const float nz = transform_axis(p.x, p.y, p.z, P.r20, P.r21, P.r22, P.off_z);
        const uint32_t gidx = voxel_index(nx, ny, nz, g);
        if (gidx != 0xFFFFFFFFu)
        {
          sum = __fadd_rn(sum, __ldg(g.prob + gidx));
          cnt += 1u;
        }
      }
    }
    __syncthreads();
  }
  if (i < n_poses)
  {
    part_sum[static_cast<size_t>(chunk) * n_poses + i] = sum;
    part_cnt[static_cast<size_t>(chunk) * n_poses + i] = cnt;
  }
}

// ------------------------------------------------------------------------------------------ v2: branch-free inner loop
// Same mapping and the same bits as the kernel above; what changes is how the work is issued:
//   * the z row of the rotation depends on roll/pitch only, so the loader folds it into the tile once per point
//     (tile.w = px*r20 + py*r21 + pz*r22) instead of every lane recomputing it for every particle;
//   * "0 <= v < ext" is one unsigned compare on the float's bits per axis;
//   * the voxel coordinate uses a two-float reciprocal (q + ql approximates v/res to ~2^-46), so the estimate
//     is ambiguous only within 2e-7 of an integer -- in practice only for coordinates that sit exactly on a
//     voxel face -- and the unrolled group tests ONE combined flag before taking the exact (double division)
//     path; no per-coordinate branches, no divergence on the hot path.
struct FastCoord
{
  int k;     // floor estimate
  float d;   // signed distance of the estimate from the nearest integer
};

__device__ __forceinline__ FastCoord fast_coord(float v, float inv_hi, float inv_lo)
{
  const float magic = 12582912.f;  // 1.5 * 2^23
  const float q = __fmul_rn(v, inv_hi);
  const float e = __fmaf_rn(v, inv_hi, -q);      // exact rounding error of q
  const float ql = __fmaf_rn(v, inv_lo, e);      // low-order part of v / res
  const float r = __fadd_rn(q, magic);
  const float kr = __fsub_rn(r, magic);          // nearest integer to q
  FastCoord c;
  c.d = __fadd_rn(__fsub_rn(q, kr), ql);
  c.k = (__float_as_int(r) - 0x4B400000) + (__float_as_int(c.d) >> 31);  // -1 when d < 0
  return c;
}

__device__ __noinline__ uint32_t voxel_index_exact(float nx, float ny, float nz, const GridView g)
{
  return voxel_index(nx, ny, nz, g);
}

template <int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK)
    weight_v2_kernel(const GridView g, const float4* __restrict__ cloud, const uint32_t n_cloud, const uint32_t chunk_len,
                     const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
                     const float* __restrict__ pa, const uint32_t n_poses, const RollPitch rp,
                     float* __restrict__ part_sum, uint32_t* __restrict__ part_cnt)
{
  __shared__ float4 tile[kTilePoints];
  const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t chunk = blockIdx.y;
  const uint32_t begin = chunk * chunk_len;
  const uint32_t end = min(begin + chunk_len, n_cloud);

  bool active = i < n_poses;
  Pose3x3 P = {};
  if (active)
  {
    const float tx = px[i], ty = py[i], tz = pz[i];
    active = is_into_map(g, tx, ty, tz);  // ParticleFilter.cpp:137
    if (active)
      P = make_pose(g, rp, tx, ty, tz, pa[i]);
  }
  const uint32_t ex = __float_as_uint(g.ext_up_x), ey = __float_as_uint(g.ext_up_y), ez = __float_as_uint(g.ext_up_z);
  const float inv_hi = g.inv_res_f, inv_lo = g.inv_res_lo;
  const float near_tol = 2e-7f;

  float sum = 0.f;
  uint32_t cnt = 0;
  for (uint32_t base = begin; base < end; base += kTilePoints)
  {
    const int len = static_cast<int>(min(static_cast<uint32_t>(kTilePoints), end - base));
    for (int j = threadIdx.x; j < kTilePoints; j += BLOCK)
    {
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < len)
      {
        p = cloud[base + j];
        // Grid3d.cpp:176 without the offset: (px*r20 + py*r21) + pz*r22, identical for every particle
        p.w = __fadd_rn(__fadd_rn(__fmul_rn(p.x, rp.r20), __fmul_rn(p.y, rp.r21)), __fmul_rn(p.z, rp.r22));
      }
      tile[j] = p;
    }
    __syncthreads();
    if (active)
    {
      // the tile is padded with zeros up to a multiple of UNROLL; padded slots are masked by (j + u < len)
      for (int j = 0; j < len; j += UNROLL)
      {
        uint32_t gi[UNROLL];
        bool ok[UNROLL];
        uint32_t redo = 0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
          const float4 p = tile[j + u];
          const float sx = __fadd_rn(__fadd_rn(__fmul_rn(p.x, P.r00), __fmul_rn(p.y, P.r01)), __fmul_rn(p.z, P.r02));
          const float sy = __fadd_rn(__fadd_rn(__fmul_rn(p.x, P.r10), __fmul_rn(p.y, P.r11)), __fmul_rn(p.z, P.r12));
          const float nx = static_cast<float>(__dadd_rn(static_cast<double>(sx), P.off_x));
          const float ny = static_cast<float>(__dadd_rn(static_cast<double>(sy), P.off_y));
          const float nz = static_cast<float>(__dadd_rn(static_cast<double>(p.w), P.off_z));
          const bool in = (__float_as_uint(nx) < ex) & (__float_as_uint(ny) < ey) & (__float_as_uint(nz) < ez) &
                          (j + u < len);
          const FastCoord cx = fast_coord(nx, inv_hi, inv_lo), cy = fast_coord(ny, inv_hi, inv_lo),
                          cz = fast_coord(nz, inv_hi, inv_lo);
          const float nearest = fminf(fminf(fabsf(cx.d), fabsf(cy.d)), fabsf(cz.d));
          gi[u] = static_cast<uint32_t>(cx.k) + static_cast<uint32_t>(cy.k) * g.step_y + static_cast<uint32_t>(cz.k) * g.step_z;
          ok[u] = in;
          redo |= (in && !(nearest > near_tol)) ? (1u << u) : 0u;
        }
        if (redo)
        {
          // exact path for the flagged points (a coordinate within 2e-7 of a voxel face, or q out of the magic range)
#pragma unroll
          for (int u = 0; u < UNROLL; ++u)
          {
            if (redo & (1u << u))
            {
              const float4 p = tile[j + u];
              const float nx = transform_axis(p.x, p.y, p.z, P.r00, P.r01, P.r02, P.off_x);
              const float ny = transform_axis(p.x, p.y, p.z, P.r10, P.r11, P.r12, P.off_y);
              const float nz = static_cast<float>(__dadd_rn(static_cast<double>(p.w), P.off_z));
              const uint32_t e = voxel_index_exact(nx, ny, nz, g);
              gi[u] = e;
              ok[u] = e != 0xFFFFFFFFu;
            }
          }
        }
        float v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          v[u] = ok[u] ? __ldg(g.prob + gi[u]) : 0.f;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
          sum = __fadd_rn(sum, v[u]);  // +0 for skipped points leaves the bits unchanged (prob >= 0, sum >= +0)
          cnt += ok[u] ? 1u : 0u;
        }
      }
    }
    __syncthreads();
  }
  if (i < n_poses)
  {
    part_sum[static_cast<size_t>(chunk) * n_poses + i] = sum;
    part_cnt[static_cast<size_t>(chunk) * n_poses + i] = cnt;
  }
}

RollPitch make_roll_pitch(float roll, float pitch)
{
  // Grid3d.cpp:139-142: sin/cos of the float-narrowed angles, double overloads
  RollPitch rp;
  rp.sr = std::sin(static_cast<double>(roll));
  rp.cr = std::cos(static_cast<double>(roll));
  rp.sp = std::sin(static_cast<double>(pitch));
  rp.cp = std::cos(static_cast<double>(pitch));
  // Grid3d.cpp:149: r20 = -sp; r21 = cp*sr; r22 = cp*cr  (double products rounded to float)
  rp.r20 = static_cast<float>(-rp.sp);
  rp.r21 = static_cast<float>(rp.cp * rp.sr);
  rp.r22 = static_cast<float>(rp.cp * rp.cr);
  return rp;
}

uint32_t choose_point_splits(const amcl3d_cuda_ctx* ctx, uint64_t n_poses, uint64_t n_cloud)
{
  if (ctx->opt_point_splits > 0)
  {
    uint64_t s = static_cast<uint64_t>(ctx->opt_point_splits);
    if (s > n_cloud)
      s = n_cloud ? n_cloud : 1;
    return static_cast<uint32_t>(s > 65535 ? 65535 : s);
  }
  // auto: size the grid to WHOLE WAVES.  All CTAs cost the same (the v2 loop is branch-free), so a grid that
  // spills a few CTAs into an extra wave pays for a full wave: pick the split count whose CTA total fills
  // m * (SMs * resident CTAs per SM) slots best, m = 1..4, with chunks never shorter than 64 points.
  const int block = ctx->opt_block_threads > 0 ? static_cast<int>(ctx->opt_block_threads) : 128;
  const int regs_per_thread = 64;  // ptxas: weight_v2_kernel<*, 4>
  int resident = 65536 / (regs_per_thread * block);
  resident = resident < 1 ? 1 : (resident > 16 ? 16 : resident);
  const uint64_t slots = static_cast<uint64_t>(ctx->sm_count) * resident;
  const uint64_t blocks_x = (n_poses + block - 1) / block;
  const uint64_t max_s = n_cloud / 64 ? n_cloud / 64 : 1;
  if (blocks_x >= 4 * slots)
    return 1;  // enough particle blocks for many waves: keep the cloud whole (bit-exact summation order)
  uint64_t best_s = 1;
  double best_fill = 0.0;
  for (uint64_t m = 1; m <= 4; ++m)
  {
    uint64_t s = (slots * m) / (blocks_x ? blocks_x : 1);
    if (s < 1)
      s = 1;
    if (s > max_s)
      s = max_s;
    if (s > 4096)
      s = 4096;
    const uint64_t total = blocks_x * s;
    const uint64_t waves = (total + slots - 1) / slots;
    const double fill = static_cast<double>(total) / static_cast<double>(waves * slots);
    if (fill > best_fill + 0.02)
    {
      best_fill = fill;
      best_s = s;
    }
  }
  return static_cast<uint32_t>(best_s);
}

int launch_weight_batch(amcl3d_cuda_ctx* ctx, const GridView& g, const float4* d_cloud, uint32_t n_cloud, const float* d_x,
                        const float* d_y, const float* d_z, const float* d_a, uint32_t n_poses, const RollPitch& rp,
                        float* d_part_sum, uint32_t* d_part_cnt, uint32_t n_splits)
{
  if (n_poses == 0)
    return 0;
  if (n_splits < 1)
    n_splits = 1;
  const uint32_t chunk_len = n_cloud ? (n_cloud + n_splits - 1) / n_splits : 1;
  const int block = ctx->opt_block_threads > 0 ? static_cast<int>(ctx->opt_block_threads) : 128;
  dim3 grid((n_poses + block - 1) / block, n_splits, 1);
  if (ctx->opt_kernel_timing)
    cudaEventRecord(ctx->ev_k0, ctx->stream);
#define A3D_LAUNCH_WEIGHT(KERNEL, BLK, UNR)                                                                          \
  KERNEL<BLK, UNR><<<dim3((n_poses + BLK - 1) / BLK, n_splits, 1), BLK, 0, ctx->stream>>>(                            \
      g, d_cloud, n_cloud, chunk_len, d_x, d_y, d_z, d_a, n_poses, rp, d_part_sum, d_part_cnt)
  // weight_variant: 0 = v2 (branch-free, unroll 4), 1 = v2 unroll 8, 2 = v1 (first kernel, kept for A/B profiling)
  const int variant = static_cast<int>(ctx->opt_weight_variant);
  (void)grid;
  if (variant == 2)
  {
    if (block == 64)
      A3D_LAUNCH_WEIGHT(weight_lane_per_particle_kernel, 64, 4);
    else if (block == 256)
      A3D_LAUNCH_WEIGHT(weight_lane_per_particle_kernel, 256, 4);
    else
      A3D_LAUNCH_WEIGHT(weight_lane_per_particle_kernel, 128, 4);
  }
  else if (variant == 1)
  {
    if (block == 64)
      A3D_LAUNCH_WEIGHT(weight_v2_kernel, 64, 8);
    else if (block == 256)
      A3D_LAUNCH_WEIGHT(weight_v2_kernel, 256, 8);
    else
      A3D_LAUNCH_WEIGHT(weight_v2_kernel, 128, 8);
  }
  else
  {
    if (block == 64)
      A3D_LAUNCH_WEIGHT(weight_v2_kernel, 64, 4);
    else if (block == 256)
      A3D_LAUNCH_WEIGHT(weight_v2_kernel, 256, 4);
    else
      A3D_LAUNCH_WEIGHT(weight_v2_kernel, 128, 4);
  }
#undef A3D_LAUNCH_WEIGHT
  if (ctx->opt_kernel_timing)
  {
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    ctx->ev_valid = true;
  }
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ single pose
// Thread per point: value (0 where skipped) and voxel index; the sum is then chained in cloud order.
__global__ void point_eval_kernel(const GridView g, const float4* __restrict__ cloud, const uint32_t n_cloud,
                                  const float tx, const float ty, const float tz, const float yaw, const RollPitch rp,
                                  float* __restrict__ vals, uint32_t* __restrict__ idx, uint32_t* __restrict__ count)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_cloud)
    return;
  const Pose3x3 P = make_pose(g, rp, tx, ty, tz, yaw);
  const float4 p = cloud[j];
  const float nx = transform_axis(p.x, p.y, p.z, P.r00, P.r01, P.r02, P.off_x);
  const float ny = transform_axis(p.x, p.y, p.z, P.r10, P.r11, P.r12, P.off_y);
  const float nz = transform_axis(p.x, p.y, p.z, P.r20, P.r21, P.r22, P.off_z);
  const uint32_t gi = voxel_index(nx, ny, nz, g);
  vals[j] = (gi != 0xFFFFFFFFu) ? g.prob[gi] : 0.f;
  if (idx)
    idx[j] = gi;
  if (gi != 0xFFFFFFFFu)
    atomicAdd(count, 1u);
}

// One block: chain the per-point values in cloud order, then Grid3d.cpp:198.
__global__ void __launch_bounds__(256) single_pose_finish_kernel(const float* __restrict__ vals, const uint32_t n_cloud,
                                                                 const uint32_t* __restrict__ count, float* __restrict__ out)
{
  __shared__ ChainSmem<1> sm;
  const float* const src[1] = { vals };
  float acc[1] = { 0.f };
  block_chain<1>(src, n_cloud, acc, nullptr, sm);
  if (threadIdx.x == 0)
  {
    const uint32_t n = *count;
    out[0] = (n <= 10u) ? 0.f : __fdiv_rn(acc[0], static_cast<float>(static_cast<int>(n)));
  }
}

// Combines the chunk partials of the batched kernel in chunk order and applies Grid3d.cpp:198.
__global__ void batch_finish_kernel(const float* __restrict__ part_sum, const uint32_t* __restrict__ part_cnt,
                                    const uint32_t n_poses, const uint32_t n_splits, float* __restrict__ weight,
                                    uint32_t* __restrict__ count)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_poses)
    return;
  float s = part_sum[i];
  uint32_t n = part_cnt[i];
  for (uint32_t c = 1; c < n_splits; ++c)
  {
    s = __fadd_rn(s, part_sum[static_cast<size_t>(c) * n_poses + i]);
    n += part_cnt[static_cast<size_t>(c) * n_poses + i];
  }
  weight[i] = (n <= 10u) ? 0.f : __fdiv_rn(s, static_cast<float>(static_cast<int>(n)));
  if (count)
    count[i] = n;
}
}  // namespace amcl3d_b200

using namespace amcl3d_b200;

namespace
{
struct DevBuf
{
  void* p{ nullptr };
  ~DevBuf()
  {
    if (p)
      cudaFree(p);
  }
  template <typename T>
  T* as()
  {
    return static_cast<T*>(p);
  }
};
}  // namespace

extern "C" {

int amcl3d_cuda_cloud_weight(const amcl3d_cuda_grid* grid, const float* cloud_xyzw, uint64_t n_cloud, float tx, float ty,
                             float tz, float roll, float pitch, float yaw, float* weight_out, uint32_t* n_out,
                             uint32_t* idx_out)
{
  if (!grid || !weight_out || (n_cloud && !cloud_xyzw))
    return fail(AMCL3D_CUDA_ERR_INVALID, "cloud_weight: NULL argument");
  *weight_out = 0.f;
  if (n_out)
    *n_out = 0;
  if (!grid->has_cells)
    return 0;  // Grid3d.cpp:136-137: not opened -> weight 0
  if (n_cloud == 0)
    return 0;
  if (n_cloud >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "cloud_weight: cloud too large");
  amcl3d_cuda_ctx* ctx = grid->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  DevBuf cloud, vals, idx, scal;
  A3D_CUDA_TRY(cudaMalloc(&cloud.p, n_cloud * sizeof(float4)));
  A3D_CUDA_TRY(cudaMalloc(&vals.p, n_cloud * sizeof(float)));
  if (idx_out)
    A3D_CUDA_TRY(cudaMalloc(&idx.p, n_cloud * sizeof(uint32_t)));
  A3D_CUDA_TRY(cudaMalloc(&scal.p, 16));
  A3D_CUDA_TRY(cudaMemcpyAsync(cloud.p, cloud_xyzw, n_cloud * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  A3D_CUDA_TRY(cudaMemsetAsync(scal.p, 0, 16, ctx->stream));
  const GridView g = grid->view();
  const RollPitch rp = make_roll_pitch(roll, pitch);
  const uint32_t n = static_cast<uint32_t>(n_cloud);
  point_eval_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(g, cloud.as<float4>(), n, tx, ty, tz, yaw, rp,
                                                             vals.as<float>(), idx.as<uint32_t>(),
                                                             scal.as<uint32_t>());
  single_pose_finish_kernel<<<1, 256, 0, ctx->stream>>>(vals.as<float>(), n, scal.as<uint32_t>(),
                                                        scal.as<float>() + 1);
  ctx->launches += 2;
  A3D_CUDA_TRY(cudaGetLastError());
  uint32_t host_scal[4] = { 0, 0, 0, 0 };
  A3D_CUDA_TRY(cudaMemcpyAsync(host_scal, scal.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (idx_out)
    A3D_CUDA_TRY(cudaMemcpyAsync(idx_out, idx.p, n_cloud * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  float w;
  std::memcpy(&w, &host_scal[1], 4);
  *weight_out = w;
  if (n_out)
    *n_out = host_scal[0];
  return 0;
}

int amcl3d_cuda_cloud_weight_batch(const amcl3d_cuda_grid* grid, const float* cloud_xyzw, uint64_t n_cloud,
                                   const float* poses_xyza, uint64_t n_poses, float roll, float pitch, float* weight_out,
                                   uint32_t* n_out)
{
  if (!grid || !weight_out || (n_cloud && !cloud_xyzw) || (n_poses && !poses_xyza))
    return fail(AMCL3D_CUDA_ERR_INVALID, "cloud_weight_batch: NULL argument");
  if (n_poses == 0)
    return 0;
  if (n_cloud >= 0xFFFFFFFFull || n_poses >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "cloud_weight_batch: too large");
  if (!grid->has_cells)
  {
    for (uint64_t i = 0; i < n_poses; ++i)
      weight_out[i] = 0.f;
    if (n_out)
      for (uint64_t i = 0; i < n_poses; ++i)
        n_out[i] = 0;
    return 0;
  }
  amcl3d_cuda_ctx* ctx = grid->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  const uint32_t splits = choose_point_splits(ctx, n_poses, n_cloud);
  DevBuf cloud, poses, soa, psum, pcnt, w, cnt;
  A3D_CUDA_TRY(cudaMalloc(&cloud.p, (n_cloud ? n_cloud : 1) * sizeof(float4)));
  A3D_CUDA_TRY(cudaMalloc(&soa.p, n_poses * 4 * sizeof(float)));
  A3D_CUDA_TRY(cudaMalloc(&psum.p, n_poses * splits * sizeof(float)));
  A3D_CUDA_TRY(cudaMalloc(&pcnt.p, n_poses * splits * sizeof(uint32_t)));
  A3D_CUDA_TRY(cudaMalloc(&w.p, n_poses * sizeof(float)));
  A3D_CUDA_TRY(cudaMalloc(&cnt.p, n_poses * sizeof(uint32_t)));
  if (n_cloud)
    A3D_CUDA_TRY(cudaMemcpyAsync(cloud.p, cloud_xyzw, n_cloud * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  // poses arrive AoS (x,y,z,a); the kernel wants planes: a strided 2-D copy per plane does the transpose
  for (int k = 0; k < 4; ++k)
    A3D_CUDA_TRY(cudaMemcpy2DAsync(soa.as<float>() + static_cast<size_t>(k) * n_poses, sizeof(float), poses_xyza + k,
                                   4 * sizeof(float), sizeof(float), n_poses, cudaMemcpyHostToDevice, ctx->stream));
  const GridView g = grid->view();
  const RollPitch rp = make_roll_pitch(roll, pitch);
  const uint32_t np = static_cast<uint32_t>(n_poses);
  float* s = soa.as<float>();
  A3D_TRY(launch_weight_batch(ctx, g, cloud.as<float4>(), static_cast<uint32_t>(n_cloud), s, s + n_poses, s + 2 * n_poses,
                              s + 3 * n_poses, np, rp, psum.as<float>(), pcnt.as<uint32_t>(), splits));
  batch_finish_kernel<<<(np + 255) / 256, 256, 0, ctx->stream>>>(psum.as<float>(), pcnt.as<uint32_t>(), np, splits,
                                                                w.as<float>(), cnt.as<uint32_t>());
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  A3D_CUDA_TRY(cudaMemcpyAsync(weight_out, w.p, n_poses * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (n_out)
    A3D_CUDA_TRY(cudaMemcpyAsync(n_out, cnt.p, n_poses * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
