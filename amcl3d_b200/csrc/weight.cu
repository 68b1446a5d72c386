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
#include <algorithm>
#include <cmath>
#include <cstring>

#include "chain.cuh"
#include "common.cuh"

namespace amcl3d_b200
{
constexpr int kTilePoints = 512;

// ------------------------------------------------------------------------------------------ v4: fused-scale estimate
// v3 still evaluates the reference's float chain s = (px*r00 + py*r01) + pz*r02 (5 non-fused operations per axis) on the
// hot path although only the VOXEL, not s, is needed there.  v4 estimates the voxel coordinate directly with the
// rotation rows pre-scaled by 1/res:
//     q = fma(px, R0, fma(py, R1, fma(pz, R2, F)))     R_j = fl32(r_j / res),  F = fl32(frac(offset/res) - 0.5)
// i.e. 3 FFMA per axis instead of 5 + 1, and because of the -0.5 folded into F the voxel is rint(q) (no sign fix-up):
// adding 1.5*2^23 leaves the integer in the low mantissa bits, k = bits(q + magic) + (int(offset/res) - 0x4B400000).
// Error of the estimate against the reference's real-valued coordinate T = v/res (u = 2^-24, a_j = p_j r_j,
// A = |a_0|+|a_1|+|a_2| <= |p|_2 for a unit rotation row, K = int(offset/res) >= 0 for an in-map particle):
//   reference side:  s carries <= 3uA (three products, two sums), v = fl32(s + offset) adds <= u|v|, and
//                    |v|/res <= K + A/res + 1                                     -> u (4 A/res + K + 1)
//   estimate side:   R_j rounding <= uA/res, F rounding <= u/2, the three FFMA roundings <= u(2A/res + 1 + |q|)
//                    with |q| <= A/res + 1 (q is relative to the particle's own voxel)  -> u (4 A/res + 2.5)
//   total            |q - (T - K - 0.5)| <= u (8 A/res + K + 3.6)
//   z axis: s_z is the reference's own float chain (folded into the tile), so only float(1/res), F, one FFMA and the
//   reference's rounding of v remain:                                  <= u (3 |s_z|/res + K_z + 2.6)
// With d = q - rint(q), floor(T) == rint(q) + int(offset/res) is proven whenever |d| < 0.5 - bound; everything else
// (1.25x safety, per-tile bounds on |p|_2 and |s_z|, per-particle K, separate bands for x/y and z; about 2e-4 of the
// evaluations on map S, 1.5e-3 on map L), plus the partially-outside last voxel of an axis, is recomputed with the reference's
// arithmetic verbatim by exact_address().  The operands of that path (exact
// rotation rows, double offsets) live in shared memory, not registers: the hot loop needs 12 pose registers
// instead of 24, which is what lets a fourth 256-thread CTA fit on an SM.
template <int LANES>
struct ExactPoseSmemT
{
  float r[6][LANES];      // r00 r01 r02 r10 r11 r12 (Grid3d.cpp:147-148)
  double off[3][LANES];   // Grid3d.cpp:155-157
};
using ExactPoseSmem = ExactPoseSmemT<256>;

// The reference's arithmetic for one point (Grid3d.cpp:174-189): the voxel's address in the physical layout, or
// 0xFFFFFFFF when the reference skips the point.  Returned by value so that no caller state becomes addressable.
template <bool BRICKED, int LANES = 256>
__device__ __noinline__ uint32_t exact_address(const GridView& g, const float4 p, const ExactPoseSmemT<LANES>& ep, const int t)
{
  const float nx = transform_axis(p.x, p.y, p.z, ep.r[0][t], ep.r[1][t], ep.r[2][t], ep.off[0][t]);
  const float ny = transform_axis(p.x, p.y, p.z, ep.r[3][t], ep.r[4][t], ep.r[5][t], ep.off[1][t]);
  const float nz = static_cast<float>(__dadd_rn(static_cast<double>(p.w), ep.off[2][t]));
  const bool in = nx >= 0.f && nx < g.ext_up_x && ny >= 0.f && ny < g.ext_up_y && nz >= 0.f && nz < g.ext_up_z;
  if (!in)
    return 0xFFFFFFFFu;
  const uint32_t kx = voxel_coord(nx, g), ky = voxel_coord(ny, g), kz = voxel_coord(nz, g);
  if (!(kx < g.size_x && ky < g.size_y && kz < g.size_z))
    return 0xFFFFFFFFu;
  const uint32_t gi = kx + ky * g.step_y + kz * g.step_z;  // uint32 arithmetic as in :187
  if (!(static_cast<uint64_t>(gi) < g.n_cells))
    return 0xFFFFFFFFu;
  return BRICKED ? phys_index(g, kx, ky, kz) : gi;
}

template <int BLOCK, bool BRICKED, bool PARTIAL>
__device__ __forceinline__ void
    weight_v4_body(const GridView& g, const float4* __restrict__ cloud, const uint32_t n_cloud,
                     const uint32_t chunk_len, const float* __restrict__ px, const float* __restrict__ py,
                     const float* __restrict__ pz, const float* __restrict__ pa, const uint32_t n_poses,
                     const RollPitch rp, const uint32_t partial_mask, void* __restrict__ part_sum_v,
                     uint32_t* __restrict__ part_cnt, const uint32_t chunk_first, const int carry,
                     const uint32_t* __restrict__ order, const uint32_t n_lanes)
{
  float* __restrict__ part_sum = static_cast<float*>(part_sum_v);  // v4: float partials only (acc_mode 0 / 1)
  constexpr int UNROLL = 4;
  static_assert(BLOCK <= 256, "ExactPoseSmem is sized for 256 lanes");
  __shared__ float4 tile[kTilePoints];
  __shared__ ExactPoseSmem ep;
  __shared__ int tile_rmax_bits, tile_zmax_bits;
  const int t = threadIdx.x;
  // lane -> particle: array order, or the pose-sorted scheduling permutation of order.cu (results go back to slot i)
  const uint32_t lane_i = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t i = lane_i < n_lanes ? (order ? order[lane_i] : lane_i) : n_poses;
  // this launch covers the points [chunk_first, n_cloud) of the staged cloud, gridDim.y consecutive sub-chunks of
  // chunk_len points each; sub-chunk y keeps its running sum in partial slot y (carried from launch to launch)
  const uint32_t slot = blockIdx.y;
  const uint32_t begin = min(chunk_first + blockIdx.y * chunk_len, n_cloud);
  const uint32_t end = min(begin + chunk_len, n_cloud);

  bool active = i < n_poses;
  float R00 = 0.f, R01 = 0.f, R02 = 0.f, R10 = 0.f, R11 = 0.f, R12 = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
  int cx = 0, cy = 0, cz = 0;
  if (active)
  {
    const float tx = px[i], ty = py[i], tz = pz[i];
    active = is_into_map(g, tx, ty, tz);  // ParticleFilter.cpp:137
    if (active)
    {
      const Pose3x3 e = make_pose(g, rp, tx, ty, tz, pa[i]);
      ep.r[0][t] = e.r00;
      ep.r[1][t] = e.r01;
      ep.r[2][t] = e.r02;
      ep.r[3][t] = e.r10;
      ep.r[4][t] = e.r11;
      ep.r[5][t] = e.r12;
      ep.off[0][t] = e.off_x;
      ep.off[1][t] = e.off_y;
      ep.off[2][t] = e.off_z;
      const double inv = 1.0 / g.res;
      R00 = static_cast<float>(e.r00 * inv);
      R01 = static_cast<float>(e.r01 * inv);
      R02 = static_cast<float>(e.r02 * inv);
      R10 = static_cast<float>(e.r10 * inv);
      R11 = static_cast<float>(e.r11 * inv);
      R12 = static_cast<float>(e.r12 * inv);
      const double dx = e.off_x * inv, dy = e.off_y * inv, dz = e.off_z * inv;
      const double ix = floor(dx), iy = floor(dy), iz = floor(dz);
      fx = static_cast<float>((dx - ix) - 0.5);
      fy = static_cast<float>((dy - iy) - 0.5);
      fz = static_cast<float>((dz - iz) - 0.5);
      cx = static_cast<int>(ix) - 0x4B400000;
      cy = static_cast<int>(iy) - 0x4B400000;
      cz = static_cast<int>(iz) - 0x4B400000;
    }
  }
  const float inv_f = g.inv_res_f;
  const float* __restrict__ prob = g.prob;
  const uint32_t sx = g.size_x, sy = g.size_y, sz = g.size_z;
  const uint32_t step_y = g.step_y, step_z = g.step_z, zero_index = g.zero_index;
  // PARTIAL: index of the last voxel of an axis when it sticks out of the metric bounds (else never matched)
  const uint32_t lastx = (partial_mask & 1u) ? sx - 1u : 0xFFFFFFFFu, lasty = (partial_mask & 2u) ? sy - 1u : 0xFFFFFFFFu,
                 lastz = (partial_mask & 4u) ? sz - 1u : 0xFFFFFFFFu;
  // bricked address = X(kx) + Y(ky) + Z(kz), each a multiply-add of k and (k & ~(brick-1)) with launch constants
  const uint32_t bsh = g.brick_shift, bmask = ~((1u << bsh) - 1u);
  const uint32_t bcx = (1u << (2 * bsh)) - 1u;
  const uint32_t bcy = (g.nbx << (2 * bsh)) - (1u << bsh);
  const uint32_t bcz = (g.nbx * g.nby - 1u) << (2 * bsh);
  auto address = [&](const uint32_t kx, const uint32_t ky, const uint32_t kz) -> uint32_t {
    if (!BRICKED)
      return kx + ky * step_y + kz * step_z;
    uint32_t a = kx + (kx & bmask) * bcx;
    a += (ky << bsh) + (ky & bmask) * bcy;
    a += (kz << (2 * bsh)) + (kz & bmask) * bcz;
    return a;
  };
  // estimate of one point: voxel address, "inside the grid", and the largest |d| of the three axes -- the z distance
  // shifted by z_shift = safe_xy - safe_z (the z band is several times narrower), so that ONE comparison against the
  // x/y band serves all three axes
  float z_shift = 0.f;
  auto estimate = [&](const float4 p, uint32_t& addr, bool& in, float& far, bool& last) {
    const float magic = 12582912.f;  // 1.5 * 2^23
    const float qx = __fmaf_rn(p.x, R00, __fmaf_rn(p.y, R01, __fmaf_rn(p.z, R02, fx)));
    const float qy = __fmaf_rn(p.x, R10, __fmaf_rn(p.y, R11, __fmaf_rn(p.z, R12, fy)));
    const float qz = __fmaf_rn(p.w, inv_f, fz);
    const float rx = __fadd_rn(qx, magic), ry = __fadd_rn(qy, magic), rz = __fadd_rn(qz, magic);
    const float dx = __fsub_rn(qx, __fsub_rn(rx, magic)), dy = __fsub_rn(qy, __fsub_rn(ry, magic)),
                dz = __fsub_rn(qz, __fsub_rn(rz, magic));
    const uint32_t kx = static_cast<uint32_t>(__float_as_int(rx) + cx), ky = static_cast<uint32_t>(__float_as_int(ry) + cy),
                   kz = static_cast<uint32_t>(__float_as_int(rz) + cz);
    in = (kx < sx) & (ky < sy) & (kz < sz);
    far = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), __fadd_rn(fabsf(dz), z_shift));
    last = PARTIAL ? ((kx == lastx) | (ky == lasty) | (kz == lastz)) : false;
    addr = address(kx, ky, kz);
  };

  float sum = 0.f;
  uint32_t cnt = 0;
  if (carry && i < n_poses)
  {
    sum = part_sum[static_cast<size_t>(slot) * n_poses + i];
    cnt = part_cnt[static_cast<size_t>(slot) * n_poses + i];
  }
  for (uint32_t base = begin; base < end; base += kTilePoints)
  {
    const int len = static_cast<int>(min(static_cast<uint32_t>(kTilePoints), end - base));
    if (threadIdx.x == 0)
    {
      tile_rmax_bits = 0;
      tile_zmax_bits = 0;
    }
    __syncthreads();
    float my_r = 0.f, my_z = 0.f;
    for (int j = threadIdx.x; j < len; j += BLOCK)
    {
      float4 p = cloud[base + j];
      // |p|_2 bounds |px r0| + |py r1| + |pz r2| for any (unit) rotation row; rounded up generously
      my_r = fmaxf(my_r, 1.0001f * sqrtf(p.x * p.x + p.y * p.y + p.z * p.z));
      if (!(fabsf(p.x) + fabsf(p.y) + fabsf(p.z) < 1e30f))
        my_r = INFINITY;  // NaN / infinite point in the tile: fmaxf would drop it -> verify the whole tile
      // Grid3d.cpp:176 without the offset: (px*r20 + py*r21) + pz*r22, identical for every particle
      p.w = __fadd_rn(__fadd_rn(__fmul_rn(p.x, rp.r20), __fmul_rn(p.y, rp.r21)), __fmul_rn(p.z, rp.r22));
      my_z = fmaxf(my_z, fabsf(p.w));
      tile[j] = p;
    }
    for (int o = 16; o > 0; o >>= 1)
    {
      my_r = fmaxf(my_r, __shfl_xor_sync(0xffffffffu, my_r, o));
      my_z = fmaxf(my_z, __shfl_xor_sync(0xffffffffu, my_z, o));
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicMax(&tile_rmax_bits, __float_as_int(my_r));  // non-negative floats order like their bit patterns
      atomicMax(&tile_zmax_bits, __float_as_int(my_z));
    }
    __syncthreads();
    const float rmax = __int_as_float(tile_rmax_bits), zmax = __int_as_float(tile_zmax_bits);
    // proven-safe bands (header comment), 1.25x safety; outside the magic-number range of the estimate (or with
    // non-finite points in the tile) everything is verified
    // per particle: K of the x/y axes and of z, recovered from the integer offsets already held for the hot loop
    const float k_xy = static_cast<float>(max(max(cx, cy) + 0x4B400000, 0)), k_z = static_cast<float>(max(cz + 0x4B400000, 0));
    float safe = 0.5f - 1.25f * 5.9604645e-8f * ((8.f * rmax) * inv_f + k_xy + 4.f);
    const float safe_z = 0.5f - 1.25f * 5.9604645e-8f * ((3.f * zmax) * inv_f + k_z + 3.f);
    z_shift = safe - safe_z;
    if (!(rmax * inv_f < 2.0e6f) || !(zmax * inv_f < 2.0e6f) || !(safe > 0.f))
    {
      safe = -1.f;
      z_shift = 0.f;
    }
    if (active)
    {
      // A skipped point gathers the always-zero padding cell (adding +0 leaves the running sum's bits unchanged:
      // prob >= 0 and the sum starts at +0), and the contributing-point count is taken from the estimate right away
      // and corrected on the verification path -- so nothing but the four addresses stays live across the branch.
      const int full = len - (len % UNROLL);
      for (int j = 0; j < full; j += UNROLL)
      {
        uint32_t gi[UNROLL];
        // Small maps: the verification path is rare, so the hot loop only keeps the largest distance of the group and
        // the path re-runs the estimate to find the flagged points.  Large maps (bricked layout, wider bands): one
        // group in four takes the path, so the hot loop records a flag bit per point instead (one more instruction
        // per point, no re-estimate).
        constexpr bool kFlagBits = BRICKED;
        float far_all = 0.f;
        uint32_t flags = 0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
          float far;
          bool last, in;
          uint32_t a;
          estimate(tile[j + u], a, in, far, last);
          gi[u] = in ? a : zero_index;
          cnt += in ? 1u : 0u;
          if (kFlagBits)
            flags |= (!(far < safe) || last) ? (1u << u) : 0u;
          else
          {
            far_all = fmaxf(far_all, far);
            flags |= last ? 1u : 0u;
          }
        }
        if (kFlagBits ? (flags != 0u) : (!(far_all < safe) || flags != 0u))
        {
          // verification path: the flagged points of this group get the reference's arithmetic verbatim
#pragma unroll
          for (int u = 0; u < UNROLL; ++u)
          {
            bool flagged;
            if (kFlagBits)
              flagged = (flags >> u) & 1u;
            else
            {
              float far;
              bool last, in;
              uint32_t a;
              estimate(tile[j + u], a, in, far, last);
              flagged = !(far < safe) || last;
            }
            if (flagged)
            {
              const uint32_t a = exact_address<BRICKED>(g, tile[j + u], ep, t);
              cnt += (a != 0xFFFFFFFFu ? 1u : 0u) - (gi[u] != zero_index ? 1u : 0u);
              gi[u] = a != 0xFFFFFFFFu ? a : zero_index;
            }
          }
        }
        float v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          v[u] = __ldg(prob + gi[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          sum = __fadd_rn(sum, v[u]);
      }
      for (int j = full; j < len; ++j)  // ragged end of the chunk
      {
        uint32_t gi;
        bool ok, last;
        float far;
        estimate(tile[j], gi, ok, far, last);
        if (!(far < safe) || last)
        {
          gi = exact_address<BRICKED>(g, tile[j], ep, t);
          ok = gi != 0xFFFFFFFFu;
        }
        if (ok)
        {
          sum = __fadd_rn(sum, __ldg(prob + gi));
          cnt += 1u;
        }
      }
    }
    __syncthreads();
  }
  if (i < n_poses)
  {
    part_sum[static_cast<size_t>(slot) * n_poses + i] = sum;
    part_cnt[static_cast<size_t>(slot) * n_poses + i] = cnt;
  }
}

}  // namespace amcl3d_b200

#include "weight_v5.cuh"
#include "weight_ordered.cuh"

namespace amcl3d_b200
{
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

// CTA width: 256 lanes measured best at 10 k particles (fewer tile loads per evaluation); small particle sets use
// narrower CTAs so that particles x point-chunks still yields at least a few CTAs per SM.
// one_piece: every particle walks the whole cloud in ONE CTA (reference summation order): the particle blocks alone
// have to fill the GPU, so they shrink down to one warp while there are fewer particles than resident lanes.
static int pick_block_threads(const amcl3d_cuda_ctx* ctx, uint64_t n_poses, bool one_piece = false)
{
  if (ctx->opt_block_threads == 64 || ctx->opt_block_threads == 128 || ctx->opt_block_threads == 256)
    return static_cast<int>(ctx->opt_block_threads);
  if (one_piece)
  {
    const uint64_t lanes = static_cast<uint64_t>(ctx->sm_count) * 1024;  // resident lanes at 64 registers
    return n_poses * 4 <= lanes ? 64 : (n_poses <= 2 * lanes ? 128 : 256);  // (measured at 131 072 and 262 144 particles: 128 best)
  }
  return n_poses <= 2048 ? 64 : (n_poses <= 8192 ? 128 : 256);
}

// weight_variant: 0 = v5 (packed fp32 pairs, default), 4 = v4 (the scalar generation, kept for A/B profiling and as a
// parity cross-check; it cannot store the value matrix).
using WeightKernel = void (*)(const GridView, const float4*, uint32_t, uint32_t, const float*, const float*, const float*,
                              const float*, uint32_t, const RollPitch, uint32_t, void*, uint32_t*, uint32_t, int,
                              const uint32_t*, float*, uint64_t, uint32_t);

// v4 behind the common signature
template <int BLOCK, bool BRICKED, bool PARTIAL>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK)
    weight_v4_entry(const __grid_constant__ GridView g, const float4* __restrict__ cloud, const uint32_t n_cloud,
                    const uint32_t chunk_len, const float* __restrict__ px, const float* __restrict__ py,
                    const float* __restrict__ pz, const float* __restrict__ pa, const uint32_t n_poses, const RollPitch rp,
                    const uint32_t partial_mask, void* __restrict__ part_sum, uint32_t* __restrict__ part_cnt,
                    const uint32_t chunk_first, const int acc_mode, const uint32_t* __restrict__ order, float*, uint64_t,
                    const uint32_t n_lanes)
{
  weight_v4_body<BLOCK, BRICKED, PARTIAL>(g, cloud, n_cloud, chunk_len, px, py, pz, pa, n_poses, rp, partial_mask, part_sum,
                                         part_cnt, chunk_first, acc_mode, order, n_lanes);
}

template <bool BRICKED, bool PARTIAL, bool STORE, int MODE>
static WeightKernel pick_weight_kernel_b(int block)
{
  return block <= 64 ? weight_v5_kernel<64, BRICKED, PARTIAL, STORE, MODE> :
                       (block == 256 ? weight_v5_kernel<256, BRICKED, PARTIAL, STORE, MODE> :
                                       weight_v5_kernel<128, BRICKED, PARTIAL, STORE, MODE>);
}

// variant: 0 = v5 (software-pipelined gathers, default), 4 = v4 (scalar generation; parity cross-check in the tests).
// (The un-pipelined and 80-register forms of v5 -- MODE 1 / 2 in weight_v5.cuh, measured slower in round 2 -- are no
// longer instantiated: they tripled the size of the library for nothing.)
template <bool BRICKED, bool PARTIAL>
static WeightKernel pick_weight_kernel_l(int variant, int block, bool store)
{
  if (variant == 4)
    return block <= 64 ? weight_v4_entry<64, BRICKED, PARTIAL> :
                         (block == 256 ? weight_v4_entry<256, BRICKED, PARTIAL> : weight_v4_entry<128, BRICKED, PARTIAL>);
  return store ? pick_weight_kernel_b<BRICKED, PARTIAL, true, 0>(block) : pick_weight_kernel_b<BRICKED, PARTIAL, false, 0>(block);
}

static WeightKernel pick_weight_kernel(int variant, int block, bool bricked, bool partial, bool store = false)
{
  if (bricked)
    return partial ? pick_weight_kernel_l<true, true>(variant, block, store) :
                     pick_weight_kernel_l<true, false>(variant, block, store);
  return partial ? pick_weight_kernel_l<false, true>(variant, block, store) :
                   pick_weight_kernel_l<false, false>(variant, block, store);
}

// CTAs of the weighting kernel one SM holds (register / shared-memory limited), asked from the runtime once per kernel.
static int resident_ctas(int variant, int block, bool bricked)
{
  static int cache[8][3][2];
  const int vi = variant == 4 ? 4 : 0;
  const int bi = block <= 64 ? 0 : (block == 256 ? 2 : 1);
  int& c = cache[vi][bi][bricked ? 1 : 0];
  if (c == 0)
  {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pick_weight_kernel(vi, block, bricked, false), block, 0) !=
            cudaSuccess ||
        n < 1)
    {
      cudaGetLastError();
      n = 65536 / (64 * block) > 0 ? 65536 / (64 * block) : 1;
    }
    c = n;
  }
  return c;
}

// Points per sequential chunk launch (0 = one launch over the whole cloud).  Option "weight_chunk_points", else on
// grids larger than L2: 512 points when the particle set alone gives several full waves per launch (measured best at
// 1 M particles), and about 2^30 / particles -- up to 8192 -- for smaller (sharded) sets, whose launches would
// otherwise be too short to amortise their prologue and wave tail (measured at 131 072 particles: 8192-point launches
// 14.2 ms, 512-point 18.1 ms; at 262 144: 512-point launches 33.2 ms).
static uint64_t auto_chunk_points(const amcl3d_cuda_ctx* ctx, uint64_t n_poses, bool large_grid)
{
  if (ctx->opt_chunk_points > 0)
    return static_cast<uint64_t>(ctx->opt_chunk_points);
  if (!large_grid)
    return 0;
  const uint64_t full = static_cast<uint64_t>(ctx->sm_count) * 1024;  // particles that fill every SM with 256-lane CTAs
  if (n_poses >= 4 * full)
    return 512;  // several full waves per launch
  uint64_t c = (1ull << 30) / (n_poses ? n_poses : 1);
  c = c / 512 * 512;
  return c < 512 ? 512 : (c > 8192 ? 8192 : c);
}

uint32_t choose_point_splits(const amcl3d_cuda_ctx* ctx, uint64_t n_poses, uint64_t n_cloud, bool large_grid, bool fast)
{
  if (ctx->opt_point_splits > 0)
  {
    uint64_t s = static_cast<uint64_t>(ctx->opt_point_splits);
    if (s > n_cloud)
      s = n_cloud ? n_cloud : 1;
    return static_cast<uint32_t>(s > 65535 ? 65535 : s);
  }
  // auto: size the grid to WHOLE WAVES.  All CTAs cost the same (the loop is branch-free), so a grid that
  // spills a few CTAs into an extra wave pays for a full wave: pick the split count whose CTA total fills
  // m * (SMs * resident CTAs per SM) slots best, m = 1..4, with chunks never shorter than 64 points.
  const int block = pick_block_threads(ctx, n_poses);
  if (ctx->opt_point_splits == 0 && ctx->opt_reference_order && !fast)
    return 1;  // the reference's summation order: one float chain per particle over the whole cloud
  const int resident = resident_ctas(static_cast<int>(ctx->opt_weight_variant), block, large_grid);
  const uint64_t slots = static_cast<uint64_t>(ctx->sm_count) * resident;
  const uint64_t blocks_x = (n_poses + block - 1) / block;
  const uint64_t max_s = n_cloud / 64 ? n_cloud / 64 : 1;
  if (blocks_x >= 4 * slots)
    return 1;  // enough particle blocks for many waves: keep the cloud whole (bit-exact summation order)
  if (large_grid)
  {
    // sequential chunk launches (launch_weight_batch) with fewer particle blocks than the GPU holds -- a sharded
    // particle set: longer chunks (auto_chunk_points) divided into 512-point sub-chunks, one CTA per particle block
    // and sub-chunk, if that fills at least half of the GPU; smaller sets take the generic single-launch split below
    const uint64_t chunk_pts = auto_chunk_points(ctx, n_poses, true);
    const uint64_t s = chunk_pts / 512 ? chunk_pts / 512 : 1;
    if (blocks_x * s * 2 >= slots && n_cloud > chunk_pts)
      return static_cast<uint32_t>(s);
  }
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

// Launches the weighting kernel(s).  `exact_order`: the caller wants every particle's sum to be ONE float chain in the
// order of `d_cloud` (the reference's own sum when that is the caller's cloud order and n_splits == 1).  Otherwise the
// partials of sequential chunk launches are accumulated in double.  *partial_kind_out: what d_part_sum holds afterwards
// (0 = float [n_splits][n_poses], 1 = double [n_splits][n_poses]); the buffer must hold 8 bytes per entry.
int launch_weight_batch(amcl3d_cuda_ctx* ctx, const GridView& g, const float4* d_cloud, uint32_t n_cloud, const float* d_x,
                        const float* d_y, const float* d_z, const float* d_a, uint32_t n_poses, const RollPitch& rp,
                        void* d_part_sum, uint32_t* d_part_cnt, uint32_t n_splits, const uint32_t* d_order,
                        bool exact_order, int* partial_kind_out, float* d_vals, uint64_t vals_stride, uint32_t n_lanes)
{
  // n_lanes: lanes to schedule (0 = one per particle).  Less than n_poses when this GPU weighs a slice of a sharded set:
  // lane l then takes particle d_order[l] of the n_poses-particle index space.
  if (partial_kind_out)
    *partial_kind_out = 0;
  if (n_lanes == 0)
    n_lanes = n_poses;
  if (n_poses == 0)
    return 0;
  if (n_splits < 1)
    n_splits = 1;
  uint32_t chunk_len = n_cloud ? (n_cloud + n_splits - 1) / n_splits : 1;
  const int block = pick_block_threads(ctx, n_lanes, n_splits == 1);
  int variant = static_cast<int>(ctx->opt_weight_variant);
  if (variant != 4)
    variant = 0;
  const uint32_t blocks_x = (n_lanes + block - 1) / block;
  // Sequential chunk launches, the large-map regime: each launch walks ONE chunk of (Morton-neighbouring) points for
  // ALL particles, so the grid footprint that is live in L2 at any time is one chunk's.  With enough particle blocks
  // to fill the GPU the chunk is not split (n_splits == 1).  With fewer particle blocks (a sharded particle set) the
  // chunk's points are divided over n_splits CTAs per particle block; sub-chunk y keeps its own partial from launch
  // to launch.  Chunk length: option "weight_chunk_points", default 512 on bricked (larger-than-L2) grids.
  uint32_t seq_chunks = 1, launch_pts = n_cloud;
  {
    uint64_t chunk_pts = auto_chunk_points(ctx, n_lanes, g.brick_shift != 0);
    // one-piece walks (reference order): launches of 2048 points keep the CTAs of a wave on the same stretch of the
    // cloud, which is what makes their gathers share L2 lines, and still amortise the CTA prologue (measured with the
    // z-weighted particle schedule at 1 048 576 / 131 072 particles: 512-point launches 74.4 / 11.9 ms, 2048-point
    // launches 72.7 / 11.3 ms)
    if (n_splits == 1 && ctx->opt_chunk_points <= 0 && chunk_pts > 0)
      chunk_pts = 2048;
    const uint64_t slots = static_cast<uint64_t>(ctx->sm_count) * resident_ctas(variant, block, g.brick_shift != 0);
    const bool fills = static_cast<uint64_t>(blocks_x) * n_splits * 2 >= slots;
    if (chunk_pts > 0 && n_cloud > chunk_pts && fills && (n_splits == 1 || chunk_pts / n_splits >= 32))
    {
      launch_pts = static_cast<uint32_t>(chunk_pts);
      chunk_len = static_cast<uint32_t>((chunk_pts + n_splits - 1) / n_splits);
      seq_chunks = static_cast<uint32_t>((n_cloud + chunk_pts - 1) / chunk_pts);
    }
  }
  // float chain carried from launch to launch (bit-exact order), or double accumulators across launches
  const bool use_double = seq_chunks > 1 && !exact_order && variant != 4;
  if (partial_kind_out)
    *partial_kind_out = use_double ? 1 : 0;
  if (ctx->opt_kernel_timing)
    cudaEventRecord(ctx->ev_k0, ctx->stream);
  // bit a set: the last voxel of axis a sticks out of the metric bounds (ext/res is not an integer)
  uint32_t partial_mask = 0;
  const double ext[3] = { g.ext_x, g.ext_y, g.ext_z };
  const uint32_t dims[3] = { g.size_x, g.size_y, g.size_z };
  for (int a = 0; a < 3; ++a)
    if (static_cast<double>(dims[a]) * g.res - ext[a] > 1e-7 * g.res)
      partial_mask |= 1u << a;
  if (d_vals && variant == 4)
    variant = 0;  // only v5 stores the value matrix
  const WeightKernel kernel = pick_weight_kernel(variant, block, g.brick_shift != 0, partial_mask != 0, d_vals != nullptr);
  for (uint32_t seq = 0; seq < seq_chunks; ++seq)
  {
    const uint32_t first = seq * launch_pts;
    const uint32_t last = static_cast<uint32_t>(std::min<uint64_t>(n_cloud, static_cast<uint64_t>(first) + launch_pts));
    const int acc_mode = use_double ? (seq > 0 ? 3 : 2) : (seq > 0 ? 1 : 0);
    kernel<<<dim3(blocks_x, n_splits, 1), block, 0, ctx->stream>>>(g, d_cloud, last, chunk_len, d_x, d_y, d_z, d_a, n_poses,
                                                                  rp, partial_mask, d_part_sum, d_part_cnt, first, acc_mode,
                                                                  d_order, d_vals, vals_stride, n_lanes);
    ctx->launches++;
  }
  if (ctx->opt_kernel_timing)
  {
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    ctx->ev_valid = true;
  }
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
  vals[j] = (gi != 0xFFFFFFFFu) ? g.prob[logical_to_phys(g, gi)] : 0.f;
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

// Combines the partials of the batched kernel (cloud_weight_from_partials) and applies Grid3d.cpp:198.
__global__ void batch_finish_kernel(const void* __restrict__ part_sum, const uint32_t* __restrict__ part_cnt,
                                    const uint32_t n_poses, const uint32_t n_splits, const int kind,
                                    float* __restrict__ weight, uint32_t* __restrict__ count)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_poses)
    return;
  uint32_t n;
  weight[i] = cloud_weight_from_partials(part_sum, part_cnt, n_poses, n_splits, i, kind, &n);
  if (count)
    count[i] = n;
}
// ------------------------------------------------------------------------------------------ replay in the caller's order
// Second half of "gather anywhere, add in order": lane = scheduled particle; the values the STORE kernel left in
// vals[lane / 32][point position][lane % 32] (stride = points per warp tile) are added in the CALLER's cloud order
// (pos_of[j] = position of caller point j in the staged, possibly Morton-ordered cloud; NULL = identity) with plain float
// adds: Grid3d.cpp:191 bit for bit (a skipped point stored +0, which leaves the sum's bits unchanged).  The loads of the
// next batch are in flight while the dependent add chain of the current one runs.
constexpr int kReplayThreads = 256;                       // warp 0 adds, warps 1..7 load
constexpr int kReplayLoaders = kReplayThreads / 32 - 1;
constexpr int kReplayDepth = 16;                          // loads in flight per loader warp
constexpr int kReplayStage = kReplayLoaders * kReplayDepth;  // points per stage

// One CTA per warp tile of the value matrix (32 scheduled particles).  Seven loader warps stream the tile's values, a
// stage of 112 points at a time and in the CALLER's order, into a double-buffered shared-memory tile [point][lane]
// (every load is one coalesced 128-byte line, 16 in flight per warp); warp 0 -- one lane per particle -- consumes the
// stages with plain dependent float adds.  The add chain (4 cycles per point) is the critical path for small particle
// sets, HBM bandwidth for large ones.
template <bool PERM>
__global__ void __launch_bounds__(kReplayThreads)
    replay_sum_kernel(const float* __restrict__ vals, const uint64_t stride, const uint32_t* __restrict__ pos_of,
                      const uint32_t n_cloud, const uint32_t n_poses, const uint32_t* __restrict__ order,
                      const uint32_t* __restrict__ part_cnt, const uint32_t n_splits, float* __restrict__ out_sum,
                      uint32_t* __restrict__ out_cnt, const uint32_t n_lanes)
{
  __shared__ float tile[2][kReplayStage][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lane_i = blockIdx.x * 32u + lane;
  const float* col = vals + static_cast<uint64_t>(blockIdx.x) * stride * 32u + lane;
  const uint32_t n_stages = (n_cloud + kReplayStage - 1) / kReplayStage;
  auto load_stage = [&](const uint32_t s) {
    // loader warp (warp - 1) takes the points  s*stage + (warp-1) + kReplayLoaders*k
    const uint32_t j0 = s * kReplayStage + static_cast<uint32_t>(warp - 1);
    uint32_t p[kReplayDepth];
#pragma unroll
    for (int k = 0; k < kReplayDepth; ++k)
    {
      const uint32_t j = j0 + kReplayLoaders * k;
      p[k] = PERM ? (j < n_cloud ? __ldg(pos_of + j) : 0u) : (j < n_cloud ? j : 0u);
    }
    float v[kReplayDepth];
#pragma unroll
    for (int k = 0; k < kReplayDepth; ++k)
      v[k] = __ldcs(col + static_cast<uint64_t>(p[k]) * 32u);
#pragma unroll
    for (int k = 0; k < kReplayDepth; ++k)
      tile[s & 1][(warp - 1) + kReplayLoaders * k][lane] = (j0 + kReplayLoaders * k < n_cloud) ? v[k] : 0.f;
  };
  if (warp > 0 && n_stages > 0)
    load_stage(0);
  __syncthreads();
  float sum = 0.f;
  for (uint32_t s = 0; s < n_stages; ++s)
  {
    if (warp > 0)
    {
      if (s + 1 < n_stages)
        load_stage(s + 1);
    }
    else
    {
      // points past the end of the cloud were stored as +0: adding them changes nothing
#pragma unroll 16
      for (int k = 0; k < kReplayStage; ++k)
        sum = __fadd_rn(sum, tile[s & 1][k][lane]);
    }
    __syncthreads();
  }
  if (warp == 0 && lane_i < n_lanes)
  {
    const uint32_t i = order ? order[lane_i] : lane_i;
    uint32_t cnt = 0;
    for (uint32_t k = 0; k < n_splits; ++k)
      cnt += part_cnt[static_cast<size_t>(k) * n_poses + i];
    // particles outside the map (and lanes that never ran) left nothing but garbage in the matrix
    out_sum[i] = cnt ? sum : 0.f;
    out_cnt[i] = cnt;
  }
}

// caller index -> position in the Morton-ordered cloud (cloud.cu leaves the original index in .w)
__global__ void cloud_pos_kernel(const float4* __restrict__ sorted, const uint32_t n, uint32_t* __restrict__ pos_of)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n)
  {
    const uint32_t orig = __float_as_uint(sorted[p].w);
    if (orig < n)
      pos_of[orig] = p;
  }
}

int launch_replay_sum(amcl3d_cuda_ctx* ctx, const float* d_vals, uint64_t stride, const uint32_t* d_pos_of, uint32_t n_cloud,
                      uint32_t n_poses, const uint32_t* d_order, const uint32_t* d_part_cnt, uint32_t n_splits,
                      float* d_out_sum, uint32_t* d_out_cnt, uint32_t n_lanes)
{
  if (n_lanes == 0)
    n_lanes = n_poses;
  if (n_poses == 0)
    return 0;
  const unsigned tiles = (n_lanes + 31) / 32;
  if (d_pos_of)
    replay_sum_kernel<true><<<tiles, kReplayThreads, 0, ctx->stream>>>(d_vals, stride, d_pos_of, n_cloud, n_poses, d_order,
                                                                      d_part_cnt, n_splits, d_out_sum, d_out_cnt, n_lanes);
  else
    replay_sum_kernel<false><<<tiles, kReplayThreads, 0, ctx->stream>>>(d_vals, stride, d_pos_of, n_cloud, n_poses,
                                                                       d_order, d_part_cnt, n_splits, d_out_sum, d_out_cnt,
                                                                       n_lanes);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ ordered sums in one kernel
// Option "ordered_mode": 0 = auto, 1 = off, 2 = on whenever the layout allows it.  Auto: linear (L2-resident) grids and
// particle sets that give every SM at least one 32-particle group but are too small for the one-lane-per-particle walk.
bool weight_ordered_applies(const amcl3d_cuda_ctx* ctx, const GridView& g, uint64_t n_lanes, uint64_t n_cloud)
{
  if (ctx->opt_ordered == 1 || n_lanes == 0 || n_cloud == 0)
    return false;
  if (ctx->opt_ordered == 2)
    return true;
  if (g.brick_shift != 0)
    return false;
  const uint64_t groups = (n_lanes + 31) / 32;
  return groups >= static_cast<uint64_t>(ctx->sm_count) && n_lanes < 2ull * ctx->sm_count * 1024;
}

template <int GW>
static void launch_weight_ordered_t(cudaStream_t stream, unsigned groups, bool bricked, bool partial, const GridView& g,
                                    const float4* d_cloud, uint32_t n_cloud, const float* d_x, const float* d_y,
                                    const float* d_z, const float* d_a, uint32_t n_poses, const RollPitch& rp,
                                    uint32_t partial_mask, float* d_out_sum, uint32_t* d_out_cnt, const uint32_t* d_order,
                                    uint32_t n_lanes)
{
  constexpr int T = 32 * (GW + 1);
#define A3D_ORD(B, P)                                                                                                  \
  weight_ordered_kernel<GW, B, P><<<groups, T, 0, stream>>>(g, d_cloud, n_cloud, d_x, d_y, d_z, d_a, n_poses, rp,      \
                                                            partial_mask, d_out_sum, d_out_cnt, d_order, n_lanes)
  if (bricked)
  {
    if (partial)
      A3D_ORD(true, true);
    else
      A3D_ORD(true, false);
  }
  else
  {
    if (partial)
      A3D_ORD(false, true);
    else
      A3D_ORD(false, false);
  }
#undef A3D_ORD
}

int launch_weight_ordered(amcl3d_cuda_ctx* ctx, const GridView& g, const float4* d_cloud, uint32_t n_cloud, const float* d_x,
                          const float* d_y, const float* d_z, const float* d_a, uint32_t n_poses, const RollPitch& rp,
                          float* d_out_sum, uint32_t* d_out_cnt, const uint32_t* d_order, uint32_t n_lanes)
{
  if (n_lanes == 0)
    n_lanes = n_poses;
  if (n_poses == 0)
    return 0;
  uint32_t partial_mask = 0;
  const double ext[3] = { g.ext_x, g.ext_y, g.ext_z };
  const uint32_t dims[3] = { g.size_x, g.size_y, g.size_z };
  for (int a = 0; a < 3; ++a)
    if (static_cast<double>(dims[a]) * g.res - ext[a] > 1e-7 * g.res)
      partial_mask |= 1u << a;
  const unsigned groups = (n_lanes + 31) / 32;
  // gatherer warps per group: 8 (three 288-thread CTAs per SM) while that keeps every group resident at once or gives
  // several waves; 4 (six 160-thread CTAs per SM) in between, where eight would leave a nearly empty second wave
  const uint64_t slots8 = static_cast<uint64_t>(ctx->sm_count) * 3, slots4 = static_cast<uint64_t>(ctx->sm_count) * 6;
  int gw = 8;
  if (ctx->opt_block_threads == 160)
    gw = 4;
  else if (ctx->opt_block_threads != 288 && groups > slots8 && groups <= slots4)
    gw = 4;
  if (ctx->opt_kernel_timing)
    cudaEventRecord(ctx->ev_k0, ctx->stream);
  if (gw == 4)
    launch_weight_ordered_t<4>(ctx->stream, groups, g.brick_shift != 0, partial_mask != 0, g, d_cloud, n_cloud, d_x, d_y, d_z,
                               d_a, n_poses, rp, partial_mask, d_out_sum, d_out_cnt, d_order, n_lanes);
  else
    launch_weight_ordered_t<8>(ctx->stream, groups, g.brick_shift != 0, partial_mask != 0, g, d_cloud, n_cloud, d_x, d_y, d_z,
                               d_a, n_poses, rp, partial_mask, d_out_sum, d_out_cnt, d_order, n_lanes);
  ctx->launches++;
  if (ctx->opt_kernel_timing)
  {
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    ctx->ev_valid = true;
  }
  A3D_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_cloud_pos(amcl3d_cuda_ctx* ctx, const float4* d_sorted, uint32_t n, uint32_t* d_pos_of)
{
  if (n == 0)
    return 0;
  cloud_pos_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_sorted, n, d_pos_of);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_batch_finish(amcl3d_cuda_ctx* ctx, const void* d_part_sum, const uint32_t* d_part_cnt, uint32_t n_poses,
                        uint32_t n_splits, int kind, float* d_weight, uint32_t* d_count)
{
  if (n_poses == 0)
    return 0;
  batch_finish_kernel<<<(n_poses + 255) / 256, 256, 0, ctx->stream>>>(d_part_sum, d_part_cnt, n_poses, n_splits, kind,
                                                                      d_weight, d_count);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace amcl3d_b200

using namespace amcl3d_b200;

namespace
{
// bump allocator over the context's scratch arena (256-byte aligned pieces)
struct Arena
{
  char* base;
  size_t used{ 0 };
  explicit Arena(void* p) : base(static_cast<char*>(p)) {}
  static size_t pad(size_t bytes) { return (bytes + 255) / 256 * 256; }
  template <typename T>
  T* take(size_t count)
  {
    T* r = reinterpret_cast<T*>(base + used);
    used += pad(count * sizeof(T));
    return r;
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
  A3D_TRY(ensure_scratch(ctx, Arena::pad(n_cloud * sizeof(float4)) + 2 * Arena::pad(n_cloud * 4) + 256));
  Arena ar(ctx->scratch);
  float4* d_cloud = ar.take<float4>(n_cloud);
  float* d_vals = ar.take<float>(n_cloud);
  uint32_t* d_idx = idx_out ? ar.take<uint32_t>(n_cloud) : nullptr;
  uint32_t* d_scal = ar.take<uint32_t>(4);
  A3D_CUDA_TRY(cudaMemcpyAsync(d_cloud, cloud_xyzw, n_cloud * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  A3D_CUDA_TRY(cudaMemsetAsync(d_scal, 0, 16, ctx->stream));
  const GridView g = grid->view();
  const RollPitch rp = make_roll_pitch(roll, pitch);
  const uint32_t n = static_cast<uint32_t>(n_cloud);
  point_eval_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(g, d_cloud, n, tx, ty, tz, yaw, rp, d_vals, d_idx, d_scal);
  single_pose_finish_kernel<<<1, 256, 0, ctx->stream>>>(d_vals, n, d_scal, reinterpret_cast<float*>(d_scal) + 1);
  ctx->launches += 2;
  A3D_CUDA_TRY(cudaGetLastError());
  uint32_t host_scal[4] = { 0, 0, 0, 0 };
  A3D_CUDA_TRY(cudaMemcpyAsync(host_scal, d_scal, 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (idx_out)
    A3D_CUDA_TRY(cudaMemcpyAsync(idx_out, d_idx, n_cloud * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
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
  const bool ordered = ctx->opt_reference_order && ctx->opt_point_splits == 0 &&
                       weight_ordered_applies(ctx, grid->view(), n_poses, n_cloud);
  const uint32_t splits = ordered ? 1u : choose_point_splits(ctx, n_poses, n_cloud, grid->brick_shift != 0, false);
  const size_t nc = n_cloud ? n_cloud : 1;
  A3D_TRY(ensure_scratch(ctx, Arena::pad(nc * sizeof(float4)) + Arena::pad(n_poses * 16) + Arena::pad(n_poses * splits * 8) +
                                  Arena::pad(n_poses * splits * 4) + 2 * Arena::pad(n_poses * 4)));
  Arena ar(ctx->scratch);
  float4* d_cloud = ar.take<float4>(nc);
  float* s = ar.take<float>(n_poses * 4);
  double* d_psum = ar.take<double>(n_poses * splits);
  uint32_t* d_pcnt = ar.take<uint32_t>(n_poses * splits);
  float* d_w = ar.take<float>(n_poses);
  uint32_t* d_cnt = ar.take<uint32_t>(n_poses);
  if (n_cloud)
    A3D_CUDA_TRY(cudaMemcpyAsync(d_cloud, cloud_xyzw, n_cloud * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  // poses arrive AoS (x,y,z,a); the kernel wants planes: a strided 2-D copy per plane does the transpose
  for (int k = 0; k < 4; ++k)
    A3D_CUDA_TRY(cudaMemcpy2DAsync(s + static_cast<size_t>(k) * n_poses, sizeof(float), poses_xyza + k, 4 * sizeof(float),
                                   sizeof(float), n_poses, cudaMemcpyHostToDevice, ctx->stream));
  const GridView g = grid->view();
  const RollPitch rp = make_roll_pitch(roll, pitch);
  const uint32_t np = static_cast<uint32_t>(n_poses);
  int kind = 0;
  // the cloud is walked in the caller's order: with one split every sum is the reference's own float chain
  if (ordered)
    A3D_TRY(launch_weight_ordered(ctx, g, d_cloud, static_cast<uint32_t>(n_cloud), s, s + n_poses, s + 2 * n_poses,
                                  s + 3 * n_poses, np, rp, reinterpret_cast<float*>(d_psum), d_pcnt, nullptr, 0));
  else
  A3D_TRY(launch_weight_batch(ctx, g, d_cloud, static_cast<uint32_t>(n_cloud), s, s + n_poses, s + 2 * n_poses,
                              s + 3 * n_poses, np, rp, d_psum, d_pcnt, splits, nullptr, splits == 1, &kind, nullptr, 0, 0));
  batch_finish_kernel<<<(np + 255) / 256, 256, 0, ctx->stream>>>(d_psum, d_pcnt, np, splits, kind, d_w, d_cnt);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  A3D_CUDA_TRY(cudaMemcpyAsync(weight_out, d_w, n_poses * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (n_out)
    A3D_CUDA_TRY(cudaMemcpyAsync(n_out, d_cnt, n_poses * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
