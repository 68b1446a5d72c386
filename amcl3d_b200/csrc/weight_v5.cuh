// weight_v5.cuh -- the weighting kernel, generation 5: v4's estimate + verify scheme issued as PACKED fp32 pairs.
//
// sm_100 has two-wide fp32 instructions on 64-bit register pairs (PTX fma.rn.f32x2 / add.rn.f32x2 / sub.rn.f32x2, SASS
// FFMA2 / FADD2), whose second source may be a single register broadcast to both halves.  The estimate of v4
//     q = fma(px, R0, fma(py, R1, fma(pz, R2, F)))      r = q + 1.5*2^23      d = q - (r - 1.5*2^23)
// is evaluated for TWO consecutive cloud points per instruction: the point tile is stored pair-interleaved
// ({xA,xB,yA,yB}, {zA,zB,wA,wB}), so that two LDS.128 deliver the four operand pairs, and the per-particle constants
// ride along as broadcast operands (no extra registers).  Every half of a packed instruction is the same IEEE
// operation as the scalar one, so the estimates, the error bound derived in weight.cu and therefore the voxels are
// bit-identical to v4; what changes is the issue count: 16 fp32 instructions per point become 8.
// The near-face test no longer adds the z band shift per point: x/y and z distances keep separate running maxima that
// are compared against their own bands (3-input FMNMX3), 1.5 instructions per point instead of 4.
//
// STORE: every gathered probability (0 for a point the reference skips) is also written to a value matrix, tiled by
// warp: vals[lane / 32][point position][lane % 32] -- the 32 lanes of a warp write 128 contiguous bytes per point, and
// one warp's values of consecutive points are contiguous, so the replay streams through memory.  That is
// the first half of the two-pass "gather anywhere, add in order" scheme: the gathers run in whatever order and split is
// fastest (Morton-ordered cloud, sub-chunk CTAs), and replay_sum_kernel (weight.cu) then adds each particle's values
// in the CALLER's cloud order -- the reference's own float chain, at the parallelism of the fast path.
//
// Accumulation (acc_mode): 0 = this launch starts at +0 and stores a float partial; 1 = it continues the float running
// sum left by the previous chunk launch (one float chain in cloud order = Grid3d.cpp:191 bit for bit); 2 / 3 = the
// launch starts at +0 and stores / adds its float sum into a DOUBLE accumulator (re-ordered clouds and split chunks:
// the order of the reference's chain is gone anyway, so the partials are combined without further rounding).
#pragma once

namespace amcl3d_b200
{
typedef unsigned long long pk64;

__device__ __forceinline__ pk64 pk2(const float lo, const float hi)
{
  pk64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(const pk64 v, float& lo, float& hi)
{
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ pk64 fma2(const pk64 a, const pk64 b, const pk64 c)
{
  pk64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ pk64 add2(const pk64 a, const pk64 b)
{
  pk64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pk64 sub2(const pk64 a, const pk64 b)
{
  pk64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// point j of a pair-interleaved tile
__device__ __forceinline__ float4 tile_point(const float4* tile, const int j)
{
  const float* f = reinterpret_cast<const float*>(tile) + 8 * (j >> 1) + (j & 1);
  return make_float4(f[0], f[2], f[4], f[6]);
}

// MODE: 0 = software-pipelined gathers at 64 registers (1024 resident lanes per SM), 1 = not pipelined (each group's
// values are added right behind its loads), 2 = pipelined at 80 registers (768 resident lanes per SM, no spills).
template <int BLOCK, bool BRICKED, bool PARTIAL, bool STORE, int MODE = 0>
__global__ void __launch_bounds__(BLOCK, (MODE == 2 ? 768 : 1024) / BLOCK)
    weight_v5_kernel(const __grid_constant__ GridView g, const float4* __restrict__ cloud, const uint32_t n_cloud,
                     const uint32_t chunk_len, const float* __restrict__ px, const float* __restrict__ py,
                     const float* __restrict__ pz, const float* __restrict__ pa, const uint32_t n_poses,
                     const RollPitch rp, const uint32_t partial_mask, void* __restrict__ part_sum,
                     uint32_t* __restrict__ part_cnt, const uint32_t chunk_first, const int acc_mode,
                     const uint32_t* __restrict__ order, float* __restrict__ vals, const uint64_t vals_stride,
                     const uint32_t n_lanes)
{
  // n_lanes scheduled lanes; lane l weighs particle order[l] (identity without `order`) out of an index space of
  // n_poses particles -- the two differ when this GPU weighs a pose-coherent SLICE of a set sharded over several GPUs
  constexpr int UNROLL = 4;  // points per group = two packed pairs
  __shared__ float4 tile[kTilePoints];  // pair-interleaved: [2k] = {xA,xB,yA,yB}, [2k+1] = {zA,zB,wA,wB}
  __shared__ ExactPoseSmemT<BLOCK> ep;
  __shared__ int tile_rmax_bits, tile_zmax_bits;
  const int t = threadIdx.x;
  const uint32_t lane_i = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t i = lane_i < n_lanes ? (order ? order[lane_i] : lane_i) : n_poses;
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
  const uint32_t lastx = (partial_mask & 1u) ? sx - 1u : 0xFFFFFFFFu, lasty = (partial_mask & 2u) ? sy - 1u : 0xFFFFFFFFu,
                 lastz = (partial_mask & 4u) ? sz - 1u : 0xFFFFFFFFu;
  // Bricked address = X(kx) + Y(ky) + Z(kz), each a multiply-add of k and (k & ~(brick - 1)).  The brick edge is a
  // compile-time constant here (kBrickShift; the host refuses anything else for this kernel), so the shifts and masks are
  // immediates and the two remaining factors are launch constants: 8 integer instructions per point, no branch.
  constexpr uint32_t bsh = kBrickShift, bmask = ~((1u << bsh) - 1u);
  constexpr uint32_t bcx = (1u << (2 * bsh)) - 1u;
  const uint32_t bcy = (g.nbx << (2 * bsh)) - (1u << bsh);
  const uint32_t bcz = (g.nbx * g.nby - 1u) << (2 * bsh);
  auto address = [&](const uint32_t kx, const uint32_t ky, const uint32_t kz) -> uint32_t {
    if (!BRICKED)
      return kx + ky * step_y + kz * step_z;
    uint32_t a = kx + (ky << bsh) + (kz << (2 * bsh));
    a += (kx & bmask) * bcx;
    a += (ky & bmask) * bcy;
    a += (kz & bmask) * bcz;
    return a;
  };
  const float magic = 12582912.f;  // 1.5 * 2^23
  float safe = -1.f, safe_z = -1.f;
  // scalar form of the estimate (verification path and ragged tails): the same IEEE operations as the packed form
  auto estimate1 = [&](const float4 p, uint32_t& addr, bool& in, bool& near) {
    const float qx = __fmaf_rn(p.x, R00, __fmaf_rn(p.y, R01, __fmaf_rn(p.z, R02, fx)));
    const float qy = __fmaf_rn(p.x, R10, __fmaf_rn(p.y, R11, __fmaf_rn(p.z, R12, fy)));
    const float qz = __fmaf_rn(p.w, inv_f, fz);
    const float rx = __fadd_rn(qx, magic), ry = __fadd_rn(qy, magic), rz = __fadd_rn(qz, magic);
    const float dx = __fsub_rn(qx, __fsub_rn(rx, magic)), dy = __fsub_rn(qy, __fsub_rn(ry, magic)),
                dz = __fsub_rn(qz, __fsub_rn(rz, magic));
    const uint32_t kx = static_cast<uint32_t>(__float_as_int(rx) + cx), ky = static_cast<uint32_t>(__float_as_int(ry) + cy),
                   kz = static_cast<uint32_t>(__float_as_int(rz) + cz);
    in = (kx < sx) & (ky < sy) & (kz < sz);
    near = !(fmaxf(fabsf(dx), fabsf(dy)) < safe) || !(fabsf(dz) < safe_z);
    if (PARTIAL)
      near |= (kx == lastx) | (ky == lasty) | (kz == lastz);
    addr = address(kx, ky, kz);
  };

  float sum = 0.f;
  uint32_t cnt = 0;
  if (acc_mode == 1 && i < n_poses)
  {
    sum = static_cast<const float*>(part_sum)[static_cast<size_t>(slot) * n_poses + i];
    cnt = part_cnt[static_cast<size_t>(slot) * n_poses + i];
  }
  const pk64 MAG = pk2(magic, magic);
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
      float* d = reinterpret_cast<float*>(tile) + 8 * (j >> 1) + (j & 1);
      d[0] = p.x;
      d[2] = p.y;
      d[4] = p.z;
      d[6] = p.w;
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
    // proven-safe bands (weight.cu header comment), 1.25x safety, per particle K; outside the magic-number range of the
    // estimate (or with non-finite points in the tile) everything is verified
    const float k_xy = static_cast<float>(max(max(cx, cy) + 0x4B400000, 0)), k_z = static_cast<float>(max(cz + 0x4B400000, 0));
    safe = 0.5f - 1.25f * 5.9604645e-8f * ((8.f * rmax) * inv_f + k_xy + 4.f);
    safe_z = 0.5f - 1.25f * 5.9604645e-8f * ((3.f * zmax) * inv_f + k_z + 3.f);
    if (!(rmax * inv_f < 2.0e6f) || !(zmax * inv_f < 2.0e6f) || !(safe > 0.f) || !(safe_z > 0.f))
    {
      safe = -1.f;
      safe_z = -1.f;
    }
    if (active)
    {
      const ulonglong2* tile2 = reinterpret_cast<const ulonglong2*>(tile);
      const int full = len - (len % UNROLL);
      // Software pipeline: a warp issues in order, so a running sum that consumes its gathers right away stalls the warp
      // on the first add until the L2 / HBM round trip is over.  Here the values of group g are only added after the
      // addresses of group g + 1 have been computed and ITS gathers issued: every lane keeps 4..8 loads in flight and
      // the round trip overlaps the lane's own arithmetic.  The adds still run in point order.  The loop is unrolled
      // twice over two value buffers (va / vb), so no register copy ever waits on a load that has just been issued.
      auto gather_group = [&](const int j, float (&v)[UNROLL]) {
        uint32_t gi[UNROLL];
        constexpr bool kFlagBits = BRICKED;  // large maps: one group in four verifies -> remember which pair
        float far_xy = 0.f, far_z = 0.f;
        uint32_t flags = 0;
#pragma unroll
        for (int h = 0; h < UNROLL / 2; ++h)
        {
          const ulonglong2 A = tile2[j + 2 * h], B = tile2[j + 2 * h + 1];  // (x2, y2), (z2, w2)
          const pk64 qx = fma2(A.x, pk2(R00, R00), fma2(A.y, pk2(R01, R01), fma2(B.x, pk2(R02, R02), pk2(fx, fx))));
          const pk64 qy = fma2(A.x, pk2(R10, R10), fma2(A.y, pk2(R11, R11), fma2(B.x, pk2(R12, R12), pk2(fy, fy))));
          const pk64 qz = fma2(B.y, pk2(inv_f, inv_f), pk2(fz, fz));
          const pk64 rx = add2(qx, MAG), ry = add2(qy, MAG), rz = add2(qz, MAG);
          const pk64 dx = sub2(qx, sub2(rx, MAG)), dy = sub2(qy, sub2(ry, MAG)), dz = sub2(qz, sub2(rz, MAG));
          float rxa, rxb, rya, ryb, rza, rzb, dxa, dxb, dya, dyb, dza, dzb;
          upk2(rx, rxa, rxb);
          upk2(ry, rya, ryb);
          upk2(rz, rza, rzb);
          upk2(dx, dxa, dxb);
          upk2(dy, dya, dyb);
          upk2(dz, dza, dzb);
          const uint32_t kxa = static_cast<uint32_t>(__float_as_int(rxa) + cx), kya = static_cast<uint32_t>(__float_as_int(rya) + cy),
                         kza = static_cast<uint32_t>(__float_as_int(rza) + cz);
          const uint32_t kxb = static_cast<uint32_t>(__float_as_int(rxb) + cx), kyb = static_cast<uint32_t>(__float_as_int(ryb) + cy),
                         kzb = static_cast<uint32_t>(__float_as_int(rzb) + cz);
          const bool ina = (kxa < sx) & (kya < sy) & (kza < sz), inb = (kxb < sx) & (kyb < sy) & (kzb < sz);
          // computed unconditionally (wrapping arithmetic on whatever the estimate gave) and then selected: no branch
          const uint32_t aa = address(kxa, kya, kza), ab = address(kxb, kyb, kzb);
          gi[2 * h] = ina ? aa : zero_index;
          gi[2 * h + 1] = inb ? ab : zero_index;
          cnt += (ina ? 1u : 0u) + (inb ? 1u : 0u);
          bool last = false;
          if (PARTIAL)
            last = (kxa == lastx) | (kya == lasty) | (kza == lastz) | (kxb == lastx) | (kyb == lasty) | (kzb == lastz);
          if (kFlagBits)
          {
            const float pxy = fmaxf(fmaxf(fabsf(dxa), fabsf(dya)), fmaxf(fabsf(dxb), fabsf(dyb)));
            const float pz2 = fmaxf(fabsf(dza), fabsf(dzb));
            flags |= (!(pxy < safe) || !(pz2 < safe_z) || last) ? (1u << h) : 0u;
          }
          else
          {
            far_xy = fmaxf(fmaxf(far_xy, fabsf(dxa)), fabsf(dya));
            far_xy = fmaxf(fmaxf(far_xy, fabsf(dxb)), fabsf(dyb));
            far_z = fmaxf(fmaxf(far_z, fabsf(dza)), fabsf(dzb));
            flags |= last ? 1u : 0u;
          }
        }
        if (kFlagBits ? (flags != 0u) : (!(far_xy < safe) || !(far_z < safe_z) || flags != 0u))
        {
          // Verification path: points whose estimate is too close to a voxel face get the reference's arithmetic.  A warp
          // that enters here does so for one or two of its lanes (1.7e-3 of the evaluations on map L, i.e. every fifth
          // group of a warp), so every instruction on this path costs a full issue slot for a single lane: the path is
          // ONE rolled loop over the candidate points (no per-point code copies) with the reference's arithmetic inlined
          // -- grid constants come straight from the parameter bank, nothing is passed through memory
          // (ncu, r2 capture: the out-of-line version took 26 % of all issued instructions).
          uint32_t todo = kFlagBits ? (((flags & 1u) ? 3u : 0u) | ((flags & 2u) ? 12u : 0u)) : 15u;
#pragma unroll 1
          while (todo)
          {
            const int u = __ffs(static_cast<int>(todo)) - 1;
            todo &= todo - 1u;
            const float4 p = tile_point(tile, j + u);
            uint32_t a;
            bool in, near;
            estimate1(p, a, in, near);
            if (near)
            {
              // Grid3d.cpp:174-189 verbatim (the operands of the exact pose live in shared memory)
              const float nx = transform_axis(p.x, p.y, p.z, ep.r[0][t], ep.r[1][t], ep.r[2][t], ep.off[0][t]);
              const float ny = transform_axis(p.x, p.y, p.z, ep.r[3][t], ep.r[4][t], ep.r[5][t], ep.off[1][t]);
              const float nz = static_cast<float>(__dadd_rn(static_cast<double>(p.w), ep.off[2][t]));
              uint32_t e = zero_index;
              if (nx >= 0.f && nx < g.ext_up_x && ny >= 0.f && ny < g.ext_up_y && nz >= 0.f && nz < g.ext_up_z)
              {
                const uint32_t kx = voxel_coord(nx, g), ky = voxel_coord(ny, g), kz = voxel_coord(nz, g);
                const uint32_t lin = kx + ky * step_y + kz * step_z;  // uint32 arithmetic as in :187
                if (kx < sx && ky < sy && kz < sz && static_cast<uint64_t>(lin) < g.n_cells)
                  e = BRICKED ? address(kx, ky, kz) : lin;
              }
              const uint32_t old = u == 0 ? gi[0] : (u == 1 ? gi[1] : (u == 2 ? gi[2] : gi[3]));
              cnt += (e != zero_index ? 1u : 0u) - (old != zero_index ? 1u : 0u);
#pragma unroll
              for (int w = 0; w < UNROLL; ++w)
                gi[w] = (w == u) ? e : gi[w];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          v[u] = __ldg(prob + gi[u]);
      };
      auto consume_group = [&](const int j, const float (&v)[UNROLL]) {
        if (STORE)
        {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u)
            __stcs(vals + (static_cast<uint64_t>(lane_i >> 5) * vals_stride + (base + j + u)) * 32u + (lane_i & 31u), v[u]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          sum = __fadd_rn(sum, v[u]);
      };
      if (MODE == 1)
      {
        for (int j = 0; j < full; j += UNROLL)
        {
          float va[UNROLL];
          gather_group(j, va);
          consume_group(j, va);
        }
      }
      else if (full > 0)
      {
        float va[UNROLL], vb[UNROLL];
        gather_group(0, va);
        int j = UNROLL;  // va holds the group that starts at j - UNROLL
        for (; j + 2 * UNROLL <= full; j += 2 * UNROLL)
        {
          gather_group(j, vb);
          consume_group(j - UNROLL, va);
          gather_group(j + UNROLL, va);
          consume_group(j, vb);
        }
        if (j < full)  // one more group
        {
          gather_group(j, vb);
          consume_group(j - UNROLL, va);
          consume_group(j, vb);
        }
        else
          consume_group(j - UNROLL, va);
      }
      for (int j = full; j < len; ++j)  // ragged end of the chunk
      {
        const float4 p = tile_point(tile, j);
        uint32_t a;
        bool ok, near;
        estimate1(p, a, ok, near);
        if (near)
        {
          a = exact_address<BRICKED, BLOCK>(g, p, ep, t);
          ok = a != 0xFFFFFFFFu;
        }
        const float v = ok ? __ldg(prob + a) : 0.f;
        if (STORE)
          __stcs(vals + (static_cast<uint64_t>(lane_i >> 5) * vals_stride + (base + j)) * 32u + (lane_i & 31u), v);
        if (ok)
        {
          sum = __fadd_rn(sum, v);
          cnt += 1u;
        }
      }
    }
    __syncthreads();
  }
  if (i < n_poses)
  {
    const size_t o = static_cast<size_t>(slot) * n_poses + i;
    if (acc_mode <= 1)
    {
      static_cast<float*>(part_sum)[o] = sum;
      part_cnt[o] = cnt;
    }
    else
    {
      double* d = static_cast<double*>(part_sum);
      d[o] = (acc_mode == 3 ? d[o] : 0.0) + static_cast<double>(sum);
      part_cnt[o] = (acc_mode == 3 ? part_cnt[o] : 0u) + cnt;
    }
  }
}
}  // namespace amcl3d_b200
