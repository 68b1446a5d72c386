// weight_ordered.cuh -- reference-order cloud sums for MID-SIZED particle sets on an L2-resident grid, in one kernel.
//
// The reference adds a particle's probabilities one by one in the caller's cloud order (Grid3d.cpp:191); at 10^4 points
// that float chain is ~1e-4 away from the exact sum, so only the same order reproduces its bits.  One lane walking the
// whole cloud (weight_v5_kernel, "direct") needs several hundred thousand particles to fill the GPU; the two-pass scheme
// (store every value, replay_sum_kernel adds them in order) moves 8 bytes of HBM traffic per evaluation.  Here the two
// passes meet in shared memory:
//
//   CTA = 32 particles (one per lane) x the whole cloud;  warp 0 = ADDER, warps 1..GW = GATHERERS.
//   The cloud is walked in stages of GW*16 points.  Gatherer w evaluates points [16 w, 16 w + 16) of the stage for the
//   32 particles -- the packed-pair estimate + verify arithmetic of weight_v5_kernel, software-pipelined -- and leaves
//   the 16 x 32 probabilities (+0 for a point the reference skips) in a double-buffered value tile.  The adder, one stage
//   behind, adds each lane's column with plain dependent float adds in point order: the reference's own chain.
//   The dependent chain costs ~4 cycles per point (10 k points: 20 us); the gathers of GW warps run beside it.
//
// Nothing but the 32 sums and counts leaves the SM.  Used while particles < 2 x resident lanes and the grid is linear
// (option "ordered_mode"; filter.cu); larger sets take the direct walk, bricked grids keep the replay path (their gathers
// need the Morton-ordered cloud, which a point-ordered ring cannot follow).
#pragma once

namespace amcl3d_b200
{
constexpr int kOrdPointsPerWarp = 16;

template <int GW, bool BRICKED, bool PARTIAL>
__global__ void __launch_bounds__(32 * (GW + 1), (GW <= 4 ? 6 : (GW <= 8 ? 3 : 1)))
    weight_ordered_kernel(const __grid_constant__ GridView g, const float4* __restrict__ cloud, const uint32_t n_cloud,
                          const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
                          const float* __restrict__ pa, const uint32_t n_poses, const RollPitch rp,
                          const uint32_t partial_mask, float* __restrict__ out_sum, uint32_t* __restrict__ out_cnt,
                          const uint32_t* __restrict__ order, const uint32_t n_lanes)
{
  constexpr int UNROLL = 4;
  constexpr int PW = kOrdPointsPerWarp;
  constexpr int STAGE = GW * PW;
  constexpr int THREADS = 32 * (GW + 1);
  static_assert(kTilePoints % STAGE == 0, "a tile must hold whole stages");
  __shared__ float4 tile[kTilePoints];  // pair-interleaved: [2k] = {xA,xB,yA,yB}, [2k+1] = {zA,zB,wA,wB}
  __shared__ ExactPoseSmemT<32> ep;
  __shared__ float vals[2][STAGE][32];
  __shared__ uint32_t cnt_s[32];
  __shared__ int tile_rmax_bits, tile_zmax_bits;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane;  // index of this lane's particle in the exact-pose table
  const uint32_t lane_i = blockIdx.x * 32u + static_cast<uint32_t>(lane);
  const uint32_t i = lane_i < n_lanes ? (order ? order[lane_i] : lane_i) : n_poses;

  // every warp derives the constants of its 32 particles itself (identical values; the table writes coincide)
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
      if (warp == 1)
      {
        ep.r[0][t] = e.r00;
        ep.r[1][t] = e.r01;
        ep.r[2][t] = e.r02;
        ep.r[3][t] = e.r10;
        ep.r[4][t] = e.r11;
        ep.r[5][t] = e.r12;
        ep.off[0][t] = e.off_x;
        ep.off[1][t] = e.off_y;
        ep.off[2][t] = e.off_z;
      }
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
  if (warp == 0)
    cnt_s[lane] = 0u;
  const float inv_f = g.inv_res_f;
  const float* __restrict__ prob = g.prob;
  const uint32_t sx = g.size_x, sy = g.size_y, sz = g.size_z;
  const uint32_t step_y = g.step_y, step_z = g.step_z, zero_index = g.zero_index;
  const uint32_t lastx = (partial_mask & 1u) ? sx - 1u : 0xFFFFFFFFu, lasty = (partial_mask & 2u) ? sy - 1u : 0xFFFFFFFFu,
                 lastz = (partial_mask & 4u) ? sz - 1u : 0xFFFFFFFFu;
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
  const pk64 MAG = pk2(magic, magic);
  float safe = -1.f, safe_z = -1.f;
  // scalar form of the estimate (weight_v5.cuh: the same IEEE operations as the packed form)
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

  float sum = 0.f;    // adder
  uint32_t cnt = 0;   // gatherers
  uint32_t stage_no = 0;  // stages issued so far (all tiles): the value buffer of stage s is s & 1
  for (uint32_t base = 0; base < n_cloud; base += kTilePoints)
  {
    const int len = static_cast<int>(min(static_cast<uint32_t>(kTilePoints), n_cloud - base));
    if (threadIdx.x == 0)
    {
      tile_rmax_bits = 0;
      tile_zmax_bits = 0;
    }
    // (the barrier at the end of the previous tile's last stage ordered every read of `tile` before these writes)
    __syncthreads();
    float my_r = 0.f, my_z = 0.f;
    for (int j = threadIdx.x; j < len; j += THREADS)
    {
      float4 p = cloud[base + j];
      my_r = fmaxf(my_r, 1.0001f * sqrtf(p.x * p.x + p.y * p.y + p.z * p.z));
      if (!(fabsf(p.x) + fabsf(p.y) + fabsf(p.z) < 1e30f))
        my_r = INFINITY;  // NaN / infinite point in the tile -> verify the whole tile
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
    if (lane == 0)
    {
      atomicMax(&tile_rmax_bits, __float_as_int(my_r));
      atomicMax(&tile_zmax_bits, __float_as_int(my_z));
    }
    __syncthreads();
    const float rmax = __int_as_float(tile_rmax_bits), zmax = __int_as_float(tile_zmax_bits);
    const float k_xy = static_cast<float>(max(max(cx, cy) + 0x4B400000, 0)), k_z = static_cast<float>(max(cz + 0x4B400000, 0));
    safe = 0.5f - 1.25f * 5.9604645e-8f * ((8.f * rmax) * inv_f + k_xy + 4.f);
    safe_z = 0.5f - 1.25f * 5.9604645e-8f * ((3.f * zmax) * inv_f + k_z + 3.f);
    if (!(rmax * inv_f < 2.0e6f) || !(zmax * inv_f < 2.0e6f) || !(safe > 0.f) || !(safe_z > 0.f))
    {
      safe = -1.f;
      safe_z = -1.f;
    }
    const ulonglong2* tile2 = reinterpret_cast<const ulonglong2*>(tile);
    // four whole points starting at tile position j (j % 4 == 0): addresses by the packed estimate, verification where a
    // coordinate is too close to a voxel face (weight_v5.cuh), then the four gathers
    auto gather_group = [&](const int j, float (&v)[UNROLL]) {
      uint32_t gi[UNROLL];
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
        const uint32_t aa = address(kxa, kya, kza), ab = address(kxb, kyb, kzb);
        gi[2 * h] = ina ? aa : zero_index;
        gi[2 * h + 1] = inb ? ab : zero_index;
        cnt += (ina ? 1u : 0u) + (inb ? 1u : 0u);
        if (PARTIAL)
          flags |= ((kxa == lastx) | (kya == lasty) | (kza == lastz) | (kxb == lastx) | (kyb == lasty) | (kzb == lastz)) ? 1u : 0u;
        far_xy = fmaxf(fmaxf(far_xy, fabsf(dxa)), fabsf(dya));
        far_xy = fmaxf(fmaxf(far_xy, fabsf(dxb)), fabsf(dyb));
        far_z = fmaxf(fmaxf(far_z, fabsf(dza)), fabsf(dzb));
      }
      if (!(far_xy < safe) || !(far_z < safe_z) || flags != 0u)
      {
        // verification path, one rolled loop over the four points (weight_v5.cuh)
        uint32_t todo = 15u;
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
    // the ragged last group of the cloud: point by point; positions past the end hold +0
    auto gather_tail = [&](const int j, float (&v)[UNROLL]) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
      {
        v[u] = 0.f;
        if (j + u < len)
        {
          const float4 p = tile_point(tile, j + u);
          uint32_t a;
          bool ok, near;
          estimate1(p, a, ok, near);
          if (near)
          {
            a = exact_address<BRICKED, 32>(g, p, ep, t);
            ok = a != 0xFFFFFFFFu;
          }
          if (ok)
          {
            v[u] = __ldg(prob + a);
            cnt += 1u;
          }
        }
      }
    };

    const int n_stages = (len + STAGE - 1) / STAGE;
    for (int s = 0; s < n_stages; ++s, ++stage_no)
    {
      if (warp > 0)
      {
        if (active)
        {
          float(*dst)[32] = vals[stage_no & 1u] + (warp - 1) * PW;
          const int j0 = s * STAGE + (warp - 1) * PW;  // first tile position of this warp's 16 points
          auto put = [&](const int k, const float (&v)[UNROLL]) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
              dst[k + u][lane] = v[u];
          };
          auto get = [&](const int k, float (&v)[UNROLL]) {
            if (j0 + k + UNROLL <= len)
              gather_group(j0 + k, v);
            else
              gather_tail(j0 + k, v);
          };
          // software pipeline over the four groups: the values of a group are stored only after the next group's
          // gathers have been issued
          float va[UNROLL], vb[UNROLL];
          get(0, va);
          get(4, vb);
          put(0, va);
          get(8, va);
          put(4, vb);
          get(12, vb);
          put(8, va);
          put(12, vb);
        }
      }
      else if (stage_no > 0 && active)
      {
        const float(*src)[32] = vals[(stage_no - 1u) & 1u];
#pragma unroll 16
        for (int k = 0; k < STAGE; ++k)
          sum = __fadd_rn(sum, src[k][lane]);
      }
      __syncthreads();
    }
  }
  // the adder is one stage behind: drain
  if (warp == 0)
  {
    if (stage_no > 0 && active)
    {
      const float(*src)[32] = vals[(stage_no - 1u) & 1u];
#pragma unroll 16
      for (int k = 0; k < STAGE; ++k)
        sum = __fadd_rn(sum, src[k][lane]);
    }
  }
  else if (active)
    atomicAdd(&cnt_s[lane], cnt);
  __syncthreads();
  if (warp == 0 && i < n_poses)
  {
    const uint32_t c = active ? cnt_s[lane] : 0u;
    out_sum[i] = c ? sum : 0.f;
    out_cnt[i] = c;
  }
}
}  // namespace amcl3d_b200
