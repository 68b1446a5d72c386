// filter_common.cuh -- per-particle math shared by the update kernels (filter.cu, filter_exact.cu).
#pragma once

#include "common.cuh"

namespace amcl3d_b200
{
// ------------------------------------------------------------------------------------------ shared per-particle math

constexpr uint32_t kInlineRanges = 16;
struct RangeParams
{
  const float* ranges;  // n_ranges x (r, ax, ay, az) in device memory; NULL when they fit `inline_ranges`
  uint32_t n_ranges;
  float k1, k2;  // ParticleFilter.cpp:231-232, evaluated on the host
  // up to 16 beacons travel in the kernel's parameter block: no host->device copy (and no pinned staging sync) at all
  float4 inline_ranges[kInlineRanges];
};

// ParticleFilter.cpp:224-244
__device__ __forceinline__ float range_weight(const RangeParams& rg, float x, float y, float z)
{
  if (rg.n_ranges == 0)
    return 0.f;
  float w = 1.f;
  for (uint32_t i = 0; i < rg.n_ranges; ++i)
  {
    const float4 b = rg.ranges ? *reinterpret_cast<const float4*>(rg.ranges + 4 * i) : rg.inline_ranges[i];  // r, ax, ay, az
    const float dx = __fsub_rn(x, b.y), dy = __fsub_rn(y, b.z), dz = __fsub_rn(z, b.w);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float r = static_cast<float>(sqrt(static_cast<double>(d2)));  // :239 double sqrt, stored to float
    const float e = __fsub_rn(r, b.x);
    const float arg = __fmul_rn(__fmul_rn(-rg.k2, e), e);  // float, left to right
    // :240  w = float( double(w) * ( double(k1) * exp(double(arg)) ) )
    w = static_cast<float>(__dmul_rn(static_cast<double>(w), __dmul_rn(static_cast<double>(rg.k1), exp(static_cast<double>(arg)))));
  }
  return w;
}

struct Planes
{
  float *x, *y, *z, *a, *w, *wp, *wr;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


}  // namespace amcl3d_b200
