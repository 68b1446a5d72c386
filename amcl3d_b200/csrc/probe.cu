// probe.cu -- the gather roofline of this device, measured instead of assumed ("L2-resident regime =
// measured L2 random-sector bandwidth on the box (a 4-B-random-gather microbenchmark over the same footprint)").
//
// Every thread issues independent 4-byte read-only loads at pseudo-random addresses inside a buffer of the requested
// footprint: each load costs the memory system one 32-byte sector, which is exactly what one in-map evaluation of the
// weighting kernel costs when the 32 lanes of a warp scatter.  `lanes_per_sector` > 1 makes groups of that many
// neighbouring lanes read the same sector (the partially coalesced case: at cfg2 a warp's request covers ~20
// sectors).  Result: sectors/s * 32 B in GB/s -- the denominator for roofline.frac_l2 in bench.py.
#include "common.cuh"

namespace amcl3d_b200
{
template <int ILP>
__global__ void __launch_bounds__(256) probe_gather_kernel(const float* __restrict__ buf, const uint32_t mask,
                                                          const uint32_t iters, const uint32_t lane_shift,
                                                          float* __restrict__ out)
{
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  // one stream of addresses per group of (1 << lane_shift) lanes
  uint32_t s = (tid >> lane_shift) * 2654435761u + 12345u;
  const uint32_t sub = tid & ((1u << lane_shift) - 1u) & 7u;  // word inside the sector
  float acc = 0.f;
  for (uint32_t it = 0; it < iters; ++it)
  {
    uint32_t a[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u)
    {
      s = s * 1664525u + 1013904223u;
      uint32_t h = s ^ (s >> 15);
      h *= 2246822519u;
      h ^= h >> 13;
      a[u] = ((h & mask) & ~7u) | sub;
    }
    float v[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u)
      v[u] = __ldg(buf + a[u]);
#pragma unroll
    for (int u = 0; u < ILP; ++u)
      acc += v[u];
  }
  if (acc == 12345.678f)  // never true for a zero-filled buffer: keeps the loads alive
    out[0] = acc;
}
}  // namespace amcl3d_b200

using namespace amcl3d_b200;

extern "C" int amcl3d_cuda_probe_gather(amcl3d_cuda_ctx* ctx, uint64_t footprint_bytes, uint32_t lanes_per_sector,
                                        double* sector_gbs, double* requests_per_s)
{
  if (!ctx || !sector_gbs)
    return fail(AMCL3D_CUDA_ERR_INVALID, "probe_gather: NULL argument");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  uint64_t words = 1ull << 10;
  while (words * 2 * sizeof(float) <= footprint_bytes && words < (1ull << 31))
    words *= 2;  // power of two <= footprint
  uint32_t lane_shift = 0;
  while ((2u << lane_shift) <= lanes_per_sector && lane_shift < 3)
    ++lane_shift;
  float* d_buf = nullptr;
  float* d_out = nullptr;
  A3D_CUDA_TRY(cudaMalloc(&d_buf, words * sizeof(float)));
  if (cudaMalloc(&d_out, sizeof(float)) != cudaSuccess)
  {
    cudaFree(d_buf);
    return fail(AMCL3D_CUDA_ERR_CUDA, "probe_gather: cudaMalloc failed");
  }
  cudaMemsetAsync(d_buf, 0, words * sizeof(float), ctx->stream);
  const uint32_t mask = static_cast<uint32_t>(words - 1);
  const int blocks = ctx->sm_count * 8;
  const uint32_t iters = 256;
  constexpr int kIlp = 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best_ms = 0.f;
  for (int rep = 0; rep < 4; ++rep)  // rep 0 warms the caches (the footprint stays L2-resident when it fits)
  {
    cudaEventRecord(e0, ctx->stream);
    probe_gather_kernel<kIlp><<<blocks, 256, 0, ctx->stream>>>(d_buf, mask, iters, lane_shift, d_out);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best_ms == 0.f || ms < best_ms))
      best_ms = ms;
    ctx->launches++;
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_buf);
  cudaFree(d_out);
  if (err != cudaSuccess)
    return fail(AMCL3D_CUDA_ERR_CUDA, std::string("probe_gather: ") + cudaGetErrorString(err));
  const double loads = static_cast<double>(blocks) * 256.0 * iters * kIlp;
  const double sectors = loads / static_cast<double>(1u << lane_shift);
  *sector_gbs = sectors * 32.0 / (best_ms * 1e-3) / 1e9;
  if (requests_per_s)
    *requests_per_s = loads / 32.0 / (best_ms * 1e-3);
  return 0;
}
