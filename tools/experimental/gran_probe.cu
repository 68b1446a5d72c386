// gran_probe.cu -- stand-alone probe: random 4-byte gathers over a footprint far larger than L2, issued with different
// load flavours, to see which one makes the memory system fetch the least DRAM data per gather (ncu on the weighting kernel
// at cfg5: 128 B of DRAM traffic per in-map gather, i.e. a full line for one float).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gran_probe tools/experimental/gran_probe.cu && ./gran_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
template <int MODE>
__device__ __forceinline__ float load(const float* p, uint64_t pol)
{
  float v;
  if (MODE == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 1) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 5) asm volatile("ld.global.nc.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 6) asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  if (MODE == 7) asm volatile("ld.global.nc.L1::evict_first.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  if (MODE == 8) asm volatile("ld.global.lu.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 9) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <int MODE>
__global__ void __launch_bounds__(256) gather(const float* __restrict__ buf, const uint64_t mask, const int iters, float* out)
{
  uint64_t pol = 0;
  if (MODE == 6 || MODE == 7)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  uint32_t h = hash32(tid * 2654435761u + 12345u);
  for (int i = 0; i < iters; ++i)
  {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      h = hash32(h + 0x9e3779b9u * (u + 1));
      const uint64_t a = ((static_cast<uint64_t>(h) << 5) ^ hash32(h ^ 0x5bd1e995u)) & mask;
      v[u] = load<MODE>(buf + a, pol);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      s += v[u];
  }
  out[tid] = s;
}
int main()
{
  const uint64_t n = 1ull << 31;  // floats: 8 GiB
  float* buf;
  CK(cudaMalloc(&buf, n * 4));
  CK(cudaMemset(buf, 0, n * 4));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int threads = prop.multiProcessorCount * 2048;
  float* out;
  CK(cudaMalloc(&out, threads * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const char* names[] = { "ld.global.nc", "ld.global.cg", "ld.global.cv", "nc.L1::no_allocate", "nc.L2::64B", "nc.L2::128B",
                          "nc.L2::cache_hint(evict_first)", "nc.L1::evict_first+L2 hint", "ld.global.lu", "ld.global.cs" };
  for (size_t lim : { (size_t)0, (size_t)32, (size_t)64, (size_t)128 })
  {
    if (lim)
    {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, lim);
      size_t got = 0;
      cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
      printf("cudaLimitMaxL2FetchGranularity <- %zu: %s, now %zu\n", lim, cudaGetErrorString(e), got);
    }
    for (int mode = 0; mode < 10; ++mode)
    {
      if (lim && mode > 1)
        continue;
      float ms = 0;
      const int iters = 64;
      for (int rep = 0; rep < 2; ++rep)
      {
        CK(cudaEventRecord(e0));
        switch (mode)
        {
          case 0: gather<0><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 1: gather<1><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 2: gather<2><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 3: gather<3><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 4: gather<4><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 5: gather<5><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 6: gather<6><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 7: gather<7><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 8: gather<8><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
          case 9: gather<9><<<threads / 256, 256>>>(buf, n - 1, iters, out); break;
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      const double g = (double)threads * iters * 8;
      printf("%-34s %8.3f ms  %.3e gathers/s  = %.2f TB/s of 32 B sectors\n", names[mode], ms, g / (ms * 1e-3), g / (ms * 1e-3) * 32 / 1e12);
    }
  }
  return 0;
}
