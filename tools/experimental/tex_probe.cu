// tex_probe.cu -- stand-alone probe (not part of the library): rate of warp-coherent 3-D gathers through (a) LDG on a
// linear plane, (b) LDG on a 32^3-bricked plane, (c) the texture unit on a block-linear cudaArray (point sampling).
// Emulates the weighting kernel's access pattern: every warp walks the same list of cloud points; for point j the 32
// lanes (32 neighbouring particles) read voxels within +-S of centre[j] + the warp's own offset (+-W).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tex_probe tools/experimental/tex_probe.cu && ./tex_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

struct Dims { int sx, sy, sz; uint32_t nbx, nby; };

template <int MODE>  // 0 linear LDG, 1 bricked LDG, 2 TEX
__global__ void __launch_bounds__(256) probe(const float* __restrict__ plane, cudaTextureObject_t tex, const int4* __restrict__ centres,
                                             const int n_points, const Dims d, const int S, const int W, float* out)
{
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t hw = hash32(warp * 2654435761u + 17u), hl = hash32(warp * 97u + lane * 7919u + 3u);
  const int ox = (int)(hw % (2 * W + 1)) - W + (int)(hl % (2 * S + 1)) - S;
  const int oy = (int)((hw >> 8) % (2 * W + 1)) - W + (int)((hl >> 8) % (2 * S + 1)) - S;
  const int oz = (int)((hw >> 16) % (2 * W + 1)) - W + (int)((hl >> 16) % (2 * S + 1)) - S;
  float sum = 0.f;
  for (int j = 0; j < n_points; j += 4)
  {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int4 c = centres[j + u];
      const int x = c.x + ox, y = c.y + oy, z = c.z + oz;
      if (MODE == 2)
        v[u] = tex3D<float>(tex, x + 0.5f, y + 0.5f, z + 0.5f);
      else if (MODE == 0)
        v[u] = __ldg(plane + ((size_t)z * d.sy + y) * d.sx + x);
      else
      {
        const uint32_t kx = x, ky = y, kz = z;
        const uint32_t brick = ((kz >> 5) * d.nby + (ky >> 5)) * d.nbx + (kx >> 5);
        v[u] = __ldg(plane + ((size_t)brick << 15) + ((kz & 31) << 10) + ((ky & 31) << 5) + (kx & 31));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      sum += v[u];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

__global__ void fill_surface(cudaSurfaceObject_t surf, Dims d)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
  if (x < d.sx)
    surf3Dwrite(1.0f + 0.001f * (float)((x + y + z) & 255), surf, x * 4, y, z);
}

int main(int argc, char** argv)
{
  const int big = argc > 1 ? atoi(argv[1]) : 1;
  Dims d;
  d.sx = big ? 2000 : 400; d.sy = big ? 2000 : 400; d.sz = big ? 400 : 100;
  d.nbx = (d.sx + 31) / 32; d.nby = (d.sy + 31) / 32;
  const uint32_t nbz = (d.sz + 31) / 32;
  const size_t n_lin = (size_t)d.sx * d.sy * d.sz, n_br = (size_t)d.nbx * d.nby * nbz << 15;
  float* plane;
  CK(cudaMalloc(&plane, n_br * 4));
  CK(cudaMemset(plane, 0, n_br * 4));
  cudaArray_t arr;
  cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float>();
  CK(cudaMalloc3DArray(&arr, &fmt, make_cudaExtent(d.sx, d.sy, d.sz), cudaArraySurfaceLoadStore));
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaSurfaceObject_t surf;
  CK(cudaCreateSurfaceObject(&surf, &rd));
  fill_surface<<<dim3((d.sx + 255) / 256, d.sy, d.sz), 256>>>(surf, d);
  CK(cudaDeviceSynchronize());
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const int n_points = 2048;
  const int n_threads = prop.multiProcessorCount * 2048 * 8;   // 8 waves of full occupancy
  float* out;
  CK(cudaMalloc(&out, (size_t)n_threads * 4));
  int4* d_c;
  CK(cudaMalloc(&d_c, n_points * sizeof(int4)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("map %dx%dx%d (%.2f GB), %d SMs, %d kHz, %d threads x %d points\n", d.sx, d.sy, d.sz, n_lin * 4 / 1e9, prop.multiProcessorCount,
         clock_khz, n_threads, n_points);
  const int R = big ? 600 : 150;           // half-width of the region the cloud points fall into (voxels)
  for (int S = 0; S <= 4; S += (S == 0 ? 1 : S))            // lane spread 0, 1, 2, 4
    for (int W = 4; W <= 64; W *= 4)                         // warp spread 4, 16, 64
    {
      std::vector<int4> c(n_points);
      srand(1234);
      for (auto& p : c)
      {
        p.x = d.sx / 2 + rand() % (2 * R) - R;
        p.y = d.sy / 2 + rand() % (2 * R) - R;
        p.z = d.sz / 2 + rand() % (d.sz / 2) - d.sz / 4;
      }
      CK(cudaMemcpy(d_c, c.data(), n_points * sizeof(int4), cudaMemcpyHostToDevice));
      float ms[3];
      for (int mode = 0; mode < 3; ++mode)
      {
        for (int rep = 0; rep < 2; ++rep)
        {
          CK(cudaEventRecord(e0));
          if (mode == 0)
            probe<0><<<n_threads / 256, 256>>>(plane, tex, d_c, n_points, d, S, W, out);
          else if (mode == 1)
            probe<1><<<n_threads / 256, 256>>>(plane, tex, d_c, n_points, d, S, W, out);
          else
            probe<2><<<n_threads / 256, 256>>>(plane, tex, d_c, n_points, d, S, W, out);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaEventElapsedTime(&ms[mode], e0, e1));
        }
      }
      const double ev = (double)n_threads * n_points;
      printf("lane spread +-%d warp spread +-%2d : linear %7.3f ms (%.2f lanes/clk/SM)  bricked %7.3f ms (%.2f)  TEX %7.3f ms (%.2f)   evals/s %.3g %.3g %.3g\n",
             S, W, ms[0], ev / (ms[0] * 1e-3) / prop.multiProcessorCount / (clock_khz * 1e3), ms[1],
             ev / (ms[1] * 1e-3) / prop.multiProcessorCount / (clock_khz * 1e3), ms[2],
             ev / (ms[2] * 1e-3) / prop.multiProcessorCount / (clock_khz * 1e3), ev / (ms[0] * 1e-3), ev / (ms[1] * 1e-3),
             ev / (ms[2] * 1e-3));
    }
  return 0;
}
