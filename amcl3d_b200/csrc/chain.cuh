// chain.cuh -- bit-exact reproduction of the reference's SEQUENTIAL float accumulations.
//
// The reference adds weights one after another in float (Grid3d.cpp:191, ParticleFilter.cpp:151-152,179,
// 190-193,214).  Float addition is not associative, so a tree reduction gives a (slightly) different number;
// to return the reference's own bits the additions must happen in the reference's order.  block_chain does
// exactly that inside one thread block: warps 1.. stream the operands through a double-buffered shared-memory
// tile (coalesced), lane 0 of warp 0 performs the dependent __fadd_rn chain on up to K planes at once
// (K independent chains interleave, so their 4-cycle add latencies overlap).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace amcl3d_b200
{
constexpr int kChainTile = 1024;  // floats per plane per stage

template <int K>
struct alignas(16) ChainSmem
{
  float in[2][K][kChainTile];
  float prefix[2][kChainTile];  // only used when the running value of plane 0 is exported
};

// All threads of the block must call this (contains __syncthreads).  Requires blockDim.x >= 64.
// src[k]: plane k (n floats, global).  acc[k]: in = starting value, out = chain result (valid in thread 0 only;
// broadcast it through shared memory if the rest of the block needs it).
// prefix_out (nullable, K must be 1 to be meaningful): inclusive running value after each element of plane 0.
template <int K>
__device__ __forceinline__ void block_chain(const float* const (&src)[K], uint64_t n, float (&acc)[K],
                                            float* prefix_out, ChainSmem<K>& sm)
{
  const int tid = threadIdx.x;
  const int loaders = blockDim.x - 32;
  const uint64_t n_tiles = (n + kChainTile - 1) / kChainTile;
  // prologue: everybody loads tile 0
  {
    const uint64_t len = n < static_cast<uint64_t>(kChainTile) ? n : kChainTile;
    for (int k = 0; k < K; ++k)
      for (uint64_t j = tid; j < len; j += blockDim.x)
        sm.in[0][k][j] = src[k][j];
  }
  __syncthreads();
  for (uint64_t t = 0; t < n_tiles; ++t)
  {
    const int b = static_cast<int>(t & 1);
    const uint64_t base = t * kChainTile;
    const int len = static_cast<int>((n - base) < static_cast<uint64_t>(kChainTile) ? (n - base) : kChainTile);
    if (tid >= 32)
    {
      // prefetch tile t+1 and flush the prefix of tile t-1 while lane 0 works
      if (t + 1 < n_tiles)
      {
        const uint64_t nbase = base + kChainTile;
        const int nlen = static_cast<int>((n - nbase) < static_cast<uint64_t>(kChainTile) ? (n - nbase) : kChainTile);
        for (int k = 0; k < K; ++k)
          for (int j = tid - 32; j < nlen; j += loaders)
            sm.in[b ^ 1][k][j] = src[k][nbase + j];
      }
      if (prefix_out && t > 0)
      {
        const uint64_t pbase = base - kChainTile;
        for (int j = tid - 32; j < kChainTile; j += loaders)
          prefix_out[pbase + j] = sm.prefix[b ^ 1][j];
      }
    }
    else if (tid == 0)
    {
      // Batches of B float4 per plane are fetched one batch AHEAD of the adds, so the ~30-cycle LDS latency
      // overlaps the dependent add chain (4 cycles per element) instead of stalling it.
      constexpr int B = (K == 1) ? 8 : (K == 2 ? 4 : 2);
      constexpr int STEP = 4 * B;
      float c[K];
#pragma unroll
      for (int k = 0; k < K; ++k)
        c[k] = acc[k];
      // Ping-pong register buffers A/B (no register-to-register copies): while the adds of one batch run,
      // the LDS of the next batch are already in flight.
      float4 bufA[K][B], bufB[K][B];
      auto fetch = [&](float4 (&dst)[K][B], int at) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int q = 0; q < B; ++q)
            dst[k][q] = *reinterpret_cast<const float4*>(&sm.in[b][k][at + 4 * q]);
      };
      auto consume = [&](const float4 (&src)[K][B], int at) {
#pragma unroll
        for (int q = 0; q < B; ++q)
        {
          float4 p;
#pragma unroll
          for (int k = 0; k < K; ++k)
            c[k] = __fadd_rn(c[k], src[k][q].x);
          p.x = c[0];
#pragma unroll
          for (int k = 0; k < K; ++k)
            c[k] = __fadd_rn(c[k], src[k][q].y);
          p.y = c[0];
#pragma unroll
          for (int k = 0; k < K; ++k)
            c[k] = __fadd_rn(c[k], src[k][q].z);
          p.z = c[0];
#pragma unroll
          for (int k = 0; k < K; ++k)
            c[k] = __fadd_rn(c[k], src[k][q].w);
          p.w = c[0];
          if (prefix_out)
            *reinterpret_cast<float4*>(&sm.prefix[b][at + 4 * q]) = p;
        }
      };
      int j = 0;
      const int n_batches = len / STEP;
      if (n_batches > 0)
        fetch(bufA, 0);
      int done = 0;
      while (done + 2 <= n_batches)
      {
        fetch(bufB, j + STEP);
        consume(bufA, j);
        if (done + 3 <= n_batches)
          fetch(bufA, j + 2 * STEP);
        consume(bufB, j + STEP);
        j += 2 * STEP;
        done += 2;
      }
      if (done < n_batches)
      {
        consume(bufA, j);
        j += STEP;
      }
      for (; j < len; ++j)
      {
#pragma unroll
        for (int k = 0; k < K; ++k)
          c[k] = __fadd_rn(c[k], sm.in[b][k][j]);
        if (prefix_out)
          sm.prefix[b][j] = c[0];
      }
#pragma unroll
      for (int k = 0; k < K; ++k)
        acc[k] = c[k];
    }
    __syncthreads();
  }
  // flush the prefix of the last tile
  if (prefix_out && n_tiles > 0)
  {
    const uint64_t t = n_tiles - 1;
    const int b = static_cast<int>(t & 1);
    const uint64_t base = t * kChainTile;
    const int len = static_cast<int>(n - base);
    for (int j = tid; j < len; j += blockDim.x)
      prefix_out[base + j] = sm.prefix[b][j];
    __syncthreads();
  }
}

}  // namespace amcl3d_b200
