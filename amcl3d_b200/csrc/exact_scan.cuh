// exact_scan.cuh -- the reference's SEQUENTIAL float accumulation, bit for bit, in O(n / threads) time.
//
// c_i = fl32(c_{i-1} + t_i) is not associative, so a tree sum returns different bits.  But while the running value
// stays inside one binade [2^E, 2^(E+1)) it is an integer multiple C of the binade's ulp q = 2^(E-23), and adding a
// non-negative term t is an INTEGER operation on C:   t/q = A + f  (A integer, 0 <= f < 1)
//      f <  1/2 :  C -> C + A
//      f >  1/2 :  C -> C + A + 1
//      f == 1/2 :  C -> (C + A + 1) & ~1          (round half to even: the only place where C's parity matters)
// Both forms belong to the family  F(C) = C + b  |  F(C) = ((C + a + 1) & ~1) + b,  which is closed under composition
// (after a tie the value is even + b, so later ties resolve to constants).  Composition is associative, so all
// prefixes F_i(C0) of a chunk come out of one block-wide scan.  The window ends at the first element that lifts the
// value to 2^24*q or beyond (values are non-decreasing for non-negative terms, so "first" is well defined): that one
// element is added with a real float add, the binade is re-read from the result, and the scan restarts behind it.
// A sum of n similar terms crosses ~log2(n) binades, most of them within the first few dozen elements, which are
// therefore added serially up front.
//
// Terms that are negative, NaN or infinite are treated as window-ending elements (added with a real float add), so
// the result is the sequential sum for ANY input; only the speed assumes non-negative data.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace amcl3d_b200
{
struct ChainFn
{
  uint32_t a;  // bit 31: tie form flag; bits 0..30: a
  uint32_t b;
};

constexpr uint32_t kChainSat = 1u << 27;  // anything >= 2^24 means "left the binade"; saturate far below 2^31

__device__ __forceinline__ uint32_t chain_sat(uint32_t v) { return v > kChainSat ? kChainSat : v; }

__device__ __forceinline__ ChainFn chain_identity()
{
  ChainFn f;
  f.a = 0;
  f.b = 0;
  return f;
}

// g after f
__device__ __forceinline__ ChainFn chain_compose(const ChainFn f, const ChainFn g)
{
  ChainFn r;
  const bool ft = f.a >> 31, gt = g.a >> 31;
  const uint32_t fa = f.a & 0x7fffffffu, ga = g.a & 0x7fffffffu;
  if (!gt)
  {
    r.a = f.a;
    r.b = chain_sat(f.b + g.b);
  }
  else if (!ft)
  {
    r.a = 0x80000000u | chain_sat(ga + f.b);
    r.b = g.b;
  }
  else
  {
    r.a = 0x80000000u | fa;
    r.b = chain_sat(((f.b + ga + 1u) & ~1u) + g.b);
  }
  return r;
}

__device__ __forceinline__ uint32_t chain_apply(const ChainFn f, const uint32_t c)
{
  if (f.a >> 31)
    return chain_sat(((c + (f.a & 0x7fffffffu) + 1u) & ~1u) + f.b);
  return chain_sat(c + f.b);
}

// The integer action of adding float `t` to a running value in the binade with biased exponent `e_run`.
__device__ __forceinline__ ChainFn chain_element(const float t, const uint32_t e_run)
{
  ChainFn f = chain_identity();
  const uint32_t u = __float_as_uint(t);
  const uint32_t et_raw = (u >> 23) & 0xffu;
  if (u == 0u)
    return f;  // + 0.0f
  if ((u >> 31) || et_raw == 0xffu)
  {
    f.b = kChainSat;  // negative / NaN / inf: end the window here, the real float add decides
    return f;
  }
  const uint32_t et = et_raw ? et_raw : 1u;
  const uint32_t m = et_raw ? ((u & 0x7fffffu) | 0x800000u) : (u & 0x7fffffu);
  if (et > e_run)
  {
    f.b = kChainSat;  // the term alone exceeds the binade
    return f;
  }
  const uint32_t s = e_run - et;
  if (s == 0u)
  {
    f.b = m;  // exact integer add
    return f;
  }
  if (s >= 26u)
    return f;  // below a quarter ulp: no effect
  const uint32_t A = m >> s, rem = m & ((1u << s) - 1u), half = 1u << (s - 1u);
  if (rem > half)
    f.b = A + 1u;
  else if (rem < half)
    f.b = A;
  else
    f.a = 0x80000000u | A;  // tie
  return f;
}

__device__ __forceinline__ float chain_make_float(const uint32_t e_run, const uint32_t c)
{
  return __uint_as_float((e_run << 23) | (c & 0x7fffffu));  // 2^23 <= c < 2^24
}

template <int THREADS>
struct ExactScanSmem
{
  ChainFn warp_fn[THREADS / 32];
  float head[96];
  float c;
  uint32_t p;
  uint32_t cross;
  uint32_t last_val;
};

// All THREADS threads of the block call this.  Returns (to every thread) the sequential float sum
// c_init + t[0] + t[1] + ... ; writes the running value after each element to prefix_out when non-null.
template <int THREADS, int ITEMS>
__device__ float block_exact_chain(const float* __restrict__ t, const uint32_t n, const float c_init,
                                   float* __restrict__ prefix_out, ExactScanSmem<THREADS>& sm)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t kSerialHead = 96;
  const uint32_t head = n < kSerialHead ? n : kSerialHead;
  // the head is staged through shared memory with one coalesced load so that the serial adds do not each wait for
  // a global-memory round trip
  if (static_cast<uint32_t>(tid) < head)
    sm.head[tid] = t[tid];
  __syncthreads();
  if (tid == 0)
  {
    float c = c_init;
    for (uint32_t i = 0; i < head; ++i)
    {
      c = __fadd_rn(c, sm.head[i]);
      sm.head[i] = c;
    }
    sm.c = c;
    sm.p = head;
  }
  __syncthreads();
  if (prefix_out && static_cast<uint32_t>(tid) < head)
    prefix_out[tid] = sm.head[tid];
  while (true)
  {
    const uint32_t p = sm.p;
    if (p >= n)
      break;
    const float c = sm.c;
    const uint32_t cu = __float_as_uint(c);
    const uint32_t e_run = (cu >> 23) & 0xffu;
    if ((cu >> 31) || e_run == 0u || e_run == 0xffu)
    {
      // zero, denormal, negative, inf or NaN running value: no integer window -- one plain float add
      __syncthreads();
      if (tid == 0)
      {
        const float v = __fadd_rn(c, t[p]);
        if (prefix_out)
          prefix_out[p] = v;
        sm.c = v;
        sm.p = p + 1;
      }
      __syncthreads();
      continue;
    }
    const uint32_t c0 = (cu & 0x7fffffu) | 0x800000u;
    // ---- this thread's ITEMS consecutive elements, composed left to right
    const uint32_t first = p + static_cast<uint32_t>(tid) * ITEMS;
    ChainFn local[ITEMS];
    ChainFn run = chain_identity();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      const uint32_t i = first + k;
      const ChainFn e = (i < n) ? chain_element(t[i], e_run) : chain_identity();
      run = chain_compose(run, e);
      local[k] = run;
    }
    // ---- block-wide exclusive scan of the per-thread totals (order-preserving composition)
    ChainFn incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      ChainFn o;
      o.a = __shfl_up_sync(0xffffffffu, incl.a, d);
      o.b = __shfl_up_sync(0xffffffffu, incl.b, d);
      if (lane >= d)
        incl = chain_compose(o, incl);
    }
    __syncthreads();  // previous iteration's readers of sm.warp_fn / sm.cross are done
    if (lane == 31)
      sm.warp_fn[warp] = incl;
    if (tid == 0)
    {
      sm.cross = 0xffffffffu;
    }
    __syncthreads();
    // warp 0 turns the per-warp totals into exclusive prefixes (one shuffle scan instead of a loop per thread)
    if (warp == 0)
    {
      ChainFn w_incl = (lane < THREADS / 32) ? sm.warp_fn[lane] : chain_identity();
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        ChainFn o;
        o.a = __shfl_up_sync(0xffffffffu, w_incl.a, d);
        o.b = __shfl_up_sync(0xffffffffu, w_incl.b, d);
        if (lane >= d)
          w_incl = chain_compose(o, w_incl);
      }
      ChainFn w_excl;
      w_excl.a = __shfl_up_sync(0xffffffffu, w_incl.a, 1);
      w_excl.b = __shfl_up_sync(0xffffffffu, w_incl.b, 1);
      if (lane == 0)
        w_excl = chain_identity();
      if (lane < THREADS / 32)
        sm.warp_fn[lane] = w_excl;
    }
    __syncthreads();
    const ChainFn before = sm.warp_fn[warp];
    // exclusive prefix of this thread = (warps before) o (lanes before in this warp)
    ChainFn lane_excl;
    lane_excl.a = __shfl_up_sync(0xffffffffu, incl.a, 1);
    lane_excl.b = __shfl_up_sync(0xffffffffu, incl.b, 1);
    if (lane == 0)
      lane_excl = chain_identity();
    const ChainFn excl = chain_compose(before, lane_excl);
    // ---- values after each of my elements; first one that leaves the binade
    uint32_t vals[ITEMS];
    uint32_t my_cross = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      vals[k] = chain_apply(chain_compose(excl, local[k]), c0);
      if (first + k < n && vals[k] >= (1u << 24) && my_cross == 0xffffffffu)
        my_cross = first + k;
    }
    if (my_cross != 0xffffffffu)
      atomicMin(&sm.cross, my_cross);
    __syncthreads();
    const uint32_t cross = sm.cross;
    const uint32_t chunk_end = min(n, p + static_cast<uint32_t>(THREADS) * ITEMS);
    const uint32_t commit_end = cross < chunk_end ? cross : chunk_end;  // elements [p, commit_end) keep the binade
    if (prefix_out)
    {
#pragma unroll
      for (int k = 0; k < ITEMS; ++k)
        if (first + k < commit_end)
          prefix_out[first + k] = chain_make_float(e_run, vals[k]);
    }
    // ---- hand the state to the next window
    if (cross < chunk_end)
    {
      // the thread that owns the crossing element performs the real float add from the value just before it
      if (cross >= first && cross < first + ITEMS)
      {
        const int k = static_cast<int>(cross - first);
        uint32_t prev = chain_apply(excl, c0);
#pragma unroll
        for (int q = 0; q < ITEMS; ++q)
          if (q < k)
            prev = vals[q];
        const float v = __fadd_rn(chain_make_float(e_run, prev), t[cross]);
        if (prefix_out)
          prefix_out[cross] = v;
        sm.c = v;
        sm.p = cross + 1;
      }
    }
    else
    {
      // whole chunk committed: the owner of the last valid element publishes the new running value
      const uint32_t last = chunk_end - 1;
      if (last >= first && last < first + ITEMS)
      {
        uint32_t v = vals[0];
#pragma unroll
        for (int q = 0; q < ITEMS; ++q)
          if (first + q == last)
            v = vals[q];
        sm.c = chain_make_float(e_run, v);
        sm.p = chunk_end;
      }
    }
    __syncthreads();
  }
  const float result = sm.c;
  __syncthreads();
  return result;
}

}  // namespace amcl3d_b200
