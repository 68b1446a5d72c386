// exact_scan.cuh -- the reference's SEQUENTIAL float accumulation, bit for bit, in O(n / threads) time.
//
// The integer algebra (ChainFn: what adding one float does to a running value inside one binade, closed under
// composition) is in chain_fn.h.  This header holds the block-level machinery built on it:
//
//   block_exact_chain   one CTA walks a plane of terms with a KNOWN starting value: windows of THREADS*ITEMS elements,
//                       all prefixes of a window from one block-wide scan of ChainFn; the window ends at the first
//                       element that takes the value out of its binade (or changes its sign): that one element is added
//                       with a real float add, the binade is re-read from the result, and the scan restarts behind it.
//                       Any input (negative terms, NaN, infinities, denormal running values) gives the sequential sum;
//                       only the speed assumes that the running value keeps its binade for many elements.
//   block_seg_build     one CTA turns a SEGMENT of terms into a SegFn: the segment's action on a starting value that is
//                       not known yet (only its binade and sign are assumed), with the certificate that proves the
//                       assumption afterwards.  Segments are built in parallel by many CTAs (and by all GPUs of a
//                       sharded particle set at once); a single thread then walks the segment summaries in order with
//                       the exact carry (seg_apply) and falls back to block_exact_chain for the few segments whose
//                       assumption failed (binade crossings).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "chain_fn.h"

namespace amcl3d_b200
{
template <int THREADS>
struct ExactScanSmem
{
  ChainFn warp_fn[THREADS / 32];
  float head[96];
  float c;
  uint32_t p;
  uint32_t cross;
  int32_t lo, hi;
  ChainFn total;
  uint32_t stopped_at;  // block_exact_chain: first element NOT processed when the window budget ran out (else n)
};

__device__ __forceinline__ ChainFn chain_shfl_up(const ChainFn f, const int d)
{
  ChainFn o;
  o.tie = __shfl_up_sync(0xffffffffu, f.tie, d);
  o.a = __shfl_up_sync(0xffffffffu, f.a, d);
  o.b = __shfl_up_sync(0xffffffffu, f.b, d);
  return o;
}

// Block-wide EXCLUSIVE scan (order-preserving composition) of one ChainFn per thread.  All threads call it; contains
// three __syncthreads.  `total_out` (nullable) receives the composition of all threads' functions (valid in every
// thread after the call).
template <int THREADS>
__device__ __forceinline__ ChainFn block_scan_chain(const ChainFn mine, ExactScanSmem<THREADS>& sm, ChainFn* total_out)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ChainFn incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const ChainFn o = chain_shfl_up(incl, d);
    if (lane >= d)
      incl = chain_compose(o, incl);
  }
  __syncthreads();  // previous users of sm.warp_fn are done
  if (lane == 31)
    sm.warp_fn[warp] = incl;
  __syncthreads();
  if (warp == 0)
  {
    ChainFn w_incl = (lane < THREADS / 32) ? sm.warp_fn[lane] : chain_identity();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const ChainFn o = chain_shfl_up(w_incl, d);
      if (lane >= d)
        w_incl = chain_compose(o, w_incl);
    }
    ChainFn w_excl = chain_shfl_up(w_incl, 1);
    if (lane == 0)
      w_excl = chain_identity();
    if (lane == THREADS / 32 - 1)
      sm.total = w_incl;
    if (lane < THREADS / 32)
      sm.warp_fn[lane] = w_excl;
  }
  __syncthreads();
  const ChainFn before = sm.warp_fn[warp];
  ChainFn lane_excl = chain_shfl_up(incl, 1);
  if (lane == 0)
    lane_excl = chain_identity();
  if (total_out)
    *total_out = sm.total;
  return chain_compose(before, lane_excl);
}

// All THREADS threads of the block call this.  Returns (to every thread) the sequential float sum
// c_init + t[0] + t[1] + ... ; writes the running value after each element to prefix_out when non-null.
// serial_head: number of leading elements added one by one up front (a chain that starts near zero crosses a binade
// every few elements at first); pass 0 when c_init is already the sum of many terms.
// max_windows: budget of window iterations (a sum that hovers around zero changes binade or sign every few elements and
// would degrade to one block-wide step per element); when it runs out the function returns early with
// sm.stopped_at < n and the value reached so far -- the caller decides what to do with such a chain.
template <int THREADS, int ITEMS>
__device__ float block_exact_chain(const float* __restrict__ t, const uint32_t n, const float c_init,
                                   float* __restrict__ prefix_out, ExactScanSmem<THREADS>& sm,
                                   const uint32_t serial_head = 96, const uint32_t max_windows = 0xffffffffu)
{
  const int tid = threadIdx.x;
  uint32_t windows = 0;
  const uint32_t head = n < serial_head ? n : (serial_head > 96u ? 96u : serial_head);
  // the head is staged through shared memory with one coalesced load so that the serial adds do not each wait for
  // a global-memory round trip
  __syncthreads();
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
    if (p >= n || windows >= max_windows)
      break;
    ++windows;
    const float c = sm.c;
    const uint32_t cu = __float_as_uint(c);
    if (!chain_windowable(cu))
    {
      // zero, denormal, inf or NaN running value: no integer window.  Terms that cannot change such a value (+-0 for a
      // zero / denormal, anything finite for an infinity, anything at all for a NaN) are skipped in parallel -- an
      // all-zero plane (no beacons: every wr is 0) must not cost one block-wide step per element -- then the first
      // term that does matter is added with a plain float add.
      const bool c_nan = (cu & 0x7fffffffu) > 0x7f800000u, c_inf = (cu & 0x7fffffffu) == 0x7f800000u;
      __syncthreads();
      if (tid == 0)
        sm.cross = 0xffffffffu;
      __syncthreads();
      const uint32_t first = p + static_cast<uint32_t>(tid) * ITEMS;
      const uint32_t chunk_end = min(n, p + static_cast<uint32_t>(THREADS) * ITEMS);
      uint32_t my_first = 0xffffffffu;
#pragma unroll
      for (int k = 0; k < ITEMS; ++k)
      {
        const uint32_t i = first + k;
        if (i < chunk_end)
        {
          const uint32_t tu = __float_as_uint(t[i]);
          // (-0) + (+0) is +0: a negative-zero running value only ignores negative zeros
          const bool inert = c_nan || (c_inf ? ((tu >> 23) & 0xffu) != 0xffu
                                             : (cu == 0x80000000u ? tu == 0x80000000u : (tu & 0x7fffffffu) == 0u));
          if (!inert && my_first == 0xffffffffu)
            my_first = i;
          if (prefix_out)
            prefix_out[i] = c;  // overwritten below from the first effective term on
        }
      }
      if (my_first != 0xffffffffu)
        atomicMin(&sm.cross, my_first);
      __syncthreads();
      const uint32_t stop = sm.cross;
      __syncthreads();
      if (tid == 0)
      {
        if (stop < chunk_end)
        {
          const float v = __fadd_rn(c, t[stop]);
          if (prefix_out)
            prefix_out[stop] = v;
          sm.c = v;
          sm.p = stop + 1;
        }
        else
          sm.p = chunk_end;
      }
      __syncthreads();
      continue;
    }
    const uint32_t e_run = (cu >> 23) & 0xffu, neg = cu >> 31;
    const int32_t c0 = static_cast<int32_t>((cu & 0x7fffffu) | 0x800000u);
    // ---- this thread's ITEMS consecutive elements, composed left to right
    const uint32_t first = p + static_cast<uint32_t>(tid) * ITEMS;
    ChainFn local[ITEMS];
    ChainFn run = chain_identity();
    uint32_t dec_mask = 0;  // bit k: element k decreases the magnitude (a result of exactly 2^23 is then not trusted)
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      const uint32_t i = first + k;
      const uint32_t tu = (i < n) ? __float_as_uint(t[i]) : 0u;
      run = chain_compose(run, chain_element(tu, e_run, neg));
      dec_mask |= chain_decreasing(tu, neg) ? (1u << k) : 0u;
      local[k] = run;
    }
    if (tid == 0)
      sm.cross = 0xffffffffu;
    const ChainFn excl = block_scan_chain<THREADS>(run, sm, nullptr);
    // ---- values after each of my elements; first one that leaves the binade
    int32_t vals[ITEMS];
    uint32_t my_cross = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      vals[k] = chain_apply(chain_compose(excl, local[k]), c0);
      if (first + k < n && !chain_step_valid(vals[k], (dec_mask >> k) & 1u) && my_cross == 0xffffffffu)
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
          prefix_out[first + k] = __uint_as_float(chain_make_bits(e_run, neg, vals[k]));
    }
    // ---- hand the state to the next window
    if (cross < chunk_end)
    {
      // the thread that owns the crossing element performs the real float add from the value just before it
      if (cross >= first && cross < first + ITEMS)
      {
        const int k = static_cast<int>(cross - first);
        int32_t prev = chain_apply(excl, c0);
#pragma unroll
        for (int q = 0; q < ITEMS; ++q)
          if (q < k)
            prev = vals[q];
        const float v = __fadd_rn(__uint_as_float(chain_make_bits(e_run, neg, prev)), t[cross]);
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
        int32_t v = vals[0];
#pragma unroll
        for (int q = 0; q < ITEMS; ++q)
          if (first + q == last)
            v = vals[q];
        sm.c = __uint_as_float(chain_make_bits(e_run, neg, v));
        sm.p = chunk_end;
      }
    }
    __syncthreads();
  }
  const float result = sm.c;
  if (tid == 0)
    sm.stopped_at = sm.p < n ? sm.p : n;
  __syncthreads();
  return result;
}

// All THREADS threads call this.  Summarises the `count` (<= THREADS*ITEMS) terms t[0..count) as a SegFn under the
// hypothesis that the incoming running value has biased exponent e_hyp (1..254) and sign `neg`.  Valid in every thread
// after the call.  e_hyp == 0 returns an "always slow path" summary without touching the terms.
template <int THREADS, int ITEMS>
__device__ SegFn block_seg_build(const float* __restrict__ t, const uint32_t count, const uint32_t e_hyp, const uint32_t neg,
                                 ExactScanSmem<THREADS>& sm)
{
  SegFn s;
  s.f = chain_identity();
  s.lo = 0;
  s.hi = 0;
  s.e_hyp = (e_hyp == 0u || e_hyp >= 0xffu) ? 0u : e_hyp;
  s.neg = neg & 1u;
  const int tid = threadIdx.x;
  const uint32_t first = static_cast<uint32_t>(tid) * ITEMS;
  ChainFn local[ITEMS];
  ChainFn run = chain_identity();
  bool any_dec = false, all_zero = true;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k)
  {
    const uint32_t i = first + k;
    const uint32_t tu = (i < count) ? __float_as_uint(t[i]) : 0u;
    all_zero &= (tu & 0x7fffffffu) == 0u;
    if (s.e_hyp)
    {
      run = chain_compose(run, chain_element(tu, s.e_hyp, neg & 1u));
      any_dec |= chain_decreasing(tu, neg & 1u);
    }
    local[k] = run;
  }
  if (tid == 0)
  {
    sm.lo = 0;
    sm.hi = 0;
  }
  const int dec_any = __syncthreads_or(any_dec ? 1 : 0);
  const int zero_all = __syncthreads_and(all_zero ? 1 : 0);
  s.neg |= (dec_any ? 2u : 0u) | (zero_all ? 4u : 0u);
  if (s.e_hyp == 0u)
    return s;  // no hypothesis: the consumer takes the slow path (unless the segment is all zeros)
  ChainFn total;
  const ChainFn excl = block_scan_chain<THREADS>(run, sm, &total);
  int32_t lo = 0, hi = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k)
  {
    if (first + k < count)
    {
      const ChainFn pj = chain_compose(excl, local[k]);
      const int32_t off = chain_offset(pj);
      lo = min(lo, off);
      hi = max(hi, chain_sat(static_cast<int64_t>(off) + (pj.tie ? 1 : 0)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((tid & 31) == 0)
  {
    atomicMin(&sm.lo, lo);
    atomicMax(&sm.hi, hi);
  }
  __syncthreads();
  s.f = total;
  s.lo = sm.lo;
  s.hi = sm.hi;
  __syncthreads();
  return s;
}

}  // namespace amcl3d_b200
