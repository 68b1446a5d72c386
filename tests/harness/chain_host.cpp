// chain_host.cpp -- CPU emulation of the exact-chain machinery (amcl3d_b200/csrc/chain_fn.h) for tests/test_chain_algebra.py.
// Everything the GPU kernels do with ChainFn / SegFn is replayed here sequentially: per-element functions, composition
// per window / per segment, certificates, fall-backs.  Compiled with g++ -ffp-contract=off (float adds are IEEE).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../amcl3d_b200/csrc/chain_fn.h"

using namespace amcl3d_b200;

static inline uint32_t bits(float f)
{
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
static inline float from_bits(uint32_t u)
{
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

extern "C" {

// the reference's loop: c += t[i]
float chain_seq_sum(const float* t, uint64_t n, float c0, float* prefix)
{
  volatile float c = c0;
  for (uint64_t i = 0; i < n; ++i)
  {
    c = c + t[i];
    if (prefix)
      prefix[i] = c;
  }
  return c;
}

// block_exact_chain replayed: windows of `window` elements, prefixes by composition, first violation ends the window
float chain_window_sum(const float* t, uint64_t n, float c0, uint32_t window, uint32_t head, float* prefix, uint64_t* n_windows)
{
  float c = c0;
  uint64_t p = 0;
  for (; p < n && p < head; ++p)
  {
    volatile float v = c + t[p];
    c = v;
    if (prefix)
      prefix[p] = c;
  }
  uint64_t wins = 0;
  while (p < n)
  {
    const uint32_t cu = bits(c);
    if (!chain_windowable(cu))
    {
      volatile float v = c + t[p];
      c = v;
      if (prefix)
        prefix[p] = c;
      ++p;
      continue;
    }
    ++wins;
    const uint32_t e_run = (cu >> 23) & 0xffu, neg = cu >> 31;
    const int32_t c0i = static_cast<int32_t>((cu & 0x7fffffu) | 0x800000u);
    const uint64_t end = std::min<uint64_t>(n, p + window);
    ChainFn run = chain_identity();
    int32_t prev = c0i;
    uint64_t i = p;
    bool crossed = false;
    for (; i < end; ++i)
    {
      run = chain_compose(run, chain_element(bits(t[i]), e_run, neg));
      const int32_t v = chain_apply(run, c0i);
      if (!chain_step_valid(v, chain_decreasing(bits(t[i]), neg)))
      {
        crossed = true;
        break;
      }
      prev = v;
      if (prefix)
        prefix[i] = from_bits(chain_make_bits(e_run, neg, v));
    }
    c = from_bits(chain_make_bits(e_run, neg, prev));
    if (crossed)
    {
      volatile float v = c + t[i];
      c = v;
      if (prefix)
        prefix[i] = c;
      p = i + 1;
    }
    else
      p = end;
  }
  if (n_windows)
    *n_windows = wins;
  return c;
}

// Segmented scheme: stage A builds one SegFn per segment from a hypothesis (binade / sign of the fp64 prefix estimate
// est_scale * sum(t[0..first)) + est_c0), stage B walks them with the exact carry and falls back to the windowed scan.
// Composition inside a segment is done as a TREE over `leaf` -element leaves (what the parallel scan does) to exercise
// associativity.  Returns the chain result; *n_fallback = segments that took the slow path.
float chain_segmented_sum(const float* t, uint64_t n, float c0, uint32_t seg, uint32_t leaf, double est_bias,
                          uint64_t* n_fallback, float* prefix)
{
  const uint64_t n_seg = (n + seg - 1) / seg;
  std::vector<SegFn> fns(n_seg);
  double est = static_cast<double>(c0);
  for (uint64_t s = 0; s < n_seg; ++s)
  {
    const uint64_t first = s * seg, end = std::min<uint64_t>(n, first + seg);
    const float est_f = static_cast<float>(est * (1.0 + est_bias));
    const uint32_t eu = bits(est_f);
    SegFn sf;
    sf.f = chain_identity();
    sf.lo = sf.hi = 0;
    sf.e_hyp = chain_windowable(eu) ? ((eu >> 23) & 0xffu) : 0u;
    sf.neg = eu >> 31;
    {
      bool all_zero = true;
      for (uint64_t i = first; i < end; ++i)
        all_zero &= (bits(t[i]) & 0x7fffffffu) == 0u;
      if (all_zero)
        sf.neg |= 4u;
    }
    if (sf.e_hyp)
    {
      // leaves composed left to right, then the leaves' totals combined pairwise (tree)
      std::vector<ChainFn> leaves;
      std::vector<ChainFn> prefix_fn;  // F_j for the certificate
      ChainFn before = chain_identity();
      for (uint64_t i = first; i < end; i += leaf)
      {
        ChainFn run = chain_identity();
        for (uint64_t j = i; j < std::min<uint64_t>(end, i + leaf); ++j)
        {
          run = chain_compose(run, chain_element(bits(t[j]), sf.e_hyp, sf.neg & 1u));
          if (chain_decreasing(bits(t[j]), sf.neg & 1u))
            sf.neg |= 2u;
          const ChainFn pj = chain_compose(before, run);
          const int32_t off = chain_offset(pj);
          sf.lo = std::min(sf.lo, off);
          sf.hi = std::max(sf.hi, chain_sat(static_cast<int64_t>(off) + (pj.tie ? 1 : 0)));
        }
        leaves.push_back(run);
        before = chain_compose(before, run);
      }
      while (leaves.size() > 1)
      {
        std::vector<ChainFn> up;
        for (size_t k = 0; k + 1 < leaves.size(); k += 2)
          up.push_back(chain_compose(leaves[k], leaves[k + 1]));
        if (leaves.size() & 1)
          up.push_back(leaves.back());
        leaves.swap(up);
      }
      sf.f = leaves.empty() ? chain_identity() : leaves[0];
    }
    fns[s] = sf;
    for (uint64_t i = first; i < end; ++i)
      est += static_cast<double>(t[i]);
  }
  float c = c0;
  uint64_t fb = 0;
  for (uint64_t s = 0; s < n_seg; ++s)
  {
    const uint64_t first = s * seg, end = std::min<uint64_t>(n, first + seg);
    uint32_t out;
    if (seg_apply(fns[s], bits(c), &out))
    {
      if (prefix)
      {
        // stage C: per-element values of a proven segment
        const uint32_t cu = bits(c);
        const int32_t c0i = static_cast<int32_t>((cu & 0x7fffffu) | 0x800000u);
        ChainFn run = chain_identity();
        for (uint64_t i = first; (fns[s].neg & 4u) && i < end; ++i)
          prefix[i] = c;
        for (uint64_t i = first; !(fns[s].neg & 4u) && i < end; ++i)
        {
          run = chain_compose(run, chain_element(bits(t[i]), fns[s].e_hyp, fns[s].neg & 1u));
          prefix[i] = from_bits(chain_make_bits(fns[s].e_hyp, fns[s].neg & 1u, chain_apply(run, c0i)));
        }
      }
      c = from_bits(out);
    }
    else
    {
      ++fb;
      c = chain_window_sum(t + first, end - first, c, 1024, s == 0 ? 96 : 0, prefix ? prefix + first : nullptr, nullptr);
    }
  }
  if (n_fallback)
    *n_fallback = fb;
  return c;
}

}  // extern "C"
