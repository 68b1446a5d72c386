// chain_fn.h -- the integer algebra behind the exact reproduction of the reference's SEQUENTIAL float sums
// (ParticleFilter.cpp:151-152,179,190-193,214) in parallel.  Plain C++ (compiles with g++ for the CPU tests in
// tests/test_chain_algebra.py and with nvcc for the kernels in exact_scan.cuh).
//
// c_i = fl32(c_{i-1} + t_i) is not associative.  But while the running value stays inside one binade
// [2^E, 2^(E+1)) and keeps its sign, its magnitude is an integer multiple C (2^23 <= C < 2^24) of the binade's ulp
// q = 2^(E-23), and adding a term t is an INTEGER operation on C.  With s = sign(c) * t / q = A + f (A = floor, integer,
// possibly negative; 0 <= f < 1), round-to-nearest-even gives
//      f <  1/2 :  C -> C + A
//      f >  1/2 :  C -> C + A + 1
//      f == 1/2 :  C -> (C + A + 1) & ~1          (the even neighbour: the only place where C's parity matters)
// Both forms belong to the family  F(C) = C + b  |  F(C) = ((C + a + 1) & ~1) + b  (a, b signed), which is closed
// under composition (after a tie the value is even + b, so later ties resolve to constants).  Composition is
// associative, so all prefixes of a chunk come out of one parallel scan, and the action of a whole SEGMENT of terms on
// an as yet unknown incoming value is one 8-byte function -- which is what lets a sharded particle set pass an exact
// carry from GPU to GPU without serialising the shards.
// The description holds while every intermediate value stays in [2^23, 2^24) -- and, for a term that DEcreases the
// magnitude, strictly above 2^23: a model value of exactly 2^23 may stand for an exact sum just below the binade, where
// the float grid is twice as fine.  ChainFn arithmetic saturates far outside that range, so a violated assumption is
// always detected (never silently wrong).
#ifndef AMCL3D_CHAIN_FN_H
#define AMCL3D_CHAIN_FN_H

#include <stdint.h>

#ifdef __CUDACC__
#define A3D_HD __host__ __device__ __forceinline__
#else
#define A3D_HD inline
#endif

namespace amcl3d_b200
{
struct ChainFn
{
  uint32_t tie;  // 1: F(C) = ((C + a + 1) & ~1) + b     0: F(C) = C + b
  int32_t a;
  int32_t b;
};

constexpr int32_t kChainSat = 1 << 27;  // |anything| >= 2^24 means "left the binade"; saturate far below 2^31

A3D_HD int32_t chain_sat(int64_t v)
{
  return v > kChainSat ? kChainSat : (v < -static_cast<int64_t>(kChainSat) ? -kChainSat : static_cast<int32_t>(v));
}

A3D_HD ChainFn chain_identity()
{
  ChainFn f;
  f.tie = 0;
  f.a = 0;
  f.b = 0;
  return f;
}

// g after f
A3D_HD ChainFn chain_compose(const ChainFn f, const ChainFn g)
{
  ChainFn r;
  if (!g.tie)
  {
    r.tie = f.tie;
    r.a = f.a;
    r.b = chain_sat(static_cast<int64_t>(f.b) + g.b);
  }
  else if (!f.tie)
  {
    // g(C + fb) = ((C + fb + ga + 1) & ~1) + gb
    r.tie = 1;
    r.a = chain_sat(static_cast<int64_t>(g.a) + f.b);
    r.b = g.b;
  }
  else
  {
    // f(C) = E + fb with E even: g(E + fb) = ((E + fb + ga + 1) & ~1) + gb = E + ((fb + ga + 1) & ~1) + gb
    r.tie = 1;
    r.a = f.a;
    const int64_t m = (static_cast<int64_t>(f.b) + g.a + 1) & ~static_cast<int64_t>(1);
    r.b = chain_sat(m + g.b);
  }
  return r;
}

// value after the function, as a saturating signed integer (valid values lie in [2^23, 2^24))
A3D_HD int32_t chain_apply(const ChainFn f, const int32_t c)
{
  if (f.tie)
    return chain_sat(((static_cast<int64_t>(c) + f.a + 1) & ~static_cast<int64_t>(1)) + f.b);
  return chain_sat(static_cast<int64_t>(c) + f.b);
}

// Offset range of F relative to its argument: F(C) - C lies in [chain_offset(F), chain_offset(F) + 1].
A3D_HD int32_t chain_offset(const ChainFn f) { return f.tie ? chain_sat(static_cast<int64_t>(f.a) + f.b) : f.b; }

// The integer action of adding the float with bit pattern `u` to a running value that lies in the binade with
// biased exponent `e_run` and has sign bit `neg`.  Terms the model cannot express (NaN / infinity, or larger than the
// binade) return a saturated function, which ends the window at that element.
A3D_HD ChainFn chain_element(const uint32_t u, const uint32_t e_run, const uint32_t neg)
{
  ChainFn f = chain_identity();
  const uint32_t et_raw = (u >> 23) & 0xffu;
  if ((u & 0x7fffffffu) == 0u)
    return f;  // +-0
  if (et_raw == 0xffu)
  {
    f.b = kChainSat;
    return f;
  }
  const uint32_t et = et_raw ? et_raw : 1u;
  const uint32_t m = et_raw ? ((u & 0x7fffffu) | 0x800000u) : (u & 0x7fffffu);
  const bool same_sign = ((u >> 31) == neg);
  if (et > e_run)
  {
    f.b = same_sign ? kChainSat : -kChainSat;  // the term alone exceeds the binade
    return f;
  }
  const uint32_t s = e_run - et;
  if (s == 0u)
  {
    f.b = same_sign ? static_cast<int32_t>(m) : -static_cast<int32_t>(m);  // exact integer add
    return f;
  }
  if (s >= 26u)
    return f;  // below a quarter ulp: no effect
  const uint32_t A = m >> s, rem = m & ((1u << s) - 1u), half = 1u << (s - 1u);
  if (same_sign)
  {
    if (rem > half)
      f.b = static_cast<int32_t>(A) + 1;
    else if (rem < half)
      f.b = static_cast<int32_t>(A);
    else
    {
      f.tie = 1;
      f.a = static_cast<int32_t>(A);
    }
  }
  else
  {
    // s' = -(A + rem/2^s): floor = -A (rem == 0) or -A-1 with fraction 1 - rem/2^s
    if (rem == 0u || rem < half)
      f.b = -static_cast<int32_t>(A);
    else if (rem > half)
      f.b = -static_cast<int32_t>(A) - 1;
    else
    {
      f.tie = 1;
      f.a = -static_cast<int32_t>(A) - 1;
    }
  }
  return f;
}

A3D_HD bool chain_in_binade(const int32_t c) { return c >= (1 << 23) && c < (1 << 24); }
// validity of a value produced by one step: `decreasing` = the step's term acts against the running value's sign
A3D_HD bool chain_step_valid(const int32_t c, const bool decreasing) { return c >= (1 << 23) + (decreasing ? 1 : 0) && c < (1 << 24); }
// true when adding the float with bit pattern u DEcreases the magnitude of a running value with sign bit `neg`
A3D_HD bool chain_decreasing(const uint32_t u, const uint32_t neg) { return (u & 0x7fffffffu) != 0u && (u >> 31) != neg; }

// float with biased exponent e_run, sign `neg` and integer significand c (2^23 <= c < 2^24)
A3D_HD uint32_t chain_make_bits(const uint32_t e_run, const uint32_t neg, const int32_t c)
{
  return (neg << 31) | (e_run << 23) | (static_cast<uint32_t>(c) & 0x7fffffu);
}

// True when a float (bit pattern) is a normal number, i.e. has an integer window at all.
A3D_HD bool chain_windowable(const uint32_t cu)
{
  const uint32_t e = (cu >> 23) & 0xffu;
  return e != 0u && e != 0xffu;
}

// ---- segment summaries -------------------------------------------------------------------------------------------
// The action of a whole segment of consecutive terms on an incoming value of binade `e_hyp` / sign `neg` (a
// HYPOTHESIS made before the incoming value is known), plus the certificate that lets the consumer prove the
// hypothesis after the fact: every prefix F_j of the segment satisfies  lo <= F_j(C) - C <= hi.
struct SegFn
{
  ChainFn f;
  int32_t lo, hi;
  uint32_t e_hyp;  // biased exponent assumed for the incoming value (0 = no hypothesis: always take the slow path)
  uint32_t neg;    // bit 0: sign assumed for the incoming value; bit 1: the segment holds magnitude-decreasing terms;
                   // bit 2: every term of the segment is +-0 (the segment leaves ANY incoming value unchanged)
};

// Applies a segment summary to the exact incoming value `c_bits` (float bit pattern).  Returns true and the exact
// outgoing value when the hypothesis is proven (same binade and sign, all prefixes inside the binade); false means the
// caller must run the terms of the segment through the windowed scan / real float adds.
A3D_HD bool seg_apply(const SegFn s, const uint32_t c_bits, uint32_t* out_bits)
{
  if ((s.neg & 4u) && c_bits != 0x80000000u)  // all-zero segment ((-0) + (+0) = +0 is left to the slow path)
  {
    *out_bits = c_bits;
    return true;
  }
  if (s.e_hyp == 0u || !chain_windowable(c_bits))
    return false;
  if (((c_bits >> 23) & 0xffu) != s.e_hyp || (c_bits >> 31) != (s.neg & 1u))
    return false;
  const bool dec = (s.neg & 2u) != 0u;
  const int32_t c0 = static_cast<int32_t>((c_bits & 0x7fffffu) | 0x800000u);
  if (!chain_step_valid(chain_sat(static_cast<int64_t>(c0) + s.lo), dec) ||
      !chain_in_binade(chain_sat(static_cast<int64_t>(c0) + s.hi)))
    return false;
  const int32_t c1 = chain_apply(s.f, c0);
  if (!chain_step_valid(c1, dec))
    return false;
  *out_bits = chain_make_bits(s.e_hyp, s.neg & 1u, c1);
  return true;
}

}  // namespace amcl3d_b200
#endif
