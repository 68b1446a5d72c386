// filter_exact.cu -- ParticleFilter::update's normalisations / mean (ParticleFilter.cpp:129-195) and
// ParticleFilter::resample (ParticleFilter.cpp:198-222) with the reference's SEQUENTIAL float sums reproduced bit for bit
// at any particle count and on any number of GPUs.
//
// The reference runs three dependent float chains over the particles (wtp / wtr, then wt, then the four mean sums) and a
// fourth one in resample (the cumulative weight).  At 10^6 particles their rounding is no longer noise: the chain value
// of sum(w) differs from the exact sum by ~1e-5 relative and the mean by ~1e-3 m, so an fp64 tree reduction does NOT
// return the reference's numbers.  Here every chain is evaluated exactly, in parallel:
//   * particles are cut into SEGMENTS of 2048; one CTA turns a segment's terms into a SegFn (exact_scan.cuh /
//     chain_fn.h): the segment's action on an incoming running value whose binade and sign are guessed from fp64 partial
//     sums, together with a certificate that proves the guess once the true value is known;
//   * one thread walks the segment summaries in order with the exact carry (an integer add and two compares per
//     segment); the few segments whose guess failed -- binade crossings, a sum hovering around zero -- are replayed with
//     the windowed scan;
//   * a particle set sharded over several GPUs passes the exact carry from rank to rank through peer-memory mailboxes
//     (PeerBox): every rank has all its summaries ready before the carry arrives, so a hop costs microseconds, not a
//     pass over the shard.  The last rank stores the finished values into every rank's box.
// One cooperative launch per update (eight grid-wide syncs) and one per resample; no host round trip, no NCCL call on
// the data path.  resample gathers its source particles straight from the owning GPU over NVLink (the cumulative
// weights of all shards are searched in place), so nothing but the particles that are actually drawn crosses the wire.
#include <cooperative_groups.h>

#include <cmath>
#include <cstring>

#include "chain.cuh"
#include "exact_scan.cuh"
#include "filter_common.cuh"

namespace cg = cooperative_groups;

namespace amcl3d_b200
{
constexpr int kSegThreads = 512;
constexpr int kSegItems = 4;
constexpr uint32_t kSeg = kSegThreads * kSegItems;  // particles per segment
constexpr int kPartCols = 20;                       // fp64 partials per segment: A, B, Px..Pa, Rx..Ra, evals, P|x|..P|a|, R|x|..R|a|, spare
constexpr int kPartUsed = 19;

struct SegArrays
{
  SegFn* fn;       // [4][seg_cap]
  double* part;    // [seg_cap][kPartCols]
  double* pre;     // [seg_cap][kPartCols]: exclusive prefix of `part` over this rank's segments
  double* tot;     // [kPartCols]: this rank's totals
  float* carry;    // [seg_cap]: resample, exact cumulative weight entering the segment
  uint32_t* slow;  // [seg_cap]: resample, 1 = the segment was replayed (its prefix values are already written)
  uint32_t seg_cap;
};

static size_t seg_bytes(uint64_t seg_cap)
{
  return seg_cap * (4 * sizeof(SegFn) + 2 * kPartCols * sizeof(double) + sizeof(float) + sizeof(uint32_t)) +
         kPartCols * sizeof(double) + 256;
}

static SegArrays seg_arrays(void* base, uint64_t seg_cap)
{
  SegArrays a;
  char* p = static_cast<char*>(base);
  a.part = reinterpret_cast<double*>(p);
  p += seg_cap * kPartCols * sizeof(double);
  a.pre = reinterpret_cast<double*>(p);
  p += seg_cap * kPartCols * sizeof(double);
  a.tot = reinterpret_cast<double*>(p);
  p += kPartCols * sizeof(double);
  a.fn = reinterpret_cast<SegFn*>(p);
  p += seg_cap * 4 * sizeof(SegFn);
  a.carry = reinterpret_cast<float*>(p);
  p += seg_cap * sizeof(float);
  a.slow = reinterpret_cast<uint32_t*>(p);
  a.seg_cap = static_cast<uint32_t>(seg_cap);
  return a;
}

// ------------------------------------------------------------------------------------------ peer-memory primitives
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spins until *f == seq; false when `timeout` clocks pass first (a peer is missing: fail loudly, never hang the GPU)
__device__ __forceinline__ bool peer_wait(const unsigned long long* f, const unsigned long long seq, const long long timeout)
{
  const long long t0 = clock64();
  for (;;)
  {
    if (ld_acquire_sys(f) == seq)
      return true;
    if (clock64() - t0 > timeout)
      return false;
    __nanosleep(40);
  }
}

struct WalkSmem
{
  float cur[4];
  uint32_t pos[4];
  uint32_t abandon;        // bit k: chain k is not evaluated (its result comes from the fp64 partials instead)
  const float* terms[4];
  double tot[kPartCols];   // global fp64 totals (all ranks)
  double ein[kPartCols];   // fp64 totals of the ranks before this one
  float bcast[4];
};

// Stage B of a chain phase, executed by ONE CTA: K chains (wk.terms[k], wk.cur[k] = exact incoming values) are carried
// over the n_seg segments of this rank; fns[k * fn_stride + s].  On return wk.cur[k] holds the exact outgoing values.
// prefix_out / seg_carry / seg_slow (K == 1, resample): replayed segments write their running values, proven ones only
// record the value that enters them (a later parallel pass expands those).
__device__ void walk_segments(const int K, const SegFn* __restrict__ fns, const uint32_t fn_stride, const uint32_t n_seg,
                              const uint64_t n, WalkSmem& wk, ExactScanSmem<kSegThreads>& xs, ChainSmem<1>& cs,
                              float* __restrict__ prefix_out, float* __restrict__ seg_carry,
                              uint32_t* __restrict__ seg_slow, const uint32_t serial_mask = 0u,
                              ChainSmem<4>* cs4 = nullptr)
{
  const int tid = threadIdx.x;
  // serial_mask: chains that hover around zero but are too long for their fp64 stand-in to stay inside the tolerance.
  // A hovering chain changes binade every few elements, so neither summaries nor windows help: all such chains of the
  // phase are added TOGETHER by the single-lane chain over this rank's whole plane (their dependent adds interleave:
  // ~4.5 ns per particle for up to four chains).
  if (serial_mask && cs4 && K == 4)
  {
    __syncthreads();
    const float* const src[4] = { wk.terms[0], wk.terms[1], wk.terms[2], wk.terms[3] };
    float acc[4] = { wk.cur[0], wk.cur[1], wk.cur[2], wk.cur[3] };
    block_chain<4>(src, n, acc, nullptr, *cs4);
    if (tid == 0)
      for (int k = 0; k < 4; ++k)
        if ((serial_mask >> k) & 1u)
          wk.cur[k] = acc[k];
    __syncthreads();
  }
  // The walk is ONE thread applying summaries in order: the carry is a true dependency, the summaries are not.  Fetching
  // them one per iteration made every step wait for an L2 round trip (~0.35 us per segment: 1.25 ms for the seven
  // chains of a 1 M-particle set, and as much again on 8 GPUs, where the ranks walk one after the other); here eight
  // summaries are loaded back to back, then applied.
  auto advance = [&](const int k) {
    uint32_t s = wk.pos[k];
    uint32_t cb = __float_as_uint(wk.cur[k]);
    const SegFn* const f = fns + static_cast<size_t>(k) * fn_stride;
    constexpr uint32_t kAhead = 8;
    bool failed = false;
    while (s < n_seg && !failed)
    {
      SegFn buf[kAhead];
#pragma unroll
      for (uint32_t q = 0; q < kAhead; ++q)
        buf[q] = f[min(s + q, n_seg - 1u)];  // independent loads, clamped into range
#pragma unroll
      for (uint32_t q = 0; q < kAhead; ++q)
      {
        if (!failed && s < n_seg)  // buf[q] is the summary of segment s: s advances by one per proven segment
        {
          uint32_t ob;
          if (!seg_apply(buf[q], cb, &ob))
            failed = true;
          else
          {
            if (seg_carry)
            {
              seg_carry[s] = __uint_as_float(cb);
              seg_slow[s] = 0u;
            }
            cb = ob;
            ++s;
          }
        }
      }
    }
    wk.pos[k] = s;
    wk.cur[k] = __uint_as_float(cb);
  };
  __syncthreads();
  if (tid < K)
  {
    wk.pos[tid] = (((wk.abandon | serial_mask) >> tid) & 1u) ? n_seg : 0u;
    advance(tid);
  }
  __syncthreads();
  for (;;)
  {
    int kk = -1;
    for (int k = 0; k < K; ++k)
      if (kk < 0 && wk.pos[k] < n_seg)
        kk = k;
    if (kk < 0)
      break;
    const uint32_t s = wk.pos[kk];
    const float c = wk.cur[kk];
    const uint64_t first = static_cast<uint64_t>(s) * kSeg;
    const uint32_t count = static_cast<uint32_t>(min(static_cast<uint64_t>(kSeg), n - first));
    __syncthreads();
    if (seg_carry && tid == 0)
    {
      seg_carry[s] = c;
      seg_slow[s] = 1u;
    }
    // Replay of one segment with the known incoming value.  A chain that is still at zero crosses a binade every few
    // elements, and so does a sum that hovers: those are added by the single-lane chain (chain.cuh, ~4.5 ns per
    // element, a predictable 9 us per segment).  Otherwise the windowed scan gets a few windows (a binade crossing in
    // the middle of a long chain costs two) before the rest of the segment goes to the single lane as well.
    uint32_t done_upto = 0;
    float r = c;
    if (c != 0.f)
    {
      r = block_exact_chain<kSegThreads, kSegItems>(wk.terms[kk] + first, count, c,
                                                    prefix_out ? prefix_out + first : nullptr, xs, 0u, 4u);
      done_upto = xs.stopped_at;
      __syncthreads();
    }
    if (done_upto < count)
    {
      const float* const src[1] = { wk.terms[kk] + first + done_upto };
      float acc[1] = { r };
      block_chain<1>(src, count - done_upto, acc, prefix_out ? prefix_out + first + done_upto : nullptr, cs);
      if (tid == 0)
        wk.bcast[3] = acc[0];
      __syncthreads();
      r = wk.bcast[3];
    }
    if (tid == 0)
    {
      wk.cur[kk] = r;
      wk.pos[kk] = s + 1;
    }
    __syncthreads();
    if (tid == kk)
      advance(kk);
    __syncthreads();
  }
}

// One chain phase of a (possibly sharded) update, executed by CTA 0: carry in from the previous rank, walk, carry out to
// the next rank, final values from the last rank.  Results land in wk.bcast[0..K) of this CTA.  Returns false on a
// peer time-out (comm_error is raised by the caller).
// give_up_mask: chains not to attempt at all (hovering mean components, decided from the GLOBAL fp64 partials so that
// every rank -- and every rank count -- takes the same decision).  wk.abandon returns the chains whose result is NOT the
// float chain.
__device__ bool chain_phase(const int phase, const int K, const SegFn* fns, const uint32_t fn_stride, const uint32_t n_seg,
                            const uint64_t n, const PeerView& pv, WalkSmem& wk, ExactScanSmem<kSegThreads>& xs,
                            ChainSmem<1>& cs, const uint32_t give_up_mask = 0u, const uint32_t serial_mask = 0u,
                            ChainSmem<4>* cs4 = nullptr)
{
  const int tid = threadIdx.x;
  const bool sharded = pv.n_ranks > 1;
  const uint32_t mp = static_cast<uint32_t>(pv.seq & 1ull);
  __shared__ int ok_sm;
  if (tid == 0)
  {
    ok_sm = 1;
    wk.abandon = give_up_mask;
    for (int k = 0; k < 4; ++k)
      wk.cur[k] = 0.f;
    if (sharded && pv.rank > 0)
    {
      PeerBox* mine = pv.box[pv.rank];
      if (!peer_wait(&mine->carry_flag[mp][phase], pv.seq, pv.timeout_clocks))
        ok_sm = 0;
      for (int k = 0; k < K; ++k)
        wk.cur[k] = *const_cast<const volatile float*>(&mine->carry[mp][phase][k]);
      wk.abandon |= *const_cast<const volatile unsigned int*>(&mine->carry_mode[mp][phase]);
    }
  }
  __syncthreads();
  walk_segments(K, fns, fn_stride, n_seg, n, wk, xs, cs, nullptr, nullptr, nullptr, serial_mask, cs4);
  if (tid == 0)
  {
    if (sharded)
    {
      if (pv.rank < pv.n_ranks - 1)
      {
        PeerBox* next = pv.box[pv.rank + 1];
        for (int k = 0; k < K; ++k)
          *const_cast<volatile float*>(&next->carry[mp][phase][k]) = wk.cur[k];
        *const_cast<volatile unsigned int*>(&next->carry_mode[mp][phase]) = wk.abandon;
        __threadfence_system();
        st_release_sys(&next->carry_flag[mp][phase], pv.seq);
      }
      else
      {
        for (int r = 0; r < pv.n_ranks; ++r)
        {
          for (int k = 0; k < K; ++k)
            *const_cast<volatile float*>(&pv.box[r]->final_[mp][phase][k]) = wk.cur[k];
          *const_cast<volatile unsigned int*>(&pv.box[r]->final_mode[mp][phase]) = wk.abandon;
        }
        __threadfence_system();
        for (int r = 0; r < pv.n_ranks; ++r)
          st_release_sys(&pv.box[r]->final_flag[mp][phase], pv.seq);
      }
      PeerBox* mine = pv.box[pv.rank];
      if (!peer_wait(&mine->final_flag[mp][phase], pv.seq, pv.timeout_clocks))
        ok_sm = 0;
      for (int k = 0; k < K; ++k)
        wk.bcast[k] = *const_cast<const volatile float*>(&mine->final_[mp][phase][k]);
      wk.abandon = *const_cast<const volatile unsigned int*>(&mine->final_mode[mp][phase]);
    }
    else
      for (int k = 0; k < K; ++k)
        wk.bcast[k] = wk.cur[k];
  }
  __syncthreads();
  return ok_sm != 0;
}

__device__ __forceinline__ void hypothesis(const double est, uint32_t* e_hyp, uint32_t* neg)
{
  const uint32_t eu = __float_as_uint(static_cast<float>(est));
  *e_hyp = chain_windowable(eu) ? ((eu >> 23) & 0xffu) : 0u;
  *neg = eu >> 31;
}

// block-wide sum of NCOL doubles per thread into out[0..NCOL) (thread 0 writes); contains two __syncthreads
template <int NCOL>
__device__ __forceinline__ void block_sum_cols(const double (&acc)[NCOL], double* __restrict__ out, double (*red)[kSegThreads / 32])
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NCOL; ++k)
  {
    const double v = warp_sum(acc[k]);
    if (lane == 0)
      red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NCOL)
  {
    double v = 0;
    for (int w = 0; w < kSegThreads / 32; ++w)
      v += red[threadIdx.x][w];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ update
struct UpdateSegParams
{
  GridView g;
  Planes p;
  uint64_t n;
  const void* part_sum;
  const uint32_t* part_cnt;
  uint32_t n_splits;
  int part_kind;
  RangeParams rg;
  double alpha;
  float* terms;
  uint64_t terms_stride;
  amcl3d_pf_scalars* scal;
  SegArrays sa;
  uint32_t n_seg;
  uint64_t n_total;  // particles of the whole (possibly sharded) set
  PeerView pv;
};

__global__ void __launch_bounds__(kSegThreads) update_seg_kernel(const __grid_constant__ UpdateSegParams P)
{
  cg::grid_group grid = cg::this_grid();
  __shared__ ExactScanSmem<kSegThreads> xs;
  __shared__ ChainSmem<4> cs4;  // four interleaved single-lane chains (hovering mean components of long particle sets)
  ChainSmem<1>& cs = *reinterpret_cast<ChainSmem<1>*>(&cs4);
  __shared__ WalkSmem wk;
  __shared__ double red[kPartCols][kSegThreads / 32];
  const int tid = threadIdx.x;
  const uint64_t n = P.n;
  const uint32_t n_seg = P.n_seg;
  const bool sharded = P.pv.n_ranks > 1;
  const uint32_t mp = static_cast<uint32_t>(P.pv.seq & 1ull);
  float* const t0 = P.terms;
  float* const t1 = P.terms + P.terms_stride;
  float* const t2 = P.terms + 2 * P.terms_stride;
  float* const t3 = P.terms + 3 * P.terms_stride;
  // peer time-outs: err[k] is written only in one interval between two grid-wide syncs and read after the next sync, so
  // all CTAs take the same early-exit decision (amcl3d_pf_scalars::err_stage); comm_error is what the host reads
  volatile unsigned int* const err = P.scal->err_stage;
  if (blockIdx.x == 0 && tid == 0)
  {
    P.scal->comm_error = 0u;  // a time-out of an earlier update has been reported by now
    for (int k = 0; k < 4; ++k)
      err[k] = 0u;
  }
  auto bail = [&](const int k) -> bool {
    if (!err[k])
      return false;
    if (blockIdx.x == 0 && tid == 0)
      P.scal->comm_error = 1u;
    return true;
  };

  // ---- phase 1 (ParticleFilter.cpp:129-153): wp, wr per particle; fp64 partial sums per segment
  for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x)
  {
    double acc[kPartUsed];
#pragma unroll
    for (int k = 0; k < kPartUsed; ++k)
      acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < kSegItems; ++k)
    {
      const uint64_t i = static_cast<uint64_t>(seg) * kSeg + static_cast<uint64_t>(tid) * kSegItems + k;
      if (i >= n)
        continue;
      const float x = P.p.x[i], y = P.p.y[i], z = P.p.z[i];
      float a0 = 0.f, a1 = 0.f;
      if (is_into_map(P.g, x, y, z))
      {
        uint32_t cnt;
        const float wp = cloud_weight_from_partials(P.part_sum, P.part_cnt, n, P.n_splits, i, P.part_kind, &cnt);
        const float wr = range_weight(P.rg, x, y, z);
        P.p.wp[i] = wp;
        P.p.wr[i] = wr;
        a0 = wp;
        a1 = wr;
        const float a = P.p.a[i];
        const double dwp = wp, dwr = wr;
        acc[0] += dwp;
        acc[1] += dwr;
        acc[2] += dwp * x;
        acc[3] += dwp * y;
        acc[4] += dwp * z;
        acc[5] += dwp * a;
        acc[6] += dwr * x;
        acc[7] += dwr * y;
        acc[8] += dwr * z;
        acc[9] += dwr * a;
        acc[10] += static_cast<double>(cnt);
        acc[11] += dwp * fabsf(x);
        acc[12] += dwp * fabsf(y);
        acc[13] += dwp * fabsf(z);
        acc[14] += dwp * fabsf(a);
        acc[15] += dwr * fabsf(x);
        acc[16] += dwr * fabsf(y);
        acc[17] += dwr * fabsf(z);
        acc[18] += dwr * fabsf(a);
      }
      else
        P.p.w[i] = 0.f;  // :140; wp / wr keep their previous values
      t0[i] = a0;
      t1[i] = a1;
    }
    block_sum_cols<kPartUsed>(acc, P.sa.part + static_cast<size_t>(seg) * kPartCols, red);
  }
  grid.sync();

  // ---- phase 1b: CTA 0 scans the segment partials (exclusive prefix per column) and publishes this rank's totals
  if (blockIdx.x == 0)
  {
    if (tid < kPartUsed)
    {
      double run = 0.0;
      for (uint32_t s = 0; s < n_seg; ++s)
      {
        P.sa.pre[static_cast<size_t>(s) * kPartCols + tid] = run;
        run += P.sa.part[static_cast<size_t>(s) * kPartCols + tid];
      }
      P.sa.tot[tid] = run;
      if (sharded)
        for (int r = 0; r < P.pv.n_ranks; ++r)
          *const_cast<volatile double*>(&P.pv.box[r]->vals[mp][P.pv.rank][tid]) = run;
      __threadfence_system();
    }
    __syncthreads();
    if (sharded && tid < P.pv.n_ranks)
      st_release_sys(&P.pv.box[tid]->flag[mp][P.pv.rank], P.pv.seq);
  }
  grid.sync();

  // ---- every CTA: global fp64 totals (all ranks, rank order) and the part that lies before this rank
  if (sharded)
  {
    const PeerBox* mine = P.pv.box[P.pv.rank];
    if (tid < P.pv.n_ranks)
      if (!peer_wait(&mine->flag[mp][tid], P.pv.seq, P.pv.timeout_clocks))
        err[0] = 1u;
    __syncthreads();
    if (tid < kPartUsed)
    {
      double gsum = 0.0, before = 0.0;
      for (int r = 0; r < P.pv.n_ranks; ++r)
      {
        const double v = *const_cast<const volatile double*>(&mine->vals[mp][r][tid]);
        if (r < P.pv.rank)
          before += v;
        gsum += v;
      }
      wk.tot[tid] = gsum;
      wk.ein[tid] = before;
    }
  }
  else if (tid < kPartUsed)
  {
    wk.tot[tid] = __ldcg(P.sa.tot + tid);
    wk.ein[tid] = 0.0;
  }
  __syncthreads();
  const double A = wk.tot[0], B = wk.tot[1];
  const double alpha = P.alpha;

  // ---- phase 2: segment summaries of the wtp / wtr chains (:151-152)
  for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x)
  {
    const uint64_t first = static_cast<uint64_t>(seg) * kSeg;
    const uint32_t count = static_cast<uint32_t>(min(static_cast<uint64_t>(kSeg), n - first));
    for (int c = 0; c < 2; ++c)
    {
      uint32_t eh, ng;
      hypothesis(wk.ein[c] + __ldcg(P.sa.pre + static_cast<size_t>(seg) * kPartCols + c), &eh, &ng);
      const SegFn f = block_seg_build<kSegThreads, kSegItems>((c ? t1 : t0) + first, count, eh, ng, xs);
      if (tid == 0)
        P.sa.fn[static_cast<size_t>(c) * P.sa.seg_cap + seg] = f;
    }
  }
  grid.sync();
  if (bail(0))
    return;

  // ---- phase 3: exact wtp, wtr
  if (blockIdx.x == 0)
  {
    if (tid == 0)
    {
      wk.terms[0] = t0;
      wk.terms[1] = t1;
    }
    if (!chain_phase(0, 2, P.sa.fn, P.sa.seg_cap, n_seg, n, P.pv, wk, xs, cs) && tid == 0)
      err[1] = 1u;
    if (tid == 0)
    {
      P.scal->wtp = wk.bcast[0];
      P.scal->wtr = wk.bcast[1];
    }
  }
  grid.sync();
  if (bail(1))
    return;
  const float wtp = __ldcg(&P.scal->wtp), wtr = __ldcg(&P.scal->wtr);

  // ---- phase 4 (:160-180): normalise wp / wr, blend; segment summaries of the wt chain
  for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x)
  {
    const uint64_t first = static_cast<uint64_t>(seg) * kSeg;
    const uint32_t count = static_cast<uint32_t>(min(static_cast<uint64_t>(kSeg), n - first));
#pragma unroll
    for (int k = 0; k < kSegItems; ++k)
    {
      const uint64_t i = first + static_cast<uint64_t>(tid) * kSegItems + k;
      if (i >= n)
        continue;
      const float wp = (wtp > 0.f) ? __fdiv_rn(P.p.wp[i], wtp) : 0.f;
      const float wr = (wtr > 0.f) ? __fdiv_rn(P.p.wr[i], wtr) : 0.f;
      P.p.wp[i] = wp;
      P.p.wr[i] = wr;
      float w = 0.f;
      if (is_into_map(P.g, P.p.x[i], P.p.y[i], P.p.z[i]))
        w = static_cast<float>(__dadd_rn(__dmul_rn(static_cast<double>(wp), alpha),
                                         __dmul_rn(static_cast<double>(wr), __dsub_rn(1.0, alpha))));  // :178
      P.p.w[i] = w;
      t0[i] = w;
    }
    __syncthreads();
    const double* pre = P.sa.pre + static_cast<size_t>(seg) * kPartCols;
    const double est = (A > 0.0 ? alpha * (wk.ein[0] + __ldcg(pre + 0)) / A : 0.0) +
                       (B > 0.0 ? (1.0 - alpha) * (wk.ein[1] + __ldcg(pre + 1)) / B : 0.0);
    uint32_t eh, ng;
    hypothesis(est, &eh, &ng);
    const SegFn f = block_seg_build<kSegThreads, kSegItems>(t0 + first, count, eh, ng, xs);
    if (tid == 0)
      P.sa.fn[seg] = f;
  }
  grid.sync();

  // ---- phase 5: exact wt (:179)
  if (blockIdx.x == 0)
  {
    if (tid == 0)
      wk.terms[0] = t0;
    if (!chain_phase(1, 1, P.sa.fn, P.sa.seg_cap, n_seg, n, P.pv, wk, xs, cs) && tid == 0)
      err[2] = 1u;
    if (tid == 0)
      P.scal->wt = wk.bcast[0];
  }
  grid.sync();
  if (bail(2))
    return;
  const float wt = __ldcg(&P.scal->wt);
  const double wt_d = (A > 0.0 ? alpha : 0.0) + (B > 0.0 ? (1.0 - alpha) : 0.0);

  // ---- phase 6 (:183-194): final normalisation; terms and segment summaries of the four mean chains
  for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x)
  {
    const uint64_t first = static_cast<uint64_t>(seg) * kSeg;
    const uint32_t count = static_cast<uint32_t>(min(static_cast<uint64_t>(kSeg), n - first));
#pragma unroll
    for (int k = 0; k < kSegItems; ++k)
    {
      const uint64_t i = first + static_cast<uint64_t>(tid) * kSegItems + k;
      if (i >= n)
        continue;
      const float w = (wt > 0.f) ? __fdiv_rn(P.p.w[i], wt) : 0.f;
      P.p.w[i] = w;
      t0[i] = __fmul_rn(w, P.p.x[i]);
      t1[i] = __fmul_rn(w, P.p.y[i]);
      t2[i] = __fmul_rn(w, P.p.z[i]);
      t3[i] = __fmul_rn(w, P.p.a[i]);
    }
    __syncthreads();
    const double* pre = P.sa.pre + static_cast<size_t>(seg) * kPartCols;
    for (int c = 0; c < 4; ++c)
    {
      double est = 0.0;
      if (wt_d > 0.0)
        est = ((A > 0.0 ? alpha * (wk.ein[2 + c] + __ldcg(pre + 2 + c)) / A : 0.0) +
               (B > 0.0 ? (1.0 - alpha) * (wk.ein[6 + c] + __ldcg(pre + 6 + c)) / B : 0.0)) /
              wt_d;
      uint32_t eh, ng;
      hypothesis(est, &eh, &ng);
      const float* tc = c == 0 ? t0 : (c == 1 ? t1 : (c == 2 ? t2 : t3));
      const SegFn f = block_seg_build<kSegThreads, kSegItems>(tc + first, count, eh, ng, xs);
      if (tid == 0)
        P.sa.fn[static_cast<size_t>(c) * P.sa.seg_cap + seg] = f;
    }
  }
  grid.sync();

  // ---- phase 7: exact mean (:190-193)
  if (blockIdx.x == 0)
  {
    if (tid == 0)
    {
      wk.terms[0] = t0;
      wk.terms[1] = t1;
      wk.terms[2] = t2;
      wk.terms[3] = t3;
    }
    // A mean component whose terms nearly cancel (|sum| far below sum |term|: a pose coordinate near zero) makes the
    // float chain hover around zero, changing binade or sign every few elements: exact only one element at a time, on
    // one GPU after the other.  Such a chain is not attempted: its component is returned as the fp64 sum -- the float
    // chain's own rounding error is proportional to the running value, so for a hovering sum it is orders of magnitude
    // below the 1e-4 m tolerance (bench.py's parity record reports the measured deviation).  The decision uses the
    // global partial sums: the same on every rank and for every rank count.
    // ... unless the set is so long that the float chain's own drift could leave the tolerance: the chain's deviation
    // from the true sum grows with the particle count (measured: 7e-7 m at 1 M particles and |mean| = 0.005 m, 2e-5 m at
    // 1 M / 0.09 m, 1e-3 m at 8 M / 0.09 m -- 0.2 .. 2 % of N * 2^-24 * |mean|).  Above 1e-3 m of that bound the
    // hovering components are evaluated exactly by the joint single-lane chain (serial_mask, ~4.5 ns per particle).
    uint32_t give_up = 0, serial = 0;
    for (int c = 0; c < 4; ++c)
    {
      const double net = (A > 0.0 ? alpha * wk.tot[2 + c] / A : 0.0) + (B > 0.0 ? (1.0 - alpha) * wk.tot[6 + c] / B : 0.0);
      const double gross = (A > 0.0 ? alpha * wk.tot[11 + c] / A : 0.0) + (B > 0.0 ? (1.0 - alpha) * wk.tot[15 + c] / B : 0.0);
      if (!(fabs(net) >= 0.125 * gross))
      {
        const double scale = wt_d > 0.0 ? 1.0 / wt_d : 0.0;
        if (static_cast<double>(P.n_total) * 5.9604645e-8 * fabs(net) * scale >= 1e-3)
          serial |= 1u << c;
        else
          give_up |= 1u << c;
      }
    }
    if (!chain_phase(2, 4, P.sa.fn, P.sa.seg_cap, n_seg, n, P.pv, wk, xs, cs, give_up, serial, &cs4) && tid == 0)
      P.scal->comm_error = 1u;
    if (tid == 0)
    {
      for (int k = 0; k < 4; ++k)
      {
        float m = wk.bcast[k];
        if ((wk.abandon >> k) & 1u)
        {
          double v = 0.0;
          if (wt_d > 0.0)
            v = ((A > 0.0 ? alpha * wk.tot[2 + k] / A : 0.0) + (B > 0.0 ? (1.0 - alpha) * wk.tot[6 + k] / B : 0.0)) / wt_d;
          m = static_cast<float>(v);
        }
        P.scal->mean[k] = m;
      }
      P.scal->mean_exact_mask = (~wk.abandon) & 0xfu;
    }
    if (tid == 0)
    {
      P.scal->evals = static_cast<unsigned long long>(__ldcg(P.sa.tot + 10));
    }
  }
}

// ------------------------------------------------------------------------------------------ resample
struct ResampleSegParams
{
  Planes src, dst;        // this rank's current / next particle planes
  const float* w;         // src.w
  float* cum;             // this rank's cumulative-weight buffer of this resample
  uint64_t n;
  SegArrays sa;
  uint32_t n_seg;
  PeerView pv;            // seq = resample step number
  ShardView sh;
  int state_index;        // which state buffer (0 / 1) holds the source planes on EVERY rank
  int cum_index;          // which cumulative-weight buffer (0 / 1) is written by this resample on every rank
  float u01;
  uint32_t* idx_out;      // nullable: global source index per local output slot
  amcl3d_pf_scalars* scal;
};

__global__ void __launch_bounds__(kSegThreads) resample_seg_kernel(const __grid_constant__ ResampleSegParams P)
{
  cg::grid_group grid = cg::this_grid();
  __shared__ ExactScanSmem<kSegThreads> xs;
  __shared__ ChainSmem<1> cs;
  __shared__ WalkSmem wk;
  __shared__ double red[1][kSegThreads / 32];
  __shared__ float ends[kMaxPeers];
  const int tid = threadIdx.x;
  const uint64_t n = P.n;
  const uint32_t n_seg = P.n_seg;
  const bool sharded = P.pv.n_ranks > 1;
  const uint32_t mp = static_cast<uint32_t>(P.pv.seq & 1ull);
  const int rank = sharded ? P.pv.rank : 0;
  const int R = sharded ? P.pv.n_ranks : 1;
  volatile unsigned int* const err = P.scal->err_stage;  // see update_seg_kernel
  if (blockIdx.x == 0 && tid == 0)
  {
    P.scal->comm_error = 0u;
    for (int k = 0; k < 4; ++k)
      err[k] = 0u;
  }
  auto bail = [&](const int k) -> bool {
    if (!err[k])
      return false;
    if (blockIdx.x == 0 && tid == 0)
      P.scal->comm_error = 1u;
    return true;
  };

  // ---- A: fp64 weight sums per segment
  for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x)
  {
    double acc[1] = { 0.0 };
#pragma unroll
    for (int k = 0; k < kSegItems; ++k)
    {
      const uint64_t i = static_cast<uint64_t>(seg) * kSeg + static_cast<uint64_t>(tid) * kSegItems + k;
      if (i < n)
        acc[0] += static_cast<double>(P.w[i]);
    }
    block_sum_cols<1>(acc, P.sa.part + static_cast<size_t>(seg) * kPartCols, red);
  }
  grid.sync();
  if (blockIdx.x == 0)
  {
    if (tid == 0)
    {
      double run = 0.0;
      for (uint32_t s = 0; s < n_seg; ++s)
      {
        P.sa.pre[static_cast<size_t>(s) * kPartCols] = run;
        run += P.sa.part[static_cast<size_t>(s) * kPartCols];
      }
      P.sa.tot[0] = run;
      if (sharded)
      {
        for (int r = 0; r < R; ++r)
          *const_cast<volatile double*>(&P.pv.box[r]->rs_total[mp][rank]) = run;
        __threadfence_system();
        for (int r = 0; r < R; ++r)
          st_release_sys(&P.pv.box[r]->rs_total_flag[mp][rank], P.pv.seq);
      }
    }
  }
  grid.sync();
  if (sharded)
  {
    const PeerBox* mine = P.pv.box[rank];
    if (tid < rank)
      if (!peer_wait(&mine->rs_total_flag[mp][tid], P.pv.seq, P.pv.timeout_clocks))
        err[0] = 1u;
    __syncthreads();
    if (tid == 0)
    {
      double before = 0.0;
      for (int r = 0; r < rank; ++r)
        before += *const_cast<const volatile double*>(&mine->rs_total[mp][r]);
      wk.ein[0] = before;
    }
  }
  else if (tid == 0)
    wk.ein[0] = 0.0;
  __syncthreads();

  // ---- B: segment summaries of the cumulative-weight chain (ParticleFilter.cpp:203,214)
  for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x)
  {
    const uint64_t first = static_cast<uint64_t>(seg) * kSeg;
    const uint32_t count = static_cast<uint32_t>(min(static_cast<uint64_t>(kSeg), n - first));
    uint32_t eh, ng;
    hypothesis(wk.ein[0] + __ldcg(P.sa.pre + static_cast<size_t>(seg) * kPartCols), &eh, &ng);
    const SegFn f = block_seg_build<kSegThreads, kSegItems>(P.w + first, count, eh, ng, xs);
    if (tid == 0)
      P.sa.fn[seg] = f;
  }
  grid.sync();
  if (bail(0))
    return;

  // ---- C: CTA 0 carries the exact chain over this rank's segments and hands it to the next rank
  if (blockIdx.x == 0)
  {
    if (tid == 0)
    {
      wk.terms[0] = P.w;
      wk.abandon = 0u;
      wk.cur[0] = 0.f;  // 0 + w_0 == w_0 exactly: starting from 0 reproduces "c = p_[0].w" (:203)
      if (sharded && rank > 0)
      {
        PeerBox* mine = P.pv.box[rank];
        if (!peer_wait(&mine->rs_carry_flag[mp], P.pv.seq, P.pv.timeout_clocks))
          err[1] = 1u;
        wk.cur[0] = *const_cast<const volatile float*>(&mine->rs_carry[mp]);
      }
    }
    __syncthreads();
    walk_segments(1, P.sa.fn, P.sa.seg_cap, n_seg, n, wk, xs, cs, P.cum, P.sa.carry, P.sa.slow);
    if (tid == 0)
    {
      wk.bcast[0] = wk.cur[0];
      if (sharded && rank < R - 1)
      {
        PeerBox* next = P.pv.box[rank + 1];
        *const_cast<volatile float*>(&next->rs_carry[mp]) = wk.cur[0];
        __threadfence_system();
        st_release_sys(&next->rs_carry_flag[mp], P.pv.seq);
      }
    }
    __syncthreads();
  }
  grid.sync();
  if (bail(1))
    return;

  // ---- D: running values inside the proven segments (the replayed ones wrote theirs in C)
  for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x)
  {
    if (__ldcg(P.sa.slow + seg))
      continue;
    const uint64_t first = static_cast<uint64_t>(seg) * kSeg;
    const uint32_t count = static_cast<uint32_t>(min(static_cast<uint64_t>(kSeg), n - first));
    block_exact_chain<kSegThreads, kSegItems>(P.w + first, count, __ldcg(P.sa.carry + seg), P.cum + first, xs, 0u);
  }
  grid.sync();
  if (blockIdx.x == 0 && tid == 0)
  {
    if (sharded)
    {
      // this rank's cumulative weights (and its source planes) may now be read by everybody
      const float end = wk.bcast[0];
      for (int r = 0; r < R; ++r)
        *const_cast<volatile float*>(&P.pv.box[r]->rs_end[mp][rank]) = end;
      __threadfence_system();
      for (int r = 0; r < R; ++r)
        st_release_sys(&P.pv.box[r]->rs_ready_flag[mp][rank], P.pv.seq);
    }
    else
      P.sa.carry[0] = wk.bcast[0];  // single GPU: the end value travels through global memory
  }
  if (sharded)
  {
    const PeerBox* mine = P.pv.box[rank];
    if (tid < R)
    {
      if (!peer_wait(&mine->rs_ready_flag[mp][tid], P.pv.seq, P.pv.timeout_clocks))
        err[2] = 1u;
      ends[tid] = *const_cast<const volatile float*>(&mine->rs_end[mp][tid]);
    }
  }
  grid.sync();  // (single GPU: orders the end value; sharded: every CTA leaves the wait before anybody returns early)
  if (bail(2))
    return;
  if (!sharded && tid == 0)
    ends[0] = __ldcg(P.sa.carry);
  __syncthreads();

  // ---- E (:207-218): every output slot finds the first source whose cumulative weight reaches u, wherever it lives
  const uint64_t n_total = P.sh.n_total;
  const float factor = __fdiv_rn(1.f, static_cast<float>(n_total));  // :201
  const float r0 = __fmul_rn(factor, P.u01);                        // :202
  int last_rank = R - 1;
  while (last_rank > 0 && P.sh.n[last_rank] == 0)
    --last_rank;
  const uint64_t m_base = P.sh.first[rank];
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + tid; k < n; k += stride)
  {
    const uint64_t m = m_base + k;
    const float u = __fadd_rn(r0, __fmul_rn(factor, static_cast<float>(static_cast<uint32_t>(m))));  // :209
    int r = 0;
    while (r < last_rank && (P.sh.n[r] == 0 || u > ends[r]))
      ++r;
    const uint64_t n_r = P.sh.n[r], cap_r = P.sh.cap[r];
    const float* blk = P.sh.block[r];
    const float* cum_r = blk + 14 * cap_r + static_cast<size_t>(P.cum_index) * cap_r;
    uint64_t lo = 0, hi = n_r;
    while (lo < hi)
    {
      const uint64_t mid = (lo + hi) >> 1;
      if (u > __ldcg(cum_r + mid))
        lo = mid + 1;
      else
        hi = mid;
    }
    const uint64_t s = lo < n_r ? lo : n_r - 1;  // the reference runs off the end here (UB); clamp
    const float* st = blk + static_cast<size_t>(P.state_index) * 7 * cap_r;
    P.dst.x[k] = __ldcg(st + s);
    P.dst.y[k] = __ldcg(st + cap_r + s);
    P.dst.z[k] = __ldcg(st + 2 * cap_r + s);
    P.dst.a[k] = __ldcg(st + 3 * cap_r + s);
    P.dst.w[k] = factor;
    P.dst.wp[k] = __ldcg(st + 5 * cap_r + s);
    P.dst.wr[k] = __ldcg(st + 6 * cap_r + s);
    if (P.idx_out)
      P.idx_out[k] = static_cast<uint32_t>(P.sh.first[r] + s);
  }
}

// ------------------------------------------------------------------------------------------ host side
static int coop_grid(const amcl3d_cuda_ctx* ctx, const void* kernel, uint32_t n_seg, int* grid_out)
{
  static int per_sm[2] = { 0, 0 };
  static const void* known[2] = { nullptr, nullptr };
  int slot = -1;
  for (int k = 0; k < 2; ++k)
    if (known[k] == kernel)
      slot = k;
  if (slot < 0)
  {
    slot = known[0] ? 1 : 0;
    known[slot] = kernel;
    int nb = 0;
    A3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kSegThreads, 0));
    per_sm[slot] = nb;
  }
  if (per_sm[slot] < 1)
    return fail(AMCL3D_CUDA_ERR_CUDA, "exact chains: the cooperative kernel does not fit on an SM");
  const uint64_t resident = static_cast<uint64_t>(per_sm[slot]) * ctx->sm_count;
  *grid_out = static_cast<int>(n_seg < resident ? (n_seg ? n_seg : 1) : resident);
  return 0;
}

static int ensure_seg(amcl3d_cuda_pf* pf, uint64_t n_seg)
{
  if (n_seg <= pf->seg_cap)
    return 0;
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (pf->d_seg)
    cudaFree(pf->d_seg);
  pf->d_seg = nullptr;
  pf->seg_cap = 0;
  const uint64_t cap = (n_seg + 63) / 64 * 64;
  A3D_CUDA_TRY(cudaMalloc(&pf->d_seg, seg_bytes(cap)));
  pf->seg_cap = cap;
  return 0;
}

int launch_update_seg(amcl3d_cuda_pf* pf, const GridView& g, const RangeParams& rg, double alpha, const void* part_sum,
                      const uint32_t* part_cnt, uint32_t n_splits, int part_kind, const PeerView& pv)
{
  amcl3d_cuda_ctx* ctx = pf->ctx;
  const uint64_t n = pf->n;
  const uint32_t n_seg = static_cast<uint32_t>((n + kSeg - 1) / kSeg);
  A3D_TRY(ensure_seg(pf, n_seg ? n_seg : 1));
  UpdateSegParams P;
  P.g = g;
  float* b = pf->d_state[pf->cur];
  const size_t c = pf->cap;
  P.p = Planes{ b, b + c, b + 2 * c, b + 3 * c, b + 4 * c, b + 5 * c, b + 6 * c };
  P.n = n;
  P.part_sum = part_sum;
  P.part_cnt = part_cnt;
  P.n_splits = n_splits;
  P.part_kind = part_kind;
  P.rg = rg;
  P.alpha = alpha;
  P.terms = pf->d_terms;
  P.terms_stride = pf->cap;
  P.scal = pf->d_scal;
  P.sa = seg_arrays(pf->d_seg, pf->seg_cap);
  P.n_seg = n_seg;
  P.n_total = ctx->n_ranks > 1 && pf->shards_valid ? pf->shards.n_total : n;
  P.pv = pv;
  int grid = 1;
  A3D_TRY(coop_grid(ctx, reinterpret_cast<const void*>(update_seg_kernel), n_seg, &grid));
  void* args[] = { &P };
  A3D_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(update_seg_kernel), dim3(grid), dim3(kSegThreads),
                                           args, 0, ctx->stream));
  ctx->launches++;
  return 0;
}

int launch_resample_seg(amcl3d_cuda_pf* pf, float u01, uint32_t* d_idx, const PeerView& pv, const ShardView& sh)
{
  amcl3d_cuda_ctx* ctx = pf->ctx;
  const uint64_t n = pf->n;
  const uint32_t n_seg = static_cast<uint32_t>((n + kSeg - 1) / kSeg);
  A3D_TRY(ensure_seg(pf, n_seg ? n_seg : 1));
  ResampleSegParams P;
  const size_t c = pf->cap;
  float* s = pf->d_state[pf->cur];
  float* d = pf->d_state[pf->cur ^ 1];
  P.src = Planes{ s, s + c, s + 2 * c, s + 3 * c, s + 4 * c, s + 5 * c, s + 6 * c };
  P.dst = Planes{ d, d + c, d + 2 * c, d + 3 * c, d + 4 * c, d + 5 * c, d + 6 * c };
  P.w = P.src.w;
  P.cum = pf->d_cum[pf->cum_cur];
  P.n = n;
  P.sa = seg_arrays(pf->d_seg, pf->seg_cap);
  P.n_seg = n_seg;
  P.pv = pv;
  P.sh = sh;
  P.state_index = pf->cur;
  P.cum_index = pf->cum_cur;
  P.u01 = u01;
  P.idx_out = d_idx;
  P.scal = pf->d_scal;
  int grid = 1;
  // the gather phase wants the whole GPU even when there are few segments
  const uint32_t want = static_cast<uint32_t>(std::max<uint64_t>(n_seg, (n + kSegThreads - 1) / kSegThreads));
  A3D_TRY(coop_grid(ctx, reinterpret_cast<const void*>(resample_seg_kernel), want, &grid));
  void* args[] = { &P };
  A3D_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(resample_seg_kernel), dim3(grid),
                                           dim3(kSegThreads), args, 0, ctx->stream));
  ctx->launches++;
  return 0;
}

}  // namespace amcl3d_b200

// Debug hook (not part of include/amcl3d_cuda.h): raw copies of the cumulative-weight buffer of the last resample and of
// the per-segment carry / replay flags.  what: 0 = cumulative weights (n floats), 1 = segment carries, 2 = replay flags.
extern "C" int amcl3d_cuda_debug_read(amcl3d_cuda_pf* pf, int what, void* out, uint64_t bytes)
{
  using namespace amcl3d_b200;
  if (!pf || !out)
    return -2;
  cudaSetDevice(pf->ctx->device);
  cudaStreamSynchronize(pf->ctx->stream);
  const SegArrays sa = seg_arrays(pf->d_seg, pf->seg_cap);
  const void* src = what == 0 ? static_cast<const void*>(pf->d_cum[pf->cum_cur ^ 1]) :
                                (what == 1 ? static_cast<const void*>(sa.carry) : static_cast<const void*>(sa.slow));
  return cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -3;
}
