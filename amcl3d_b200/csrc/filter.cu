// filter.cu -- particle-filter state and the predict / update / resample loops of ParticleFilter
// (ParticleFilter.cpp:46-256) on device-resident SoA particles.
//
// Particle planes (each `cap` floats): 0 x, 1 y, 2 z, 3 a, 4 w, 5 wp, 6 wr   (Particle, ParticleFilter.h:35-49)
//
// update() comes in two numerically different flavours, selected by option "sum_mode":
//   exact : every sum over particles is the reference's sequential float chain (bit-exact wtp/wtr/wt/mean);
//           one single-block kernel after the weighting kernel does finalize -> chain -> normalise -> chain ->
//           normalise -> chain.
//   fast  : fp64 block reductions; the three dependent sums of the reference are folded into ONE reduction of
//           10 partials (sum wp, sum wr, sum wp*pose, sum wr*pose), which is also the only thing a multi-GPU
//           update has to all-reduce.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>

#include "chain.cuh"
#include "exact_scan.cuh"
#include "filter_common.cuh"

namespace amcl3d_b200
{
constexpr float kTwoPi = 6.283185307179586f;

// ------------------------------------------------------------------------------------------ update, exact mode
// One block.  Reproduces ParticleFilter.cpp:129-195 after the per-particle cloud sums are known.  The three sums over
// non-negative weights (wtp, wtr, wt) run through the windowed exact scan (exact_scan.cuh): bit-identical to the
// reference's sequential float sums at any particle count.  EXACT_MEAN: the four signed mean sums use the single-lane
// chain (bit-exact, O(n) serial); otherwise they are reduced in fp64 (mean pose tolerance 1e-4 m).
template <bool EXACT_MEAN, int THREADS>
__global__ void __launch_bounds__(THREADS)
    update_exact_kernel(const GridView g, Planes p, const uint64_t n, const void* __restrict__ part_sum,
                        const uint32_t* __restrict__ part_cnt, const uint32_t n_splits, const int part_kind,
                        const RangeParams rg,
                        const double alpha, float* __restrict__ terms, const uint64_t terms_stride,
                        amcl3d_pf_scalars* __restrict__ scal, const int serial_chain)
{
  __shared__ ChainSmem<EXACT_MEAN ? 4 : 2> sm;
  __shared__ ExactScanSmem<THREADS> xs;
  __shared__ float bcast[4];
  __shared__ double red[4][THREADS / 32];
  __shared__ unsigned long long evals_sm;
  float* t0 = terms;
  float* t1 = terms + terms_stride;
  float* t2 = terms + 2 * terms_stride;
  float* t3 = terms + 3 * terms_stride;
  const uint32_t n32 = static_cast<uint32_t>(n);
  if (threadIdx.x == 0)
    evals_sm = 0ull;
  __syncthreads();

  // loop 1 (:129-153): wp, wr for in-map particles; terms = what gets added to wtp / wtr (0 for skipped ones)
  unsigned long long my_evals = 0;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
  {
    const float x = p.x[i], y = p.y[i], z = p.z[i];
    float a0 = 0.f, a1 = 0.f;
    if (is_into_map(g, x, y, z))
    {
      uint32_t cnt;
      const float wp = cloud_weight_from_partials(part_sum, part_cnt, n, n_splits, i, part_kind, &cnt);
      const float wr = range_weight(rg, x, y, z);
      p.wp[i] = wp;
      p.wr[i] = wr;
      a0 = wp;
      a1 = wr;
      my_evals += cnt;
    }
    else
      p.w[i] = 0.f;  // :140; wp / wr keep their previous values
    t0[i] = a0;
    t1[i] = a1;
  }
  atomicAdd(&evals_sm, my_evals);
  __syncthreads();
  // below ~2 k particles the single-lane chain (4.5 ns per element) beats the scan's fixed cost per window
  const bool use_serial = serial_chain || n <= 2048;
  float wtp, wtr;
  if (use_serial)
  {
    const float* const src[2] = { t0, t1 };
    float acc[2] = { 0.f, 0.f };
    block_chain<2>(src, n, acc, nullptr, *reinterpret_cast<ChainSmem<2>*>(&sm));
    if (threadIdx.x == 0)
    {
      bcast[0] = acc[0];
      bcast[1] = acc[1];
    }
    __syncthreads();
    wtp = bcast[0];
    wtr = bcast[1];
  }
  else
  {
    wtp = block_exact_chain<THREADS, 4096 / THREADS>(t0, n32, 0.f, nullptr, xs);
    wtr = block_exact_chain<THREADS, 4096 / THREADS>(t1, n32, 0.f, nullptr, xs);
  }

  // loop 2 (:160-180)
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
  {
    const float wp = (wtp > 0.f) ? __fdiv_rn(p.wp[i], wtp) : 0.f;
    const float wr = (wtr > 0.f) ? __fdiv_rn(p.wr[i], wtr) : 0.f;
    p.wp[i] = wp;
    p.wr[i] = wr;
    float w = 0.f;
    if (is_into_map(g, p.x[i], p.y[i], p.z[i]))
      w = static_cast<float>(__dadd_rn(__dmul_rn(static_cast<double>(wp), alpha),
                                       __dmul_rn(static_cast<double>(wr), __dsub_rn(1.0, alpha))));  // :178
    p.w[i] = w;
    t0[i] = w;
  }
  __syncthreads();
  float wt;
  if (use_serial)
  {
    const float* const src[1] = { t0 };
    float acc[1] = { 0.f };
    block_chain<1>(src, n, acc, nullptr, *reinterpret_cast<ChainSmem<1>*>(&sm));
    if (threadIdx.x == 0)
      bcast[2] = acc[0];
    __syncthreads();
    wt = bcast[2];
  }
  else
    wt = block_exact_chain<THREADS, 4096 / THREADS>(t0, n32, 0.f, nullptr, xs);

  // loop 3 (:183-194)
  double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
  {
    const float w = (wt > 0.f) ? __fdiv_rn(p.w[i], wt) : 0.f;
    p.w[i] = w;
    if (EXACT_MEAN)
    {
      t0[i] = __fmul_rn(w, p.x[i]);
      t1[i] = __fmul_rn(w, p.y[i]);
      t2[i] = __fmul_rn(w, p.z[i]);
      t3[i] = __fmul_rn(w, p.a[i]);
    }
    else
    {
      const double dw = w;
      m0 += dw * p.x[i];
      m1 += dw * p.y[i];
      m2 += dw * p.z[i];
      m3 += dw * p.a[i];
    }
  }
  __syncthreads();
  float mean[4] = { 0.f, 0.f, 0.f, 0.f };
  if (EXACT_MEAN)
  {
    const float* const src[4] = { t0, t1, t2, t3 };
    block_chain<4>(src, n, mean, nullptr, *reinterpret_cast<ChainSmem<4>*>(&sm));
  }
  else
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    m0 = warp_sum(m0);
    m1 = warp_sum(m1);
    m2 = warp_sum(m2);
    m3 = warp_sum(m3);
    if (lane == 0)
    {
      red[0][warp] = m0;
      red[1][warp] = m1;
      red[2][warp] = m2;
      red[3][warp] = m3;
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int k = 0; k < 4; ++k)
      {
        double s = 0.0;
        for (int w = 0; w < THREADS / 32; ++w)
          s += red[k][w];
        mean[k] = static_cast<float>(s);
      }
  }
  if (threadIdx.x == 0)
  {
    scal->wtp = wtp;
    scal->wtr = wtr;
    scal->wt = wt;
    scal->mean[0] = mean[0];
    scal->mean[1] = mean[1];
    scal->mean[2] = mean[2];
    scal->mean[3] = mean[3];
    scal->mean_exact_mask = EXACT_MEAN ? 0xfu : 0u;
    scal->evals = evals_sm;
  }
}

// ------------------------------------------------------------------------------------------ update, fast mode
// Stage 1: finalize per-particle weights and reduce the 10 partials (+ in-map evaluation count) into the parity
// buffer `par`.  Sharded (pv.n_ranks > 1): the last CTA to finish also PUBLISHES this rank's partials -- it stores
// them into slot [rank] of every rank's PeerBox over NVLink and raises the step flag there; no NCCL call, no extra
// launch, the exchange rides on the reduction kernel itself.
__global__ void __launch_bounds__(256)
    update_fast_stage1_kernel(const GridView g, Planes p, const uint64_t n, const void* __restrict__ part_sum,
                              const uint32_t* __restrict__ part_cnt, const uint32_t n_splits, const int part_kind,
                              const RangeParams rg,
                              amcl3d_pf_scalars* __restrict__ scal, const uint32_t par, const PeerView pv)
{
  double acc[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  unsigned long long evals = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0)
    scal->comm_error = 0u;  // a time-out of an earlier update has been reported by now (stage 2 may raise it again)
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float x = p.x[i], y = p.y[i], z = p.z[i];
    if (is_into_map(g, x, y, z))
    {
      uint32_t cnt;
      const float wp = cloud_weight_from_partials(part_sum, part_cnt, n, n_splits, i, part_kind, &cnt);
      const float wr = range_weight(rg, x, y, z);
      p.wp[i] = wp;
      p.wr[i] = wr;
      const float a = p.a[i];
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
      evals += cnt;
    }
    else
      p.w[i] = 0.f;
  }
  __shared__ double red[10][8];
  __shared__ unsigned long long red_e[8];
  __shared__ bool last_cta;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 10; ++k)
  {
    const double v = warp_sum(acc[k]);
    if (lane == 0)
      red[k][warp] = v;
  }
  {
    unsigned long long e = evals;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0)
      red_e[warp] = e;
  }
  __syncthreads();
  if (threadIdx.x < 10)
  {
    double v = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w)
      v += red[threadIdx.x][w];
    atomicAdd(&scal->dsum[par][threadIdx.x], v);
  }
  if (threadIdx.x == 32)
  {
    unsigned long long e = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w)
      e += red_e[w];
    atomicAdd(&scal->evals_acc[par], e);
  }
  if (pv.n_ranks > 1)
  {
    __threadfence();  // this CTA's additions are visible device-wide before its ticket is
    __syncthreads();
    if (threadIdx.x == 0)
      last_cta = atomicAdd(&scal->ticket, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (last_cta)
    {
      const uint32_t mp = static_cast<uint32_t>(pv.seq & 1ull);
      if (threadIdx.x < 10)
      {
        const double v = atomicAdd(&scal->dsum[par][threadIdx.x], 0.0);  // device-coherent read of the finished total
        for (int r = 0; r < pv.n_ranks; ++r)
          *const_cast<volatile double*>(&pv.box[r]->vals[mp][pv.rank][threadIdx.x]) = v;
        __threadfence_system();
      }
      __syncthreads();
      if (threadIdx.x < pv.n_ranks)
      {
        unsigned long long* f = &pv.box[threadIdx.x]->flag[mp][pv.rank];
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(pv.seq) : "memory");
      }
      if (threadIdx.x == 0)
        scal->ticket = 0;
    }
  }
}

// Stage 2: normalise, blend, final normalise; thread 0 of CTA 0 writes the mean.  The ten totals come from the parity
// buffer (one GPU, or after ncclAllReduce) or -- sharded with peer memory -- from this rank's PeerBox: every CTA waits
// for the step flags of all ranks and adds the slots in rank order, so all ranks hold identical bits.
__global__ void __launch_bounds__(256)
    update_fast_stage2_kernel(const GridView g, Planes p, const uint64_t n, const double alpha,
                              amcl3d_pf_scalars* __restrict__ scal, const uint32_t par, const PeerView pv)
{
  __shared__ double tot[10];
  __shared__ int timed_out;
  if (threadIdx.x == 0)
    timed_out = 0;
  __syncthreads();
  if (pv.n_ranks > 1)
  {
    const uint32_t mp = static_cast<uint32_t>(pv.seq & 1ull);
    const PeerBox* mine = pv.box[pv.rank];
    if (threadIdx.x < pv.n_ranks)
    {
      const unsigned long long* f = &mine->flag[mp][threadIdx.x];
      const long long t0 = clock64();
      unsigned long long seen = 0;
      for (;;)
      {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
        if (seen == pv.seq)
          break;
        if (clock64() - t0 > pv.timeout_clocks)  // a peer is missing: fail loudly instead of hanging the GPU
        {
          atomicExch(&scal->comm_error, 1u);
          timed_out = 1;
          break;
        }
        __nanosleep(64);
      }
    }
    __syncthreads();
    if (timed_out)
      return;  // incomplete totals: leave the particles as stage 1 left them (the host reports the error)
    if (threadIdx.x < 10)
    {
      double v = 0.0;
      for (int r = 0; r < pv.n_ranks; ++r)
        v += *const_cast<const volatile double*>(&mine->vals[mp][r][threadIdx.x]);
      tot[threadIdx.x] = v;
    }
  }
  else if (threadIdx.x < 10)
    tot[threadIdx.x] = scal->dsum[par][threadIdx.x];
  __syncthreads();
  const double A = tot[0], B = tot[1];
  const float wtp = static_cast<float>(A), wtr = static_cast<float>(B);
  // sum over in-map particles of wp/wtp is 1 whenever wtp > 0 (same set, same divisor): wt is known in closed form
  const double wt_d = (A > 0.0 ? alpha : 0.0) + (B > 0.0 ? (1.0 - alpha) : 0.0);
  const float wt = static_cast<float>(wt_d);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float wp = (wtp > 0.f) ? __fdiv_rn(p.wp[i], wtp) : 0.f;
    const float wr = (wtr > 0.f) ? __fdiv_rn(p.wr[i], wtr) : 0.f;
    p.wp[i] = wp;
    p.wr[i] = wr;
    float w = 0.f;
    if (is_into_map(g, p.x[i], p.y[i], p.z[i]))
      w = static_cast<float>(static_cast<double>(wp) * alpha + static_cast<double>(wr) * (1.0 - alpha));
    p.w[i] = (wt > 0.f) ? __fdiv_rn(w, wt) : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    scal->wtp = wtp;
    scal->wtr = wtr;
    scal->wt = wt;
    for (int k = 0; k < 4; ++k)
    {
      double m = 0.0;
      if (wt_d > 0.0)
      {
        if (A > 0.0)
          m += alpha * tot[2 + k] / A;
        if (B > 0.0)
          m += (1.0 - alpha) * tot[6 + k] / B;
        m /= wt_d;
      }
      scal->mean[k] = static_cast<float>(m);
    }
    scal->mean_exact_mask = 0u;  // fp64 sums
    scal->evals = scal->evals_acc[par];
    // clear the other parity buffer for the next update (nobody touches it until that update's stage 1)
    for (int k = 0; k < 12; ++k)
      scal->dsum[par ^ 1u][k] = 0.0;
    scal->evals_acc[par ^ 1u] = 0ull;
  }
}

// ------------------------------------------------------------------------------------------ predict
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
#pragma unroll
  for (int r = 0; r < 10; ++r)
  {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// four independent N(0,1) draws for (global particle index, step) under key `seed`
__device__ __forceinline__ void philox_normal4(uint64_t index, uint64_t step, uint64_t seed, float out[4])
{
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), static_cast<uint32_t>(step),
                static_cast<uint32_t>(step >> 32), static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  // 24-bit uniforms strictly inside (0,1)
  const float u0 = (static_cast<float>(r[0] >> 8) + 0.5f) * 5.9604644775390625e-8f;
  const float u1 = (static_cast<float>(r[1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
  const float u2 = (static_cast<float>(r[2] >> 8) + 0.5f) * 5.9604644775390625e-8f;
  const float u3 = (static_cast<float>(r[3] >> 8) + 0.5f) * 5.9604644775390625e-8f;
  const float m0 = sqrtf(-2.f * logf(u0)), m1 = sqrtf(-2.f * logf(u2));
  float s0, c0, s1, c1;
  sincosf(kTwoPi * u1, &s0, &c0);
  sincosf(kTwoPi * u3, &s1, &c1);
  out[0] = m0 * c0;
  out[1] = m0 * s0;
  out[2] = m1 * c1;
  out[3] = m1 * s1;
}

struct PredictParams
{
  double delta[4];
  float dev[4];  // float(|delta*mod|): std::normal_distribution<float>'s stddev (ParticleFilter.cpp:101-104,248)
};

// ParticleFilter.cpp:108-118
__global__ void __launch_bounds__(256) predict_kernel(Planes p, const uint64_t n, const PredictParams pp,
                                                      const float* __restrict__ noise, const uint64_t seed,
                                                      const uint64_t step, const uint64_t index_base)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    float gx, gy, gz, ga;
    if (noise)
    {
      const float4 v = *reinterpret_cast<const float4*>(noise + 4 * i);
      gx = v.x;
      gy = v.y;
      gz = v.z;
      ga = v.w;
    }
    else
    {
      float z4[4];
      philox_normal4(index_base + i, step, seed, z4);
      gx = z4[0] * pp.dev[0];
      gy = z4[1] * pp.dev[1];
      gz = z4[2] * pp.dev[2];
      ga = z4[3] * pp.dev[3];
    }
    const float a = p.a[i];
    double sd, cd;
    sincos(static_cast<double>(a), &sd, &cd);
    const float sa = static_cast<float>(sd), ca = static_cast<float>(cd);                          // :110-111
    const float rand_x = static_cast<float>(__dadd_rn(pp.delta[0], static_cast<double>(gx)));      // :112
    const float rand_y = static_cast<float>(__dadd_rn(pp.delta[1], static_cast<double>(gy)));      // :113
    p.x[i] = __fadd_rn(p.x[i], __fsub_rn(__fmul_rn(ca, rand_x), __fmul_rn(sa, rand_y)));           // :114
    p.y[i] = __fadd_rn(p.y[i], __fadd_rn(__fmul_rn(sa, rand_x), __fmul_rn(ca, rand_y)));           // :115
    p.z[i] = static_cast<float>(__dadd_rn(static_cast<double>(p.z[i]), __dadd_rn(pp.delta[2], static_cast<double>(gz))));  // :116
    p.a[i] = static_cast<float>(__dadd_rn(static_cast<double>(a), __dadd_rn(pp.delta[3], static_cast<double>(ga))));      // :117
  }
}

// ------------------------------------------------------------------------------------------ init
struct InitParams
{
  float pose[4];
  float dev[4];
  float g1, g2;  // ParticleFilter.cpp:55-56, evaluated on the host
};

// ParticleFilter.cpp:58-80 (particle values and unnormalised weights); terms[0] = w for the chain
__global__ void __launch_bounds__(256) init_kernel(Planes p, const uint64_t n, const InitParams ip,
                                                   const float* __restrict__ noise, const uint64_t seed,
                                                   const uint64_t index_base, float* __restrict__ t0)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    float x = ip.pose[0], y = ip.pose[1], z = ip.pose[2], a = ip.pose[3], w = ip.g1;
    if (index_base + i != 0)
    {
      float g[4];
      if (noise)
      {
        const float4 v = *reinterpret_cast<const float4*>(noise + 4 * i);
        g[0] = v.x;
        g[1] = v.y;
        g[2] = v.z;
        g[3] = v.w;
      }
      else
      {
        philox_normal4(index_base + i, 0, seed, g);
        for (int k = 0; k < 4; ++k)
          g[k] *= ip.dev[k];
      }
      x = __fadd_rn(ip.pose[0], g[0]);
      y = __fadd_rn(ip.pose[1], g[1]);
      z = __fadd_rn(ip.pose[2], g[2]);
      a = __fadd_rn(ip.pose[3], g[3]);
      const float dx = __fsub_rn(x, ip.pose[0]), dy = __fsub_rn(y, ip.pose[1]), dz = __fsub_rn(z, ip.pose[2]);
      const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const float dist = static_cast<float>(sqrt(static_cast<double>(s)));                   // :74-75
      const float arg = __fmul_rn(__fmul_rn(-dist, dist), ip.g2);                            // :77 float argument
      w = static_cast<float>(__dmul_rn(static_cast<double>(ip.g1), exp(static_cast<double>(arg))));
    }
    p.x[i] = x;
    p.y[i] = y;
    p.z[i] = z;
    p.a[i] = a;
    p.w[i] = w;
    p.wp[i] = 0.f;
    p.wr[i] = 0.f;
    t0[i] = w;
  }
}

// One block: wt chain, normalise, mean chains (ParticleFilter.cpp:64,79,82-92).
__global__ void __launch_bounds__(512) init_finish_kernel(Planes p, const uint64_t n, float* __restrict__ terms,
                                                           const uint64_t terms_stride, amcl3d_pf_scalars* __restrict__ scal)
{
  __shared__ ChainSmem<4> sm;
  __shared__ float bcast;
  float* t0 = terms;
  float* t1 = terms + terms_stride;
  float* t2 = terms + 2 * terms_stride;
  float* t3 = terms + 3 * terms_stride;
  {
    const float* const src[1] = { t0 };
    float acc[1] = { 0.f };
    block_chain<1>(src, n, acc, nullptr, *reinterpret_cast<ChainSmem<1>*>(&sm));
    if (threadIdx.x == 0)
      bcast = acc[0];
  }
  __syncthreads();
  const float wt = bcast;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
  {
    const float w = __fdiv_rn(p.w[i], wt);
    p.w[i] = w;
    t0[i] = __fmul_rn(w, p.x[i]);
    t1[i] = __fmul_rn(w, p.y[i]);
    t2[i] = __fmul_rn(w, p.z[i]);
    t3[i] = __fmul_rn(w, p.a[i]);
  }
  __syncthreads();
  const float* const src[4] = { t0, t1, t2, t3 };
  float acc[4] = { 0.f, 0.f, 0.f, 0.f };
  block_chain<4>(src, n, acc, nullptr, sm);
  if (threadIdx.x == 0)
  {
    scal->wt = wt;
    for (int k = 0; k < 4; ++k)
      scal->mean[k] = acc[k];
  }
}

// ------------------------------------------------------------------------------------------ resample
// Exact mode, step 1 (one block): the float cumulative chain c_i of ParticleFilter.cpp:203,214.
// `serial` != 0 selects the plain single-lane chain (kept as the cross-check of the windowed scan in exact_scan.cuh).
__global__ void __launch_bounds__(1024) resample_chain_kernel(const float* __restrict__ w, const uint64_t n,
                                                              float* __restrict__ chain, const int serial)
{
  if (serial)
  {
    __shared__ ChainSmem<1> sm;
    const float* const src[1] = { w };
    float acc[1] = { 0.f };  // 0 + w_0 == w_0 exactly, so starting from 0 reproduces "c = p_[0].w"
    block_chain<1>(src, n, acc, chain, sm);
  }
  else
  {
    __shared__ ExactScanSmem<1024> sm;
    block_exact_chain<1024, 4>(w, static_cast<uint32_t>(n), 0.f, chain, sm);
  }
}

// Step 2: for every output slot m, the first source index whose cumulative weight reaches
// u = r + factor*m (ParticleFilter.cpp:209-216), then the copy of ParticleFilter.cpp:216-217.
__global__ void __launch_bounds__(256)
    resample_gather_kernel(const float* __restrict__ chain, const uint64_t n_src, const Planes src, Planes dst,
                           const uint64_t m_base, const uint64_t m_count, const uint64_t n_total, const float u01,
                           uint32_t* __restrict__ idx_out)
{
  const float factor = __fdiv_rn(1.f, static_cast<float>(n_total));  // :201
  const float r = __fmul_rn(factor, u01);                            // :202
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < m_count; k += stride)
  {
    const uint64_t m = m_base + k;
    const float u = __fadd_rn(r, __fmul_rn(factor, static_cast<float>(static_cast<uint32_t>(m))));  // :209
    const float uu = u;
    // smallest i with !(u > c_i); the chain is non-decreasing (weights >= 0)
    uint64_t lo = 0, hi = n_src;
    while (lo < hi)
    {
      const uint64_t mid = (lo + hi) >> 1;
      if (uu > chain[mid])
        lo = mid + 1;
      else
        hi = mid;
    }
    const uint64_t s = lo < n_src ? lo : n_src - 1;  // the reference runs off the end here (UB); clamp
    dst.x[k] = src.x[s];
    dst.y[k] = src.y[s];
    dst.z[k] = src.z[s];
    dst.a[k] = src.a[s];
    dst.w[k] = factor;
    dst.wp[k] = src.wp[s];
    dst.wr[k] = src.wr[s];
    if (idx_out)
      idx_out[k] = static_cast<uint32_t>(s);
  }
}

// AoS <-> SoA particle conversion
__global__ void aos_to_soa_kernel(const float* __restrict__ aos, Planes p, const uint64_t n)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float* q = aos + 7 * i;
    p.x[i] = q[0];
    p.y[i] = q[1];
    p.z[i] = q[2];
    p.a[i] = q[3];
    p.w[i] = q[4];
    p.wp[i] = q[5];
    p.wr[i] = q[6];
  }
}
__global__ void soa_to_aos_kernel(float* __restrict__ aos, const Planes p, const uint64_t n)
{
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    float* q = aos + 7 * i;
    q[0] = p.x[i];
    q[1] = p.y[i];
    q[2] = p.z[i];
    q[3] = p.a[i];
    q[4] = p.w[i];
    q[5] = p.wp[i];
    q[6] = p.wr[i];
  }
}


static Planes planes_of(const amcl3d_cuda_pf* pf, int which)
{
  float* b = pf->d_state[which];
  const size_t c = pf->cap;
  Planes p = { b, b + c, b + 2 * c, b + 3 * c, b + 4 * c, b + 5 * c, b + 6 * c };
  return p;
}

static int grid_for(const amcl3d_cuda_ctx* ctx, uint64_t n, int block)
{
  uint64_t blocks = (n + block - 1) / block;
  const uint64_t cap = static_cast<uint64_t>(ctx->sm_count) * 16;
  if (blocks > cap)
    blocks = cap;
  return static_cast<int>(blocks ? blocks : 1);
}

// (re)allocates everything that scales with the particle count.  The two state buffers and the two cumulative-weight
// buffers share ONE allocation (amcl3d_cuda_pf::d_block): state[0] | state[1] | cum[0] | cum[1].
static int reserve_particles(amcl3d_cuda_pf* pf, uint64_t n)
{
  if (n <= pf->cap && pf->d_block)
    return 0;
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  comm_release_shards(ctx, &pf->shards);
  pf->shards_valid = false;
  const uint64_t cap = ((n ? n : 1) + 255) / 256 * 256;
  if (pf->d_block)
    cudaFree(pf->d_block);
  pf->d_block = nullptr;
  pf->d_state[0] = pf->d_state[1] = pf->d_cum[0] = pf->d_cum[1] = nullptr;
  pf->cap = 0;
  A3D_CUDA_TRY(cudaMalloc(&pf->d_block, cap * 16 * sizeof(float)));
  A3D_CUDA_TRY(cudaMemsetAsync(pf->d_block, 0, cap * 16 * sizeof(float), ctx->stream));
  pf->d_state[0] = pf->d_block;
  pf->d_state[1] = pf->d_block + 7 * cap;
  pf->d_cum[0] = pf->d_block + 14 * cap;
  pf->d_cum[1] = pf->d_block + 15 * cap;
  if (pf->d_terms)
    cudaFree(pf->d_terms);
  pf->d_terms = nullptr;
  A3D_CUDA_TRY(cudaMalloc(&pf->d_terms, cap * 4 * sizeof(float)));
  if (pf->d_idx)
    cudaFree(pf->d_idx);
  pf->d_idx = nullptr;
  A3D_CUDA_TRY(cudaMalloc(&pf->d_idx, cap * sizeof(uint32_t)));
  pf->cap = cap;
  return 0;
}

static int reserve_bytes(void** p, uint64_t* cap, uint64_t want)
{
  if (want <= *cap)
    return 0;
  if (*p)
    cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const uint64_t bytes = (want + 4095) / 4096 * 4096;
  A3D_CUDA_TRY(cudaMalloc(p, bytes));
  *cap = bytes;
  return 0;
}

// A new particle set starts in state buffer 0 / cumulative buffer 0 on EVERY rank (peers index each other's blocks by
// these two numbers, which then flip in lock step with every resample).  With a communicator attached this is a
// collective: counts and particle blocks are exchanged (comm.cu: comm_exchange_shards), so shards may differ in size.
static int begin_particle_set(amcl3d_cuda_pf* pf, uint64_t n)
{
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_TRY(reserve_particles(pf, n));
  pf->n = n;
  pf->cur = 0;
  pf->cum_cur = 0;
  pf->order_valid = pf->gorder_valid = false;
  A3D_TRY(comm_exchange_shards(ctx, pf->d_block, pf->cap, n, &pf->shards));
  pf->shards_valid = true;
  return 0;
}

// launch helpers of filter_exact.cu
int launch_update_seg(amcl3d_cuda_pf* pf, const GridView& g, const RangeParams& rg, double alpha, const void* part_sum,
                      const uint32_t* part_cnt, uint32_t n_splits, int part_kind, const PeerView& pv);
int launch_resample_seg(amcl3d_cuda_pf* pf, float u01, uint32_t* d_idx, const PeerView& pv, const ShardView& sh);

}  // namespace amcl3d_b200

using namespace amcl3d_b200;

extern "C" {

int amcl3d_cuda_pf_create(amcl3d_cuda_ctx* ctx, amcl3d_cuda_pf** out)
{
  if (!ctx || !out)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_create: NULL argument");
  *out = nullptr;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  amcl3d_cuda_pf* pf = new amcl3d_cuda_pf();
  pf->ctx = ctx;
  cudaError_t e = cudaMalloc(&pf->d_scal, sizeof(amcl3d_pf_scalars));
  if (e == cudaSuccess)
    e = cudaMemsetAsync(pf->d_scal, 0, sizeof(amcl3d_pf_scalars), ctx->stream);
  if (e != cudaSuccess)
  {
    delete pf;
    return fail(AMCL3D_CUDA_ERR_CUDA, std::string("pf_create: ") + cudaGetErrorString(e));
  }
  *out = pf;
  return 0;
}

int amcl3d_cuda_pf_destroy(amcl3d_cuda_pf* pf)
{
  if (!pf)
    return 0;
  cudaSetDevice(pf->ctx->device);
  cudaStreamSynchronize(pf->ctx->stream);
  comm_release_shards(pf->ctx, &pf->shards);
  void* bufs[] = { pf->d_block,  pf->d_cloud, pf->d_part_sum,  pf->d_part_cnt,   pf->d_terms, pf->d_chain,     pf->d_idx,
                   pf->d_ranges, pf->d_scal,  pf->d_noise,     pf->d_cloud_tmp,  pf->d_cloud_work, pf->d_order,
                   pf->d_order_work, pf->d_seg, pf->d_vals, pf->d_pos_of, pf->d_rep_sum, pf->d_rep_cnt,
                   pf->d_gpose, pf->d_gorder, pf->d_gorder_work, pf->d_gex, pf->d_gstage, pf->d_gorder_tmp };
  for (void* b : bufs)
    if (b)
      cudaFree(b);
  delete pf;
  return 0;
}

int amcl3d_cuda_pf_size(const amcl3d_cuda_pf* pf, uint64_t* n)
{
  if (!pf || !n)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_size: NULL argument");
  *n = pf->n;
  return 0;
}

int amcl3d_cuda_pf_upload_particles(amcl3d_cuda_pf* pf, const float* particles7, uint64_t n)
{
  if (!pf || (n && !particles7))
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_upload_particles: NULL argument");
  if (n >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_upload_particles: too many particles");
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  A3D_TRY(begin_particle_set(pf, n));
  if (n == 0)
    return 0;
  // stage the AoS block in the context's scratch arena (never in the alternate state buffer: a peer may still be
  // gathering from it), transpose into the current planes
  A3D_TRY(ensure_scratch(ctx, n * 7 * sizeof(float)));
  float* stage = static_cast<float*>(ctx->scratch);
  A3D_CUDA_TRY(cudaMemcpyAsync(stage, particles7, n * 7 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  aos_to_soa_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(stage, planes_of(pf, pf->cur), n);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the host buffer may be reused by the caller
  return 0;
}

int amcl3d_cuda_pf_download_particles(amcl3d_cuda_pf* pf, float* particles7)
{
  if (!pf || (pf->n && !particles7))
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_download_particles: NULL argument");
  if (pf->n == 0)
    return 0;
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  A3D_TRY(ensure_scratch(ctx, pf->n * 7 * sizeof(float)));
  float* stage = static_cast<float*>(ctx->scratch);
  soa_to_aos_kernel<<<grid_for(ctx, pf->n, 256), 256, 0, ctx->stream>>>(stage, planes_of(pf, pf->cur), pf->n);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  A3D_CUDA_TRY(cudaMemcpyAsync(particles7, stage, pf->n * 7 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static_assert(offsetof(amcl3d_pf_scalars, dsum) == kPfScalarsHeadBytes, "host read-back covers the head of the block");

static int read_mean(amcl3d_cuda_pf* pf, float* mean4_out)
{
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_TRY(ensure_pinned(ctx, sizeof(amcl3d_pf_scalars)));
  A3D_CUDA_TRY(cudaMemcpyAsync(ctx->pinned, pf->d_scal, kPfScalarsHeadBytes, cudaMemcpyDeviceToHost, ctx->stream));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  const amcl3d_pf_scalars* s = static_cast<const amcl3d_pf_scalars*>(ctx->pinned);
  std::memcpy(pf->mean, s->mean, sizeof(pf->mean));
  pf->mean_exact_mask = s->mean_exact_mask;
  pf->last_evals = s->evals;
  if (s->comm_error)
    return fail(AMCL3D_CUDA_ERR_NCCL,
                "the peer-memory exchange of the last update / resample timed out (a rank is missing or more than "
                "peer_timeout_ms behind); the particle weights of that step are not normalised");
  if (mean4_out)
    std::memcpy(mean4_out, s->mean, 4 * sizeof(float));
  return 0;
}

int amcl3d_cuda_pf_get_mean(amcl3d_cuda_pf* pf, float mean4_out[4])
{
  if (!pf || !mean4_out)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_get_mean: NULL argument");
  A3D_CUDA_TRY(cudaSetDevice(pf->ctx->device));
  return read_mean(pf, mean4_out);
}

int amcl3d_cuda_pf_mean_exact_mask(amcl3d_cuda_pf* pf, uint32_t* mask)
{
  if (!pf || !mask)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_mean_exact_mask: NULL argument");
  A3D_CUDA_TRY(cudaSetDevice(pf->ctx->device));
  A3D_TRY(read_mean(pf, nullptr));
  *mask = pf->mean_exact_mask;
  return 0;
}

int amcl3d_cuda_pf_last_cloud_weights(amcl3d_cuda_pf* pf, float* weight_out, uint32_t* n_out)
{
  if (!pf || !weight_out)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_last_cloud_weights: NULL argument");
  if (pf->n == 0)
    return 0;
  if (pf->last_n != pf->n || pf->last_splits == 0)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_last_cloud_weights: no update has run on this particle set");
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  const uint32_t n = static_cast<uint32_t>(pf->n);
  A3D_TRY(ensure_scratch(ctx, static_cast<size_t>(n) * 8 + 512));
  float* d_w = static_cast<float*>(ctx->scratch);
  uint32_t* d_c = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->scratch) + (static_cast<size_t>(n) * 4 + 255) / 256 * 256);
  A3D_TRY(launch_batch_finish(ctx, pf->last_w_sum, pf->last_w_cnt, n, pf->last_splits, pf->last_kind, d_w, d_c));
  A3D_CUDA_TRY(cudaMemcpyAsync(weight_out, d_w, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_out)
    A3D_CUDA_TRY(cudaMemcpyAsync(n_out, d_c, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  // the reference never evaluates particles outside the map (ParticleFilter.cpp:137-142): the kernel leaves their
  // partials at zero, so weight 0 / count 0 is what comes back for them
  return 0;
}

int amcl3d_cuda_pf_last_in_map_evals(amcl3d_cuda_pf* pf, uint64_t* evals)
{
  if (!pf || !evals)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_last_in_map_evals: NULL argument");
  A3D_CUDA_TRY(cudaSetDevice(pf->ctx->device));
  A3D_TRY(read_mean(pf, nullptr));
  *evals = pf->last_evals;
  return 0;
}

static int stage_noise(amcl3d_cuda_pf* pf, const float* noise_n4, uint64_t n)
{
  amcl3d_cuda_ctx* ctx = pf->ctx;
  uint64_t cap_bytes = pf->noise_cap * 16;
  A3D_TRY(reserve_bytes(reinterpret_cast<void**>(&pf->d_noise), &cap_bytes, n * 16));
  pf->noise_cap = cap_bytes / 16;
  A3D_CUDA_TRY(cudaMemcpyAsync(pf->d_noise, noise_n4, n * 16, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int amcl3d_cuda_pf_init(amcl3d_cuda_pf* pf, uint64_t n, const float pose4[4], const float devs4[4],
                        const float* noise_n4, uint64_t seed, float* mean4_out)
{
  if (!pf || !pose4 || !devs4)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_init: NULL argument");
  if (n == 0 || n >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_init: particle count must be in [1, 2^32)");
  amcl3d_cuda_ctx* ctx = pf->ctx;
  if (ctx->n_ranks > 1)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_init: initialise on one rank and upload shards (multi-GPU init is host-driven)");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  A3D_TRY(begin_particle_set(pf, n));
  InitParams ip;
  std::memcpy(ip.pose, pose4, 16);
  std::memcpy(ip.dev, devs4, 16);
  // ParticleFilter.cpp:54-56
  const float dev = std::fmax(std::fmax(devs4[0], devs4[1]), devs4[2]);
  ip.g1 = static_cast<float>(1. / (static_cast<double>(dev) * std::sqrt(2 * M_PI)));
  const float two_dev_dev = 2 * dev * dev;
  ip.g2 = static_cast<float>(1. / static_cast<double>(two_dev_dev));
  if (noise_n4)
    A3D_TRY(stage_noise(pf, noise_n4, n));
  const Planes p = planes_of(pf, pf->cur);
  init_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(p, n, ip, noise_n4 ? pf->d_noise : nullptr, seed, 0,
                                                              pf->d_terms);
  init_finish_kernel<<<1, 512, 0, ctx->stream>>>(p, n, pf->d_terms, pf->cap, pf->d_scal);
  ctx->launches += 2;
  A3D_CUDA_TRY(cudaGetLastError());
  return read_mean(pf, mean4_out);
}

int amcl3d_cuda_pf_predict(amcl3d_cuda_pf* pf, const double mods4[4], const double deltas4[4], const float* noise_n4,
                           uint64_t seed, uint64_t step)
{
  if (!pf || !mods4 || !deltas4)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_predict: NULL argument");
  if (pf->n == 0)
    return 0;
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  PredictParams pp;
  for (int k = 0; k < 4; ++k)
  {
    pp.delta[k] = deltas4[k];
    pp.dev[k] = static_cast<float>(std::fabs(deltas4[k] * mods4[k]));  // ParticleFilter.cpp:101-104 then :248
  }
  if (noise_n4)
    A3D_TRY(stage_noise(pf, noise_n4, pf->n));
  // Philox counter = GLOBAL particle index: the first index of this rank's shard comes from the exchanged shard table
  const uint64_t base = pf->shards_valid ? pf->shards.first[pf->shards.n_ranks > 1 ? pf->shards.rank : 0] : 0;
  predict_kernel<<<grid_for(ctx, pf->n, 256), 256, 0, ctx->stream>>>(planes_of(pf, pf->cur), pf->n, pp,
                                                                    noise_n4 ? pf->d_noise : nullptr, seed, step, base);
  ctx->launches++;
  pf->order_valid = pf->gorder_valid = false;
  A3D_CUDA_TRY(cudaGetLastError());
  if (noise_n4)
    A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // caller may reuse the host noise buffer
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ pose-balanced weighting
// A particle set sharded over several GPUs by INDEX gives every GPU a thinned-out copy of the whole pose distribution:
// the map region one cloud point can reach is as large as for the whole set, but only 1/R of the particles share it, so
// every fetched sector serves 1/R of the gathers (measured: 131 072 i.i.d. particles 18.3 ms, a pose-coherent slice of the
// same size 14.5 ms).  With the "global schedule" the weighting WORK is dealt out by pose instead: the poses of all
// shards are gathered once per pose change (ncclAllGather), rank 0 computes the scheduling permutation of the whole set
// (order.cu) and broadcasts it, and rank r weighs the r-th contiguous slice of that permutation -- particles that are
// neighbours in pose, owned by whichever rank.  Cloud sums and counts go back to the owners through one in-place
// ncclAllReduce of uint32 words over arrays in which every entry is written by exactly one rank (bits + 0 = bits: exact).
// Ownership, indices and every sum over particles are untouched: the result is bit-identical to the per-shard schedule.
namespace amcl3d_b200
{
__global__ void pack_pose_planes_kernel(const Planes p, const uint64_t n, const uint64_t max_n, float* __restrict__ out)
{
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
  {
    out[i] = p.x[i];
    out[max_n + i] = p.y[i];
    out[2 * max_n + i] = p.z[i];
    out[3 * max_n + i] = p.a[i];
  }
}

// recv: [rank][4][max_n]  ->  all: [4][n_total] in global particle order
__global__ void unpack_pose_planes_kernel(const float* __restrict__ recv, const ShardView sh, const uint64_t max_n,
                                          float* __restrict__ all)
{
  for (uint64_t g = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; g < sh.n_total;
       g += static_cast<uint64_t>(gridDim.x) * blockDim.x)
  {
    int r = 0;
    while (r + 1 < sh.n_ranks && g >= sh.first[r] + sh.n[r])
      ++r;
    const uint64_t i = g - sh.first[r];
    const float* src = recv + static_cast<uint64_t>(r) * 4 * max_n + i;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      all[static_cast<uint64_t>(k) * sh.n_total + g] = src[static_cast<uint64_t>(k) * max_n];
  }
}

// sorted[i] -> dealt[...]: chunk c = i / chunk of the sorted permutation goes to region c % n_ranks, position c / n_ranks
struct DealPlan
{
  uint32_t chunk;
  int n_ranks;
  uint32_t start[kMaxPeers];
};
__global__ void deal_chunks_kernel(const uint32_t* __restrict__ sorted, uint32_t* __restrict__ dealt, const uint64_t n,
                                   const DealPlan plan)
{
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
  {
    const uint64_t c = i / plan.chunk, o = i % plan.chunk;
    const uint64_t r = c % plan.n_ranks, k = c / plan.n_ranks;
    dealt[plan.start[r] + k * plan.chunk + o] = sorted[i];
  }
}

// lane l of this rank's slice: combined cloud sum (bit pattern) and count of particle order[l] into the exchange arrays
__global__ void pack_slice_results_kernel(const void* __restrict__ part_sum, const uint32_t* __restrict__ part_cnt,
                                          const uint64_t n_total, const uint32_t n_splits, const int kind,
                                          const uint32_t* __restrict__ order, const uint32_t n_lanes,
                                          uint32_t* __restrict__ ex_sum, uint32_t* __restrict__ ex_cnt)
{
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lanes)
    return;
  const uint32_t i = order[l];
  uint32_t c;
  const float sum = combine_partials(part_sum, part_cnt, n_total, n_splits, i, kind, &c);
  ex_sum[i] = __float_as_uint(sum);
  ex_cnt[i] = c;
}
}  // namespace amcl3d_b200

// (Re)builds the gathered poses and the scheduling permutation of the whole sharded set.  Collective.
static int refresh_global_schedule(amcl3d_cuda_pf* pf)
{
  amcl3d_cuda_ctx* ctx = pf->ctx;
  const ShardView& sh = pf->shards;
  const uint64_t nt = sh.n_total;
  uint64_t max_n = 1;
  for (int r = 0; r < sh.n_ranks; ++r)
    max_n = std::max<uint64_t>(max_n, sh.n[r]);
  max_n = (max_n + 63) / 64 * 64;
  if (pf->g_cap < nt || pf->gstage_cap < max_n)
  {
    A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    void* old[] = { pf->d_gpose, pf->d_gorder, pf->d_gorder_work, pf->d_gex, pf->d_gstage, pf->d_gorder_tmp };
    for (void* b : old)
      if (b)
        cudaFree(b);
    pf->d_gpose = pf->d_gstage = nullptr;
    pf->d_gorder = pf->d_gorder_work = pf->d_gex = pf->d_gorder_tmp = nullptr;
    pf->g_cap = pf->gstage_cap = 0;
    const uint64_t cap = (nt + 4095) / 4096 * 4096;
    A3D_CUDA_TRY(cudaMalloc(&pf->d_gpose, cap * 4 * sizeof(float)));
    A3D_CUDA_TRY(cudaMalloc(&pf->d_gorder, cap * sizeof(uint32_t)));
    A3D_CUDA_TRY(cudaMalloc(&pf->d_gorder_work, order_work_words(cap) * sizeof(uint32_t)));
    A3D_CUDA_TRY(cudaMalloc(&pf->d_gorder_tmp, cap * sizeof(uint32_t)));
    A3D_CUDA_TRY(cudaMalloc(&pf->d_gex, cap * 2 * sizeof(uint32_t)));
    A3D_CUDA_TRY(cudaMalloc(&pf->d_gstage, static_cast<size_t>(sh.n_ranks + 1) * 4 * max_n * sizeof(float)));
    pf->g_cap = cap;
    pf->gstage_cap = max_n;
  }
  // staging layout: [send: 4 x max_n][recv: n_ranks x 4 x max_n]
  float* send = pf->d_gstage;
  float* recv = pf->d_gstage + 4 * max_n;
  if (pf->n)
  {
    float* b = pf->d_state[pf->cur];
    const size_t c = pf->cap;
    const Planes p = { b, b + c, b + 2 * c, b + 3 * c, b + 4 * c, b + 5 * c, b + 6 * c };
    pack_pose_planes_kernel<<<grid_for(ctx, pf->n, 256), 256, 0, ctx->stream>>>(p, pf->n, max_n, send);
    ctx->launches++;
  }
  A3D_TRY(comm_all_gather(ctx, send, recv, 4 * max_n * sizeof(float)));
  unpack_pose_planes_kernel<<<grid_for(ctx, nt, 256), 256, 0, ctx->stream>>>(recv, sh, max_n, pf->d_gpose);
  ctx->launches++;
  A3D_CUDA_TRY(cudaGetLastError());
  // one rank orders (ties inside a bucket are broken by atomics: two ranks would not produce the same permutation)
  if (ctx->rank == 0)
  {
    A3D_TRY(order_particles(ctx, pf->d_gpose, pf->d_gpose + nt, pf->d_gpose + 2 * nt, pf->d_gpose + 3 * nt,
                            static_cast<uint32_t>(nt), pf->cloud_r_eff, pf->d_gorder_tmp, pf->d_gorder_work));
    // Deal the sorted permutation out in chunks, round-robin over the ranks: contiguous slices of the pose order cost
    // different amounts (the sparse fringes of the pose cloud gather less coherently than its core), and the update
    // ends when the slowest rank is done.  Every rank then weighs every n_ranks-th chunk -- a representative sample --
    // while the lanes of a warp / CTA stay neighbours in pose.  Option "global_schedule_chunk" (particles; 0 = slices).
    // Measured at cfg4 on 8 GPUs (weighting kernel per rank, ms): slices 8.7 .. 11.96 (mean 9.8, the fringes of the pose
    // cloud are the slow ones), 2048-particle chunks 11.3 .. 12.8 (balanced, but every rank's voxel footprint per launch
    // is now the whole cloud's: mean 11.8), 16384-particle chunks 9.6 .. 11.25 -- the default; update 13.59 / 14.39 /
    // 12.88 ms.  On 2 GPUs the three are within 2 %.
    DealPlan plan;
    plan.chunk = ctx->opt_deal_chunk > 0 ? static_cast<uint32_t>(std::min<int64_t>(ctx->opt_deal_chunk, 1 << 24)) : 0u;
    plan.n_ranks = sh.n_ranks;
    if (plan.chunk == 0 || nt <= plan.chunk * static_cast<uint64_t>(sh.n_ranks))
      A3D_CUDA_TRY(cudaMemcpyAsync(pf->d_gorder, pf->d_gorder_tmp, nt * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    else
    {
      const uint64_t n_chunks = (nt + plan.chunk - 1) / plan.chunk;
      uint64_t at = 0;
      for (int r = 0; r < sh.n_ranks; ++r)
      {
        plan.start[r] = static_cast<uint32_t>(at);
        // chunks c with c % n_ranks == r; all of them are full except the very last chunk of the permutation
        const uint64_t mine = n_chunks > static_cast<uint64_t>(r) ? (n_chunks - 1 - r) / sh.n_ranks + 1 : 0;
        uint64_t particles = mine * plan.chunk;
        if (mine && (n_chunks - 1) % sh.n_ranks == static_cast<uint64_t>(r))
          particles -= n_chunks * plan.chunk - nt;
        at += particles;
      }
      deal_chunks_kernel<<<grid_for(ctx, nt, 256), 256, 0, ctx->stream>>>(pf->d_gorder_tmp, pf->d_gorder, nt, plan);
      ctx->launches++;
      A3D_CUDA_TRY(cudaGetLastError());
    }
  }
  A3D_TRY(comm_broadcast(ctx, pf->d_gorder, nt * sizeof(uint32_t), 0));
  return 0;
}

extern "C" {

int amcl3d_cuda_pf_stage_cloud(amcl3d_cuda_pf* pf, const float* cloud_xyzw, uint64_t n_cloud)
{
  if (!pf || (n_cloud && !cloud_xyzw))
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_stage_cloud: NULL argument");
  if (n_cloud >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_stage_cloud: cloud too large");
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  uint64_t cap_bytes = pf->cloud_cap * 16;
  A3D_TRY(reserve_bytes(reinterpret_cast<void**>(&pf->d_cloud), &cap_bytes, (n_cloud ? n_cloud : 1) * 16));
  pf->cloud_cap = cap_bytes / 16;
  pf->n_cloud = n_cloud;
  pf->cloud_sorted = false;
  if (n_cloud)
  {
    A3D_CUDA_TRY(cudaMemcpyAsync(pf->d_cloud, cloud_xyzw, n_cloud * 16, cudaMemcpyHostToDevice, ctx->stream));
    // mean range of a strided sample of the cloud: how many metres a yaw step moves a point (order.cu)
    const uint64_t stride = n_cloud > 256 ? n_cloud / 256 : 1;
    double acc = 0.0;
    uint64_t m = 0;
    for (uint64_t i = 0; i < n_cloud; i += stride, ++m)
    {
      const float* q = cloud_xyzw + 4 * i;
      const double r2 = static_cast<double>(q[0]) * q[0] + static_cast<double>(q[1]) * q[1];
      acc += (r2 == r2 && r2 < 1e30) ? std::sqrt(r2) : 0.0;
    }
    const float r_eff = m ? static_cast<float>(acc / static_cast<double>(m)) : 1.f;
    // the particle schedule depends on r_eff only through its bit budget: keep it while the range stays similar
    if (!(r_eff > 0.8f * pf->cloud_r_eff && r_eff < 1.25f * pf->cloud_r_eff))
    {
      pf->cloud_r_eff = r_eff;
      pf->order_valid = pf->gorder_valid = false;
    }
  }
  return 0;
}

int amcl3d_cuda_pf_update_staged(amcl3d_cuda_pf* pf, const amcl3d_cuda_grid* grid, const float* ranges4,
                                 uint32_t n_ranges, double alpha, double sigma, double roll, double pitch,
                                 float* mean4_out)
{
  if (!pf || !grid || (n_ranges && !ranges4))
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_update: NULL argument");
  if (grid->ctx != pf->ctx)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_update: grid and filter belong to different contexts");
  if (!grid->has_cells)
    return fail(AMCL3D_CUDA_ERR_NOT_OPEN, "pf_update: grid has no cells");
  amcl3d_cuda_ctx* ctx = pf->ctx;
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  const uint64_t n = pf->n;
  const bool sharded = ctx->n_ranks > 1;
  if (sharded && !pf->shards_valid)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_update: upload this rank's particle shard first (collective)");
  if (n == 0 && !sharded)
  {
    std::memset(pf->mean, 0, sizeof(pf->mean));
    if (mean4_out)
      std::memset(mean4_out, 0, 16);
    return 0;
  }
  if (!pf->d_block)
    A3D_TRY(reserve_particles(pf, 1));  // an empty shard still takes part in the exchange
  const uint32_t n_cloud = static_cast<uint32_t>(pf->n_cloud);
  const bool large_grid = grid->brick_shift != 0;
  // Pose-balanced weighting of a sharded set (see refresh_global_schedule): n_w lanes of this rank weigh particles of an
  // index space of n_idx particles.  Option "global_schedule": 0 = auto (on for sharded sets of >= 4096 particles),
  // 1 = off (every rank weighs its own shard).  Must be equal on all ranks, like the cloud and the particle operations.
  const bool gsched = sharded && ctx->opt_global_schedule != 1 && ctx->opt_particle_order != 1 &&
                      pf->shards.n_total >= 4096 && n_cloud > 0;
  uint64_t n_w = n, n_idx = n, slice_first = 0;
  if (gsched)
  {
    const uint64_t nt = pf->shards.n_total, per = (nt + ctx->n_ranks - 1) / ctx->n_ranks;
    slice_first = std::min<uint64_t>(nt, per * static_cast<uint64_t>(ctx->rank));
    n_w = std::min<uint64_t>(per, nt - slice_first);
    n_idx = nt;
  }
  // ---- how the per-particle cloud sums are formed
  // reference order (default): the reference adds a particle's probabilities one by one in the caller's cloud order; at
  // 10^4 points that float chain is ~1e-4 away from the exact sum, so only the same order reproduces its numbers.
  //   direct  : one lane walks the whole cloud (sequential chunk launches on large maps) -- needs >= ~2 waves of particle
  //             CTAs to be efficient;
  //   replay  : "gather anywhere, add in order" -- the gathers run at the fast path's parallelism and locality (point
  //             splits, Morton-ordered cloud) and store every value; replay_sum_kernel adds them in the caller's order.
  //             8 extra bytes of HBM traffic per evaluation, so it pays while the particle set is too small for `direct`.
  // fast (reference_order = 0, or an explicit split count / cloud_order = 2): re-associated sums, partials in double.
  const bool ref_order = ctx->opt_reference_order && ctx->opt_point_splits == 0 && ctx->opt_cloud_order != 2;
  const uint64_t vals_stride = n_cloud;                      // points per warp tile of the value matrix
  const uint64_t vals_lanes = (n_w + 31) / 32 * 32;
  bool replay = false;
  if (ref_order && n_w && n_cloud)
  {
    const uint64_t lanes = static_cast<uint64_t>(ctx->sm_count) * 1024;
    const uint64_t bytes = vals_lanes * n_cloud * sizeof(float);
    // (large maps: measured at 131 072 / 262 144 particles the direct walk wins, 18.3 / 34.6 ms against 19.4 / 36.9 ms)
    replay = ctx->opt_replay == 2 || (ctx->opt_replay == 0 && n_w < (large_grid ? lanes / 2 : 2 * lanes) &&
                                      bytes <= (static_cast<uint64_t>(ctx->opt_replay_max_mb) << 20));
  }
  // ordered: the two passes of `replay` fused in one kernel through shared memory (weight_ordered.cuh) -- gatherer warps
  // and one adder warp per 32 particles, nothing but the sums leaves the SM.  Needs the caller's cloud order in place.
  const bool ordered = ref_order && n_w && n_cloud && !pf->cloud_sorted && weight_ordered_applies(ctx, grid->view(), n_w, n_cloud);
  if (ordered)
    replay = false;
  const uint32_t splits = (n_w && !ordered) ? choose_point_splits(ctx, n_w, n_cloud, large_grid, replay) : 1;
  {
    // 8 bytes per (particle, split): float partials or double accumulators (launch_weight_batch decides)
    uint64_t cap_bytes = pf->part_cap * 8;
    const uint64_t want = (n_idx ? n_idx : 1) * splits * 8;
    if (want > cap_bytes)
    {
      A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      uint64_t c2 = pf->part_cap * 4;
      A3D_TRY(reserve_bytes(reinterpret_cast<void**>(&pf->d_part_sum), &cap_bytes, want));
      A3D_TRY(reserve_bytes(reinterpret_cast<void**>(&pf->d_part_cnt), &c2, want / 2));
      pf->part_cap = cap_bytes / 8;
    }
  }
  // beacons
  RangeParams rg;
  rg.n_ranges = n_ranges;
  rg.ranges = nullptr;
  rg.k1 = static_cast<float>(1.f / (sigma * std::sqrt(2 * M_PI)));  // ParticleFilter.cpp:231
  rg.k2 = static_cast<float>(0.5f / (sigma * sigma));               // :232
  if (n_ranges && n_ranges <= kInlineRanges)
    std::memcpy(rg.inline_ranges, ranges4, static_cast<size_t>(n_ranges) * 16);
  else if (n_ranges)
  {
    if (n_ranges > pf->ranges_cap)
    {
      A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      if (pf->d_ranges)
        cudaFree(pf->d_ranges);
      pf->d_ranges = nullptr;
      const uint32_t cap = n_ranges < 64 ? 64 : n_ranges;
      A3D_CUDA_TRY(cudaMalloc(&pf->d_ranges, static_cast<size_t>(cap) * 16));
      pf->ranges_cap = cap;
    }
    A3D_TRY(ensure_pinned(ctx, static_cast<size_t>(n_ranges) * 16 + 4096));
    // stage through pinned memory (offset 2048 keeps clear of the scalar read-back area)
    float* stage = reinterpret_cast<float*>(static_cast<char*>(ctx->pinned) + 2048);
    A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::memcpy(stage, ranges4, static_cast<size_t>(n_ranges) * 16);
    A3D_CUDA_TRY(cudaMemcpyAsync(pf->d_ranges, stage, static_cast<size_t>(n_ranges) * 16, cudaMemcpyHostToDevice, ctx->stream));
    rg.ranges = pf->d_ranges;
  }

  const GridView g = grid->view();
  // Large-map regime: re-order the cloud along a Morton curve once per staged cloud (cloud.cu).  Option "cloud_order":
  // 0 = auto (re-order when the grid is bricked and the caller did not pin the summation order with
  // weight_point_splits = 1), 1 = keep the caller's order, 2 = always re-order.
  {
    const bool want = ctx->opt_cloud_order == 2 ||
                      (ctx->opt_cloud_order == 0 && g.brick_shift != 0 && ctx->opt_point_splits != 1 && (!ref_order || replay));
    if (want && !pf->cloud_sorted && n_cloud > 1024)
    {
      if (pf->cloud_tmp_cap < n_cloud)
      {
        A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (pf->d_cloud_tmp)
          cudaFree(pf->d_cloud_tmp);
        if (pf->d_cloud_work)
          cudaFree(pf->d_cloud_work);
        pf->d_cloud_tmp = nullptr;
        pf->d_cloud_work = nullptr;
        const uint64_t cap = (static_cast<uint64_t>(n_cloud) + 4095) / 4096 * 4096;
        A3D_CUDA_TRY(cudaMalloc(&pf->d_cloud_tmp, cap * sizeof(float4)));
        A3D_CUDA_TRY(cudaMalloc(&pf->d_cloud_work, (cap + 2 * 32768 + 8) * sizeof(uint32_t)));
        pf->cloud_tmp_cap = cap;
      }
      A3D_TRY(sort_cloud_morton(ctx, pf->d_cloud, pf->d_cloud_tmp, pf->d_cloud_work, n_cloud));
      pf->cloud_sorted = true;
    }
  }
  if (replay || ordered)
  {
    // value matrix, caller index -> staged position (replay only), sums in the caller's order
    const uint64_t want_vals = replay ? vals_lanes * n_cloud : 0;
    if (want_vals > pf->vals_cap || (replay && n_cloud > pf->pos_cap) || n_idx > pf->rep_cap)
    {
      A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      if (want_vals > pf->vals_cap)
      {
        if (pf->d_vals)
          cudaFree(pf->d_vals);
        pf->d_vals = nullptr;
        pf->vals_cap = 0;
        A3D_CUDA_TRY(cudaMalloc(&pf->d_vals, want_vals * sizeof(float)));
        pf->vals_cap = want_vals;
      }
      if (replay && n_cloud > pf->pos_cap)
      {
        if (pf->d_pos_of)
          cudaFree(pf->d_pos_of);
        pf->d_pos_of = nullptr;
        const uint64_t cap = (static_cast<uint64_t>(n_cloud) + 4095) / 4096 * 4096;
        A3D_CUDA_TRY(cudaMalloc(&pf->d_pos_of, cap * sizeof(uint32_t)));
        pf->pos_cap = cap;
      }
      if (n_idx > pf->rep_cap)
      {
        if (pf->d_rep_sum)
          cudaFree(pf->d_rep_sum);
        if (pf->d_rep_cnt)
          cudaFree(pf->d_rep_cnt);
        pf->d_rep_sum = nullptr;
        pf->d_rep_cnt = nullptr;
        const uint64_t cap = (n_idx + 4095) / 4096 * 4096;
        A3D_CUDA_TRY(cudaMalloc(&pf->d_rep_sum, cap * sizeof(float)));
        A3D_CUDA_TRY(cudaMalloc(&pf->d_rep_cnt, cap * sizeof(uint32_t)));
        pf->rep_cap = cap;
      }
    }
    if (replay && pf->cloud_sorted)
      A3D_TRY(launch_cloud_pos(ctx, pf->d_cloud, n_cloud, pf->d_pos_of));
  }
  // ParticleFilter.cpp:145 narrows roll/pitch to float at the call
  const RollPitch rp = make_roll_pitch(static_cast<float>(roll), static_cast<float>(pitch));
  const Planes p = planes_of(pf, pf->cur);
  // Scheduling permutation (order.cu): lanes of a warp get neighbouring poses.  Option "particle_order".  The
  // permutation only depends on the poses: it is kept until predict / resample / upload changes them.
  const uint32_t* d_order = nullptr;
  const float *wx = p.x, *wy = p.y, *wz = p.z, *wa = p.a;  // poses the weighting kernel indexes
  if (gsched)
  {
    if (!pf->gorder_valid || pf->g_cap < n_idx)
    {
      A3D_TRY(refresh_global_schedule(pf));
      pf->gorder_valid = true;
    }
    d_order = pf->d_gorder + slice_first;
    wx = pf->d_gpose;
    wy = pf->d_gpose + n_idx;
    wz = pf->d_gpose + 2 * n_idx;
    wa = pf->d_gpose + 3 * n_idx;
    // exchange arrays: every entry is written by the one rank that weighs the particle
    A3D_CUDA_TRY(cudaMemsetAsync(pf->d_gex, 0, n_idx * 2 * sizeof(uint32_t), ctx->stream));
  }
  else if ((ctx->opt_particle_order == 2 || (ctx->opt_particle_order == 0 && n >= 4096)) && n_cloud > 0 && n > 0)
  {
    if (pf->order_cap < n)
    {
      A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      if (pf->d_order)
        cudaFree(pf->d_order);
      if (pf->d_order_work)
        cudaFree(pf->d_order_work);
      pf->d_order = pf->d_order_work = nullptr;
      pf->order_cap = 0;
      pf->order_valid = false;  // (the local permutation only: the gathered one must stay in step on all ranks)
      const uint64_t cap = (n + 4095) / 4096 * 4096;
      A3D_CUDA_TRY(cudaMalloc(&pf->d_order, cap * sizeof(uint32_t)));
      A3D_CUDA_TRY(cudaMalloc(&pf->d_order_work, order_work_words(cap) * sizeof(uint32_t)));
      pf->order_cap = cap;
    }
    if (!pf->order_valid)
    {
      A3D_TRY(order_particles(ctx, p.x, p.y, p.z, p.a, static_cast<uint32_t>(n), pf->cloud_r_eff, pf->d_order,
                              pf->d_order_work));
      pf->order_valid = true;
    }
    d_order = pf->d_order;
  }
  // the running sums stay the reference's own float chains only when every particle walks the caller's cloud order in
  // one piece; otherwise chunk partials are accumulated in double (launch_weight_batch)
  int part_kind = 0;
  if (n_w && !ordered)
    A3D_TRY(launch_weight_batch(ctx, g, pf->d_cloud, n_cloud, wx, wy, wz, wa, static_cast<uint32_t>(n_idx), rp,
                                pf->d_part_sum, pf->d_part_cnt, splits, d_order, splits == 1 && !pf->cloud_sorted,
                                &part_kind, replay ? pf->d_vals : nullptr, vals_stride, static_cast<uint32_t>(n_w)));
  // what the post kernels (and amcl3d_cuda_pf_last_cloud_weights) read: the partials of the weighting kernel, or the
  // sums replayed in the caller's order
  const void* w_sum = pf->d_part_sum;
  const uint32_t* w_cnt = pf->d_part_cnt;
  uint32_t w_splits = splits;
  if (ordered)
  {
    // (global schedule: the sums of this rank's slice go straight into the exchange arrays)
    float* r_sum = gsched ? reinterpret_cast<float*>(pf->d_gex) : pf->d_rep_sum;
    uint32_t* r_cnt = gsched ? pf->d_gex + n_idx : pf->d_rep_cnt;
    A3D_TRY(launch_weight_ordered(ctx, g, pf->d_cloud, n_cloud, wx, wy, wz, wa, static_cast<uint32_t>(n_idx), rp, r_sum, r_cnt,
                                  d_order, static_cast<uint32_t>(n_w)));
    w_sum = r_sum;
    w_cnt = r_cnt;
    w_splits = 1;
    part_kind = 0;
  }
  if (replay && n_w)
  {
    // (global schedule: the replayed sums of this rank's slice go straight into the exchange arrays)
    float* r_sum = gsched ? reinterpret_cast<float*>(pf->d_gex) : pf->d_rep_sum;
    uint32_t* r_cnt = gsched ? pf->d_gex + n_idx : pf->d_rep_cnt;
    A3D_TRY(launch_replay_sum(ctx, pf->d_vals, vals_stride, pf->cloud_sorted ? pf->d_pos_of : nullptr, n_cloud,
                              static_cast<uint32_t>(n_idx), d_order, pf->d_part_cnt, splits, r_sum, r_cnt,
                              static_cast<uint32_t>(n_w)));
    if (ctx->opt_kernel_timing)
      cudaEventRecord(ctx->ev_k1, ctx->stream);  // the replay belongs to the weighting step
    w_sum = r_sum;
    w_cnt = r_cnt;
    w_splits = 1;
    part_kind = 0;
  }
  if (gsched)
  {
    if (!(replay && n_w) && !ordered && n_w)
    {
      pack_slice_results_kernel<<<static_cast<unsigned>((n_w + 255) / 256), 256, 0, ctx->stream>>>(
          pf->d_part_sum, pf->d_part_cnt, n_idx, splits, part_kind, d_order, static_cast<uint32_t>(n_w), pf->d_gex,
          pf->d_gex + n_idx);
      ctx->launches++;
    }
    A3D_TRY(comm_all_reduce_u32(ctx, pf->d_gex, n_idx * 2));
    const uint64_t mine = pf->shards.first[ctx->rank];
    w_sum = reinterpret_cast<const float*>(pf->d_gex) + mine;
    w_cnt = pf->d_gex + n_idx + mine;
    w_splits = 1;
    part_kind = 0;
  }
  if (ctx->opt_kernel_timing && n_w)
    cudaEventRecord(ctx->ev_k2, ctx->stream);  // cloud sums are where the particles live
  pf->last_replayed = (replay || ordered) && n_w;
  pf->last_splits = w_splits;
  pf->last_kind = part_kind;
  pf->last_n = n;
  pf->last_w_sum = w_sum;
  pf->last_w_cnt = w_cnt;

  // sum_mode: 0 = auto, 1 = exact, one CTA (small sets on one GPU), 2 = fast fp64 sums, 3 = exact, segmented (any
  // particle count, any number of GPUs)
  int mode = static_cast<int>(ctx->opt_sum_mode);
  if (mode == 0)
    mode = (!sharded && n <= 2048) ? 1 : 3;
  if (mode == 1 && sharded)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_update: sum_mode 1 (single-CTA chains) is single-GPU only; use 3");
  if (mode == 1)
  {
    const int serial = ctx->opt_serial_chain ? 1 : 0;
    if (n <= 2048)  // small sets: 512 threads (128 registers each) and the single-lane chains
      update_exact_kernel<true, 512><<<1, 512, 0, ctx->stream>>>(g, p, n, w_sum, w_cnt, w_splits, part_kind, rg, alpha,
                                                                 pf->d_terms, pf->cap, pf->d_scal, serial);
    else
      update_exact_kernel<true, 1024><<<1, 1024, 0, ctx->stream>>>(g, p, n, w_sum, w_cnt, w_splits, part_kind, rg, alpha,
                                                                   pf->d_terms, pf->cap, pf->d_scal, serial);
    ctx->launches++;
  }
  else if (mode == 3)
  {
    PeerView pv;
    const int peer = comm_peer_view(ctx, &pv);
    if (sharded && !peer)
      return fail(AMCL3D_CUDA_ERR_NCCL, "pf_update: exact sums on a sharded particle set need the peer-memory mailboxes "
                                        "(CUDA IPC); use sum_mode 2 for the fp64 / ncclAllReduce path");
    A3D_TRY(launch_update_seg(pf, g, rg, alpha, w_sum, w_cnt, w_splits, part_kind, pv));
  }
  else
  {
    // fast: fp64 sums.  Sharded: the ten partials are exchanged through peer memory inside the two kernels
    // (PeerView); without a usable peer mapping the same totals come from one ncclAllReduce between them.
    PeerView pv;
    const int peer = comm_peer_view(ctx, &pv);
    const uint32_t par = pf->fast_parity;
    pf->fast_parity ^= 1u;
    // small particle sets: narrow CTAs so that the two latency-bound passes spread over all SMs (10 k particles:
    // 157 CTAs of 64 threads instead of 40 of 256)
    const int fb = n <= 32768 ? 64 : 256;
    update_fast_stage1_kernel<<<grid_for(ctx, n, fb), fb, 0, ctx->stream>>>(g, p, n, w_sum, w_cnt, w_splits, part_kind, rg,
                                                                           pf->d_scal, par, pv);
    ctx->launches++;
    if (sharded && !peer)
    {
      const int rc = comm_all_reduce_f64(ctx, pf->d_scal->dsum[par], 10);
      if (rc != 0)
      {
        // leave both accumulator buffers clean: the next update starts from zero whatever its parity
        cudaMemsetAsync(pf->d_scal->dsum, 0, sizeof(pf->d_scal->dsum) + sizeof(pf->d_scal->evals_acc), ctx->stream);
        return rc;
      }
    }
    update_fast_stage2_kernel<<<grid_for(ctx, n, fb), fb, 0, ctx->stream>>>(g, p, n, alpha, pf->d_scal, par, pv);
    ctx->launches++;
  }
  A3D_CUDA_TRY(cudaGetLastError());
  if (ctx->opt_kernel_timing && n_w)
  {
    cudaEventRecord(ctx->ev_k3, ctx->stream);
    ctx->ev_phases_valid = true;
  }
  if (mean4_out)
    return read_mean(pf, mean4_out);
  return 0;
}

int amcl3d_cuda_pf_update(amcl3d_cuda_pf* pf, const amcl3d_cuda_grid* grid, const float* cloud_xyzw, uint64_t n_cloud,
                          const float* ranges4, uint32_t n_ranges, double alpha, double sigma, double roll, double pitch,
                          float* mean4_out)
{
  A3D_TRY(amcl3d_cuda_pf_stage_cloud(pf, cloud_xyzw, n_cloud));
  float mean[4];
  A3D_TRY(amcl3d_cuda_pf_update_staged(pf, grid, ranges4, n_ranges, alpha, sigma, roll, pitch, mean));
  if (mean4_out)
    std::memcpy(mean4_out, mean, 16);
  return 0;
}

int amcl3d_cuda_pf_resample(amcl3d_cuda_pf* pf, float u01, uint32_t* idx_out)
{
  if (!pf)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_resample: NULL argument");
  const uint64_t n = pf->n;
  amcl3d_cuda_ctx* ctx = pf->ctx;
  const bool sharded = ctx->n_ranks > 1;
  if (sharded && !pf->shards_valid)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_resample: upload this rank's particle shard first (collective)");
  const uint64_t n_total = sharded ? pf->shards.n_total : n;
  if (n_total == 0)
    return 0;
  if (n_total >= 0xFFFFFFFFull)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_resample: too many particles");
  A3D_CUDA_TRY(cudaSetDevice(ctx->device));
  // resample_mode: 0 = auto (3), 1 = one CTA (windowed exact scan, or the single-lane chain with serial_chain = 1;
  // single GPU: kept as the cross-check), 3 = segmented exact chain + gather over peer memory (any size, any ranks).
  // Every mode returns the reference's indices bit for bit.
  int mode = static_cast<int>(ctx->opt_resample_mode);
  if (mode != 1)
    mode = 3;
  if (mode == 1 && sharded)
    return fail(AMCL3D_CUDA_ERR_INVALID, "pf_resample: resample_mode 1 (single-CTA chain) is single-GPU only");
  uint32_t* d_idx = idx_out ? pf->d_idx : nullptr;
  if (mode == 1)
  {
    const Planes src = planes_of(pf, pf->cur), dst = planes_of(pf, pf->cur ^ 1);
    uint64_t cap_bytes = pf->chain_cap;
    if (n * sizeof(float) > cap_bytes)
    {
      A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      A3D_TRY(reserve_bytes(reinterpret_cast<void**>(&pf->d_chain), &cap_bytes, n * sizeof(float)));
      pf->chain_cap = cap_bytes;
    }
    resample_chain_kernel<<<1, 1024, 0, ctx->stream>>>(src.w, n, pf->d_chain, ctx->opt_serial_chain ? 1 : 0);
    resample_gather_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(pf->d_chain, n, src, dst, 0, n, n, u01, d_idx);
    ctx->launches += 2;
  }
  else
  {
    PeerView pv;
    std::memset(&pv, 0, sizeof(pv));
    pv.n_ranks = 1;
    ShardView sh = pf->shards;
    if (sharded)
    {
      if (!ctx->peer_ok)
        return fail(AMCL3D_CUDA_ERR_NCCL, "pf_resample: a sharded particle set needs the peer-memory mailboxes (CUDA IPC)");
      for (int r = 0; r < ctx->n_ranks; ++r)
        pv.box[r] = static_cast<PeerBox*>(ctx->peer_box[r]);
      pv.n_ranks = ctx->n_ranks;
      pv.rank = ctx->rank;
      pv.seq = ++ctx->peer_rs_seq;
      pv.timeout_clocks = ctx->opt_peer_timeout_ms * ctx->clock_khz;
    }
    else
    {
      std::memset(&sh, 0, sizeof(sh));
      sh.n_ranks = 1;
      sh.n[0] = n;
      sh.cap[0] = pf->cap;
      sh.block[0] = pf->d_block;
      sh.n_total = n;
    }
    A3D_TRY(launch_resample_seg(pf, u01, d_idx, pv, sh));
    pf->cum_cur ^= 1;
  }
  A3D_CUDA_TRY(cudaGetLastError());
  pf->cur ^= 1;  // ParticleFilter.cpp:221  p_ = new_p
  pf->order_valid = pf->gorder_valid = false;
  if (idx_out && n)
  {
    A3D_CUDA_TRY(cudaMemcpyAsync(idx_out, pf->d_idx, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    A3D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

}  // extern "C"
