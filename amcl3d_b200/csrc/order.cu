// order.cu -- scheduling permutation of the particle set for the weighting kernel.
//
// The weighting kernel maps one lane to one particle and lets the 32 lanes of a warp gather the SAME cloud point.
// How many distinct 128-byte lines (L1 tag look-ups) and 32-byte sectors (L2 requests) such a warp request costs is
// set by how far apart the 32 poses are -- and the particle array is in arrival order, i.e. random with respect to
// pose.  Measured on B200 (cfg2, 10 k x 10 k): 20 sectors per request, L1TEX tag stage saturated
// (requests x lines ~ elapsed cycles), instruction issue irrelevant.
//
// This file computes a permutation that puts neighbouring poses into neighbouring lanes: a counting sort over a
// Morton-style key of (yaw, x, y, z), with the 13 / 16 key bits handed to the four axes by their effect on the
// transformed points (a yaw step moves a point by range x dyaw, so yaw usually earns the most bits).  The kernel
// reads pose order[i] in lane i and writes its result back to slot order[i]: the particle arrays, every
// per-particle result and the order of all sums over particles are untouched -- the permutation is scheduling only,
// results are bit-identical with and without it, and it need not be deterministic (ties are broken by atomics).
#include <cooperative_groups.h>

#include "common.cuh"

namespace amcl3d_b200
{
namespace
{
constexpr int kSmallBits = 13;            // single-CTA path: 8192 buckets in shared memory
constexpr int kLargeBitsMax = 20;         // multi-kernel path: up to 2^20 buckets in global memory (about one per particle)
constexpr uint32_t kSmallMax = 32768;
constexpr int kMaxKeyBits = 20;

__device__ __forceinline__ uint32_t f2o(float f)  // monotone float -> uint
{
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o)
{
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct KeyPlan
{
  float lo[4], scale[4], qmax[4];  // q = clamp((v - lo) * scale, 0, qmax)
  int n_bits;                      // key bits in use
  uint8_t src_dim[kMaxKeyBits];    // key bit b (0 = least significant) is bit src_bit[b] of axis src_dim[b]
  uint8_t src_bit[kMaxKeyBits];
};

// Hands `total_bits` key bits to the axes (x, y, z, yaw): always to the axis whose cells are currently the most
// EXPENSIVE, i.e. the largest in metres of point displacement (yaw cells are multiplied by the effective point range)
// times the axis weight `axis_w`.  The weights express what a metre of spread costs in the memory system: the grid is
// stored x-fastest, so neighbours in x share a 32-byte sector / 128-byte line, neighbours in y are 128 B apart (one DRAM
// page inside a brick) and neighbours in z 4 KB apart -- a warp (and a wave of CTAs) that is tight in z touches far fewer
// lines and pages than one that is tight in x.  Measured on B200, cfg4 1 M x 32 k, reference order: weights 1:1:1:1
// 99.5 ms, z x4 85.3, z x16 80.0, (x, y, z, yaw) = (0.5, 4, 32, 1) 74.2 ms (tools/specs/r2x..r3a, profiles/r2_order_weights.md).
// The bits are then interleaved Morton-style, an axis with more bits contributing its extra bits at the coarse end.
// `clip_sigma` > 0 additionally clips the quantisation range of every axis to mean +- clip_sigma standard deviations
// (`mom`: count, sums, sums of squares); measured neutral to slightly negative here, so it is off by default.
__device__ void make_plan(const uint32_t* box, const double* mom, const float r_eff, const int total_bits, KeyPlan& kp,
                          const float clip_sigma, const float4 axis_w4)
{
  float span[4], cell[4];
  const float axis_w[4] = { axis_w4.x, axis_w4.y, axis_w4.z, axis_w4.w };
  int bits[4];
  for (int d = 0; d < 4; ++d)
  {
    float lo = o2f(box[d]), hi = o2f(box[4 + d]);
    if (mom && clip_sigma > 0.f && mom[0] >= 2.0 && box[d] <= box[4 + d])
    {
      const double mean = mom[1 + d] / mom[0];
      const double var = mom[5 + d] / mom[0] - mean * mean;
      if (var > 0.0 && var < 1e30)
      {
        const double sd = sqrt(var);
        lo = fmaxf(lo, static_cast<float>(mean - clip_sigma * sd));
        hi = fminf(hi, static_cast<float>(mean + clip_sigma * sd));
      }
    }
    kp.lo[d] = lo;
    span[d] = (box[d] <= box[4 + d] && hi > lo) ? hi - lo : 0.f;
    cell[d] = span[d] * (d == 3 ? r_eff : 1.f) * axis_w[d];
    bits[d] = 0;
  }
  int used = 0;
  for (int b = 0; b < total_bits; ++b)
  {
    int best = 0;
    for (int d = 1; d < 4; ++d)
      if (cell[d] > cell[best])
        best = d;
    if (!(cell[best] > 0.f) || bits[best] >= 12)
      break;
    bits[best]++;
    cell[best] *= 0.5f;
    ++used;
  }
  int max_bits = 0;
  for (int d = 0; d < 4; ++d)
  {
    kp.scale[d] = span[d] > 0.f ? static_cast<float>(1u << bits[d]) / span[d] : 0.f;
    kp.qmax[d] = static_cast<float>((1u << bits[d]) - 1u);
    max_bits = max(max_bits, bits[d]);
  }
  kp.n_bits = used;
  int pos = used;  // fill from the most significant key bit downwards
  for (int r = max_bits - 1; r >= 0; --r)
    for (int d = 3; d >= 0; --d)
      if (bits[d] > r)
      {
        --pos;
        kp.src_dim[pos] = static_cast<uint8_t>(d);
        kp.src_bit[pos] = static_cast<uint8_t>(r);
      }
}

__device__ __forceinline__ uint32_t pose_key(const KeyPlan& kp, const float v[4])
{
  uint32_t q[4];
#pragma unroll
  for (int d = 0; d < 4; ++d)
  {
    float t = (v[d] - kp.lo[d]) * kp.scale[d];
    t = (t == t) ? fminf(fmaxf(t, 0.f), kp.qmax[d]) : 0.f;
    q[d] = static_cast<uint32_t>(t);
  }
  uint32_t key = 0;
  for (int b = 0; b < kp.n_bits; ++b)
  {
    const uint32_t d = kp.src_dim[b];
    const uint32_t qd = d == 0 ? q[0] : (d == 1 ? q[1] : (d == 2 ? q[2] : q[3]));
    key |= ((qd >> kp.src_bit[b]) & 1u) << b;
  }
  return key;
}

__device__ __forceinline__ void box_accumulate(const float v[4], uint32_t lo[4], uint32_t hi[4])
{
#pragma unroll
  for (int d = 0; d < 4; ++d)
    if (v[d] == v[d] && fabsf(v[d]) < 1e30f)
    {
      const uint32_t o = f2o(v[d]);
      lo[d] = min(lo[d], o);
      hi[d] = max(hi[d], o);
    }
}

// moments of the finite poses: m[0] = count, m[1..4] = sums, m[5..8] = sums of squares
__device__ __forceinline__ void moment_accumulate(const float v[4], double m[9])
{
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 4; ++d)
    ok = ok && v[d] == v[d] && fabsf(v[d]) < 1e15f;
  if (!ok)
    return;
  m[0] += 1.0;
#pragma unroll
  for (int d = 0; d < 4; ++d)
  {
    m[1 + d] += static_cast<double>(v[d]);
    m[5 + d] += static_cast<double>(v[d]) * static_cast<double>(v[d]);
  }
}

// warp-reduces the nine moments and adds lane 0's totals to `dst` (shared or global memory)
__device__ __forceinline__ void moment_flush(double m[9], double* dst)
{
#pragma unroll
  for (int k = 0; k < 9; ++k)
  {
    for (int s = 16; s > 0; s >>= 1)
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], s);
    if ((threadIdx.x & 31) == 0 && m[k] != 0.0)
      atomicAdd(dst + k, m[k]);
  }
}

// exclusive scan of 1024 per-thread totals inside one 1024-thread CTA; returns this thread's offset
__device__ __forceinline__ uint32_t block_exclusive_1024(const uint32_t tot, uint32_t* warp_tot)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t incl = tot;
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o)
      incl += t;
  }
  if (lane == 31)
    warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0)
  {
    const uint32_t w = warp_tot[lane];
    uint32_t wi = w;
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o)
        wi += t;
    }
    warp_tot[lane] = wi - w;
  }
  __syncthreads();
  return warp_tot[warp] + incl - tot;
}

// ---- small sets (<= 32768 particles): ONE launch of one 8-CTA thread-block cluster.  A single CTA takes ~30 us for
// 10 k particles (one SM's issue rate), separate launches pay a launch gap per phase; a cluster gets 8 SMs and
// hardware cluster barriers, and the CTAs read each other's histograms through distributed shared memory:
//   A  every thread keeps its (<= 4) particles in registers; CTA bounding boxes -> cluster.sync -> combined box, plan
//   B  keys; per-CTA histogram over all 8192 buckets in its own shared memory                      -> cluster.sync
//   C  CTA c owns buckets [1024 c, 1024 c + 1024): it reads the 8 counts of each of its buckets over DSMEM, scans
//      its bucket totals, publishes its total -> cluster.sync -> adds the totals of the CTAs before it and writes
//      every CTA's start cursor for the bucket back into that CTA's shared memory                   -> cluster.sync
//   D  scatter through the CTA's own cursors
constexpr int kClusterCtas = 8;
constexpr int kPerThread = static_cast<int>(kSmallMax / (kClusterCtas * 1024));  // 4

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(1024)
    order_small_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                       const float* __restrict__ a, const uint32_t n, const float r_eff, uint32_t* __restrict__ order,
                       const float clip_sigma, const float4 axis_w)
{
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ uint32_t hist[1u << kSmallBits];
  __shared__ uint32_t box[8];
  __shared__ double mom[9], mom_all[9];
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t cta_total;
  __shared__ KeyPlan kp;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t cta = cluster.block_rank();
  if (tid < 4)
  {
    box[tid] = 0xFFFFFFFFu;
    box[4 + tid] = 0u;
  }
  if (tid < 9)
    mom[tid] = 0.0;
  for (uint32_t b = tid; b < (1u << kSmallBits); b += 1024)
    hist[b] = 0;
  __syncthreads();
  // A: this thread's particles are cta*1024 + tid + u * 8192
  float v[kPerThread][4];
  uint32_t lo[4] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
  for (int u = 0; u < kPerThread; ++u)
  {
    const uint32_t i = cta * 1024u + tid + u * (kClusterCtas * 1024u);
    const bool in = i < n;
    v[u][0] = in ? x[i] : NAN;
    v[u][1] = in ? y[i] : NAN;
    v[u][2] = in ? z[i] : NAN;
    v[u][3] = in ? a[i] : NAN;
  }
  double mo[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
  for (int u = 0; u < kPerThread; ++u)
  {
    box_accumulate(v[u], lo, hi);  // NaN is ignored
    moment_accumulate(v[u], mo);
  }
  for (int d = 0; d < 4; ++d)
  {
    for (int s = 16; s > 0; s >>= 1)
    {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
    if (lane == 0)
    {
      atomicMin(&box[d], lo[d]);
      atomicMax(&box[4 + d], hi[d]);
    }
  }
  moment_flush(mo, mom);
  cluster.sync();
  if (tid < 8)
  {
    uint32_t m = tid < 4 ? 0xFFFFFFFFu : 0u;
    for (int c = 0; c < kClusterCtas; ++c)
    {
      const uint32_t o = cluster.map_shared_rank(box, c)[tid];
      m = tid < 4 ? min(m, o) : max(m, o);
    }
    warp_tot[tid] = m;  // staging: the CTA's own box is still being read by the other CTAs
  }
  else if (tid >= 32 && tid < 41)
  {
    double m = 0.0;
    for (int c = 0; c < kClusterCtas; ++c)  // fixed order: every CTA gets the same sums, hence the same plan
      m += cluster.map_shared_rank(mom, c)[tid - 32];
    mom_all[tid - 32] = m;
  }
  __syncthreads();
  if (tid == 0)
    make_plan(warp_tot, mom_all, r_eff, kSmallBits, kp, clip_sigma, axis_w);
  __syncthreads();
  // B
  uint32_t key[kPerThread];
#pragma unroll
  for (int u = 0; u < kPerThread; ++u)
  {
    const uint32_t i = cta * 1024u + tid + u * (kClusterCtas * 1024u);
    key[u] = pose_key(kp, v[u]);
    if (i < n)
      atomicAdd(&hist[key[u]], 1u);
  }
  cluster.sync();
  // C: bucket b = cta * 1024 + tid
  const uint32_t b = cta * 1024u + tid;
  uint32_t cnt[kClusterCtas], tot = 0;
#pragma unroll
  for (int c = 0; c < kClusterCtas; ++c)
  {
    cnt[c] = cluster.map_shared_rank(hist, c)[b];
    tot += cnt[c];
  }
  const uint32_t local_off = block_exclusive_1024(tot, warp_tot);
  if (tid == 1023)
    cta_total = local_off + tot;
  cluster.sync();
  uint32_t run = local_off;
  for (uint32_t c = 0; c < cta; ++c)
    run += *cluster.map_shared_rank(&cta_total, c);
#pragma unroll
  for (int c = 0; c < kClusterCtas; ++c)
  {
    cluster.map_shared_rank(hist, c)[b] = run;  // CTA c's start cursor for bucket b (only this thread touches it now)
    run += cnt[c];
  }
  cluster.sync();
  // D
#pragma unroll
  for (int u = 0; u < kPerThread; ++u)
  {
    const uint32_t i = cta * 1024u + tid + u * (kClusterCtas * 1024u);
    if (i < n)
      order[atomicAdd(&hist[key[u]], 1u)] = i;
  }
}

// ---- large sets: the same steps as separate launches over up to 2^20 global buckets
// work layout (words): [0..7] box, [8] bucket-block totals (1024), [2048 ...) histogram / cursors (2^bits), then n keys
constexpr uint32_t kTotOff = 8, kMomOff = 1040, kHistOff = 2048;  // [1040..1058): nine doubles (pose moments)

__global__ void order_init_kernel(uint32_t* __restrict__ work, const uint32_t n_buckets)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4)
    work[i] = 0xFFFFFFFFu;
  else if (i < 8)
    work[i] = 0u;
  else if (i >= kMomOff && i < kMomOff + 18)
    work[i] = 0u;
  if (i < n_buckets)
    work[kHistOff + i] = 0u;
}

__global__ void order_bbox_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                  const float* __restrict__ a, const uint32_t n, uint32_t* __restrict__ work)
{
  uint32_t lo[4] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[4] = { 0u, 0u, 0u, 0u };
  double mo[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float v[4] = { x[i], y[i], z[i], a[i] };
    box_accumulate(v, lo, hi);
    moment_accumulate(v, mo);
  }
  moment_flush(mo, reinterpret_cast<double*>(work + kMomOff));
  for (int d = 0; d < 4; ++d)
  {
    for (int s = 16; s > 0; s >>= 1)
    {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicMin(&work[d], lo[d]);
      atomicMax(&work[4 + d], hi[d]);
    }
  }
}

__global__ void order_hist_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                  const float* __restrict__ a, const uint32_t n, const float r_eff, const int bits,
                                  uint32_t* __restrict__ work, uint32_t* __restrict__ keys, const float clip_sigma,
                                  const float4 axis_w)
{
  __shared__ KeyPlan kp;
  if (threadIdx.x == 0)
    make_plan(work, reinterpret_cast<const double*>(work + kMomOff), r_eff, bits, kp, clip_sigma, axis_w);  // same plan in every CTA
  __syncthreads();
  uint32_t* hist = work + kHistOff;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float v[4] = { x[i], y[i], z[i], a[i] };
    const uint32_t k = pose_key(kp, v);
    keys[i] = k;
    atomicAdd(&hist[k], 1u);
  }
}

// CTA b scans buckets [1024 b, 1024 b + 1024) in place (exclusive, local) and publishes their total
__global__ void __launch_bounds__(1024) order_scan_local_kernel(uint32_t* __restrict__ work)
{
  __shared__ uint32_t warp_tot[32];
  uint32_t* hist = work + kHistOff + blockIdx.x * 1024u;
  const uint32_t c = hist[threadIdx.x];
  const uint32_t off = block_exclusive_1024(c, warp_tot);
  hist[threadIdx.x] = off;
  if (threadIdx.x == 1023)
    work[kTotOff + blockIdx.x] = off + c;
}

// one CTA: exclusive scan of the (<= 1024) bucket-block totals
__global__ void __launch_bounds__(1024) order_scan_totals_kernel(uint32_t* __restrict__ work, const uint32_t n_blocks)
{
  __shared__ uint32_t warp_tot[32];
  const uint32_t c = threadIdx.x < n_blocks ? work[kTotOff + threadIdx.x] : 0u;
  const uint32_t off = block_exclusive_1024(c, warp_tot);
  if (threadIdx.x < n_blocks)
    work[kTotOff + threadIdx.x] = off;
}

__global__ void order_scatter_kernel(const uint32_t n, uint32_t* __restrict__ work, const uint32_t* __restrict__ keys,
                                     uint32_t* __restrict__ order)
{
  uint32_t* cursor = work + kHistOff;
  const uint32_t* block_off = work + kTotOff;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const uint32_t k = keys[i];
    order[atomicAdd(&cursor[k], 1u) + block_off[k >> 10]] = i;
  }
}

int large_bits(uint64_t n)
{
  int b = 14;
  while (b < kLargeBitsMax && (1ull << b) < n)
    ++b;
  return b;
}

}  // namespace

uint64_t order_work_words(uint64_t n)
{
  return n <= kSmallMax ? 8 : kHistOff + (1ull << large_bits(n)) + n;
}

// d_order[0..n) becomes a permutation of 0..n-1 that groups neighbouring poses; d_work holds order_work_words(n) words.
int order_particles(amcl3d_cuda_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, const float* d_a, uint32_t n,
                    float r_eff, uint32_t* d_order, uint32_t* d_work)
{
  if (n == 0)
    return 0;
  if (!(r_eff > 0.f))
    r_eff = 1.f;
  // option "order_clip_sigma_x10": half-width of the key range in tenths of a standard deviation (0 = bounding box)
  const float clip = 0.1f * static_cast<float>(ctx->opt_order_clip);
  // options "order_weight_x/y/z/yaw" (percent): relative cost of a metre of displacement along the axis
  const float4 aw = make_float4(0.01f * ctx->opt_order_w[0], 0.01f * ctx->opt_order_w[1], 0.01f * ctx->opt_order_w[2],
                                0.01f * ctx->opt_order_w[3]);
  if (n <= kSmallMax)
  {
    order_small_kernel<<<kClusterCtas, 1024, 0, ctx->stream>>>(d_x, d_y, d_z, d_a, n, r_eff, d_order, clip, aw);
    ctx->launches++;
  }
  else
  {
    const int bits = (ctx->opt_order_bits >= 14 && ctx->opt_order_bits <= large_bits(n)) ? static_cast<int>(ctx->opt_order_bits) : large_bits(n);
    const uint32_t n_buckets = 1u << bits;
    uint32_t* keys = d_work + kHistOff + n_buckets;
    const int blocks = static_cast<int>(std::min<uint32_t>((n + 255) / 256, static_cast<uint32_t>(ctx->sm_count) * 8));
    order_init_kernel<<<n_buckets / 256, 256, 0, ctx->stream>>>(d_work, n_buckets);
    order_bbox_kernel<<<blocks, 256, 0, ctx->stream>>>(d_x, d_y, d_z, d_a, n, d_work);
    order_hist_kernel<<<blocks, 256, 0, ctx->stream>>>(d_x, d_y, d_z, d_a, n, r_eff, bits, d_work, keys, clip, aw);
    order_scan_local_kernel<<<n_buckets / 1024, 1024, 0, ctx->stream>>>(d_work);
    order_scan_totals_kernel<<<1, 1024, 0, ctx->stream>>>(d_work, n_buckets / 1024);
    order_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(n, d_work, keys, d_order);
    ctx->launches += 6;
  }
  A3D_CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace amcl3d_b200
