// common.cuh -- shared declarations for the sm_100a implementation behind include/amcl3d_cuda.h.
//
// Numerical contract: everything that feeds a voxel index or a weight is evaluated
// with the reference's own operand types and rounding points -- float products/sums without fused
// multiply-add, "+ offset" in double, index = floor(float / double).  The translation units are compiled
// with -fmad=false and the hot expressions additionally use the explicit _rn intrinsics, so no compiler
// setting can silently fuse them.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/amcl3d_cuda.h"

namespace amcl3d_b200
{
// ------------------------------------------------------------------------------------------ error plumbing
void set_last_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define A3D_CUDA_TRY(expr)                                                                                             \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t a3d_e_ = (expr);                                                                                       \
    if (a3d_e_ != cudaSuccess)                                                                                         \
      return ::amcl3d_b200::fail(AMCL3D_CUDA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(a3d_e_));        \
  } while (0)

#define A3D_TRY(expr)                                                                                                  \
  do                                                                                                                   \
  {                                                                                                                    \
    int a3d_r_ = (expr);                                                                                               \
    if (a3d_r_ != 0)                                                                                                   \
      return a3d_r_;                                                                                                   \
  } while (0)

// ------------------------------------------------------------------------------------------ device views
// Everything a kernel needs to know about the grid; passed by value (lives in the constant bank).
struct GridView
{
  const float* prob;  // size_x*size_y*size_z probabilities, x-fastest (Grid3dCell::prob, split out of the AoS cell)
  uint32_t size_x, size_y, size_z;
  uint32_t step_y, step_z;  // uint32 products exactly as PointCloudTools.cpp:100-101
  uint64_t n_cells;
  double min_x, min_y, min_z;     // octo_min_*
  double max_x, max_y, max_z;     // octo_max_*
  double ext_x, ext_y, ext_z;     // octo_max - octo_min (Grid3d.cpp:151-153)
  double res;                     // octo_resol
  float inv_res_f;                // float(1/res): fast-path quotient estimate
  float inv_res_lo;               // float(1/res - double(inv_res_f)): second term of the two-float reciprocal
  float ext_up_x, ext_up_y, ext_up_z;  // smallest float >= ext_*: (float v < double ext)  <=>  (v < ext_up)
  // Physical layout of `prob`.  brick_shift == 0: linear, address == the reference's linear index.
  // brick_shift == b > 0: the grid is stored as bricks of (2^b)^3 voxels (x fastest inside a brick, bricks x fastest),
  // axes padded up to whole bricks.  The LOGICAL index (ix + iy*step_y + iz*step_z) stays the parity quantity; only
  // the address differs.  Large grids use bricks so that the voxels a chunk of points can reach for all particles
  // live in a few 2 MB pages instead of one page per z-layer (linear layout: z stride = size_x*size_y*4 B).
  uint32_t brick_shift;
  uint32_t nbx, nby;  // bricks per axis (x, y)
  // address (in floats, physical layout) of a padding cell behind the plane that always holds 0.f: a point the
  // reference skips gathers this cell instead of being predicated off (adding +0 leaves a sum's bits unchanged)
  uint32_t zero_index;
};
constexpr uint64_t kZeroCellPad = 8;
// Edge of a brick of the bricked layout, log2: 32^3 voxels = 128 KB per brick.  The weighting kernel compiles it in.
constexpr uint32_t kBrickShift = 5;

// Rotation inputs shared by all particles of one update: sin/cos of roll and pitch, evaluated on the host in
// double from the float-narrowed angles exactly as Grid3d.cpp:139-142 does.
struct RollPitch
{
  double sr, cr, sp, cp;
  // third rotation row (Grid3d.cpp:149): depends on roll/pitch only, i.e. it is the same for every particle
  float r20, r21, r22;
};

struct Pose3x3
{
  float r00, r01, r02, r10, r11, r12, r20, r21, r22;
  double off_x, off_y, off_z;
};

// ------------------------------------------------------------------------------------------ exact device arithmetic
#ifdef __CUDACC__
// Grid3d.cpp:146-149 + :155-157 for one pose.  Products and sums in double in the reference's association
// order, entries rounded to float on assignment.
__device__ __forceinline__ Pose3x3 make_pose(const GridView& g, const RollPitch& rp, float tx, float ty, float tz,
                                             float yaw)
{
  double sy, cy;
  sincos(static_cast<double>(yaw), &sy, &cy);
  Pose3x3 p;
  const double cysp = __dmul_rn(cy, rp.sp), sysp = __dmul_rn(sy, rp.sp);
  p.r00 = static_cast<float>(__dmul_rn(cy, rp.cp));
  p.r01 = static_cast<float>(__dsub_rn(__dmul_rn(cysp, rp.sr), __dmul_rn(sy, rp.cr)));
  p.r02 = static_cast<float>(__dadd_rn(__dmul_rn(cysp, rp.cr), __dmul_rn(sy, rp.sr)));
  p.r10 = static_cast<float>(__dmul_rn(sy, rp.cp));
  p.r11 = static_cast<float>(__dadd_rn(__dmul_rn(sysp, rp.sr), __dmul_rn(cy, rp.cr)));
  p.r12 = static_cast<float>(__dsub_rn(__dmul_rn(sysp, rp.cr), __dmul_rn(cy, rp.sr)));
  p.r20 = static_cast<float>(-rp.sp);
  p.r21 = static_cast<float>(__dmul_rn(rp.cp, rp.sr));
  p.r22 = static_cast<float>(__dmul_rn(rp.cp, rp.cr));
  p.off_x = __dsub_rn(static_cast<double>(tx), g.min_x);
  p.off_y = __dsub_rn(static_cast<double>(ty), g.min_y);
  p.off_z = __dsub_rn(static_cast<double>(tz), g.min_z);
  return p;
}

// Address (in floats) of voxel (kx, ky, kz) in the physical layout described by GridView::brick_shift.
__device__ __forceinline__ uint32_t phys_index(const GridView& g, uint32_t kx, uint32_t ky, uint32_t kz)
{
  if (g.brick_shift == 0)
    return kx + ky * g.step_y + kz * g.step_z;
  const uint32_t b = g.brick_shift, m = (1u << b) - 1u;
  const uint32_t brick = ((kz >> b) * g.nby + (ky >> b)) * g.nbx + (kx >> b);
  return (brick << (3 * b)) | ((kz & m) << (2 * b)) | ((ky & m) << b) | (kx & m);
}

// Same, from the reference's linear index (two integer divisions: only for verification paths and copies).
__device__ __forceinline__ uint32_t logical_to_phys(const GridView& g, uint32_t gi)
{
  if (g.brick_shift == 0)
    return gi;
  const uint32_t kz = gi / g.step_z, rem = gi - kz * g.step_z;
  const uint32_t ky = rem / g.step_y, kx = rem - ky * g.step_y;
  return phys_index(g, kx, ky, kz);
}

// Grid3d.cpp:201-208
__device__ __forceinline__ bool is_into_map(const GridView& g, float x, float y, float z)
{
  const double dx = x, dy = y, dz = z;
  return dx >= g.min_x && dx < g.max_x && dy >= g.min_y && dy < g.max_y && dz >= g.min_z && dz < g.max_z;
}

// One coordinate of Grid3d.cpp:174-176: ((px*ra + py*rb) + pz*rc) in float, "+ offset" in double, to float.
__device__ __forceinline__ float transform_axis(float px, float py, float pz, float ra, float rb, float rc, double off)
{
  const float s = __fadd_rn(__fadd_rn(__fmul_rn(px, ra), __fmul_rn(py, rb)), __fmul_rn(pz, rc));
  return static_cast<float>(__dadd_rn(static_cast<double>(s), off));
}

// (uint32) floor(double(v) / res) for a float 0 <= v (Grid3d.cpp:181-183) without paying for an IEEE double
// division per coordinate: with the two-float reciprocal (inv_res_f + inv_res_lo = 1/res to ~2^-48) the quotient
// q + ql is known to a relative 2^-45, so whenever it is farther than that from an integer its floor IS the
// reference's result.  The rare value that lands inside the band of an integer takes the exact double division.
__device__ __forceinline__ uint32_t voxel_coord(float v, const GridView& g)
{
  const float q = __fmul_rn(v, g.inv_res_f);
  const float e = __fmaf_rn(v, g.inv_res_f, -q);   // exact rounding error of q
  const float ql = __fmaf_rn(v, g.inv_res_lo, e);  // low-order part of v / res
  const float magic = 12582912.f;                 // 1.5 * 2^23: adding it rounds q to the nearest integer
  const float r = __fadd_rn(q, magic);
  const float kr = __fsub_rn(r, magic);           // nearest integer to q, as a float
  const float d = __fadd_rn(__fsub_rn(q, kr), ql);  // signed distance of v/res from that integer
  const float tol = __fmaf_rn(q, 1.2e-7f * 1.2e-7f * 64.f, 1e-12f);  // ~2^-40 relative, far above the 2^-45 error
  if (!(fabsf(d) > tol) || !(q < 4.0e6f))
    return static_cast<uint32_t>(floor(static_cast<double>(v) / g.res));
  const int k = __float_as_int(r) - 0x4B400000;   // integer value of kr
  return static_cast<uint32_t>(d < 0.f ? k - 1 : k);
}

// Linear voxel index of a transformed point or 0xFFFFFFFF when the reference would skip it
// (Grid3d.cpp:178-189).
__device__ __forceinline__ uint32_t voxel_index(float nx, float ny, float nz, const GridView& g)
{
  const bool in = nx >= 0.f && nx < g.ext_up_x && ny >= 0.f && ny < g.ext_up_y && nz >= 0.f && nz < g.ext_up_z;
  if (!in)
    return 0xFFFFFFFFu;
  const uint32_t ix = voxel_coord(nx, g), iy = voxel_coord(ny, g), iz = voxel_coord(nz, g);
  if (!(ix < g.size_x && iy < g.size_y && iz < g.size_z))
    return 0xFFFFFFFFu;
  const uint32_t gi = ix + iy * g.step_y + iz * g.step_z;  // uint32 arithmetic as in :187
  return (static_cast<uint64_t>(gi) < g.n_cells) ? gi : 0xFFFFFFFFu;
}
// Combines the partial sums the weighting kernel left for particle i and applies Grid3d.cpp:198.  kind 0: float
// partials [n_splits][n]; one partial = the particle's own float chain (the reference's sum, untouched), several are
// added in double.  kind 1: double accumulators [n_splits][n] (sequential chunk launches of a re-ordered / split cloud).
// the particle's cloud sum (as the float the division of Grid3d.cpp:198 sees) and its contributing-point count
__device__ __forceinline__ float combine_partials(const void* part_sum, const uint32_t* part_cnt, uint64_t n, uint32_t n_splits,
                                                  uint64_t i, int kind, uint32_t* cnt_out)
{
  uint32_t c = part_cnt[i];
  float s;
  if (kind == 0 && n_splits == 1)
    s = static_cast<const float*>(part_sum)[i];
  else
  {
    const float* pf = static_cast<const float*>(part_sum);
    const double* pd = static_cast<const double*>(part_sum);
    double d = kind ? pd[i] : static_cast<double>(pf[i]);
#pragma unroll 4
    for (uint32_t k = 1; k < n_splits; ++k)
    {
      const size_t o = static_cast<size_t>(k) * n + i;
      d += kind ? pd[o] : static_cast<double>(pf[o]);
      c += part_cnt[o];
    }
    s = static_cast<float>(d);
  }
  *cnt_out = c;
  return s;
}

__device__ __forceinline__ float cloud_weight_from_partials(const void* part_sum, const uint32_t* part_cnt, uint64_t n,
                                                            uint32_t n_splits, uint64_t i, int kind, uint32_t* cnt_out)
{
  uint32_t c;
  const float s = combine_partials(part_sum, part_cnt, n, n_splits, i, kind, &c);
  *cnt_out = c;
  return (c <= 10u) ? 0.f : __fdiv_rn(s, static_cast<float>(static_cast<int>(c)));
}
#endif  // __CUDACC__

}  // namespace amcl3d_b200

// ------------------------------------------------------------------------------------------ handle definitions
struct NcclApi;  // comm.cu

// Device-resident scalars of one particle filter (results of the reductions inside update/resample).
struct amcl3d_pf_scalars
{
  // ---- head: what the host reads back after an update (kPfScalarsHeadBytes)
  float wtp, wtr, wt;          // ParticleFilter.cpp:126,159 running totals
  float mean[4];               // mean_ x, y, z, a (ParticleFilter.cpp:190-195)
  unsigned int mean_exact_mask;  // bit k: mean component k is the reference's sequential float sum bit for bit
  unsigned long long evals;    // sum of contributing-point counts (in-map evaluations)
  unsigned int comm_error;     // set when the peer-memory exchange of a sharded update timed out
  unsigned int ticket;         // "last block done" counter of update_fast_stage1_kernel
  // ---- device-only accumulators of the fast path, double-buffered by update parity: stage 1 of update k adds into
  // buffer k & 1, stage 2 of update k reads it and clears buffer (k + 1) & 1 for the next update (no memset launch)
  double dsum[2][12];          // fp64 partials: A, B, Px,Py,Pz,Pa, Rx,Ry,Rz,Ra, spare
  unsigned long long evals_acc[2];
  // cooperative kernels (filter_exact.cu): peer time-outs noticed between two grid-wide syncs.  Word k is written only
  // between sync k and sync k + 1 and read only after sync k + 1, so every CTA takes the same early-exit decision.
  unsigned int err_stage[4];
};
constexpr size_t kPfScalarsHeadBytes = 48;

// Peer-memory mailboxes of a sharded particle set (comm.cu sets them up, filter.cu / filter_exact.cu use them).  Every
// rank owns one PeerBox in its HBM and maps all the others through CUDA IPC over NVLink; data is stored into the
// DESTINATION's box, followed by a release store of the step number into the matching flag; readers spin on their own
// box with acquire loads.  Everything is double-buffered by step parity (a rank can be at most one step ahead).
//   vals / flag           the ten fp64 partial sums of an update, rank r -> slot [r] of every box (fast sums, and the
//                         binade hypotheses of the exact chains)
//   carry / carry_flag    exact-chain carry of phase p (0: wtp,wtr  1: wt  2: mean x,y,z,a) entering THIS rank, written
//                         by rank - 1: the running float values after the last particle of the previous shard
//   final_ / final_flag   the finished chain values of phase p, written by the last rank into every box
//   rs_*                  the same for resample: fp64 shard totals, the exact cumulative weight entering this rank, the
//                         exact cumulative weight at the end of every shard, and "chain + source planes ready" flags
constexpr int kMaxPeers = 8;
constexpr int kChainPhases = 3;
struct PeerBox
{
  double vals[2][kMaxPeers][20];
  unsigned long long flag[2][kMaxPeers];
  float carry[2][kChainPhases][4];
  unsigned int carry_mode[2][kChainPhases];   // bit k: chain k was given up (hovering sum) by an earlier rank
  unsigned long long carry_flag[2][kChainPhases];
  float final_[2][kChainPhases][4];
  unsigned int final_mode[2][kChainPhases];
  unsigned long long final_flag[2][kChainPhases];
  double rs_total[2][kMaxPeers];
  unsigned long long rs_total_flag[2][kMaxPeers];
  float rs_carry[2];
  unsigned long long rs_carry_flag[2];
  float rs_end[2][kMaxPeers];
  unsigned long long rs_ready_flag[2][kMaxPeers];
};
struct PeerView
{
  PeerBox* box[kMaxPeers];   // box[r] = rank r's box as mapped into this process (box[rank] = the local one)
  int n_ranks, rank;
  unsigned long long seq;    // step number, starts at 1 (boxes are zero-initialised)
  long long timeout_clocks;  // spin-wait limit (option "peer_timeout_ms")
};

// Particle shards of a communicator (filled by comm_exchange_shards): counts, first global indices, and every rank's
// particle block as mapped into this process.  Block layout (floats): state[0] (7 planes of cap), state[1], chain[0]
// (cap), chain[1] (cap).
struct ShardView
{
  uint64_t n[kMaxPeers];
  uint64_t first[kMaxPeers];
  uint64_t cap[kMaxPeers];
  float* block[kMaxPeers];
  uint64_t n_total;
  int n_ranks, rank;
};

struct amcl3d_cuda_ctx
{
  int device{ 0 };
  cudaStream_t stream{ nullptr };
  bool own_stream{ false };
  int sm_count{ 0 };
  int64_t l2_bytes{ 0 }, l2_persist_max{ 0 };
  int cc{ 0 };
  // options
  int64_t opt_point_splits{ 0 }, opt_sum_mode{ 0 }, opt_resample_mode{ 0 }, opt_kernel_timing{ 0 }, opt_l2_persist{ 0 },
      opt_max_cells{ 0 }, opt_block_threads{ 0 }, opt_weight_variant{ 0 }, opt_l2_fetch{ 0 }, opt_chunk_points{ 0 }, opt_grid_layout{ 0 }, opt_cloud_order{ 0 }, opt_serial_chain{ 0 }, opt_particle_order{ 0 }, opt_reference_order{ 1 }, opt_replay{ 0 }, opt_replay_max_mb{ 40960 }, opt_ordered{ 0 }, opt_global_schedule{ 0 }, opt_deal_chunk{ 16384 }, opt_order_clip{ 0 }, opt_order_bits{ 0 };
  // relative cost of a metre of pose displacement along x, y, z and of a metre of yaw-induced point motion (order.cu)
  int64_t opt_order_w[4]{ 50, 400, 3200, 100 };
  cudaEvent_t ev_k0{ nullptr }, ev_k1{ nullptr }, ev_k2{ nullptr }, ev_k3{ nullptr };
  bool ev_phases_valid{ false };
  bool ev_valid{ false };
  uint64_t launches{ 0 };
  // pinned staging for small host<->device exchanges
  void* pinned{ nullptr };
  size_t pinned_bytes{ 0 };
  // device scratch arena of the context-level entry points (cloud_weight, cloud_weight_batch, voxel_grid): grown on
  // demand, never freed per call
  void* scratch{ nullptr };
  size_t scratch_bytes{ 0 };
  // multi-GPU
  void* nccl_comm{ nullptr };
  int rank{ 0 }, n_ranks{ 1 };
  // peer-memory mailboxes (PeerBox) of all ranks, mapped with CUDA IPC; peer_ok = the fused exchange is usable
  void* peer_box[8]{ nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  bool peer_ok{ false };
  unsigned long long peer_seq{ 0 };      // update step number (PeerView::seq)
  unsigned long long peer_rs_seq{ 0 };   // resample step number
  int64_t opt_peer_reduce{ 0 };  // option "peer_reduce": 0 = auto (use when available), 1 = NCCL all-reduce
  int64_t opt_peer_timeout_ms{ 3000 };
  int64_t clock_khz{ 0 };
};

struct amcl3d_cuda_grid
{
  amcl3d_cuda_ctx* ctx{ nullptr };
  double bounds[7]{};
  uint32_t dims[3]{};
  uint64_t n_cells{ 0 };
  uint32_t brick_shift{ 0 };  // 0 = linear storage, b = bricks of (2^b)^3 voxels (GridView::brick_shift)
  uint32_t nb[3]{ 0, 0, 0 };  // bricks per axis
  uint64_t n_phys{ 0 };       // floats per plane in the physical layout (>= n_cells)
  double sensor_dev{ 0 };
  float* d_prob{ nullptr };
  float* d_dist{ nullptr };  // optional plane (only needed for .grid export)
  bool has_cells{ false };
  amcl3d_b200::GridView view() const;
};

struct amcl3d_cuda_pf
{
  amcl3d_cuda_ctx* ctx{ nullptr };
  uint64_t n{ 0 }, cap{ 0 };
  // SoA particle state, double-buffered for resample: [x y z a w wp wr] planes of `cap` floats each.  d_state[0],
  // d_state[1] and the two cumulative-weight buffers d_cum[0], d_cum[1] live in ONE allocation (d_block), so that a
  // sharded set can hand the whole block to its peers with a single CUDA-IPC handle (ShardView).
  float* d_block{ nullptr };
  float* d_state[2]{ nullptr, nullptr };
  float* d_cum[2]{ nullptr, nullptr };
  int cur{ 0 };
  int cum_cur{ 0 };
  // sharded particle set: counts / offsets / peer mappings (valid when shards.n_ranks > 1)
  ShardView shards{};
  bool shards_valid{ false };
  // segment summaries of the exact chains (filter_exact.cu)
  void* d_seg{ nullptr };
  uint64_t seg_cap{ 0 };
  // two-pass reference-order scheme: value matrix [point position][scheduled lane], caller index -> position, and the
  // replayed per-particle sums / counts
  float* d_vals{ nullptr };
  uint64_t vals_cap{ 0 };      // floats
  uint32_t* d_pos_of{ nullptr };
  uint64_t pos_cap{ 0 };
  float* d_rep_sum{ nullptr };
  uint32_t* d_rep_cnt{ nullptr };
  uint64_t rep_cap{ 0 };
  bool last_replayed{ false };
  // what the weighting step of the last update left in d_part_sum / d_part_cnt (amcl3d_cuda_pf_last_cloud_weights)
  uint32_t last_splits{ 0 };
  int last_kind{ 0 };
  uint64_t last_n{ 0 };
  bool order_valid{ false };  // d_order matches the current poses (cleared by predict / resample / upload / init)
  // Pose-balanced weighting of a sharded set (filter.cu, "global schedule"): poses of ALL ranks' particles, the scheduling
  // permutation of the whole set (computed on rank 0, broadcast), the exchange arrays (cloud sum bits | counts, one entry
  // per particle of the whole set) and the staging buffer of the pose all-gather
  float* d_gpose{ nullptr };
  uint32_t* d_gorder{ nullptr };
  uint32_t* d_gorder_work{ nullptr };
  uint32_t* d_gorder_tmp{ nullptr };  // the sorted permutation before its chunks are dealt out to the ranks
  uint32_t* d_gex{ nullptr };
  float* d_gstage{ nullptr };
  uint64_t g_cap{ 0 }, gstage_cap{ 0 };
  bool gorder_valid{ false };  // d_gpose / d_gorder match the current poses of all shards
  // what the post kernels of the last update consumed (amcl3d_cuda_pf_last_cloud_weights)
  const void* last_w_sum{ nullptr };
  const uint32_t* last_w_cnt{ nullptr };
  // staged sensor cloud
  float4* d_cloud{ nullptr };
  uint64_t n_cloud{ 0 }, cloud_cap{ 0 };
  // Morton re-ordering of the staged cloud (cloud.cu): scratch + "already re-ordered" flag
  float4* d_cloud_tmp{ nullptr };
  uint32_t* d_cloud_work{ nullptr };
  uint64_t cloud_tmp_cap{ 0 };
  bool cloud_sorted{ false };
  float cloud_r_eff{ 1.f };   // mean point range of the staged cloud (bit budget of the particle ordering, order.cu)
  // scheduling permutation of the particles for the weighting kernel (order.cu)
  uint32_t* d_order{ nullptr };
  uint32_t* d_order_work{ nullptr };
  uint64_t order_cap{ 0 };
  // scratch
  void* d_part_sum{ nullptr };   // float partials or double accumulators, 8 bytes per (particle, split)
  uint32_t* d_part_cnt{ nullptr };
  uint64_t part_cap{ 0 };
  float* d_terms{ nullptr };  // 4 planes of cap floats (chain inputs)
  float* d_chain{ nullptr };  // cap floats (resample cumulative chain) -- also reused as double scratch
  uint64_t chain_cap{ 0 };
  uint32_t* d_idx{ nullptr };
  float* d_ranges{ nullptr };
  uint32_t ranges_cap{ 0 };
  struct amcl3d_pf_scalars* d_scal{ nullptr };  // small device scalar block (see filter.cu)
  float* d_noise{ nullptr };
  uint64_t noise_cap{ 0 };
  float mean[4]{ 0, 0, 0, 0 };
  uint32_t mean_exact_mask{ 0 };
  uint64_t last_evals{ 0 };
  uint32_t fast_parity{ 0 };  // which amcl3d_pf_scalars::dsum buffer the next fast update accumulates into
  float* plane(int k) const { return d_state[cur] + static_cast<size_t>(k) * cap; }
  float* plane_alt(int k) const { return d_state[cur ^ 1] + static_cast<size_t>(k) * cap; }
};

namespace amcl3d_b200
{
// weight.cu
int launch_weight_batch(amcl3d_cuda_ctx* ctx, const GridView& g, const float4* d_cloud, uint32_t n_cloud, const float* d_x,
                        const float* d_y, const float* d_z, const float* d_a, uint32_t n_poses, const RollPitch& rp,
                        void* d_part_sum, uint32_t* d_part_cnt, uint32_t n_splits, const uint32_t* d_order,
                        bool exact_order, int* partial_kind_out, float* d_vals, uint64_t vals_stride, uint32_t n_lanes = 0);
// "gather anywhere, add in order" (weight_v5.cuh STORE + weight.cu replay_sum_kernel)
int launch_replay_sum(amcl3d_cuda_ctx* ctx, const float* d_vals, uint64_t stride, const uint32_t* d_pos_of, uint32_t n_cloud,
                      uint32_t n_poses, const uint32_t* d_order, const uint32_t* d_part_cnt, uint32_t n_splits,
                      float* d_out_sum, uint32_t* d_out_cnt, uint32_t n_lanes = 0);
int launch_cloud_pos(amcl3d_cuda_ctx* ctx, const float4* d_sorted, uint32_t n, uint32_t* d_pos_of);
// reference-order sums in one kernel: gatherer warps + one adder warp per 32 particles (weight_ordered.cuh)
int launch_weight_ordered(amcl3d_cuda_ctx* ctx, const GridView& g, const float4* d_cloud, uint32_t n_cloud, const float* d_x,
                          const float* d_y, const float* d_z, const float* d_a, uint32_t n_poses, const RollPitch& rp,
                          float* d_out_sum, uint32_t* d_out_cnt, const uint32_t* d_order, uint32_t n_lanes = 0);
bool weight_ordered_applies(const amcl3d_cuda_ctx* ctx, const GridView& g, uint64_t n_lanes, uint64_t n_cloud);
// combines the partials of launch_weight_batch into per-particle weights / counts (d_count nullable)
int launch_batch_finish(amcl3d_cuda_ctx* ctx, const void* d_part_sum, const uint32_t* d_part_cnt, uint32_t n_poses,
                        uint32_t n_splits, int kind, float* d_weight, uint32_t* d_count);
// order.cu
uint64_t order_work_words(uint64_t n);
int order_particles(amcl3d_cuda_ctx* ctx, const float* d_x, const float* d_y, const float* d_z, const float* d_a, uint32_t n,
                    float r_eff, uint32_t* d_order, uint32_t* d_work);
// fast = false: the configured policy (reference order -> 1 unless the caller set a split count); true: the split count
// that fills the GPU best (re-associated sums; also the gather pass of the two-pass reference-order scheme)
uint32_t choose_point_splits(const amcl3d_cuda_ctx* ctx, uint64_t n_poses, uint64_t n_cloud, bool large_grid, bool fast);
RollPitch make_roll_pitch(float roll, float pitch);
// comm.cu
int comm_all_reduce_f64(amcl3d_cuda_ctx* ctx, double* d_buf, size_t count);
int comm_all_reduce_u32(amcl3d_cuda_ctx* ctx, uint32_t* d_buf, size_t count);
int comm_broadcast(amcl3d_cuda_ctx* ctx, void* d_buf, size_t bytes, int root);
// Collective: all ranks publish their particle count and particle block (CUDA IPC) and map everybody else's.
int comm_exchange_shards(amcl3d_cuda_ctx* ctx, float* local_block, uint64_t cap, uint64_t n, ShardView* out);
void comm_release_shards(amcl3d_cuda_ctx* ctx, ShardView* sv);
// fills *pv for the next sharded fast update and returns 1 when the peer-memory exchange is in use (else 0: NCCL)
int comm_peer_view(amcl3d_cuda_ctx* ctx, PeerView* pv);
// cloud.cu
int sort_cloud_morton(amcl3d_cuda_ctx* ctx, float4* d_cloud, float4* d_tmp, uint32_t* d_work, uint32_t n);
int comm_all_gather(amcl3d_cuda_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank);
// distance_field.cu: exclusive prefix sum of n uint32 (three launches, synchronises the stream)
int scan_u32(amcl3d_cuda_ctx* ctx, const uint32_t* d_in, uint64_t n, uint32_t* d_out);
// api.cu
int ensure_pinned(amcl3d_cuda_ctx* ctx, size_t bytes);
// grows the context's device scratch arena to at least `bytes` (synchronises the stream when it has to re-allocate)
int ensure_scratch(amcl3d_cuda_ctx* ctx, size_t bytes);
}  // namespace amcl3d_b200
