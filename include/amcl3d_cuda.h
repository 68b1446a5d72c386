/* include/amcl3d_cuda.h -- C-ABI of the B200 (sm_100a) implementation of amcl3d's measurement-update hot path.
 *
 * This is the drop-in boundary: plain C, opaque handles, caller-owned host buffers, no exceptions, no
 * C++/torch types.  It is what the host-side C++ classes (amcl3d_b200/host/{Grid3d,ParticleFilter,
 * PointCloudTools}.cpp, which keep the reference's class API) call, and what a maintainer of the
 * reference would bind if they kept their own classes (INTEGRATION.md shows that patch).
 * Every entry point names the reference code it replaces (paths relative to /root/reference/amcl3d/src).
 *
 * Conventions
 *   - return value: 0 = ok, <0 = amcl3d_cuda_status; amcl3d_cuda_last_error() gives the text (per thread).
 *   - one host thread per context; calls are stream-ordered on the context's stream and return after the
 *     results they hand back to the host are complete (they synchronise only when they return host data).
 *   - there is NO CPU fallback: without a CUDA device every call fails with AMCL3D_CUDA_ERR_NO_DEVICE.
 *   - layouts:  point    = 4 floats x,y,z,pad          (pcl::PointXYZ, 16 B)
 *               cell     = 2 floats dist,prob          (Grid3dCell, PointCloudTools.h:28-33), x-fastest
 *               particle = 7 floats x,y,z,a,w,wp,wr    (Particle, ParticleFilter.h:35-49)
 *               range    = 4 floats r,ax,ay,az         (Range, ParticleFilter.h:53-63)
 */
#ifndef AMCL3D_CUDA_H
#define AMCL3D_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMCL3D_CUDA_ABI_VERSION 1

typedef enum amcl3d_cuda_status
{
  AMCL3D_CUDA_OK = 0,
  AMCL3D_CUDA_ERR_NO_DEVICE = -1,  /* no usable CUDA device / driver */
  AMCL3D_CUDA_ERR_INVALID = -2,    /* bad argument or handle state */
  AMCL3D_CUDA_ERR_CUDA = -3,       /* a CUDA runtime call failed */
  AMCL3D_CUDA_ERR_TOO_BIG = -4,    /* grid exceeds the configured cell cap (PointCloudTools.cpp:103-105) */
  AMCL3D_CUDA_ERR_NCCL = -5,       /* NCCL missing or a collective failed */
  AMCL3D_CUDA_ERR_NOT_OPEN = -6    /* grid has no cells yet */
} amcl3d_cuda_status;

typedef struct amcl3d_cuda_ctx amcl3d_cuda_ctx;   /* device + stream + scratch (+ optional NCCL communicator) */
typedef struct amcl3d_cuda_grid amcl3d_cuda_grid; /* Grid3dInfo + PointCloudInfo bounds, resident in HBM */
typedef struct amcl3d_cuda_pf amcl3d_cuda_pf;     /* ParticleFilter state (SoA particles), resident in HBM */

/* ---------------------------------------------------------------------------------------------- context */

int amcl3d_cuda_abi_version(void);
const char* amcl3d_cuda_last_error(void);

/* Creates a context on `device`.  `stream` is a cudaStream_t to launch on (e.g. the caller's own stream so it
 * can bracket calls with its own events) or NULL to let the context create a private non-blocking stream. */
int amcl3d_cuda_ctx_create(int device, void* stream, amcl3d_cuda_ctx** out);
int amcl3d_cuda_ctx_destroy(amcl3d_cuda_ctx* ctx);
int amcl3d_cuda_ctx_set_stream(amcl3d_cuda_ctx* ctx, void* stream);
int amcl3d_cuda_ctx_synchronize(amcl3d_cuda_ctx* ctx);
/* info: [0] SM count, [1] L2 bytes, [2] persisting-L2 max bytes, [3] compute capability major*10+minor */
int amcl3d_cuda_ctx_device_info(amcl3d_cuda_ctx* ctx, int64_t info[4]);

/* Tuning / parity options (all have working defaults).
 *   "weight_point_splits"  0 = auto; k >= 1 = split the cloud into k sequential chunks per particle.
 *                          With 1 every per-particle sum runs in cloud order and is BIT-EXACT w.r.t.
 *                          Grid3d.cpp:191; with k > 1 chunk partials are added in chunk order.
 *   "reference_order"      1 (default) = every per-particle cloud sum is the reference's own float chain in the CALLER's
 *                          cloud order (Grid3d.cpp:191), bit for bit, by one of three implementations chosen by size
 *                          (direct walk / ordered kernel / gather + replay, see "ordered_mode", "replay");
 *                          0 = re-associated sums (Morton-ordered cloud, point splits, partials in double): within 2e-6 of
 *                          the exact sum, up to ~1e-4 from the reference's chain on a 3*10^4-point cloud.
 *   "replay"               0 = auto, 1 = never, 2 = always: gather anywhere + add in the caller's order through a value
 *                          matrix in HBM (8 B per evaluation; "replay_max_mb" caps the matrix, default 40960).
 *   "global_schedule"      sharded sets: 0 = auto (sets of >= 4096 particles: the weighting work is dealt out by pose over
 *                          all ranks, cloud sums return through one exact uint32 all-reduce), 1 = every rank weighs its
 *                          own shard.  Must be equal on all ranks.  Same bits either way.
 *   "sum_mode"             How the sums over particles of ParticleFilter::update (ParticleFilter.cpp:151-152,179,190-193)
 *                          are formed.  0 = auto (1 up to 2048 particles on one GPU, else 3);
 *                          1 = exact, one CTA: the reference's sequential float sums bit for bit (single GPU);
 *                          3 = exact, segmented (filter_exact.cu): the same bits at ANY particle count and on any number
 *                              of GPUs -- segment summaries built in parallel, the exact carry handed from GPU to GPU
 *                              through peer memory.  wtp, wtr, wt, every normalised weight and the mean are the
 *                              reference's (given the same per-particle weights);
 *                          2 = fast: fp64 reductions folded into one 10-value reduction.  NOT the reference's numbers
 *                              at large particle counts (its float chains are off by ~1e-5 relative at 10^6 particles);
 *                              kept as the cheapest path when only tolerance-level agreement is wanted.
 *   "resample_mode"        0 / 3 = segmented exact cumulative-weight chain + gather over peer memory (any particle
 *                          count, any number of GPUs: indices are the reference's bit for bit, ParticleFilter.cpp:207-218);
 *                          1 = the same chain in one CTA (single GPU; serial_chain = 1 makes it a single-lane loop).
 *   "peer_timeout_ms"      sharded sets: how long a kernel waits for a peer's mailbox flag before it gives up and the
 *                          next synchronising call reports AMCL3D_CUDA_ERR_NCCL (default 3000).
 *   "serial_chain"         1 = use the single-lane float chain instead of the windowed scan (cross-check).
 *   "cloud_order"          0 = auto: re-order the staged cloud along a Morton curve when the grid is larger than L2
 *                          (bricked) and weight_point_splits != 1; 1 = keep the caller's order (the summation
 *                          order of Grid3d.cpp:191); 2 = always re-order.
 *   "particle_order"       0 = auto: from 4096 particles on, the weighting kernel walks the particles in a pose-sorted
 *                          order (order.cu: neighbouring poses in neighbouring lanes => fewer cache lines per warp
 *                          request); 1 = array order; 2 = always sorted.  Scheduling only: results are bit-identical.
 *   "peer_reduce"          sharded updates: 0 = auto -- the ten partial sums are exchanged through CUDA-IPC peer memory
 *                          inside the two reduction kernels (comm.cu PeerBox; falls back to ncclAllReduce when the
 *                          mapping is not available), 1 = always ncclAllReduce.  Must be equal on all ranks.
 *   "weight_chunk_points"  points per sequential chunk launch for large particle sets (0 = auto on bricked grids: 2048 for
 *                          one-piece walks in the reference's order, 512 .. 8192 otherwise; unchunked on linear grids).  Chunk launches carry the running sums: same bits as one launch.
 *   "grid_layout"          0 = auto (linear while the probability plane fits L2, else 32^3-voxel bricks), 1 = linear,
 *                          2 = bricked.  Read at amcl3d_cuda_grid_create.  Invisible through this ABI.
 *   "weight_block_threads" 0 = auto, 64 / 128 / 256 = CTA width of the weighting kernel; 160 / 288 = 4 / 8 gatherer warps
 *                          (+ one adder warp) per 32-particle CTA of the ordered kernel.
 *   "weight_variant"       0 = v5: estimate+verify issued as packed fp32 pairs (FFMA2 / FADD2), software-pipelined
 *                          gathers (default); 4 = v4, the scalar generation (bit-identical; kept as a cross-check).
 *   "kernel_timing"        1 = record CUDA events around the weighting kernel (amcl3d_cuda_ctx_last_kernel_ms).
 *   "ordered_mode"         reference-order sums of mid-sized particle sets in ONE kernel (gatherer warps + an adder warp per
 *                          32 particles, weight_ordered.cuh): 0 = auto (linear grids, >= one group per SM, below the
 *                          one-lane-per-particle threshold), 1 = off, 2 = whenever possible.  Same bits either way.
 *   "global_schedule_chunk" sharded sets: particles per chunk of the pose-sorted schedule dealt round-robin to the ranks
 *                          (default 16384; 0 = contiguous slices).  Same bits either way.
 *   "l2_fetch_granularity" 32 / 64 / 128: cudaLimitMaxL2FetchGranularity (device-wide; no measurable effect on B200).
 *   "max_cells"            cell cap for grid creation; 0 = unlimited (reference: 250000000). Default 0.
 */
int amcl3d_cuda_ctx_set_option(amcl3d_cuda_ctx* ctx, const char* name, int64_t value);
int amcl3d_cuda_ctx_get_option(amcl3d_cuda_ctx* ctx, const char* name, int64_t* value);
/* Device time of the most recent weighting kernel (needs option kernel_timing = 1); synchronises. */
int amcl3d_cuda_ctx_last_kernel_ms(amcl3d_cuda_ctx* ctx, float* ms);
/* Device time of the three steps of the most recent amcl3d_cuda_pf_update* (option kernel_timing = 1; synchronises):
 * ms3[0] weighting kernels (= last_kernel_ms), ms3[1] what follows them until the cloud sums are where the particles
 * live (sharded sets: the exchange of the pose-balanced schedule, which also absorbs the wait for the slowest rank;
 * else ~0), ms3[2] the sums over particles (ParticleFilter.cpp:151-195: normalisations, blend, mean). */
int amcl3d_cuda_ctx_last_update_phases_ms(amcl3d_cuda_ctx* ctx, float ms3[3]);
/* Number of kernels this context has launched since creation (bench.py reports it as gpu_launches). */
int amcl3d_cuda_ctx_launch_count(amcl3d_cuda_ctx* ctx, uint64_t* count);
/* Diagnostic: the device's random-gather roofline.  Independent 4-byte read-only loads at random addresses inside
 * a `footprint_bytes` buffer (smaller than L2 -> the L2 sector rate, larger -> the HBM sector rate); groups of
 * `lanes_per_sector` (1, 2, 4, 8) neighbouring lanes share a 32-byte sector.  Returns sectors/s * 32 B in GB/s and,
 * optionally, warp-level load requests per second.  bench.py reports the weighting kernel against this figure. */
int amcl3d_cuda_probe_gather(amcl3d_cuda_ctx* ctx, uint64_t footprint_bytes, uint32_t lanes_per_sector,
                             double* sector_gbs, double* requests_per_s);

/* ---------------------------------------------------------------------------------------------- grid
 * Replaces Grid3dInfo / the grid half of Grid3d (Grid3d.h:156-158, PointCloudTools.h:37-50). */

/* Dimensions follow PointCloudTools.cpp:93-101: ceil((max - min) / resolution) per axis, in double.
 * bounds7 = min xyz, max xyz, resolution (PointCloudInfo, PointCloudTools.h:61-67). */
int amcl3d_cuda_grid_create(amcl3d_cuda_ctx* ctx, const double bounds7[7], amcl3d_cuda_grid** out);
int amcl3d_cuda_grid_destroy(amcl3d_cuda_grid* grid);
int amcl3d_cuda_grid_dims(const amcl3d_cuda_grid* grid, uint32_t dims3[3]);
int amcl3d_cuda_grid_bounds(const amcl3d_cuda_grid* grid, double bounds7[7]);

/* Installs externally computed cells (what Grid3d::loadGrid reads, Grid3d.cpp:239-275).
 * cells = size_x*size_y*size_z (dist, prob) pairs on the host. */
int amcl3d_cuda_grid_upload_cells(amcl3d_cuda_grid* grid, const float* cells, double sensor_dev);
/* Copies the cells back (what Grid3d::saveGrid writes, Grid3d.cpp:210-237). */
int amcl3d_cuda_grid_download_cells(const amcl3d_cuda_grid* grid, float* cells);
/* Probability plane only (size_x*size_y*size_z floats) -- enough for the hot path and the slice message. */
int amcl3d_cuda_grid_download_prob(const amcl3d_cuda_grid* grid, float* prob);
/* `count` probabilities starting at linear voxel index `first` (what Grid3d::buildGridSliceMsg scans,
 * Grid3d.cpp:100-118); indices past the end of the grid read as 0. */
int amcl3d_cuda_grid_download_prob_range(const amcl3d_cuda_grid* grid, uint64_t first, uint64_t count, float* prob);
/* Probabilities at `n` LOGICAL voxel indices (the reference's ix + iy*step_y + iz*step_z, Grid3d.cpp:187; 0xFFFFFFFF and
 * indices past the end read as 0): what computeCloudWeight gathers for one pose, without downloading the plane.  Used by
 * the parity checks on maps too large to copy back. */
int amcl3d_cuda_grid_gather_prob(const amcl3d_cuda_grid* grid, const uint32_t* idx, uint64_t n, float* prob_out);
/* 1 when the grid holds cells (after upload_cells / compute), else 0. */
int amcl3d_cuda_grid_has_cells(const amcl3d_cuda_grid* grid, int* has_cells);

/* computeGrid (PointCloudTools.cpp:84-149): exact nearest-map-point squared distance and
 * prob = k1*expf(-dist*dist*k2) for every voxel, on the device.  points = n_points x 4 floats on the host.
 * keep_dist = 0 skips materialising the `dist` plane (halves the footprint; hot path needs prob only). */
int amcl3d_cuda_grid_compute(amcl3d_cuda_grid* grid, const float* points_xyzw, uint64_t n_points, double sensor_dev,
                             int keep_dist);

/* ---------------------------------------------------------------------------------------------- weighting
 * Replaces Grid3d::computeCloudWeight (Grid3d.cpp:133-199) and Grid3d::isIntoMap (Grid3d.cpp:201-208). */

/* One pose.  weight_out: the return value of computeCloudWeight.  n_out (nullable): contributing points.
 * idx_out (nullable, n_cloud entries): linear voxel index per cloud point, 0xFFFFFFFF where none. */
int amcl3d_cuda_cloud_weight(const amcl3d_cuda_grid* grid, const float* cloud_xyzw, uint64_t n_cloud, float tx, float ty,
                             float tz, float roll, float pitch, float yaw, float* weight_out, uint32_t* n_out,
                             uint32_t* idx_out);

/* Many poses, one cloud (the inner loop of ParticleFilter::update, ParticleFilter.cpp:129-153, without the
 * range term).  poses_xyza = n_poses x 4 floats.  weight_out / n_out (nullable) = n_poses entries. */
int amcl3d_cuda_cloud_weight_batch(const amcl3d_cuda_grid* grid, const float* cloud_xyzw, uint64_t n_cloud,
                                   const float* poses_xyza, uint64_t n_poses, float roll, float pitch, float* weight_out,
                                   uint32_t* n_out);

int amcl3d_cuda_is_into_map(const amcl3d_cuda_grid* grid, float x, float y, float z, int* inside);

/* ---------------------------------------------------------------------------------------------- particle filter
 * Replaces ParticleFilter's particle state and its predict / update / resample loops. */

int amcl3d_cuda_pf_create(amcl3d_cuda_ctx* ctx, amcl3d_cuda_pf** out);
int amcl3d_cuda_pf_destroy(amcl3d_cuda_pf* pf);
/* std::vector<Particle> p_ (ParticleFilter.h:208) <-> device SoA. */
int amcl3d_cuda_pf_upload_particles(amcl3d_cuda_pf* pf, const float* particles7, uint64_t n);
int amcl3d_cuda_pf_download_particles(amcl3d_cuda_pf* pf, float* particles7);
int amcl3d_cuda_pf_size(const amcl3d_cuda_pf* pf, uint64_t* n);

/* ParticleFilter::init (ParticleFilter.cpp:46-95).  noise_n4 (nullable): the n x 4 Gaussian draws
 * (row 0 unused) in the reference's order; NULL = Philox4x32-10 keyed by (seed, 0). mean4_out nullable. */
int amcl3d_cuda_pf_init(amcl3d_cuda_pf* pf, uint64_t n, const float pose4[4], const float devs4[4],
                        const float* noise_n4, uint64_t seed, float* mean4_out);

/* ParticleFilter::predict (ParticleFilter.cpp:97-119).  mods4 = odom_{x,y,z,a}_mod, deltas4 = delta_{x,y,z,a}.
 * noise_n4 (nullable): injected draws N(0,|delta*mod|) per particle in x,y,z,a order (bit parity with the
 * reference's mt19937 stream); NULL = Philox4x32-10, key (seed, step), counter = global particle index. */
int amcl3d_cuda_pf_predict(amcl3d_cuda_pf* pf, const double mods4[4], const double deltas4[4], const float* noise_n4,
                           uint64_t seed, uint64_t step);

/* Stages a sensor cloud on the device (what ParticleFilter::update receives as `cloud`). */
int amcl3d_cuda_pf_stage_cloud(amcl3d_cuda_pf* pf, const float* cloud_xyzw, uint64_t n_cloud);

/* ParticleFilter::update (ParticleFilter.cpp:121-196) on the staged cloud: weighting, range likelihood
 * (computeRangeWeight, :224-244), both normalisations, blend and mean.  ranges4 = n_ranges x 4 host floats
 * (nullable when n_ranges = 0).  mean4_out (nullable): x,y,z,a of mean_; passing it makes the call
 * synchronise, NULL keeps it asynchronous (read later with amcl3d_cuda_pf_get_mean). */
int amcl3d_cuda_pf_update_staged(amcl3d_cuda_pf* pf, const amcl3d_cuda_grid* grid, const float* ranges4,
                                 uint32_t n_ranges, double alpha, double sigma, double roll, double pitch,
                                 float* mean4_out);
/* Host-buffer form: stage_cloud + update_staged + mean read-back in one call (the end-to-end path). */
int amcl3d_cuda_pf_update(amcl3d_cuda_pf* pf, const amcl3d_cuda_grid* grid, const float* cloud_xyzw, uint64_t n_cloud,
                          const float* ranges4, uint32_t n_ranges, double alpha, double sigma, double roll, double pitch,
                          float* mean4_out);
int amcl3d_cuda_pf_get_mean(amcl3d_cuda_pf* pf, float mean4_out[4]);
/* Which components of the last update's mean (bit 0 x, 1 y, 2 z, 3 yaw) are the reference's sequential float sums
 * (ParticleFilter.cpp:190-193) bit for bit.  With the exact sum modes that is every component whose terms do not nearly
 * cancel; a component that hovers around zero (|sum| < 1/8 of the sum of |terms|, e.g. a pose coordinate at the origin)
 * is returned as the fp64 sum instead -- its float chain would have to be evaluated one element at a time, while its
 * own rounding error, proportional to the running value, is orders of magnitude below the 1e-4 m tolerance. */
int amcl3d_cuda_pf_mean_exact_mask(amcl3d_cuda_pf* pf, uint32_t* mask);
/* The RAW per-particle results of the weighting step of the last update, before any normalisation: what
 * computeCloudWeight returned for each of this rank's particles (ParticleFilter.cpp:145; 0 for the particles outside the
 * map, which the reference skips) and, optionally, the number of cloud points that contributed.  weight_out / n_out
 * (nullable) hold amcl3d_cuda_pf_size entries.  Valid until the next update. */
int amcl3d_cuda_pf_last_cloud_weights(amcl3d_cuda_pf* pf, float* weight_out, uint32_t* n_out);
/* Sum over this rank's particles of the contributing-point counts of the last update (in-map evaluations). */
int amcl3d_cuda_pf_last_in_map_evals(amcl3d_cuda_pf* pf, uint64_t* evals);

/* ParticleFilter::resample (ParticleFilter.cpp:198-222).  u01 = the single uniform draw in [0,1) (:202).
 * idx_out (nullable, n entries): source index of each output particle. */
int amcl3d_cuda_pf_resample(amcl3d_cuda_pf* pf, float u01, uint32_t* idx_out);

/* ---------------------------------------------------------------------------------------------- sensor cloud
 * Replaces the pcl::VoxelGrid<pcl::PointXYZ> down-sampling the node applies to every incoming cloud right before
 * ParticleFilter::update (Node.cpp:131-137: setLeafSize(voxel_size x 3), filter): one output point per occupied
 * leaf -- the centroid of its points -- in ascending leaf-index order (x fastest).  cloud_xyzw / out_xyzw are
 * pcl::PointXYZ arrays (16 bytes per point); out_capacity in points (n_cloud always suffices).  Non-finite points are
 * dropped; when the leaf is too small for 32-bit leaf indices the input is returned unchanged, as PCL does. */
int amcl3d_cuda_voxel_grid(amcl3d_cuda_ctx* ctx, const float* cloud_xyzw, uint64_t n_cloud, float leaf_x, float leaf_y,
                           float leaf_z, float* out_xyzw, uint64_t out_capacity, uint64_t* n_out);

/* ---------------------------------------------------------------------------------------------- multi-GPU
 * Particles are block-partitioned across ranks (one process per GPU; shards may differ in size, empty ones included),
 * the grid is replicated.  With a communicator attached the filter calls become COLLECTIVE: every rank makes the same
 * sequence of upload_particles / predict / update / resample calls.
 *   upload_particles  exchanges the shard table (counts, first global indices) and maps every rank's particle block into
 *                     every other rank (CUDA IPC over NVLink);
 *   update            exchanges ten fp64 partial sums and the exact float carries of the reference's sequential sums
 *                     through peer-memory mailboxes inside its kernels (no NCCL call on the data path): weights and mean
 *                     are the single-GPU / reference bits on every rank;
 *   resample          one global low-variance resample: per-GPU exact cumulative weights, carry passed rank to rank,
 *                     every output slot searches the global cumulative weights and copies its source particle straight
 *                     from the owning GPU;
 *   predict           no communication (Philox counter = global particle index).
 * NCCL (loaded with dlopen("libnccl.so.2")) bootstraps the mailboxes and carries sum_mode 2's all-reduce when peer memory
 * is unavailable. */
int amcl3d_cuda_comm_unique_id(uint8_t id_out[128]);
int amcl3d_cuda_comm_init(amcl3d_cuda_ctx* ctx, const uint8_t id[128], int rank, int n_ranks);
int amcl3d_cuda_comm_destroy(amcl3d_cuda_ctx* ctx);
int amcl3d_cuda_comm_rank(const amcl3d_cuda_ctx* ctx, int* rank, int* n_ranks);
/* 1 when the ranks' partial sums travel through CUDA-IPC peer memory inside the update kernels, 0 when through NCCL */
int amcl3d_cuda_comm_peer_active(const amcl3d_cuda_ctx* ctx, int* active);

#ifdef __cplusplus
}
#endif
#endif /* AMCL3D_CUDA_H */
