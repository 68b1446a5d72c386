/* oracle/amcl3d_oracle.h -- CPU restatement of amcl3d's measurement-update hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may load it,
 * and only as the checker.  The CUDA path never calls into this file.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py) against
 *   (1) the reference's own known-answer vectors (KAT-1 computeCloudWeight golden
 *       3.8109049797058105, tests/Grid3dTest.cpp:128-132; KAT-2 nav_msg.bin probability slice;
 *       KAT-3 isIntoMap, tests/Grid3dTest.cpp:245-266), and
 *   (2) bit-for-bit against the UNMODIFIED reference sources compiled here into
 *       oracle/_ref/libamcl3d_ref.so (oracle/Makefile, target `ref`).
 * ONE EXCEPTION -- oracle_voxel_grid (the pcl::VoxelGrid step of Node.cpp:131-137): PARITY UNPINNED.  PCL is a
 * third-party dependency that is neither in the reference tree (package.xml:27 pins only `pcl_ros`; ROS Kinetic ships
 * PCL 1.7.2) nor in this image, and the reference's tests hold no fixture for the filter; the function restates the
 * published algorithm of pcl/filters/impl/voxel_grid.hpp and is cross-checked only against an independent numpy
 * restatement (tests/test_oracle_voxel_grid.py).
 *
 * Layouts (shared with the C-ABI in include/amcl3d_cuda.h):
 *   point    : 4 floats  x, y, z, pad          (pcl::PointXYZ)
 *   cell     : 2 floats  dist, prob            (Grid3dCell, PointCloudTools.h:28-33), x-fastest
 *   particle : 7 floats  x, y, z, a, w, wp, wr (Particle, ParticleFilter.h:35-49)
 *   range    : 4 floats  r, ax, ay, az         (Range, ParticleFilter.h:53-63)
 *   bounds7  : 7 doubles min xyz, max xyz, resolution (PointCloudInfo, PointCloudTools.h:54-68)
 */
#ifndef AMCL3D_ORACLE_H
#define AMCL3D_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PointCloudTools.cpp:93-101 -- grid dimensions ceil((max-min)/res) per axis, in double. */
void oracle_grid_dims(const double* bounds7, uint32_t* dims3);

/* PointCloudTools.cpp:84-149 -- computeGrid for the z-layers [iz0, iz1) (whole grid: 0, dims[2]).
 * `cells` is the FULL grid array (2 floats per voxel); only the requested layers are written.
 * `max_cells` = 0 lifts the reference's 250 M-cell cap (PointCloudTools.cpp:103-105), otherwise the
 * call fails (returns -1) when the grid exceeds it.  Returns 0 on success. */
int oracle_compute_grid(const float* points_xyzw, uint64_t n_points, const double* bounds7, double sensor_dev,
                        float* cells, uint32_t iz0, uint32_t iz1, uint64_t max_cells);

/* Squared float distance from the search point of voxel (ix,iy,iz) to its exact nearest map point,
 * by brute force over all points (slow; cross-check for the bucketed search above). */
float oracle_nn_dist2_bruteforce(const float* points_xyzw, uint64_t n_points, const double* bounds7, uint32_t ix,
                                 uint32_t iy, uint32_t iz);

/* Grid3d.cpp:133-199 -- computeCloudWeight for one pose.  idx_out (nullable, n entries) receives the
 * linear voxel index of every cloud point, 0xFFFFFFFF for points that do not contribute.
 * n_out (nullable) receives the number of contributing points. */
float oracle_cloud_weight(const float* cells, const uint32_t* dims3, const double* bounds7, const float* cloud_xyzw,
                          uint64_t n_cloud, float tx, float ty, float tz, float roll, float pitch, float yaw,
                          uint32_t* idx_out, uint32_t* n_out);

/* Grid3d.cpp:201-208 */
int oracle_is_into_map(const double* bounds7, float x, float y, float z);

/* ParticleFilter.cpp:224-244 */
float oracle_range_weight(float x, float y, float z, const float* ranges4, uint32_t n_ranges, double sigma);

/* ParticleFilter.cpp:121-196 -- update(): weights, normalisation, mean (mean4 = x, y, z, a). */
void oracle_update(float* particles7, uint64_t n, const float* cells, const uint32_t* dims3, const double* bounds7,
                   const float* cloud_xyzw, uint64_t n_cloud, const float* ranges4, uint32_t n_ranges, double alpha,
                   double sigma, double roll, double pitch, float* mean4);

/* ParticleFilter.cpp:198-222 -- resample() given the single uniform draw u01 in [0,1).
 * idx_out (nullable, n entries) receives the source index of every output particle.  Where the
 * reference would read p_[i] with i == n (undefined behaviour), the source index is clamped to n-1. */
void oracle_resample(float* particles7, uint64_t n, float u01, uint32_t* idx_out);

/* ParticleFilter.cpp:151-195 -- the three sequential float chains and normalisations of update() for particles whose
 * wp / wr fields already hold the RAW per-particle weights of the in-map particles (parity checks at sizes where the
 * weighting itself is only verified on a subsample). */
void oracle_update_from_weights(float* particles7, uint64_t n, const double* bounds7, double alpha, float* mean4);

/* Grid3d.cpp:133-199 for many poses (x, y, z, yaw; shared roll / pitch), OpenMP over poses.  cells == NULL: counts
 * only. */
void oracle_cloud_weight_batch(const float* cells, const uint32_t* dims3, const double* bounds7, const float* cloud_xyzw,
                               uint64_t n_cloud, const float* poses4, uint64_t n_poses, float roll, float pitch,
                               float* w_out, uint32_t* n_out);

/* ParticleFilter.cpp:97-119 -- predict() with the Gaussian draws supplied (noise_n4: x, y, z, a per particle,
 * exactly the values ranGaussian(0, |delta*mod|) returned). */
void oracle_predict(float* particles7, uint64_t n, const double* mods4, const double* deltas4, const float* noise_n4);

/* ParticleFilter.cpp:46-95 -- init() with the Gaussian draws supplied (noise_n4 row 0 unused). */
void oracle_init(float* particles7, uint64_t n, float x, float y, float z, float a, float x_dev, float y_dev,
                 float z_dev, float a_dev, const float* noise_n4, float* mean4);

/* Grid3d.cpp:80-121,277-282 -- buildGridSliceMsg payload.  Returns -1 when the reference returns false,
 * else the payload length (writes min(cap, length) bytes). */
int64_t oracle_grid_slice(const float* cells, const uint32_t* dims3, const double* bounds7, double z, int8_t* out,
                          uint64_t cap);

/* Node.cpp:131-137 -- pcl::VoxelGrid<pcl::PointXYZ>::filter, restated from the published PCL algorithm (PCL itself is
 * absent: parity unpinned).  out_xyzw must hold n points.  Returns the output count, or -1 for PCL's "leaf size too
 * small" pass-through. */
int64_t oracle_voxel_grid(const float* pts_xyzw, uint64_t n, float leaf_x, float leaf_y, float leaf_z, float* out_xyzw);

#ifdef __cplusplus
}
#endif
#endif
