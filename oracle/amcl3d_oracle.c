/* oracle/amcl3d_oracle.c -- plain-C restatement of amcl3d's measurement-update hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see amcl3d_oracle.h).  Parity: PINNED against the reference's goldens and
 * against the unmodified reference compiled in oracle/_ref (tests/test_oracle_*.py) -- except oracle_voxel_grid
 * at the end of this file (pcl::VoxelGrid, a third-party algorithm absent from the reference tree: parity unpinned).
 *
 * Arithmetic notes.  The reference is C++ whose unqualified sin/cos/exp/sqrt/fabs resolve to the
 * DOUBLE overloads on its platform; in C those names are double by definition, so
 * the expression types below follow the reference by construction as long as every float/double
 * promotion is written out.  Build with -ffp-contract=off (no fused multiply-add anywhere).
 */
#include "amcl3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------ grid geometry */

/* PointCloudTools.cpp:93-98: size = (uint32) ceil((max - min) / resolution), all in double. */
void oracle_grid_dims(const double* b, uint32_t* dims3)
{
  for (int a = 0; a < 3; ++a)
  {
    const double extent = b[3 + a] - b[a];
    dims3[a] = (uint32_t)ceil(extent / b[6]);
  }
}

/* FLANN L2_Simple<float> as called by KdTreeFLANN::nearestKSearch (PointCloudTools.cpp:133):
 * result = 0; for each dim: diff = query - point; result += diff*diff   -- all float. */
static inline float sq_dist_f(const float* q, const float* p)
{
  float acc = 0.f;
  for (int a = 0; a < 3; ++a)
  {
    const float diff = q[a] - p[a];
    const float sq = diff * diff;
    acc = acc + sq;
  }
  return acc;
}

/* PointCloudTools.cpp:127-129: search point = min + i*res in double, stored to float. */
static inline void voxel_search_point(const double* b, uint32_t ix, uint32_t iy, uint32_t iz, float* q)
{
  q[0] = (float)(b[0] + ((double)ix * b[6]));
  q[1] = (float)(b[1] + ((double)iy * b[6]));
  q[2] = (float)(b[2] + ((double)iz * b[6]));
}

float oracle_nn_dist2_bruteforce(const float* pts, uint64_t n, const double* b, uint32_t ix, uint32_t iy, uint32_t iz)
{
  float q[3];
  voxel_search_point(b, ix, iy, iz, q);
  float best = INFINITY;
  for (uint64_t i = 0; i < n; ++i)
  {
    const float d = sq_dist_f(q, pts + 4 * i);
    if (d < best)
      best = d;
  }
  return best;
}

/* Exact 1-NN accelerator: points bucketed into coarse blocks of BLK voxels per side; a query visits the
 * blocks in growing Chebyshev rings and stops once no unvisited block can hold a closer point.  The stop
 * test is done in double with a relative margin far above float rounding, so the float minimum it
 * returns is the same number an exhaustive scan returns (oracle_nn_dist2_bruteforce cross-checks it). */
typedef struct
{
  int nb[3];          /* blocks per axis */
  double origin[3];   /* metric origin of block (0,0,0) */
  double bsize;       /* block edge, metres */
  uint32_t* start;    /* CSR offsets, nb[0]*nb[1]*nb[2] + 1 */
  float* pts;         /* points reordered by block, 3 floats each */
} BlockIndex;

#define ORACLE_BLK 8

static int block_of(const BlockIndex* bi, int a, float c)
{
  int k = (int)floor(((double)c - bi->origin[a]) / bi->bsize);
  if (k < 0)
    k = 0;
  if (k >= bi->nb[a])
    k = bi->nb[a] - 1;
  return k;
}

static int block_index_build(BlockIndex* bi, const float* pts, uint64_t n, const double* b)
{
  /* cover the map bounds and any point outside them */
  double lo[3], hi[3];
  for (int a = 0; a < 3; ++a)
  {
    lo[a] = b[a];
    hi[a] = b[3 + a];
  }
  for (uint64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a)
    {
      const double c = pts[4 * i + a];
      if (c < lo[a])
        lo[a] = c;
      if (c > hi[a])
        hi[a] = c;
    }
  bi->bsize = b[6] * ORACLE_BLK;
  uint64_t total = 1;
  for (int a = 0; a < 3; ++a)
  {
    bi->origin[a] = lo[a];
    bi->nb[a] = (int)floor((hi[a] - lo[a]) / bi->bsize) + 1;
    total *= (uint64_t)bi->nb[a];
  }
  bi->start = (uint32_t*)calloc(total + 1, sizeof(uint32_t));
  bi->pts = (float*)malloc((n ? n : 1) * 3 * sizeof(float));
  if (!bi->start || !bi->pts)
    return -1;
  uint32_t* which = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  if (!which)
    return -1;
  for (uint64_t i = 0; i < n; ++i)
  {
    const int kx = block_of(bi, 0, pts[4 * i]), ky = block_of(bi, 1, pts[4 * i + 1]), kz = block_of(bi, 2, pts[4 * i + 2]);
    which[i] = (uint32_t)(((uint64_t)kz * bi->nb[1] + ky) * bi->nb[0] + kx);
    bi->start[which[i] + 1]++;
  }
  for (uint64_t k = 0; k < total; ++k)
    bi->start[k + 1] += bi->start[k];
  uint32_t* fill = (uint32_t*)malloc((total ? total : 1) * sizeof(uint32_t));
  if (!fill)
    return -1;
  memcpy(fill, bi->start, total * sizeof(uint32_t));
  for (uint64_t i = 0; i < n; ++i)
  {
    const uint32_t slot = fill[which[i]]++;
    bi->pts[3 * slot + 0] = pts[4 * i + 0];
    bi->pts[3 * slot + 1] = pts[4 * i + 1];
    bi->pts[3 * slot + 2] = pts[4 * i + 2];
  }
  free(fill);
  free(which);
  return 0;
}

static void block_index_free(BlockIndex* bi)
{
  free(bi->start);
  free(bi->pts);
}

static inline void scan_block(const BlockIndex* bi, int kx, int ky, int kz, const float* q, float* best)
{
  const uint64_t k = ((uint64_t)kz * bi->nb[1] + ky) * bi->nb[0] + kx;
  for (uint32_t s = bi->start[k]; s < bi->start[k + 1]; ++s)
  {
    const float d = sq_dist_f(q, bi->pts + 3 * s);
    if (d < *best)
      *best = d;
  }
}

/* `best` may carry a valid upper bound on entry (distance to a real point). */
static float nn_dist2(const BlockIndex* bi, const float* q, float best)
{
  const int c[3] = { block_of(bi, 0, q[0]), block_of(bi, 1, q[1]), block_of(bi, 2, q[2]) };
  int max_ring = 0;
  for (int a = 0; a < 3; ++a)
  {
    if (c[a] > max_ring)
      max_ring = c[a];
    if (bi->nb[a] - 1 - c[a] > max_ring)
      max_ring = bi->nb[a] - 1 - c[a];
  }
  for (int r = 0; r <= max_ring; ++r)
  {
    /* every point of ring r lies farther than (r-1)*bsize from q along some axis
       (q may sit outside the index by clamping, which only makes ring points farther) */
    if (r >= 2)
    {
      const double reach = (double)(r - 1) * bi->bsize;
      if (reach * reach > (double)best * 1.0001)
        break;
    }
    const int z0 = c[2] - r, z1 = c[2] + r, y0 = c[1] - r, y1 = c[1] + r, x0 = c[0] - r, x1 = c[0] + r;
    for (int kz = z0; kz <= z1; ++kz)
    {
      if (kz < 0 || kz >= bi->nb[2])
        continue;
      const int zface = (kz == z0 || kz == z1);
      for (int ky = y0; ky <= y1; ++ky)
      {
        if (ky < 0 || ky >= bi->nb[1])
          continue;
        const int yface = (ky == y0 || ky == y1);
        if (zface || yface)
        {
          for (int kx = x0; kx <= x1; ++kx)
            if (kx >= 0 && kx < bi->nb[0])
              scan_block(bi, kx, ky, kz, q, &best);
        }
        else
        {
          if (x0 >= 0)
            scan_block(bi, x0, ky, kz, q, &best);
          if (x1 < bi->nb[0] && x1 != x0)
            scan_block(bi, x1, ky, kz, q, &best);
        }
      }
    }
  }
  return best;
}

/* PointCloudTools.cpp:84-149 */
int oracle_compute_grid(const float* pts, uint64_t n, const double* b, double sensor_dev, float* cells, uint32_t iz0,
                        uint32_t iz1, uint64_t max_cells)
{
  uint32_t dims[3];
  oracle_grid_dims(b, dims);
  const uint64_t total = (uint64_t)dims[0] * dims[1] * dims[2];
  if (max_cells && total > max_cells)
    return -1; /* :103-105 "Octomap size is too big" */
  if (iz1 > dims[2])
    iz1 = dims[2];

  /* :114-115 -- constants evaluated in double, stored as float */
  const float gauss_c1 = (float)(1. / (sensor_dev * sqrt(2 * M_PI)));
  const float gauss_c2 = (float)(1. / (2. * sensor_dev * sensor_dev));

  BlockIndex bi;
  if (n > 0 && block_index_build(&bi, pts, n, b) != 0)
    return -2;

  const uint32_t step_y = dims[0];
  const uint32_t step_z = dims[0] * dims[1]; /* uint32 arithmetic as in :100-101 */
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t izs = (int64_t)iz0; izs < (int64_t)iz1; ++izs)
  {
    const uint32_t iz = (uint32_t)izs;
    for (uint32_t iy = 0; iy < dims[1]; ++iy)
    {
      float carry = INFINITY; /* bound inherited from the previous voxel of the row */
      float prev_q[3] = { 0, 0, 0 };
      for (uint32_t ix = 0; ix < dims[0]; ++ix)
      {
        float q[3];
        voxel_search_point(b, ix, iy, iz, q);
        const uint64_t index = (uint64_t)ix + (uint64_t)iy * step_y + (uint64_t)iz * step_z;
        float* cell = cells + 2 * index;
        if (n == 0)
        { /* :139-143 no neighbour found */
          cell[0] = -1.0f;
          cell[1] = 0.0f;
          continue;
        }
        /* a safe starting bound: the previous voxel's nearest point is at most sqrt(carry)+step away */
        float bound = INFINITY;
        if (isfinite(carry))
        {
          const double step = fabs((double)q[0] - (double)prev_q[0]);
          const double r = sqrt((double)carry) + step;
          bound = (float)(r * r * 1.001 + 1e-12);
        }
        float d = nn_dist2(&bi, q, bound);
        if (!(d < bound)) /* the bound was never beaten by a real point: redo without it */
          d = nn_dist2(&bi, q, INFINITY);
        carry = d;
        prev_q[0] = q[0];
        cell[0] = d;                                   /* :136 dist = SQUARED distance */
        cell[1] = gauss_c1 * expf(-d * d * gauss_c2);  /* :137 */
      }
    }
  }
  if (n > 0)
    block_index_free(&bi);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ weighting */

/* Grid3d.cpp:133-199 */
float oracle_cloud_weight(const float* cells, const uint32_t* dims, const double* b, const float* cloud, uint64_t n_cloud,
                          float tx, float ty, float tz, float roll, float pitch, float yaw, uint32_t* idx_out,
                          uint32_t* n_out)
{
  /* :139-144 trig of float arguments in double */
  const double sr = sin((double)roll), cr = cos((double)roll);
  const double sp = sin((double)pitch), cp = cos((double)pitch);
  const double sy = sin((double)yaw), cy = cos((double)yaw);

  /* :146-149 rotation entries: double expressions rounded to float on assignment */
  const float r00 = (float)(cy * cp), r01 = (float)(cy * sp * sr - sy * cr), r02 = (float)(cy * sp * cr + sy * sr);
  const float r10 = (float)(sy * cp), r11 = (float)(sy * sp * sr + cy * cr), r12 = (float)(sy * sp * cr - cy * sr);
  const float r20 = (float)(-sp), r21 = (float)(cp * sr), r22 = (float)(cp * cr);

  /* :151-157 */
  const double size_x = b[3] - b[0], size_y = b[4] - b[1], size_z = b[5] - b[2];
  const double off_x = (double)tx - b[0], off_y = (double)ty - b[1], off_z = (double)tz - b[2];
  const double res = b[6];

  const uint32_t step_y = dims[0], step_z = dims[0] * dims[1];
  const uint64_t grid_size = (uint64_t)dims[0] * dims[1] * dims[2];

  float weight = 0.f;
  int n = 0;
  const float error_z = 0;
  for (uint64_t k = 0; k < n_cloud; ++k)
  {
    const float px = cloud[4 * k], py = cloud[4 * k + 1], pz = cloud[4 * k + 2];
    if (idx_out)
      idx_out[k] = 0xFFFFFFFFu;
    /* :174-176 float products and sums left to right, then "+ offset" in double, stored to float */
    const float sx = px * r00 + py * r01 + (pz + error_z) * r02;
    const float sy_ = px * r10 + py * r11 + (pz + error_z) * r12;
    const float sz = px * r20 + py * r21 + (pz + error_z) * r22;
    const float nx = (float)((double)sx + off_x);
    const float ny = (float)((double)sy_ + off_y);
    const float nz = (float)((double)sz + off_z);
    /* :178-179 */
    if (nx >= 0 && (double)nx < size_x && ny >= 0 && (double)ny < size_y && nz >= 0 && (double)nz < size_z)
    {
      /* :181-183 */
      const uint32_t ix = (uint32_t)floor((double)nx / res);
      const uint32_t iy = (uint32_t)floor((double)ny / res);
      const uint32_t iz = (uint32_t)floor((double)nz / res);
      if (ix < dims[0] && iy < dims[1] && iz < dims[2]) /* :185 */
      {
        const uint32_t gi = ix + iy * step_y + iz * step_z; /* :187 uint32 arithmetic */
        if ((uint64_t)gi < grid_size)                       /* :189 */
        {
          if (cells) /* NULL: indices / count only (the caller gathers the probabilities itself) */
            weight += cells[2 * (uint64_t)gi + 1]; /* :191 prob */
          n += 1;
          if (idx_out)
            idx_out[k] = gi;
        }
      }
    }
  }
  if (n_out)
    *n_out = (uint32_t)n;
  return (n <= 10) ? 0 : weight / n; /* :198 */
}

/* Grid3d.cpp:201-208 */
int oracle_is_into_map(const double* b, float x, float y, float z)
{
  return (double)x >= b[0] && (double)x < b[3] && (double)y >= b[1] && (double)y < b[4] && (double)z >= b[2] &&
         (double)z < b[5];
}

/* ParticleFilter.cpp:224-244 */
float oracle_range_weight(float x, float y, float z, const float* ranges, uint32_t n_ranges, double sigma)
{
  if (n_ranges == 0)
    return 0;
  float w = 1;
  const float k1 = (float)(1.f / (sigma * sqrt(2 * M_PI))); /* :231 */
  const float k2 = (float)(0.5f / (sigma * sigma));         /* :232 */
  for (uint32_t i = 0; i < n_ranges; ++i)
  {
    const float ri = ranges[4 * i], ax = ranges[4 * i + 1], ay = ranges[4 * i + 2], az = ranges[4 * i + 3];
    const float d2 = (x - ax) * (x - ax) + (y - ay) * (y - ay) + (z - az) * (z - az);
    const float r = (float)sqrt((double)d2);   /* :239 */
    const float arg = -k2 * (r - ri) * (r - ri); /* float, left to right */
    w = (float)((double)w * ((double)k1 * exp((double)arg))); /* :240 */
  }
  return w;
}

/* ParticleFilter.cpp:121-196 */
/* loops 2 and 3 of ParticleFilter::update (:160-195) given the two chain totals of loop 1 */
static void oracle_normalise(float* p, uint64_t n, const double* b, double alpha, float wtp, float wtr, float* mean4)
{
  float wt = 0;
  for (uint64_t i = 0; i < n; ++i) /* :160-180 */
  {
    float* q = p + 7 * i;
    if (wtp > 0)
      q[5] /= wtp;
    else
      q[5] = 0;
    if (wtr > 0)
      q[6] /= wtr;
    else
      q[6] = 0;
    if (!oracle_is_into_map(b, q[0], q[1], q[2]))
      q[4] = 0;
    else
      q[4] = (float)((double)q[5] * alpha + (double)q[6] * (1 - alpha)); /* :178 */
    wt += q[4];
  }
  float mx = 0, my = 0, mz = 0, ma = 0;
  for (uint64_t i = 0; i < n; ++i) /* :183-194 */
  {
    float* q = p + 7 * i;
    if (wt > 0)
      q[4] /= wt;
    else
      q[4] = 0;
    mx += q[4] * q[0];
    my += q[4] * q[1];
    mz += q[4] * q[2];
    ma += q[4] * q[3];
  }
  mean4[0] = mx;
  mean4[1] = my;
  mean4[2] = mz;
  mean4[3] = ma;
}

void oracle_update(float* p, uint64_t n, const float* cells, const uint32_t* dims, const double* b, const float* cloud,
                   uint64_t n_cloud, const float* ranges, uint32_t n_ranges, double alpha, double sigma, double roll,
                   double pitch, float* mean4)
{
  float wtp = 0, wtr = 0;
  for (uint64_t i = 0; i < n; ++i) /* :129-153 */
  {
    float* q = p + 7 * i;
    const float tx = q[0], ty = q[1], tz = q[2];
    if (!oracle_is_into_map(b, tx, ty, tz))
    {
      q[4] = 0; /* wp, wr keep their old values */
      continue;
    }
    q[5] = oracle_cloud_weight(cells, dims, b, cloud, n_cloud, tx, ty, tz, (float)roll, (float)pitch, q[3], 0, 0);
    q[6] = oracle_range_weight(tx, ty, tz, ranges, n_ranges, sigma);
    wtp += q[5];
    wtr += q[6];
  }
  oracle_normalise(p, n, b, alpha, wtp, wtr, mean4);
}

/* ParticleFilter.cpp:129-195 for particles whose wp / wr (fields 5, 6) already hold the RAW computeCloudWeight /
 * computeRangeWeight results of the in-map particles (however they were obtained): the three sequential float chains
 * and the normalisations.  Test infrastructure for the parity checks at sizes where the weighting itself is only
 * checked on a subsample. */
void oracle_update_from_weights(float* p, uint64_t n, const double* b, double alpha, float* mean4)
{
  float wtp = 0, wtr = 0;
  for (uint64_t i = 0; i < n; ++i) /* :129-153 */
  {
    float* q = p + 7 * i;
    if (!oracle_is_into_map(b, q[0], q[1], q[2]))
    {
      q[4] = 0;
      continue;
    }
    wtp += q[5];
    wtr += q[6];
  }
  oracle_normalise(p, n, b, alpha, wtp, wtr, mean4);
}

/* computeCloudWeight for many poses (x, y, z, yaw) with shared roll / pitch: what loop 1 of update() evaluates.
 * OpenMP over poses (each pose is the reference's sequential loop). */
void oracle_cloud_weight_batch(const float* cells, const uint32_t* dims, const double* b, const float* cloud,
                               uint64_t n_cloud, const float* poses4, uint64_t n_poses, float roll, float pitch,
                               float* w_out, uint32_t* n_out)
{
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < (int64_t)n_poses; ++i)
  {
    uint32_t cnt = 0;
    w_out[i] = oracle_cloud_weight(cells, dims, b, cloud, n_cloud, poses4[4 * i], poses4[4 * i + 1], poses4[4 * i + 2],
                                   roll, pitch, poses4[4 * i + 3], 0, &cnt);
    if (n_out)
      n_out[i] = cnt;
  }
}

/* ParticleFilter.cpp:198-222 */
void oracle_resample(float* p, uint64_t n, float u01, uint32_t* idx_out)
{
  if (n == 0)
    return;
  float* fresh = (float*)malloc(n * 7 * sizeof(float));
  const float factor = 1.f / (float)n; /* :201 */
  const float r = factor * u01;        /* :202 */
  float c = p[4];                      /* :203 w of particle 0 */
  uint64_t i = 0;
  for (uint64_t m = 0; m < n; ++m)
  {
    const float fm = factor * (float)(uint32_t)m; /* m is uint32_t in the reference */
    const float u = r + fm;                       /* :209 */
    while (u > c)
    {
      if (++i >= n)
        break;
      c += p[7 * i + 4];
    }
    const uint64_t src = i < n ? i : n - 1; /* reference reads out of bounds here; clamp */
    memcpy(fresh + 7 * m, p + 7 * src, 7 * sizeof(float));
    fresh[7 * m + 4] = factor;
    if (idx_out)
      idx_out[m] = (uint32_t)src;
  }
  memcpy(p, fresh, n * 7 * sizeof(float));
  free(fresh);
}

/* ParticleFilter.cpp:97-119 */
void oracle_predict(float* p, uint64_t n, const double* mods, const double* d, const float* noise)
{
  (void)mods; /* the deviations are already folded into the supplied draws */
  const double delta_x = d[0], delta_y = d[1], delta_z = d[2], delta_a = d[3];
  for (uint64_t i = 0; i < n; ++i)
  {
    float* q = p + 7 * i;
    const float sa = (float)sin((double)q[3]);
    const float ca = (float)cos((double)q[3]);
    const float rand_x = (float)(delta_x + (double)noise[4 * i + 0]);
    const float rand_y = (float)(delta_y + (double)noise[4 * i + 1]);
    q[0] = q[0] + (ca * rand_x - sa * rand_y);
    q[1] = q[1] + (sa * rand_x + ca * rand_y);
    q[2] = (float)((double)q[2] + (delta_z + (double)noise[4 * i + 2]));
    q[3] = (float)((double)q[3] + (delta_a + (double)noise[4 * i + 3]));
  }
}

/* ParticleFilter.cpp:46-95 */
void oracle_init(float* p, uint64_t n, float x, float y, float z, float a, float x_dev, float y_dev, float z_dev,
                 float a_dev, const float* noise, float* mean4)
{
  (void)a_dev;
  if (n == 0)
    return;
  memset(p, 0, n * 7 * sizeof(float));
  float dev = x_dev > y_dev ? x_dev : y_dev;
  dev = dev > z_dev ? dev : z_dev;                              /* :54 */
  const float g1 = (float)(1. / ((double)dev * sqrt(2 * M_PI))); /* :55 */
  const float two_dev_dev = 2 * dev * dev;                       /* int*float*float in float */
  const float g2 = (float)(1. / (double)two_dev_dev);            /* :56 */
  p[0] = x;
  p[1] = y;
  p[2] = z;
  p[3] = a;
  p[4] = g1;
  float wt = p[4];
  for (uint64_t i = 1; i < n; ++i)
  {
    float* q = p + 7 * i;
    q[0] = p[0] + noise[4 * i + 0];
    q[1] = p[1] + noise[4 * i + 1];
    q[2] = p[2] + noise[4 * i + 2];
    q[3] = p[3] + noise[4 * i + 3];
    const float s = (q[0] - p[0]) * (q[0] - p[0]) + (q[1] - p[1]) * (q[1] - p[1]) + (q[2] - p[2]) * (q[2] - p[2]);
    const float dist = (float)sqrt((double)s);       /* :74-75 */
    const float arg = -dist * dist * g2;             /* float */
    q[4] = (float)((double)g1 * exp((double)arg));   /* :77 */
    wt += q[4];
  }
  float mx = 0, my = 0, mz = 0, ma = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    float* q = p + 7 * i;
    q[4] /= wt;
    mx += q[4] * q[0];
    my += q[4] * q[1];
    mz += q[4] * q[2];
    ma += q[4] * q[3];
  }
  mean4[0] = mx;
  mean4[1] = my;
  mean4[2] = mz;
  mean4[3] = ma;
}

/* Grid3d.cpp:277-282 with float arguments */
static uint32_t point_to_grid(const double* b, const uint32_t* dims, float x, float y, float z)
{
  const uint32_t step_y = dims[0], step_z = dims[0] * dims[1];
  return (uint32_t)(((double)x - b[0]) / b[6]) + (uint32_t)(((double)y - b[1]) / b[6]) * step_y +
         (uint32_t)(((double)z - b[2]) / b[6]) * step_z;
}

/* Grid3d.cpp:80-121 */
int64_t oracle_grid_slice(const float* cells, const uint32_t* dims, const double* b, double z, int8_t* out, uint64_t cap)
{
  if (z < b[2] || z > b[5])
    return -1;
  const uint32_t init = point_to_grid(b, dims, (float)b[0], (float)b[1], (float)z);
  const uint32_t end = point_to_grid(b, dims, (float)b[3], (float)b[4], (float)z);
  float max_prob = -1.0f;
  for (uint32_t i = init; i < end; ++i)
  {
    const float t = cells[2 * (uint64_t)i + 1];
    if (t > max_prob)
      max_prob = t;
  }
  if (max_prob < 0.000001f)
    max_prob = 0.000001f;
  max_prob = 100.f / max_prob;
  const uint32_t len = end - init;
  for (uint32_t i = 0; i < len && i < cap; ++i)
    out[i] = (int8_t)(cells[2 * (uint64_t)(init + i) + 1] * max_prob);
  return (int64_t)len;
}

/* ------------------------------------------------------------------------------------------------ voxel grid filter
 * The step before the hot path, Node.cpp:131-137: pcl::VoxelGrid<pcl::PointXYZ> with setLeafSize(v, v, v).
 * PCL is NOT in the reference tree (package.xml:27 pins only `pcl_ros`; ROS Kinetic ships PCL 1.7.2) and not in this
 * image: this is a restatement of the PUBLISHED algorithm of pcl/filters/impl/voxel_grid.hpp (applyFilter, default
 * settings) -- PARITY UNPINNED: the reference's tests hold no fixture for it.
 * std::sort leaves the order of the points inside one cell unspecified; this restatement (and the CUDA path) use the
 * input order, i.e. the sort key is (cell index, input index).
 * Returns the number of output points, or -1 when the leaf size is too small for int32 cell indices (PCL then returns
 * the input unchanged). */
typedef struct
{
  uint32_t idx, i;
} vg_item;

static int vg_cmp(const void* a, const void* b)
{
  const vg_item* x = (const vg_item*)a;
  const vg_item* y = (const vg_item*)b;
  if (x->idx != y->idx)
    return x->idx < y->idx ? -1 : 1;
  return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

int64_t oracle_voxel_grid(const float* pts_xyzw, uint64_t n, float leaf_x, float leaf_y, float leaf_z, float* out_xyzw)
{
  const float leaf[3] = { leaf_x, leaf_y, leaf_z };
  float min_p[3] = { 3.4028234e38f, 3.4028234e38f, 3.4028234e38f }, max_p[3] = { -3.4028234e38f, -3.4028234e38f, -3.4028234e38f };
  uint64_t n_finite = 0;
  for (uint64_t i = 0; i < n; ++i) /* getMinMax3D, non-dense branch: skip non-finite points */
  {
    const float* p = pts_xyzw + 4 * i;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2]))
      continue;
    for (int a = 0; a < 3; ++a)
    {
      if (p[a] < min_p[a])
        min_p[a] = p[a];
      if (p[a] > max_p[a])
        max_p[a] = p[a];
    }
    ++n_finite;
  }
  if (n_finite == 0)
    return 0;
  float inv[3];
  int min_b[3], div_b[3];
  int64_t d[3];
  for (int a = 0; a < 3; ++a)
  {
    inv[a] = 1.0f / leaf[a];                                  /* inverse_leaf_size_ = 1 / leaf_size_ (Array4f) */
    d[a] = (int64_t)((max_p[a] - min_p[a]) * inv[a]) + 1;     /* overflow check */
    min_b[a] = (int)floorf(min_p[a] * inv[a]);
    const int max_b = (int)floorf(max_p[a] * inv[a]);
    div_b[a] = max_b - min_b[a] + 1;
  }
  if (d[0] * d[1] * d[2] > (int64_t)2147483647)
    return -1;
  const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];       /* divb_mul_ */
  vg_item* items = (vg_item*)malloc((size_t)(n_finite ? n_finite : 1) * sizeof(vg_item));
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    const float* p = pts_xyzw + 4 * i;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2]))
      continue;
    const int i0 = (int)(floorf(p[0] * inv[0]) - (float)min_b[0]);
    const int i1 = (int)(floorf(p[1] * inv[1]) - (float)min_b[1]);
    const int i2 = (int)(floorf(p[2] * inv[2]) - (float)min_b[2]);
    items[m].idx = (uint32_t)(i0 + i1 * mul1 + i2 * mul2);
    items[m].i = (uint32_t)i;
    ++m;
  }
  qsort(items, (size_t)m, sizeof(vg_item), vg_cmp);
  int64_t n_out = 0;
  for (uint64_t first = 0; first < m;)
  {
    uint64_t last = first;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    while (last < m && items[last].idx == items[first].idx)
    {
      const float* p = pts_xyzw + 4 * (uint64_t)items[last].i;
      sx += p[0];
      sy += p[1];
      sz += p[2];
      ++last;
    }
    const float cnt = (float)(last - first);
    float* o = out_xyzw + 4 * n_out;
    o[0] = sx / cnt;
    o[1] = sy / cnt;
    o[2] = sz / cnt;
    o[3] = 1.0f;
    ++n_out;
    first = last;
  }
  free(items);
  return n_out;
}
