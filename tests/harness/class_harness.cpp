// Test harness: a flat extern "C" view of the amcl3d C++ class API (Grid3d / ParticleFilter /
// PointCloudTools) so that pytest can drive it through ctypes.
//
// The SAME source is compiled twice (that it compiles against both is itself the drop-in check):
//   * -DAMCL3D_HARNESS_REFERENCE : against the UNMODIFIED reference sources in
//     /root/reference/amcl3d/src (built by oracle/Makefile into oracle/_ref/libamcl3d_ref.so).
//     White-box access (particle vector, mt19937) is obtained with `#define private public`
//     around the reference headers only.
//   * otherwise: against this repo's B200 host classes in amcl3d_b200/host/ (built by
//     amcl3d_b200/build.py into amcl3d_b200/lib/libamcl3d_host.so), using their additive
//     public accessors.
//
// This file is test infrastructure.  It is not part of the product and is never on a timed path
// except as the `--impl reference` / cpu_baseline arm of bench.py (reference build only).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <geometry_msgs/PoseArray.h>
#include <nav_msgs/OccupancyGrid.h>
#include <octomap/OcTree.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>

#ifdef AMCL3D_HARNESS_REFERENCE
#define private public
#include "Grid3d.h"
#include "ParticleFilter.h"
#undef private
#else
#include "Grid3d.h"
#include "ParticleFilter.h"
#include <pcl/filters/voxel_grid.h>  // the device-backed stand-in (the reference build has no PCL filters at all)
#endif

using amcl3d::Grid3d;
using amcl3d::Grid3dCell;
using amcl3d::Grid3dInfo;
using amcl3d::Particle;
using amcl3d::ParticleFilter;
using amcl3d::PointCloudInfo;
using amcl3d::Range;

namespace
{
typedef pcl::PointCloud<pcl::PointXYZ> Cloud;

// cloud arrives as n x 4 floats (x, y, z, pad) -- the float4 layout of pcl::PointXYZ.
Cloud::Ptr makeCloud(const float* xyzw, uint64_t n)
{
  Cloud::Ptr c(new Cloud());
  c->points.resize(n);
  for (uint64_t i = 0; i < n; ++i)
  {
    c->points[i].x = xyzw[4 * i + 0];
    c->points[i].y = xyzw[4 * i + 1];
    c->points[i].z = xyzw[4 * i + 2];
  }
  c->width = static_cast<uint32_t>(n);
  c->height = 1;
  return c;
}

PointCloudInfo::Ptr makePcInfo(const float* xyzw, uint64_t n, const double* bounds7)
{
  PointCloudInfo::Ptr pc(new PointCloudInfo());
  pc->cloud = makeCloud(xyzw, n);
  pc->octo_min_x = bounds7[0];
  pc->octo_min_y = bounds7[1];
  pc->octo_min_z = bounds7[2];
  pc->octo_max_x = bounds7[3];
  pc->octo_max_y = bounds7[4];
  pc->octo_max_z = bounds7[5];
  pc->octo_resol = bounds7[6];
  return pc;
}

struct GridBox
{
  Grid3d grid;
  // cached cloud object for repeated computeCloudWeight / update calls on the same sensor cloud
  Cloud::Ptr cloud;
};
struct FilterBox
{
  ParticleFilter pf;
};
}  // namespace

extern "C" {

// 1 when the unqualified libm calls inside namespace amcl3d resolve to the double overloads
// (SURVEY.md App. C): the oracle build asserts this.
int h_math_overloads_are_double()
{
  using namespace amcl3d;
  const float one = 1.f;
  return sizeof(sin(one)) == 8 && sizeof(exp(one)) == 8 && sizeof(sqrt(one)) == 8 && sizeof(fabs(one)) == 8;
}

const char* h_impl_name()
{
#ifdef AMCL3D_HARNESS_REFERENCE
  return "reference";
#else
  return "amcl3d_b200";
#endif
}

// ------------------------------------------------------------------------------------------ Grid3d
void* h_grid_new() { return new GridBox(); }
void h_grid_free(void* g) { delete static_cast<GridBox*>(g); }

int h_grid_open(void* g, const char* map_path, double sensor_dev)
{
  return static_cast<GridBox*>(g)->grid.open(map_path, sensor_dev) ? 1 : 0;
}

// Builds the grid from an in-memory map cloud + bounds (bounds7 = min xyz, max xyz, resolution):
// the PointCloudInfo that computePointCloud would have produced, then computeGrid.
// Returns 1 on success, 0 on exception (message to stderr when AMCL3D_LOG=1).
int h_grid_open_from_cloud(void* g, const float* xyzw, uint64_t n, const double* bounds7, double sensor_dev)
{
  GridBox* b = static_cast<GridBox*>(g);
  try
  {
    PointCloudInfo::Ptr pc = makePcInfo(xyzw, n, bounds7);
#ifdef AMCL3D_HARNESS_REFERENCE
    b->grid.pc_info_ = pc;
    b->grid.grid_info_ = amcl3d::computeGrid(pc, sensor_dev);
    return 1;
#else
    return b->grid.openFromPointCloud(pc, sensor_dev) ? 1 : 0;
#endif
  }
  catch (const std::exception& e)
  {
    ROS_ERROR("h_grid_open_from_cloud: %s", e.what());
    return 0;
  }
}

// Installs externally computed cells ((dist, prob) pairs, x-fastest) together with the map info.
int h_grid_set_cells(void* g, const float* xyzw, uint64_t n, const double* bounds7, double sensor_dev,
                     const uint32_t* dims3, const float* cells)
{
  GridBox* b = static_cast<GridBox*>(g);
  PointCloudInfo::Ptr pc = makePcInfo(xyzw, n, bounds7);
  Grid3dInfo::Ptr gi(new Grid3dInfo());
  gi->sensor_dev = sensor_dev;
  gi->size_x = dims3[0];
  gi->size_y = dims3[1];
  gi->size_z = dims3[2];
  gi->step_y = dims3[0];
  gi->step_z = dims3[0] * dims3[1];
  const uint64_t cells_n = static_cast<uint64_t>(dims3[0]) * dims3[1] * dims3[2];
  gi->grid.resize(cells_n);
  std::memcpy(static_cast<void*>(gi->grid.data()), cells, cells_n * sizeof(Grid3dCell));
#ifdef AMCL3D_HARNESS_REFERENCE
  b->grid.pc_info_ = pc;
  b->grid.grid_info_ = gi;
  return 1;
#else
  return b->grid.setGrid(pc, gi) ? 1 : 0;
#endif
}

int h_grid_dims(void* g, uint32_t* dims3)
{
  GridBox* b = static_cast<GridBox*>(g);
#ifdef AMCL3D_HARNESS_REFERENCE
  Grid3dInfo::Ptr gi = b->grid.grid_info_;
#else
  Grid3dInfo::ConstPtr gi = b->grid.gridInfo();
#endif
  if (!gi)
    return 0;
  dims3[0] = gi->size_x;
  dims3[1] = gi->size_y;
  dims3[2] = gi->size_z;
  return 1;
}

// Copies the (dist, prob) cells out; `cells` must hold 2 * size_x*size_y*size_z floats.
int h_grid_get_cells(void* g, float* cells)
{
  GridBox* b = static_cast<GridBox*>(g);
#ifdef AMCL3D_HARNESS_REFERENCE
  Grid3dInfo::Ptr gi = b->grid.grid_info_;
#else
  Grid3dInfo::ConstPtr gi = b->grid.gridInfo();
#endif
  if (!gi)
    return 0;
  std::memcpy(cells, static_cast<const void*>(gi->grid.data()), gi->grid.size() * sizeof(Grid3dCell));
  return 1;
}

void h_grid_set_cloud(void* g, const float* xyzw, uint64_t n)
{
  static_cast<GridBox*>(g)->cloud = makeCloud(xyzw, n);
}

float h_grid_cloud_weight(void* g, float tx, float ty, float tz, float roll, float pitch, float yaw)
{
  GridBox* b = static_cast<GridBox*>(g);
  return b->grid.computeCloudWeight(b->cloud, tx, ty, tz, roll, pitch, yaw);
}

int h_grid_is_into_map(void* g, float x, float y, float z)
{
  return static_cast<GridBox*>(g)->grid.isIntoMap(x, y, z) ? 1 : 0;
}

// buildGridSliceMsg: returns -1 when the call reports false, else the payload length; writes at most
// `cap` bytes to `out` and the message geometry to info4 = (width, height, resolution, origin z).
int64_t h_grid_slice(void* g, double z, int8_t* out, uint64_t cap, double* info4)
{
  nav_msgs::OccupancyGrid msg;
  if (!static_cast<GridBox*>(g)->grid.buildGridSliceMsg(z, msg))
    return -1;
  const uint64_t n = std::min<uint64_t>(cap, msg.data.size());
  if (n)
    std::memcpy(out, msg.data.data(), n);
  if (info4)
  {
    info4[0] = msg.info.width;
    info4[1] = msg.info.height;
    info4[2] = msg.info.resolution;
    info4[3] = msg.info.origin.position.z;
  }
  return static_cast<int64_t>(msg.data.size());
}

// buildMapPointCloudMsg: returns -1 on false, else number of points; copies up to cap points (16 B each).
int64_t h_grid_map_cloud(void* g, float* xyzw_out, uint64_t cap)
{
  sensor_msgs::PointCloud2 msg;
  if (!static_cast<GridBox*>(g)->grid.buildMapPointCloudMsg(msg))
    return -1;
  const uint64_t n = static_cast<uint64_t>(msg.width) * msg.height;
  const uint64_t m = std::min<uint64_t>(cap, n);
  if (m)
    std::memcpy(xyzw_out, msg.data.data(), m * 16);
  return static_cast<int64_t>(n);
}

// Map info read back from the opened grid: bounds7 = min xyz, max xyz, resolution. Returns #map points or -1.
int64_t h_grid_map_info(void* g, double* bounds7)
{
  GridBox* b = static_cast<GridBox*>(g);
#ifdef AMCL3D_HARNESS_REFERENCE
  PointCloudInfo::Ptr pc = b->grid.pc_info_;
#else
  PointCloudInfo::ConstPtr pc = b->grid.pointCloudInfo();
#endif
  if (!pc)
    return -1;
  bounds7[0] = pc->octo_min_x;
  bounds7[1] = pc->octo_min_y;
  bounds7[2] = pc->octo_min_z;
  bounds7[3] = pc->octo_max_x;
  bounds7[4] = pc->octo_max_y;
  bounds7[5] = pc->octo_max_z;
  bounds7[6] = pc->octo_resol;
  return pc->cloud ? static_cast<int64_t>(pc->cloud->size()) : 0;
}

// ------------------------------------------------------------------------------------------ free functions
// openOcTree + computePointCloud; returns number of occupied points, or -1 with the exception text in err.
int64_t h_tools_load_octomap(const char* path, double* bounds7, float* xyzw_out, uint64_t cap, char* err, uint64_t err_cap)
{
  try
  {
    boost::shared_ptr<octomap::OcTree> tree = amcl3d::openOcTree(path);
    PointCloudInfo::Ptr pc = amcl3d::computePointCloud(tree);
    bounds7[0] = pc->octo_min_x;
    bounds7[1] = pc->octo_min_y;
    bounds7[2] = pc->octo_min_z;
    bounds7[3] = pc->octo_max_x;
    bounds7[4] = pc->octo_max_y;
    bounds7[5] = pc->octo_max_z;
    bounds7[6] = pc->octo_resol;
    const uint64_t n = pc->cloud->size();
    const uint64_t m = std::min<uint64_t>(cap, n);
    for (uint64_t i = 0; i < m; ++i)
    {
      xyzw_out[4 * i + 0] = pc->cloud->points[i].x;
      xyzw_out[4 * i + 1] = pc->cloud->points[i].y;
      xyzw_out[4 * i + 2] = pc->cloud->points[i].z;
      xyzw_out[4 * i + 3] = 1.f;
    }
    return static_cast<int64_t>(n);
  }
  catch (const std::exception& e)
  {
    if (err && err_cap)
    {
      std::strncpy(err, e.what(), err_cap - 1);
      err[err_cap - 1] = 0;
    }
    return -1;
  }
}

// Writes an octomap file (".bt" when as_ot == 0, else ".ot") whose occupied leaves are the voxels containing the
// given points; depths (nullable) gives the tree depth of each leaf (16 = finest, 15 = 2x2x2 voxels, ...), and
// free_xyzw adds free leaves (they only widen the metric bounds).  Uses the compat octomap implementation.
int h_tools_write_octomap(const char* path, const float* xyzw, uint64_t n, const uint8_t* depths, const float* free_xyzw,
                          uint64_t n_free, double res, int as_ot)
{
  octomap::OcTree tree(res);
  for (uint64_t i = 0; i < n; ++i)
    tree.setLeaf(tree.coordToKey(xyzw[4 * i]), tree.coordToKey(xyzw[4 * i + 1]), tree.coordToKey(xyzw[4 * i + 2]),
                 depths ? depths[i] : 16, true);
  for (uint64_t i = 0; i < n_free; ++i)
    tree.setLeaf(tree.coordToKey(free_xyzw[4 * i]), tree.coordToKey(free_xyzw[4 * i + 1]),
                 tree.coordToKey(free_xyzw[4 * i + 2]), 16, false);
  tree.updateInnerOccupancy();
  return (as_ot ? tree.write(path) : tree.writeBinary(path)) ? 1 : 0;
}

// Library options of the B200 build (no-op for the reference build, which has none).
int h_set_option(const char* name, int64_t value)
{
#ifdef AMCL3D_HARNESS_REFERENCE
  (void)name;
  (void)value;
  return 0;
#else
  return amcl3d_cuda_ctx_set_option(amcl3d::cuda::context(), name, value);
#endif
}

// computePointCloud(nullptr) must throw (PointCloudToolsTest.cpp:136-154): returns 1 if it did.
int h_tools_null_tree_throws()
{
  try
  {
    amcl3d::computePointCloud(nullptr);
  }
  catch (const std::exception&)
  {
    return 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ ParticleFilter
void* h_pf_new() { return new FilterBox(); }
void h_pf_free(void* p) { delete static_cast<FilterBox*>(p); }

// Re-seeds the filter's mt19937 (the reference seeds from std::random_device and has no seed API).
void h_pf_seed(void* p, uint32_t seed)
{
#ifdef AMCL3D_HARNESS_REFERENCE
  static_cast<FilterBox*>(p)->pf.generator_.seed(seed);
#else
  static_cast<FilterBox*>(p)->pf.seed(seed);
#endif
}

int h_pf_is_initialized(void* p) { return static_cast<FilterBox*>(p)->pf.isInitialized() ? 1 : 0; }

void h_pf_init(void* p, int n, float x, float y, float z, float a, float xd, float yd, float zd, float ad)
{
  static_cast<FilterBox*>(p)->pf.init(n, x, y, z, a, xd, yd, zd, ad);
}

uint64_t h_pf_size(void* p)
{
#ifdef AMCL3D_HARNESS_REFERENCE
  return static_cast<FilterBox*>(p)->pf.p_.size();
#else
  return static_cast<FilterBox*>(p)->pf.size();
#endif
}

// particles as n x 7 floats: x, y, z, a, w, wp, wr (the layout of amcl3d::Particle)
void h_pf_set_particles(void* p, const float* aos7, uint64_t n)
{
  std::vector<Particle> v(n);
  static_assert(sizeof(Particle) == 28, "Particle must be 7 floats");
  if (n)
    std::memcpy(static_cast<void*>(v.data()), aos7, n * sizeof(Particle));
#ifdef AMCL3D_HARNESS_REFERENCE
  static_cast<FilterBox*>(p)->pf.p_ = v;
  static_cast<FilterBox*>(p)->pf.initialized_ = true;
#else
  static_cast<FilterBox*>(p)->pf.setParticles(v);
#endif
}

void h_pf_get_particles(void* p, float* aos7)
{
#ifdef AMCL3D_HARNESS_REFERENCE
  const std::vector<Particle>& v = static_cast<FilterBox*>(p)->pf.p_;
#else
  const std::vector<Particle> v = static_cast<FilterBox*>(p)->pf.getParticles();
#endif
  if (!v.empty())
    std::memcpy(aos7, static_cast<const void*>(v.data()), v.size() * sizeof(Particle));
}

void h_pf_get_mean(void* p, float* out7)
{
  const Particle m = static_cast<FilterBox*>(p)->pf.getMean();
  std::memcpy(out7, static_cast<const void*>(&m), sizeof(Particle));
}

void h_pf_predict(void* p, const double* mods4, const double* deltas4)
{
  static_cast<FilterBox*>(p)->pf.predict(mods4[0], mods4[1], mods4[2], mods4[3], deltas4[0], deltas4[1], deltas4[2],
                                         deltas4[3]);
}

// ranges as n_ranges x 4 floats (r, ax, ay, az); uses the grid's cached cloud (h_grid_set_cloud).
void h_pf_update(void* p, void* g, const float* ranges4, uint32_t n_ranges, double alpha, double sigma, double roll,
                 double pitch)
{
  GridBox* gb = static_cast<GridBox*>(g);
  std::vector<Range> ranges;
  ranges.reserve(n_ranges);
  for (uint32_t i = 0; i < n_ranges; ++i)
    ranges.push_back(Range(ranges4[4 * i], ranges4[4 * i + 1], ranges4[4 * i + 2], ranges4[4 * i + 3]));
  static_cast<FilterBox*>(p)->pf.update(gb->grid, gb->cloud, ranges, alpha, sigma, roll, pitch);
}

void h_pf_resample(void* p) { static_cast<FilterBox*>(p)->pf.resample(); }

// buildParticlesPoseMsg: n x 7 doubles (position xyz, orientation xyzw)
uint64_t h_pf_pose_msg(void* p, double* out7, uint64_t cap)
{
  geometry_msgs::PoseArray msg;
  static_cast<FilterBox*>(p)->pf.buildParticlesPoseMsg(msg);
  const uint64_t n = std::min<uint64_t>(cap, msg.poses.size());
  for (uint64_t i = 0; i < n; ++i)
  {
    out7[7 * i + 0] = msg.poses[i].position.x;
    out7[7 * i + 1] = msg.poses[i].position.y;
    out7[7 * i + 2] = msg.poses[i].position.z;
    out7[7 * i + 3] = msg.poses[i].orientation.x;
    out7[7 * i + 4] = msg.poses[i].orientation.y;
    out7[7 * i + 5] = msg.poses[i].orientation.z;
    out7[7 * i + 6] = msg.poses[i].orientation.w;
  }
  return msg.poses.size();
}

// ------------------------------------------------------------------------------------------ RNG replay helpers
// Exactly the draws ParticleFilter::ranGaussian / rngUniform make (ParticleFilter.cpp:246-256): a fresh
// std::normal_distribution<float>(mean, sigma) / std::uniform_real_distribution<float>(0, 1) per call on
// one mt19937.  Used to inject the reference's own noise into the CUDA path (never re-implemented).
void* h_rng_new(uint32_t seed) { return new std::mt19937(seed); }
void h_rng_free(void* r) { delete static_cast<std::mt19937*>(r); }
float h_rng_gaussian(void* r, double mean, double sigma)
{
  std::normal_distribution<float> d(mean, sigma);
  return d(*static_cast<std::mt19937*>(r));
}
float h_rng_uniform01(void* r)
{
  std::uniform_real_distribution<float> d(0, 1);
  return d(*static_cast<std::mt19937*>(r));
}
// predict()'s draw order for n particles: x, y, z, a per particle with sigmas |delta*mod| (ParticleFilter.cpp:101-117)
void h_rng_predict_noise(void* r, uint64_t n, const double* mods4, const double* deltas4, float* noise_n4)
{
  double dev[4];
  for (int k = 0; k < 4; ++k)
    dev[k] = std::fabs(deltas4[k] * mods4[k]);
  for (uint64_t i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k)
      noise_n4[4 * i + k] = h_rng_gaussian(r, 0, dev[k]);
}

// ------------------------------------------------------------------------------------------ timing (CPU baseline arm)
// Runs `reps` update() calls and returns the best wall time of one call in seconds.
double h_time_update(void* p, void* g, const float* ranges4, uint32_t n_ranges, double alpha, double sigma, double roll,
                     double pitch, int reps)
{
  double best = 1e300;
  for (int i = 0; i < reps; ++i)
  {
    const auto t0 = std::chrono::steady_clock::now();
    h_pf_update(p, g, ranges4, n_ranges, alpha, sigma, roll, pitch);
    const auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  return best;
}

// Node.cpp:131-137 verbatim (VoxelGrid in front of the update), B200 build only: the reference build has no PCL.
// Returns the number of output points (written to out_xyzw, capacity n), -1 on failure, -2 in the reference build.
int64_t h_node_voxel_filter(const float* xyzw, uint64_t n, double voxel_size, float* out_xyzw)
{
#ifdef AMCL3D_HARNESS_REFERENCE
  (void)xyzw;
  (void)n;
  (void)voxel_size;
  (void)out_xyzw;
  return -2;
#else
  try
  {
    Cloud::Ptr cloud_src = makeCloud(xyzw, n);
    Cloud::Ptr cloud_down(new Cloud());
    pcl::VoxelGrid<pcl::PointXYZ> sor;
    sor.setInputCloud(cloud_src);
    sor.setLeafSize(voxel_size, voxel_size, voxel_size);
    sor.filter(*cloud_down);
    std::memcpy(out_xyzw, cloud_down->points.data(), cloud_down->points.size() * 16);
    return static_cast<int64_t>(cloud_down->points.size());
  }
  catch (const std::exception& e)
  {
    std::fprintf(stderr, "h_node_voxel_filter: %s\n", e.what());
    return -1;
  }
#endif
}

}  // extern "C"
