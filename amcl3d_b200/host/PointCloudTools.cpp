// PointCloudTools.cpp -- host side of the map preprocessing (reference: amcl3d/src/PointCloudTools.cpp).
#include "PointCloudTools.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include <boost/filesystem.hpp>

#include "CudaRuntime.h"

namespace amcl3d
{
namespace
{
bool endsWith(const std::string& s, const char* suffix)
{
  const std::size_t n = std::strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
}  // namespace

boost::shared_ptr<octomap::OcTree> openOcTree(const std::string& file_path)
{
  if (!boost::filesystem::exists(file_path))
    throw std::runtime_error(std::string("Cannot find file ") + file_path);

  boost::shared_ptr<octomap::OcTree> tree;
  if (endsWith(file_path, ".bt"))
  {
    tree.reset(new octomap::OcTree(0.1));
    if (!tree->readBinary(file_path))
      throw std::runtime_error("OcTree cannot be read");
  }
  else if (endsWith(file_path, ".ot"))
  {
    // a tree of another type (or an unreadable file) leaves the pointer empty
    tree.reset(dynamic_cast<octomap::OcTree*>(octomap::AbstractOcTree::read(file_path)));
  }
  if (!tree)
    throw std::runtime_error(std::string("OcTree cannot be created from file ") + file_path);
  return tree;
}

PointCloudInfo::Ptr computePointCloud(boost::shared_ptr<octomap::OcTree> octo_tree)
{
  if (!octo_tree)
    throw std::runtime_error("OcTree is NULL");
  if (octo_tree->size() <= 1)
    throw std::runtime_error("OcTree is empty");

  PointCloudInfo::Ptr info(new PointCloudInfo());
  octo_tree->getMetricMin(info->octo_min_x, info->octo_min_y, info->octo_min_z);
  octo_tree->getMetricMax(info->octo_max_x, info->octo_max_y, info->octo_max_z);
  info->octo_resol = octo_tree->getResolution();
  info->cloud.reset(new pcl::PointCloud<pcl::PointXYZ>());
  for (octomap::OcTree::leaf_iterator leaf = octo_tree->begin_leafs(), last = octo_tree->end_leafs(); leaf != last; ++leaf)
  {
    if (!octo_tree->isNodeOccupied(*leaf))
      continue;
    // leaf centre narrowed to float, at whatever depth the leaf sits (pruned leaves yield one point)
    info->cloud->push_back(pcl::PointXYZ(static_cast<float>(leaf.getX()), static_cast<float>(leaf.getY()),
                                         static_cast<float>(leaf.getZ())));
  }
  return info;
}

Grid3dInfo::Ptr computeGrid(PointCloudInfo::Ptr pc_info, const double sensor_dev)
{
  if (!pc_info)
    throw std::runtime_error("PointCloudInfo is NULL");

  const double bounds[7] = { pc_info->octo_min_x, pc_info->octo_min_y, pc_info->octo_min_z, pc_info->octo_max_x,
                             pc_info->octo_max_y, pc_info->octo_max_z, pc_info->octo_resol };
  cuda::GridHandle grid = cuda::makeGrid(bounds);  // throws "Octomap size is too big..." past the cap (see cuda::context)

  static_assert(sizeof(pcl::PointXYZ) == 16, "map points are handed to the device as float4");
  const float* pts = (pc_info->cloud && !pc_info->cloud->points.empty()) ?
                         reinterpret_cast<const float*>(pc_info->cloud->points.data()) :
                         nullptr;
  const uint64_t n = pc_info->cloud ? pc_info->cloud->points.size() : 0;
  cuda::check(amcl3d_cuda_grid_compute(grid.get(), pts, n, sensor_dev, /*keep_dist=*/1), "computeGrid");

  Grid3dInfo::Ptr out(new Grid3dInfo());
  uint32_t dims[3];
  cuda::check(amcl3d_cuda_grid_dims(grid.get(), dims), "grid_dims");
  out->sensor_dev = sensor_dev;
  out->size_x = dims[0];
  out->size_y = dims[1];
  out->size_z = dims[2];
  out->step_y = dims[0];
  out->step_z = dims[0] * dims[1];
  static_assert(sizeof(Grid3dCell) == 8, "Grid3dCell must be the (dist, prob) float pair of the .grid format");
  out->grid.resize(static_cast<std::size_t>(dims[0]) * dims[1] * dims[2]);
  cuda::check(amcl3d_cuda_grid_download_cells(grid.get(), reinterpret_cast<float*>(out->grid.data())), "download_cells");
  return out;
}

// ---------------------------------------------------------------------------------------------- shared device context
namespace cuda
{
namespace
{
struct ContextHolder
{
  amcl3d_cuda_ctx* ctx{ nullptr };
  ContextHolder()
  {
    int device = 0;
    if (const char* d = std::getenv("AMCL3D_CUDA_DEVICE"))
      device = std::atoi(d);
    check(amcl3d_cuda_ctx_create(device, nullptr, &ctx), "amcl3d_cuda_ctx_create");
    // Drop-in defaults.  The reference refuses grids over 250 M cells (PointCloudTools.cpp:103-105) and so do these
    // classes; AMCL3D_MAX_CELLS=0 lifts the cap (the device holds far larger grids), any other value replaces it.
    long long cap = 250000000ll;
    if (const char* c = std::getenv("AMCL3D_MAX_CELLS"))
      cap = std::atoll(c);
    check(amcl3d_cuda_ctx_set_option(ctx, "max_cells", cap), "max_cells");
    // AMCL3D_EXACT=1: every per-particle cloud sum is ONE float chain in the caller's cloud order (Grid3d.cpp:191 bit for
    // bit) instead of split / re-ordered chunks whose partials are added in double (a few 1e-7 relative away).  The sums
    // over particles (wtp, wtr, wt, mean, resample) are the reference's sequential float sums in either setting.
    if (const char* e = std::getenv("AMCL3D_EXACT"))
      if (std::atoi(e) != 0)
      {
        check(amcl3d_cuda_ctx_set_option(ctx, "weight_point_splits", 1), "weight_point_splits");
        check(amcl3d_cuda_ctx_set_option(ctx, "cloud_order", 1), "cloud_order");
      }
  }
  ~ContextHolder()
  {
    // handles may outlive static destruction order; leave the context to process teardown
  }
};
}  // namespace

amcl3d_cuda_ctx* context()
{
  static ContextHolder holder;
  return holder.ctx;
}

GridHandle makeGrid(const double bounds7[7])
{
  amcl3d_cuda_grid* g = nullptr;
  check(amcl3d_cuda_grid_create(context(), bounds7, &g), "amcl3d_cuda_grid_create");
  return GridHandle(g, [](amcl3d_cuda_grid* p) { amcl3d_cuda_grid_destroy(p); });
}

FilterHandle makeFilter()
{
  amcl3d_cuda_pf* f = nullptr;
  check(amcl3d_cuda_pf_create(context(), &f), "amcl3d_cuda_pf_create");
  return FilterHandle(f, [](amcl3d_cuda_pf* p) { amcl3d_cuda_pf_destroy(p); });
}
}  // namespace cuda

}  // namespace amcl3d
