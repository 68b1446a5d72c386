// PointCloudTools.h -- B200 host-side counterpart of the reference's PointCloudTools.h
// (reference: amcl3d/src/PointCloudTools.h:28-99).  Same public PODs and free functions; computeGrid runs on
// the GPU through the C-ABI (include/amcl3d_cuda.h), openOcTree/computePointCloud stay host code.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <octomap/OcTree.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace amcl3d
{
// One voxel of the likelihood grid (reference PointCloudTools.h:28-33): squared distance to the nearest
// occupied map point and the sensor-model probability derived from it.
class Grid3dCell
{
public:
  float dist{ -1 };
  float prob{ 0 };
};

// Host-side mirror of the likelihood grid (reference PointCloudTools.h:37-50).  The device copy used by the hot
// path lives in an amcl3d_cuda_grid; this object is what the public API hands out / accepts.
class Grid3dInfo
{
public:
  typedef boost::shared_ptr<Grid3dInfo> Ptr;
  typedef boost::shared_ptr<const Grid3dInfo> ConstPtr;

  std::vector<Grid3dCell> grid;  // x-fastest: index = ix + iy*step_y + iz*step_z
  double sensor_dev{ 0 };
  uint32_t size_x{ 0 };
  uint32_t size_y{ 0 };
  uint32_t size_z{ 0 };
  uint32_t step_y{ 0 };
  uint32_t step_z{ 0 };
};

// Occupied-leaf centres of the octomap plus its metric bounds and resolution (reference PointCloudTools.h:54-68).
class PointCloudInfo
{
public:
  typedef boost::shared_ptr<PointCloudInfo> Ptr;
  typedef boost::shared_ptr<const PointCloudInfo> ConstPtr;

  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud;
  double octo_min_x{ 0 };
  double octo_min_y{ 0 };
  double octo_min_z{ 0 };
  double octo_max_x{ 0 };
  double octo_max_y{ 0 };
  double octo_max_z{ 0 };
  double octo_resol{ 0 };
};

// Loads a ".bt" (binary) or ".ot" (full) octomap.  Throws std::runtime_error when the file is missing, unreadable
// or of an unknown kind (reference PointCloudTools.cpp:26-49).
boost::shared_ptr<octomap::OcTree> openOcTree(const std::string& file_path);

// Occupied leaf centres + bounds + resolution.  Throws on a null or empty tree (reference PointCloudTools.cpp:51-82).
PointCloudInfo::Ptr computePointCloud(boost::shared_ptr<octomap::OcTree> octo_tree);

// Builds the likelihood grid on the GPU and returns its host mirror (reference PointCloudTools.cpp:84-149).
// Throws std::runtime_error on a null input, on a CUDA failure, or -- when the environment variable
// AMCL3D_MAX_CELLS is set -- when the grid exceeds that many cells (the reference hard-codes 250 000 000).
Grid3dInfo::Ptr computeGrid(PointCloudInfo::Ptr pc_info, const double sensor_dev);

}  // namespace amcl3d
