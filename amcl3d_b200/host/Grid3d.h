// Grid3d.h -- B200 host-side counterpart of the reference's Grid3d (reference: amcl3d/src/Grid3d.h:34-159).
//
// Public interface = the reference's, unchanged, so Node / tests / ParticleFilter compile against it as they do
// against the original.  The likelihood grid lives in HBM (amcl3d_cuda_grid); a host mirror (Grid3dInfo) is
// materialised only when somebody asks for it.  Additive members (not in the reference) are grouped at the end
// of the public section.
#pragma once

#include <geometry_msgs/PoseArray.h>
#include <nav_msgs/OccupancyGrid.h>
#include <octomap/OcTree.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf/tf.h>

#include "CudaRuntime.h"
#include "PointCloudTools.h"

namespace amcl3d
{
class Grid3d
{
public:
  explicit Grid3d() {}
  virtual ~Grid3d() {}

  // --- reference API --------------------------------------------------------------------------------------------
  // Loads the octomap, then its ".grid" cache if present and computed for the same sensor_dev, else builds the
  // likelihood grid on the GPU and writes the cache.  false on any failure (reference Grid3d.cpp:25-78).
  bool open(const std::string& map_path, const double sensor_dev);

  // Horizontal slice of the probability field at height z, scaled to [0, 100] (reference Grid3d.cpp:80-121).
  bool buildGridSliceMsg(const double z, nav_msgs::OccupancyGrid& msg) const;

  // The map's occupied points as a PointCloud2 (reference Grid3d.cpp:123-131).
  bool buildMapPointCloudMsg(sensor_msgs::PointCloud2& msg) const;

  // Mean grid probability of the cloud transformed by the pose; 0 when fewer than 11 points hit the map or the
  // grid is not open (reference Grid3d.cpp:133-199).  Evaluated on the GPU.
  float computeCloudWeight(const pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud, const float tx, const float ty,
                           const float tz, const float roll, const float pitch, const float yaw) const;

  // min <= v < max per axis against the octomap bounds (reference Grid3d.cpp:201-208).
  bool isIntoMap(const float x, const float y, const float z) const;

  // --- additive (B200 build only) -------------------------------------------------------------------------------
  // open() without the octomap file: takes what computePointCloud would have produced.
  bool openFromPointCloud(PointCloudInfo::Ptr pc_info, const double sensor_dev);
  // Installs an externally computed grid (e.g. read from a ".grid" cache) together with its map info.
  bool setGrid(PointCloudInfo::Ptr pc_info, Grid3dInfo::Ptr grid_info);
  // Host mirror of the grid (downloaded on first use) / the map info; null before open.
  Grid3dInfo::ConstPtr gridInfo() const;
  PointCloudInfo::ConstPtr pointCloudInfo() const { return pc_info_; }
  // Device handle for ParticleFilter::update; null before open.
  const amcl3d_cuda_grid* deviceGrid() const { return device_.get(); }
  // ".grid" cache I/O (byte-compatible with the reference's private saveGrid/loadGrid, Grid3d.cpp:210-275).
  bool saveGrid(const std::string& grid_path);
  bool loadGrid(const std::string& grid_path, const double sensor_dev);

private:
  inline uint32_t point2grid(const float x, const float y, const float z) const;
  bool deviceReady() const;

  PointCloudInfo::Ptr pc_info_;
  mutable Grid3dInfo::Ptr grid_info_;  // lazy host mirror
  cuda::GridHandle device_;
  double sensor_dev_{ 0 };
  uint32_t dims_[3]{ 0, 0, 0 };
};

}  // namespace amcl3d
