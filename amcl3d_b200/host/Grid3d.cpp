// Grid3d.cpp -- host side of the likelihood grid (reference: amcl3d/src/Grid3d.cpp).  All arithmetic on grid
// cells happens on the device behind include/amcl3d_cuda.h; this file is orchestration, file I/O and messages.
#include "Grid3d.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>

namespace amcl3d
{
namespace
{
const char* nodeName() { return ros::this_node::getName().data(); }

void boundsOf(const PointCloudInfo& pc, double b[7])
{
  b[0] = pc.octo_min_x;
  b[1] = pc.octo_min_y;
  b[2] = pc.octo_min_z;
  b[3] = pc.octo_max_x;
  b[4] = pc.octo_max_y;
  b[5] = pc.octo_max_z;
  b[6] = pc.octo_resol;
}

// map_path with its ".bt"/".ot" suffix swapped for ".grid" (empty when it has neither)
std::string cachePathFor(const std::string& map_path)
{
  if (map_path.size() < 3)
    return std::string();
  const std::string ext = map_path.substr(map_path.size() - 3);
  if (ext != ".bt" && ext != ".ot")
    return std::string();
  return map_path.substr(0, map_path.find(ext)) + ".grid";
}
}  // namespace

bool Grid3d::deviceReady() const
{
  int has = 0;
  return device_ && amcl3d_cuda_grid_has_cells(device_.get(), &has) == 0 && has;
}

bool Grid3d::open(const std::string& map_path, const double sensor_dev)
{
  PointCloudInfo::Ptr pc;
  try
  {
    boost::shared_ptr<octomap::OcTree> tree = openOcTree(map_path);
    ROS_INFO("[%s] Octomap loaded", nodeName());
    pc = computePointCloud(tree);
    ROS_INFO("[%s] Map size: X %lf..%lf  Y %lf..%lf  Z %lf..%lf  Res %lf", nodeName(), pc->octo_min_x, pc->octo_max_x,
             pc->octo_min_y, pc->octo_max_y, pc->octo_min_z, pc->octo_max_z, pc->octo_resol);
  }
  catch (std::exception& e)
  {
    ROS_ERROR("[%s] %s", nodeName(), e.what());
    return false;
  }
  pc_info_ = pc;  // loadGrid checks the cache against the new map's bounds

  const std::string grid_path = cachePathFor(map_path);
  if (!grid_path.empty() && loadGrid(grid_path, sensor_dev))
    return true;

  ROS_INFO("[%s] Computing 3D occupancy grid on the GPU", nodeName());
  if (!openFromPointCloud(pc, sensor_dev))
  {
    // neither the cache nor the device build produced a grid for the NEW map: do not keep answering isIntoMap with the
    // new bounds while computeCloudWeight still gathers from an older grid -- the object goes back to "not opened"
    device_.reset();
    grid_info_.reset();
    pc_info_.reset();
    return false;
  }
  ROS_INFO("[%s] Computing 3D occupancy grid done!", nodeName());
  if (!grid_path.empty())
    saveGrid(grid_path);
  return true;
}

bool Grid3d::openFromPointCloud(PointCloudInfo::Ptr pc_info, const double sensor_dev)
{
  if (!pc_info)
    return false;
  try
  {
    double b[7];
    boundsOf(*pc_info, b);
    cuda::GridHandle g = cuda::makeGrid(b);
    const uint64_t n = pc_info->cloud ? pc_info->cloud->points.size() : 0;
    const float* pts = n ? reinterpret_cast<const float*>(pc_info->cloud->points.data()) : nullptr;
    cuda::check(amcl3d_cuda_grid_compute(g.get(), pts, n, sensor_dev, /*keep_dist=*/1), "computeGrid");
    cuda::check(amcl3d_cuda_grid_dims(g.get(), dims_), "grid_dims");
    device_ = g;
    pc_info_ = pc_info;
    grid_info_.reset();
    sensor_dev_ = sensor_dev;
    return true;
  }
  catch (const std::exception& e)
  {
    ROS_ERROR("[%s] %s", nodeName(), e.what());
    return false;
  }
}

bool Grid3d::setGrid(PointCloudInfo::Ptr pc_info, Grid3dInfo::Ptr grid_info)
{
  if (!pc_info || !grid_info)
    return false;
  try
  {
    double b[7];
    boundsOf(*pc_info, b);
    cuda::GridHandle g = cuda::makeGrid(b);
    uint32_t d[3];
    cuda::check(amcl3d_cuda_grid_dims(g.get(), d), "grid_dims");
    if (d[0] != grid_info->size_x || d[1] != grid_info->size_y || d[2] != grid_info->size_z ||
        grid_info->grid.size() != static_cast<std::size_t>(d[0]) * d[1] * d[2])
    {
      ROS_WARN("[%s] Grid dimensions do not match the map bounds", nodeName());
      return false;
    }
    cuda::check(amcl3d_cuda_grid_upload_cells(g.get(), reinterpret_cast<const float*>(grid_info->grid.data()),
                                              grid_info->sensor_dev),
                "upload_cells");
    std::memcpy(dims_, d, sizeof(d));
    device_ = g;
    pc_info_ = pc_info;
    grid_info_ = grid_info;
    sensor_dev_ = grid_info->sensor_dev;
    return true;
  }
  catch (const std::exception& e)
  {
    ROS_ERROR("[%s] %s", nodeName(), e.what());
    return false;
  }
}

Grid3dInfo::ConstPtr Grid3d::gridInfo() const
{
  if (grid_info_ || !deviceReady())
    return grid_info_;
  Grid3dInfo::Ptr gi(new Grid3dInfo());
  gi->sensor_dev = sensor_dev_;
  gi->size_x = dims_[0];
  gi->size_y = dims_[1];
  gi->size_z = dims_[2];
  gi->step_y = dims_[0];
  gi->step_z = dims_[0] * dims_[1];
  gi->grid.resize(static_cast<std::size_t>(dims_[0]) * dims_[1] * dims_[2]);
  cuda::check(amcl3d_cuda_grid_download_cells(device_.get(), reinterpret_cast<float*>(gi->grid.data())), "download_cells");
  grid_info_ = gi;
  return grid_info_;
}

bool Grid3d::buildGridSliceMsg(const double z, nav_msgs::OccupancyGrid& msg) const
{
  if (!deviceReady() || !pc_info_)
    return false;
  if (z < pc_info_->octo_min_z || z > pc_info_->octo_max_z)
    return false;

  msg.info.map_load_time = ros::Time::now();
  msg.info.resolution = pc_info_->octo_resol;
  msg.info.width = dims_[0];
  msg.info.height = dims_[1];
  msg.info.origin.position.x = 0.;
  msg.info.origin.position.y = 0.;
  msg.info.origin.position.z = z;
  msg.info.origin.orientation.x = 0.;
  msg.info.origin.orientation.y = 0.;
  msg.info.origin.orientation.z = 0.;
  msg.info.origin.orientation.w = 1.;

  // The payload spans [first, last) in LINEAR index space, both ends computed from float-narrowed corners
  // exactly as the reference does (Grid3d.cpp:101-102): its length equals width*height only through rounding.
  const uint32_t first = point2grid(pc_info_->octo_min_x, pc_info_->octo_min_y, z);
  const uint32_t last = point2grid(pc_info_->octo_max_x, pc_info_->octo_max_y, z);
  const uint32_t count = last > first ? last - first : 0;
  std::vector<float> prob(count);
  if (count)
    cuda::check(amcl3d_cuda_grid_download_prob_range(device_.get(), first, count, prob.data()), "download_prob_range");

  float peak = -1.0f;
  for (uint32_t i = 0; i < count; ++i)
    if (prob[i] > peak)
      peak = prob[i];
  if (peak < 0.000001f)
    peak = 0.000001f;
  const float scale = 100.f / peak;
  msg.data.resize(count);
  for (uint32_t i = 0; i < count; ++i)
    msg.data[i] = static_cast<int8_t>(prob[i] * scale);
  return true;
}

bool Grid3d::buildMapPointCloudMsg(sensor_msgs::PointCloud2& msg) const
{
  if (!pc_info_ || !pc_info_->cloud)
    return false;
  pcl::toROSMsg(*pc_info_->cloud, msg);
  return true;
}

float Grid3d::computeCloudWeight(const pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud, const float tx, const float ty,
                                 const float tz, const float roll, const float pitch, const float yaw) const
{
  if (!deviceReady() || !pc_info_ || !cloud)
    return 0;
  float weight = 0.f;
  const uint64_t n = cloud->points.size();
  const float* pts = n ? reinterpret_cast<const float*>(cloud->points.data()) : nullptr;
  cuda::check(amcl3d_cuda_cloud_weight(device_.get(), pts, n, tx, ty, tz, roll, pitch, yaw, &weight, nullptr, nullptr),
              "computeCloudWeight");
  return weight;
}

bool Grid3d::isIntoMap(const float x, const float y, const float z) const
{
  if (!pc_info_)
    return false;
  const PointCloudInfo& m = *pc_info_;
  return x >= m.octo_min_x && x < m.octo_max_x && y >= m.octo_min_y && y < m.octo_max_y && z >= m.octo_min_z &&
         z < m.octo_max_z;
}

// ".grid" layout: uint32 size_x, size_y, size_z; double sensor_dev; then size_x*size_y*size_z (dist, prob) float pairs.
bool Grid3d::saveGrid(const std::string& grid_path)
{
  if (!deviceReady())
    return false;
  Grid3dInfo::ConstPtr gi = gridInfo();
  FILE* f = std::fopen(grid_path.c_str(), "wb");
  if (!f)
  {
    ROS_ERROR("[%s] Error opening file %s for writing", nodeName(), grid_path.c_str());
    return false;
  }
  bool ok = std::fwrite(&gi->size_x, sizeof(uint32_t), 1, f) == 1 && std::fwrite(&gi->size_y, sizeof(uint32_t), 1, f) == 1 &&
            std::fwrite(&gi->size_z, sizeof(uint32_t), 1, f) == 1 && std::fwrite(&gi->sensor_dev, sizeof(double), 1, f) == 1;
  ok = ok && std::fwrite(gi->grid.data(), sizeof(Grid3dCell), gi->grid.size(), f) == gi->grid.size();
  std::fclose(f);
  if (ok)
    ROS_INFO("[%s] Grid map successfully saved on %s", nodeName(), grid_path.c_str());
  return ok;
}

bool Grid3d::loadGrid(const std::string& grid_path, const double sensor_dev)
{
  if (!pc_info_)
    return false;
  FILE* f = std::fopen(grid_path.c_str(), "rb");
  if (!f)
  {
    ROS_WARN("[%s] Error opening file %s for reading", nodeName(), grid_path.c_str());
    return false;
  }
  Grid3dInfo::Ptr gi(new Grid3dInfo());
  bool ok = std::fread(&gi->size_x, sizeof(uint32_t), 1, f) == 1 && std::fread(&gi->size_y, sizeof(uint32_t), 1, f) == 1 &&
            std::fread(&gi->size_z, sizeof(uint32_t), 1, f) == 1 && std::fread(&gi->sensor_dev, sizeof(double), 1, f) == 1;
  if (ok && std::fabs(gi->sensor_dev - sensor_dev) >= std::numeric_limits<double>::epsilon())
  {
    ROS_WARN("[%s] Loaded sensorDev is different", nodeName());
    ok = false;
  }
  if (ok)
  {
    gi->step_y = gi->size_x;
    gi->step_z = gi->size_x * gi->size_y;
    const std::size_t cells = static_cast<std::size_t>(gi->size_x) * gi->size_y * gi->size_z;
    gi->grid.resize(cells);
    ok = std::fread(gi->grid.data(), sizeof(Grid3dCell), cells, f) == cells;  // a truncated cache is rejected
  }
  std::fclose(f);
  if (!ok || !setGrid(pc_info_, gi))
    return false;
  ROS_INFO("[%s] Grid map successfully loaded from %s", nodeName(), grid_path.c_str());
  return true;
}

inline uint32_t Grid3d::point2grid(const float x, const float y, const float z) const
{
  const PointCloudInfo& m = *pc_info_;
  const uint32_t ix = static_cast<uint32_t>((x - m.octo_min_x) / m.octo_resol);
  const uint32_t iy = static_cast<uint32_t>((y - m.octo_min_y) / m.octo_resol);
  const uint32_t iz = static_cast<uint32_t>((z - m.octo_min_z) / m.octo_resol);
  return ix + iy * dims_[0] + iz * (dims_[0] * dims_[1]);
}

}  // namespace amcl3d
