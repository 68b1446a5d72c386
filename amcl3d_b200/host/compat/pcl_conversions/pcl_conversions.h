// compat stand-in: pcl::toROSMsg / pcl::fromROSMsg for PointCloud<PointXYZ> (Grid3d.cpp:128).
#pragma once
#include <cstring>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl
{
inline void toROSMsg(const PointCloud<PointXYZ>& cloud, sensor_msgs::PointCloud2& msg)
{
  msg.header.seq = cloud.header.seq;
  msg.header.frame_id = cloud.header.frame_id;
  msg.header.stamp = ros::Time(static_cast<uint32_t>(cloud.header.stamp / 1000000ULL),
                               static_cast<uint32_t>((cloud.header.stamp % 1000000ULL) * 1000ULL));
  if (cloud.width == 0 && cloud.height == 0)
  {
    msg.width = static_cast<uint32_t>(cloud.points.size());
    msg.height = 1;
  }
  else
  {
    msg.width = cloud.width;
    msg.height = cloud.height;
  }
  static const char* names[3] = { "x", "y", "z" };
  msg.fields.resize(3);
  for (uint32_t f = 0; f < 3; ++f)
  {
    msg.fields[f].name = names[f];
    msg.fields[f].offset = 4 * f;
    msg.fields[f].datatype = sensor_msgs::PointField::FLOAT32;
    msg.fields[f].count = 1;
  }
  msg.is_bigendian = false;
  msg.point_step = sizeof(PointXYZ);
  msg.row_step = msg.point_step * msg.width;
  msg.is_dense = cloud.is_dense;
  msg.data.resize(cloud.points.size() * sizeof(PointXYZ));
  if (!cloud.points.empty())
    std::memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
}

inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZ>& cloud)
{
  uint32_t off[3] = { 0, 4, 8 };
  for (const auto& f : msg.fields)
  {
    if (f.name == "x") off[0] = f.offset;
    if (f.name == "y") off[1] = f.offset;
    if (f.name == "z") off[2] = f.offset;
  }
  const std::size_t n = static_cast<std::size_t>(msg.width) * msg.height;
  cloud.points.resize(n);
  cloud.width = msg.width;
  cloud.height = msg.height;
  cloud.is_dense = msg.is_dense;
  cloud.header.frame_id = msg.header.frame_id;
  cloud.header.seq = msg.header.seq;
  for (std::size_t i = 0; i < n; ++i)
  {
    const uint8_t* p = msg.data.data() + i * msg.point_step;
    std::memcpy(&cloud.points[i].x, p + off[0], 4);
    std::memcpy(&cloud.points[i].y, p + off[1], 4);
    std::memcpy(&cloud.points[i].z, p + off[2], 4);
  }
}
}  // namespace pcl
