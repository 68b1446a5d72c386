#pragma once
#include <array>
#include <geometry_msgs/Point.h>
#include <std_msgs/Header.h>
namespace geometry_msgs
{
struct PoseWithCovariance
{
  Pose pose;
  std::array<double, 36> covariance{};
};
struct PoseWithCovarianceStamped
{
  std_msgs::Header header;
  PoseWithCovariance pose;
};
}  // namespace geometry_msgs
