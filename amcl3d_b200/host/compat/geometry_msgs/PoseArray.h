#pragma once
#include <vector>
#include <geometry_msgs/Point.h>
#include <std_msgs/Header.h>
namespace geometry_msgs
{
struct PoseArray
{
  std_msgs::Header header;
  std::vector<Pose> poses;
};
}  // namespace geometry_msgs
