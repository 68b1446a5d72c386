// compat stand-in for <ros/ros.h>: printf-style logging macros, node name, ros::Time.
// Logging goes to stderr and is silent unless AMCL3D_LOG=1 (errors always print).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace ros
{
struct Time
{
  uint32_t sec{ 0 };
  uint32_t nsec{ 0 };
  Time() = default;
  Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
  static void init() {}
  static Time now()
  {
    const auto t = std::chrono::system_clock::now().time_since_epoch();
    const auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t).count();
    return Time(static_cast<uint32_t>(ns / 1000000000LL), static_cast<uint32_t>(ns % 1000000000LL));
  }
  double toSec() const { return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec); }
};
namespace this_node
{
inline const std::string& getName()
{
  static const std::string name("/amcl3d");
  return name;
}
}  // namespace this_node
namespace compat_detail
{
inline bool verbose()
{
  static const bool v = [] {
    const char* e = std::getenv("AMCL3D_LOG");
    return e && e[0] && e[0] != '0';
  }();
  return v;
}
}  // namespace compat_detail
}  // namespace ros

#define AMCL3D_COMPAT_LOG(level, always, ...)                                                                          \
  do                                                                                                                   \
  {                                                                                                                    \
    if ((always) || ::ros::compat_detail::verbose())                                                                   \
    {                                                                                                                  \
      std::fprintf(stderr, "[" level "] ");                                                                            \
      std::fprintf(stderr, __VA_ARGS__);                                                                               \
      std::fprintf(stderr, "\n");                                                                                      \
    }                                                                                                                  \
  } while (0)
#define ROS_DEBUG(...)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
  } while (0)
#define ROS_INFO(...) AMCL3D_COMPAT_LOG("INFO", false, __VA_ARGS__)
#define ROS_WARN(...) AMCL3D_COMPAT_LOG("WARN", false, __VA_ARGS__)
#define ROS_ERROR(...) AMCL3D_COMPAT_LOG("ERROR", ::ros::compat_detail::verbose(), __VA_ARGS__)
