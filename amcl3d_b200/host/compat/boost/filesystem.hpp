// compat stand-in: only boost::filesystem::exists(path) is needed (PointCloudTools.cpp:28).
#pragma once
#include <string>
#include <sys/stat.h>
namespace boost
{
namespace filesystem
{
inline bool exists(const std::string& path)
{
  struct stat st;
  return ::stat(path.c_str(), &st) == 0;
}
}  // namespace filesystem
}  // namespace boost
