// compat stand-in: pcl::PointXYZ with PCL's 16-byte layout (x, y, z, padding).
#pragma once
namespace pcl
{
struct alignas(16) PointXYZ
{
  float x{ 0.f };
  float y{ 0.f };
  float z{ 0.f };
  float padding_{ 1.f };  // PCL stores 1.0f in data[3]
  PointXYZ() = default;
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ must be float4-sized");
}  // namespace pcl
