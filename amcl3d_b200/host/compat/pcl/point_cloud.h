// compat stand-in: the slice of pcl::PointCloud<T> used by the amcl3d library TUs and tests.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <boost/shared_ptr.hpp>
namespace pcl
{
struct PCLHeader
{
  uint32_t seq{ 0 };
  uint64_t stamp{ 0 };  // microseconds, as in PCL
  std::string frame_id;
};

template <typename PointT>
class PointCloud
{
public:
  typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
  typedef std::vector<PointT> VectorType;
  typedef typename VectorType::iterator iterator;
  typedef typename VectorType::const_iterator const_iterator;
  typedef PointT PointType;

  PCLHeader header;
  VectorType points;
  uint32_t width{ 0 };
  uint32_t height{ 0 };
  bool is_dense{ true };

  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void reserve(std::size_t n) { points.reserve(n); }
  void resize(std::size_t n)
  {
    points.resize(n);
    width = static_cast<uint32_t>(n);
    height = 1;
  }
  void clear()
  {
    points.clear();
    width = height = 0;
  }
  void push_back(const PointT& p)
  {
    points.push_back(p);
    width = static_cast<uint32_t>(points.size());
    height = 1;
  }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};
}  // namespace pcl
