// compat stand-in: the slice of pcl::VoxelGrid<PointT> the amcl3d node uses (Node.cpp:131-137:
// setInputCloud / setLeafSize / filter), executed on the device through amcl3d_cuda_voxel_grid.
//
// With a real PCL on the include path this header is simply not added and the node keeps pcl::VoxelGrid; with this
// stand-in the SAME node source down-samples on the GPU: same leaf indices and output order as PCL's applyFilter
// (ascending leaf index, x fastest), centroid per leaf, non-finite points dropped, "leaf too small" returns the input.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "CudaRuntime.h"

namespace pcl
{
template <typename PointT>
class VoxelGrid
{
public:
  typedef typename PointCloud<PointT>::Ptr PointCloudPtr;
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;

  VoxelGrid() : leaf_{ 0.f, 0.f, 0.f } {}

  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  void setLeafSize(float lx, float ly, float lz)
  {
    leaf_[0] = lx;
    leaf_[1] = ly;
    leaf_[2] = lz;
  }

  void filter(PointCloud<PointT>& output)
  {
    static_assert(sizeof(PointT) == 16, "pcl::PointXYZ-like 16-byte points expected");
    output.header = input_ ? input_->header : PCLHeader();
    if (!input_ || input_->points.empty())
    {
      output.clear();
      return;
    }
    if (!(leaf_[0] > 0.f && leaf_[1] > 0.f && leaf_[2] > 0.f))
    {
      // PCL: "Leaf size not set" -> empty output
      output.clear();
      return;
    }
    typename PointCloud<PointT>::VectorType out(input_->points.size());
    uint64_t n_out = 0;
    amcl3d::cuda::check(amcl3d_cuda_voxel_grid(amcl3d::cuda::context(),
                                               reinterpret_cast<const float*>(input_->points.data()),
                                               input_->points.size(), leaf_[0], leaf_[1], leaf_[2],
                                               reinterpret_cast<float*>(out.data()), out.size(), &n_out),
                        "pcl::VoxelGrid::filter");
    out.resize(static_cast<std::size_t>(n_out));
    output.points.swap(out);
    output.width = static_cast<uint32_t>(output.points.size());
    output.height = 1;
    output.is_dense = true;
  }

private:
  PointCloudConstPtr input_;
  float leaf_[3];
};
}  // namespace pcl
