// ORACLE-ONLY stand-in for <pcl/kdtree/kdtree_flann.h> (test infrastructure, never shipped).
//
// The reference's computeGrid (amcl3d/src/PointCloudTools.cpp:110-111,133-135) asks
// pcl::KdTreeFLANN for the single exact nearest neighbour of every voxel corner and uses the
// returned SQUARED distance.  PCL/FLANN are not vendored under /root/reference (pinned only as
// `pcl_ros`, amcl3d/package.xml:27; ROS Kinetic = PCL 1.7.2 / FLANN 1.8.4), so this header
// restates the published behaviour of that call: exact (eps = 0) 1-NN with FLANN's
// L2_Simple<float> metric, i.e. squared distance accumulated in float as
// ((0 + dx*dx) + dy*dy) + dz*dz with diff = query - point and no fused multiply-add.
// The minimum of that float quantity over all points is unique, so any exact search returns the
// same distance; only the winning index may differ between equidistant points.
//
// Implementation: median-split kd-tree, leaves of <= 12 points, box pruning in double with a
// relative safety margin (so a point whose float distance could still win is never pruned), and
// the previous answer as the starting bound (queries arrive in raster order).
#pragma once

#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace pcl
{
template <typename PointT>
class KdTreeFLANN
{
public:
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;

  void setInputCloud(const PointCloudConstPtr& cloud)
  {
    cloud_ = cloud;
    nodes_.clear();
    order_.clear();
    last_ = -1;
    if (!cloud_ || cloud_->points.empty())
      return;
    const std::size_t n = cloud_->points.size();
    order_.resize(n);
    for (std::size_t i = 0; i < n; ++i)
      order_[i] = static_cast<int>(i);
    nodes_.reserve(2 * n / kLeaf + 16);
    build(0, static_cast<int>(n));
    // leaf-ordered copy of the coordinates: contiguous scans at the leaves
    pts_.resize(n);
    for (std::size_t i = 0; i < n; ++i)
    {
      const PointT& p = cloud_->points[order_[i]];
      pts_[i].v[0] = p.x;
      pts_[i].v[1] = p.y;
      pts_[i].v[2] = p.z;
    }
  }

  int nearestKSearch(const PointT& q, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const
  {
    if (k != 1 || nodes_.empty())
      return 0;
    Query s;
    s.q[0] = q.x;
    s.q[1] = q.y;
    s.q[2] = q.z;
    s.best = std::numeric_limits<float>::max();
    s.best_i = -1;
    if (last_ >= 0)
    {
      s.best = dist2(s.q, pts_[last_].v);
      s.best_i = last_;
    }
    search(0, s);
    last_ = s.best_i;
    k_indices.resize(1);
    k_sqr_distances.resize(1);
    k_indices[0] = order_[s.best_i];
    k_sqr_distances[0] = s.best;
    return 1;
  }

private:
  enum
  {
    kLeaf = 12
  };
  struct P3
  {
    float v[3];
  };
  struct Node
  {
    int lo, hi;       // point range [lo, hi) in leaf order
    int left, right;  // children (-1 for a leaf)
    float bmin[3], bmax[3];
  };
  struct Query
  {
    float q[3];
    float best;
    int best_i;
  };

  // FLANN L2_Simple<float>: result starts at 0 and accumulates diff*diff per dimension, all in float.
  static inline float dist2(const float* q, const float* p)
  {
    float r = 0.f;
    for (int a = 0; a < 3; ++a)
    {
      const float d = q[a] - p[a];
      const float sq = d * d;
      r = r + sq;
    }
    return r;
  }

  int build(int lo, int hi)
  {
    Node nd;
    nd.lo = lo;
    nd.hi = hi;
    nd.left = nd.right = -1;
    for (int a = 0; a < 3; ++a)
    {
      nd.bmin[a] = std::numeric_limits<float>::max();
      nd.bmax[a] = -std::numeric_limits<float>::max();
    }
    for (int i = lo; i < hi; ++i)
    {
      const PointT& p = cloud_->points[order_[i]];
      const float c[3] = { p.x, p.y, p.z };
      for (int a = 0; a < 3; ++a)
      {
        nd.bmin[a] = std::min(nd.bmin[a], c[a]);
        nd.bmax[a] = std::max(nd.bmax[a], c[a]);
      }
    }
    const int self = static_cast<int>(nodes_.size());
    nodes_.push_back(nd);
    if (hi - lo > kLeaf)
    {
      int axis = 0;
      float ext = nd.bmax[0] - nd.bmin[0];
      for (int a = 1; a < 3; ++a)
        if (nd.bmax[a] - nd.bmin[a] > ext)
        {
          ext = nd.bmax[a] - nd.bmin[a];
          axis = a;
        }
      const int mid = lo + (hi - lo) / 2;
      const PointCloud<PointT>& c = *cloud_;
      std::nth_element(order_.begin() + lo, order_.begin() + mid, order_.begin() + hi, [&c, axis](int a, int b) {
        const float va = axis == 0 ? c.points[a].x : (axis == 1 ? c.points[a].y : c.points[a].z);
        const float vb = axis == 0 ? c.points[b].x : (axis == 1 ? c.points[b].y : c.points[b].z);
        return va < vb;
      });
      const int l = build(lo, mid);
      const int r = build(mid, hi);
      nodes_[self].left = l;
      nodes_[self].right = r;
    }
    return self;
  }

  // Exact squared distance from the query to the node's bounding box, in double.
  static inline double boxDist2(const Node& nd, const float* q)
  {
    double s = 0.0;
    for (int a = 0; a < 3; ++a)
    {
      double d = 0.0;
      if (q[a] < nd.bmin[a])
        d = static_cast<double>(nd.bmin[a]) - static_cast<double>(q[a]);
      else if (q[a] > nd.bmax[a])
        d = static_cast<double>(q[a]) - static_cast<double>(nd.bmax[a]);
      s += d * d;
    }
    return s;
  }

  void search(int ni, Query& s) const
  {
    const Node& nd = nodes_[ni];
    // prune only when the box is farther than the current best by a margin that dwarfs float rounding
    if (boxDist2(nd, s.q) > static_cast<double>(s.best) * 1.00001 + 1e-30)
      return;
    if (nd.left < 0)
    {
      for (int i = nd.lo; i < nd.hi; ++i)
      {
        const float d = dist2(s.q, pts_[i].v);
        if (d < s.best)
        {
          s.best = d;
          s.best_i = i;
        }
      }
      return;
    }
    const double dl = boxDist2(nodes_[nd.left], s.q);
    const double dr = boxDist2(nodes_[nd.right], s.q);
    if (dl <= dr)
    {
      search(nd.left, s);
      search(nd.right, s);
    }
    else
    {
      search(nd.right, s);
      search(nd.left, s);
    }
  }

  PointCloudConstPtr cloud_;
  std::vector<Node> nodes_;
  std::vector<int> order_;
  std::vector<P3> pts_;
  mutable int last_{ -1 };
};
}  // namespace pcl
