// ParticleFilter.h -- B200 host-side counterpart of the reference's ParticleFilter
// (reference: amcl3d/src/ParticleFilter.h:35-213).
//
// Same public interface.  The particle set lives in HBM as SoA planes (amcl3d_cuda_pf) and never crosses PCIe
// during predict/update/resample; only the sensor cloud goes down and the 16-byte mean comes back per update.
// Additive members (not in the reference) are grouped at the end of the public section.
#pragma once

#include <cstdint>
#include <random>
#include <vector>

#include <ros/ros.h>

#include <geometry_msgs/Point32.h>
#include <geometry_msgs/PoseArray.h>
#include <geometry_msgs/PoseWithCovarianceStamped.h>

#include "Grid3d.h"

namespace amcl3d
{
// One pose hypothesis (reference ParticleFilter.h:35-49): position, yaw, total weight and the two sensor weights.
struct Particle
{
  float x, y, z, a;
  float w, wp, wr;
  Particle() : x(0), y(0), z(0), a(0), w(0), wp(0), wr(0) {}
};

// One radio-range measurement (reference ParticleFilter.h:53-63): range and anchor position.
struct Range
{
  float r, ax, ay, az;
  Range(const float r_, const float ax_, const float ay_, const float az_) : r(r_), ax(ax_), ay(ay_), az(az_) {}
};

class ParticleFilter
{
public:
  explicit ParticleFilter();
  virtual ~ParticleFilter();

  // --- reference API --------------------------------------------------------------------------------------------
  bool isInitialized() const { return initialized_; }
  Particle getMean() const { return mean_; }
  void buildParticlesPoseMsg(geometry_msgs::PoseArray& msg) const;

  void init(const int num_particles, const float x_init, const float y_init, const float z_init, const float a_init,
            const float x_dev, const float y_dev, const float z_dev, const float a_dev);

  void predict(const double odom_x_mod, const double odom_y_mod, const double odom_z_mod, const double odom_a_mod,
               const double delta_x, const double delta_y, const double delta_z, const double delta_a);

  void update(const Grid3d& grid3d, const pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud,
              const std::vector<Range>& range_data, const double alpha, const double sigma, const double roll,
              const double pitch);
  // declared by the reference as well (ParticleFilter.h:157) and defined nowhere; kept for source compatibility
  void update(const std::vector<Range>& range_data, const double alpha, const double sigma);

  void resample();

  // --- additive (B200 build only) -------------------------------------------------------------------------------
  // Where the Gaussian draws of init/predict come from.
  //   HostMt19937 : std::mt19937 + a fresh std::normal_distribution<float> per draw on the host, i.e. the
  //                 reference's own stream (ParticleFilter.cpp:246-250), uploaded as injected noise.
  //   DevicePhilox: Philox4x32-10 on the device, counter = particle index, no host traffic.
  //   Auto        : HostMt19937 up to 65 536 particles, DevicePhilox above.
  enum RngMode
  {
    Auto = 0,
    HostMt19937 = 1,
    DevicePhilox = 2
  };
  void setRngMode(RngMode mode) { rng_mode_ = mode; }
  // Re-seeds the host generator (the reference seeds from std::random_device and offers no seed API).
  void seed(uint32_t s);
  std::size_t size() const;
  void setParticles(const std::vector<Particle>& particles);
  std::vector<Particle> getParticles() const;
  amcl3d_cuda_pf* deviceFilter() const { return device_.get(); }

private:
  float ranGaussian(const double mean, const double sigma);
  float rngUniform(const float range_from, const float range_to);
  bool useHostRng(std::size_t n) const;

  bool initialized_{ false };
  Particle mean_;
  cuda::FilterHandle device_;
  RngMode rng_mode_{ Auto };
  uint64_t philox_seed_{ 0 };
  uint64_t step_{ 0 };

  std::random_device rd_;
  std::mt19937 generator_;
};

}  // namespace amcl3d
