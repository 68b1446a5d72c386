// ParticleFilter.cpp -- host side of the particle filter (reference: amcl3d/src/ParticleFilter.cpp).
// Every loop over particles runs on the device behind include/amcl3d_cuda.h.
#include "ParticleFilter.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace amcl3d
{
static_assert(sizeof(Particle) == 7 * sizeof(float), "Particle is handed to the device as 7 packed floats");
static_assert(sizeof(Range) == 4 * sizeof(float), "Range is handed to the device as 4 packed floats");

ParticleFilter::ParticleFilter() : generator_(rd_())
{
  device_ = cuda::makeFilter();  // throws when no CUDA device is usable: there is no CPU path
  philox_seed_ = (static_cast<uint64_t>(rd_()) << 32) | rd_();
}

ParticleFilter::~ParticleFilter() {}

void ParticleFilter::seed(uint32_t s)
{
  generator_.seed(s);
  philox_seed_ = 0x9E3779B97F4A7C15ull ^ s;
  step_ = 0;
}

std::size_t ParticleFilter::size() const
{
  uint64_t n = 0;
  cuda::check(amcl3d_cuda_pf_size(device_.get(), &n), "pf_size");
  return static_cast<std::size_t>(n);
}

void ParticleFilter::setParticles(const std::vector<Particle>& particles)
{
  cuda::check(amcl3d_cuda_pf_upload_particles(device_.get(), reinterpret_cast<const float*>(particles.data()),
                                              particles.size()),
              "pf_upload_particles");
  initialized_ = true;
}

std::vector<Particle> ParticleFilter::getParticles() const
{
  std::vector<Particle> out(size());
  if (!out.empty())
    cuda::check(amcl3d_cuda_pf_download_particles(device_.get(), reinterpret_cast<float*>(out.data())),
                "pf_download_particles");
  return out;
}

bool ParticleFilter::useHostRng(std::size_t n) const
{
  return rng_mode_ == HostMt19937 || (rng_mode_ == Auto && n <= 65536);
}

void ParticleFilter::buildParticlesPoseMsg(geometry_msgs::PoseArray& msg) const
{
  const std::vector<Particle> p = getParticles();
  msg.poses.resize(p.size());
  for (std::size_t i = 0; i < p.size(); ++i)
  {
    msg.poses[i].position.x = static_cast<double>(p[i].x);
    msg.poses[i].position.y = static_cast<double>(p[i].y);
    msg.poses[i].position.z = static_cast<double>(p[i].z);
    // yaw-only quaternion; the half angle is formed in float first, as the reference does
    const float half = p[i].a * 0.5f;
    msg.poses[i].orientation.x = 0.;
    msg.poses[i].orientation.y = 0.;
    msg.poses[i].orientation.z = std::sin(static_cast<double>(half));
    msg.poses[i].orientation.w = std::cos(static_cast<double>(half));
  }
}

void ParticleFilter::init(const int num_particles, const float x_init, const float y_init, const float z_init,
                          const float a_init, const float x_dev, const float y_dev, const float z_dev, const float a_dev)
{
  const std::size_t n = static_cast<std::size_t>(std::abs(num_particles));
  if (n == 0)
    return;  // the reference indexes p_[0] of an empty vector here (undefined behaviour)
  const float pose[4] = { x_init, y_init, z_init, a_init };
  const float devs[4] = { x_dev, y_dev, z_dev, a_dev };
  float mean[4] = { 0, 0, 0, 0 };
  if (useHostRng(n))
  {
    // the reference's draw order: particles 1..n-1, x y z a each (ParticleFilter.cpp:69-72)
    std::vector<float> noise(4 * n, 0.f);
    for (std::size_t i = 1; i < n; ++i)
      for (int k = 0; k < 4; ++k)
        noise[4 * i + k] = ranGaussian(0, devs[k]);
    cuda::check(amcl3d_cuda_pf_init(device_.get(), n, pose, devs, noise.data(), 0, mean), "pf_init");
  }
  else
    cuda::check(amcl3d_cuda_pf_init(device_.get(), n, pose, devs, nullptr, philox_seed_, mean), "pf_init");
  Particle m;
  m.x = mean[0];
  m.y = mean[1];
  m.z = mean[2];
  m.a = mean[3];
  mean_ = m;
  initialized_ = true;
}

void ParticleFilter::predict(const double odom_x_mod, const double odom_y_mod, const double odom_z_mod,
                             const double odom_a_mod, const double delta_x, const double delta_y, const double delta_z,
                             const double delta_a)
{
  const double mods[4] = { odom_x_mod, odom_y_mod, odom_z_mod, odom_a_mod };
  const double deltas[4] = { delta_x, delta_y, delta_z, delta_a };
  const std::size_t n = size();
  ++step_;
  if (n == 0)
    return;
  if (useHostRng(n))
  {
    double dev[4];
    for (int k = 0; k < 4; ++k)
      dev[k] = std::fabs(deltas[k] * mods[k]);
    // x, y, z, a per particle, in particle order (ParticleFilter.cpp:112-117)
    std::vector<float> noise(4 * n);
    for (std::size_t i = 0; i < n; ++i)
      for (int k = 0; k < 4; ++k)
        noise[4 * i + k] = ranGaussian(0, dev[k]);
    cuda::check(amcl3d_cuda_pf_predict(device_.get(), mods, deltas, noise.data(), 0, step_), "pf_predict");
  }
  else
    cuda::check(amcl3d_cuda_pf_predict(device_.get(), mods, deltas, nullptr, philox_seed_, step_), "pf_predict");
}

void ParticleFilter::update(const Grid3d& grid3d, const pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud,
                            const std::vector<Range>& range_data, const double alpha, const double sigma,
                            const double roll, const double pitch)
{
  const amcl3d_cuda_grid* grid = grid3d.deviceGrid();
  int open = 0;
  if (grid)
    cuda::check(amcl3d_cuda_grid_has_cells(grid, &open), "grid_has_cells");
  if (!grid || !open)
  {
    // an unopened Grid3d reports every particle as outside the map: all weights 0, mean 0
    // (reference ParticleFilter.cpp:137-142,172-176,185-188)
    std::vector<Particle> p = getParticles();
    for (std::size_t i = 0; i < p.size(); ++i)
    {
      p[i].w = 0;
      p[i].wp = 0;
      p[i].wr = 0;
    }
    if (!p.empty())
      setParticles(p);
    mean_ = Particle();
    return;
  }
  const uint64_t n_cloud = cloud ? cloud->points.size() : 0;
  const float* pts = n_cloud ? reinterpret_cast<const float*>(cloud->points.data()) : nullptr;
  const float* ranges = range_data.empty() ? nullptr : reinterpret_cast<const float*>(range_data.data());
  float mean[4] = { 0, 0, 0, 0 };
  cuda::check(amcl3d_cuda_pf_update(device_.get(), grid, pts, n_cloud, ranges, static_cast<uint32_t>(range_data.size()),
                                    alpha, sigma, roll, pitch, mean),
              "pf_update");
  Particle m;
  m.x = mean[0];
  m.y = mean[1];
  m.z = mean[2];
  m.a = mean[3];
  mean_ = m;
}

void ParticleFilter::resample()
{
  if (size() == 0)
    return;
  const float u = rngUniform(0, 1);  // the single draw of ParticleFilter.cpp:202
  cuda::check(amcl3d_cuda_pf_resample(device_.get(), u, nullptr), "pf_resample");
}

float ParticleFilter::ranGaussian(const double mean, const double sigma)
{
  std::normal_distribution<float> d(mean, sigma);
  return d(generator_);
}

float ParticleFilter::rngUniform(const float range_from, const float range_to)
{
  std::uniform_real_distribution<float> d(range_from, range_to);
  return d(generator_);
}

}  // namespace amcl3d
