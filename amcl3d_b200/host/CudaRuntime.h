// CudaRuntime.h -- process-wide device context shared by the host-side Grid3d / ParticleFilter classes.
//
// The reference is single-threaded and owns no device; the B200 host classes all talk to ONE
// amcl3d_cuda_ctx (device AMCL3D_CUDA_DEVICE, default 0) so that a Grid3d and the ParticleFilter that borrows
// it (ParticleFilter::update takes `const Grid3d&`) live on the same device and stream.
// There is no CPU fallback: if no context can be created every user of this header throws std::runtime_error.
#pragma once

#include <memory>
#include <stdexcept>
#include <string>

#include "amcl3d_cuda.h"

namespace amcl3d
{
namespace cuda
{
// Throws std::runtime_error carrying amcl3d_cuda_last_error() when rc != 0.
inline void check(int rc, const char* what)
{
  if (rc != 0)
    throw std::runtime_error(std::string(what) + ": " + amcl3d_cuda_last_error());
}

// The shared context (created on first use, destroyed at process exit).
amcl3d_cuda_ctx* context();

typedef std::shared_ptr<amcl3d_cuda_grid> GridHandle;
typedef std::shared_ptr<amcl3d_cuda_pf> FilterHandle;

// bounds7 = min xyz, max xyz, resolution
GridHandle makeGrid(const double bounds7[7]);
FilterHandle makeFilter();
}  // namespace cuda
}  // namespace amcl3d
