"""GPU parity: amcl3d_cuda_voxel_grid (pcl::VoxelGrid down-sampling of the sensor cloud, Node.cpp:131-137) against the
C restatement of the published PCL algorithm -- same leaf indices, same output order, bit-identical centroids."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check(ctx, port, cloud, leaf):
    got = ctx.voxel_grid(cloud, leaf)
    want = port.voxel_grid(cloud, leaf)
    if want is None:                       # PCL's "leaf size too small" case: input returned unchanged
        want = np.ascontiguousarray(cloud, np.float32)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    return got


@pytest.mark.parametrize("n", [1, 5, 2047, 2048, 2049, 10000, 100000, 300001])
def test_random_clouds(cuda_ctx, port, n):
    rng = np.random.default_rng(n)
    cloud = np.zeros((n, 4), np.float32)
    cloud[:, :3] = rng.normal(0, 6.0, (n, 3))
    cloud[:, 2] *= 0.2
    out = check(cuda_ctx, port, cloud, 0.1)
    assert 0 < len(out) <= n


def test_lidar_like_cloud_feeds_the_update(cuda_ctx, port, cfg1):
    """A dense raw scan (many points per leaf) down-sampled at the node's voxel size: the output is the cloud the
    weighting kernel walks, in the order PCL would hand it over."""
    from amcl3d_b200 import synth
    raw = synth.sensor_cloud(cfg1["map_points"], cfg1["pose"], 60000, 8.0, seed=77)
    rng = np.random.default_rng(78)
    raw = np.repeat(raw, 3, axis=0)
    raw[:, :3] += rng.normal(0, 0.01, (len(raw), 3)).astype(np.float32)
    out = check(cuda_ctx, port, raw, 0.1)
    assert len(out) < len(raw) // 3


def test_anisotropic_leaf_single_cell_and_non_finite(cuda_ctx, port):
    rng = np.random.default_rng(5)
    cloud = np.zeros((5000, 4), np.float32)
    cloud[:, :3] = rng.uniform(-1, 1, (5000, 3))
    check(cuda_ctx, port, cloud, (0.1, 0.25, 0.5))
    out = check(cuda_ctx, port, cloud, 100.0)           # everything in one leaf: one long sequential float sum
    assert len(out) in (1, 2, 4, 8)
    cloud[17, 0] = np.nan
    cloud[99, 2] = np.inf
    cloud[4000, 1] = -np.inf
    check(cuda_ctx, port, cloud, 0.2)
    assert len(cuda_ctx.voxel_grid(np.full((10, 4), np.nan, np.float32), 0.1)) == 0
    assert len(cuda_ctx.voxel_grid(np.zeros((0, 4), np.float32), 0.1)) == 0


def test_leaf_too_small_returns_the_input(cuda_ctx, port):
    cloud = np.zeros((4, 4), np.float32)
    cloud[:, :3] = np.array([[0, 0, 0], [1000, 1000, 1000], [2000, 2000, 2000], [3000, 3000, 3000]], np.float32)
    check(cuda_ctx, port, cloud, 0.001)
