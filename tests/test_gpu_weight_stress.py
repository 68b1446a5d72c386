"""GPU parity, hostile inputs for the BATCHED weighting kernel (weight_v4_kernel: fused-scale estimate + exact
verification, and the older generations): voxel faces hit exactly, a last voxel that sticks out of the metric
bounds, large coordinates, poses on the map border.  The grid cells are random numbers -- the kernel only gathers
them -- so every sum is sensitive to a single wrong voxel.  Oracle: the C restatement of Grid3d.cpp:133-199."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def random_grid(ctx, bounds, seed):
    import amcl3d_b200
    g = amcl3d_b200.Grid(ctx, bounds)
    dims = [int(d) for d in g.dims]
    n = dims[0] * dims[1] * dims[2]
    rng = np.random.default_rng(seed)
    cells = np.empty((n, 2), np.float32)
    cells[:, 0] = rng.random(n, dtype=np.float32)
    cells[:, 1] = rng.random(n, dtype=np.float32) + np.float32(0.01)
    g.upload_cells(cells, 0.05)
    return g, cells, dims


def check_batch(ctx, port, g, cells, dims, bounds, cloud, poses, roll, pitch, variants=(0, 4, 5), stride=1):
    cloud = np.ascontiguousarray(cloud, np.float32)
    poses = np.ascontiguousarray(poses, np.float32)
    ref = {}
    for i in range(0, len(poses), stride):
        p = poses[i]
        ref[i] = port.cloud_weight(cells, dims, bounds, cloud, (p[0], p[1], p[2], roll, pitch, p[3]))
    # variants >= 100: the fused gather + ordered-add kernel (weight_ordered.cuh), 4 / 8 gatherer warps per 32 particles
    for variant in tuple(variants) + (104, 108):
        if variant >= 100:
            ctx.set_option("ordered_mode", 2)
            ctx.set_option("weight_block_threads", 160 if variant == 104 else 288)
        else:
            ctx.set_option("ordered_mode", 1)
            ctx.set_option("weight_variant", variant)
            ctx.set_option("weight_point_splits", 1)
        try:
            w_g, n_g = g.cloud_weight_batch(cloud, poses, roll, pitch)
        finally:
            ctx.set_option("weight_variant", 0)
            ctx.set_option("weight_point_splits", 0)
            ctx.set_option("ordered_mode", 0)
            ctx.set_option("weight_block_threads", 0)
        for i, (w_o, n_o) in ref.items():
            p = poses[i]
            if not port.is_into_map(bounds, p[0], p[1], p[2]):
                continue  # the batched kernel gates on isIntoMap like ParticleFilter::update does
            assert n_g[i] == n_o, (variant, i, p)
            assert bits(w_g[i]) == bits(w_o), (variant, i, p)


def test_partial_last_voxel_and_border_poses(cuda_ctx, port):
    # extents that are not a multiple of the resolution on any axis; many points fall outside
    bounds = np.array([-3.07, -2.01, -0.33, 4.13, 3.9, 2.2, 0.07])
    g, cells, dims = random_grid(cuda_ctx, bounds, 11)
    rng = np.random.default_rng(12)
    cloud = np.zeros((3001, 4), np.float32)
    cloud[:, :3] = rng.normal(0, 2.5, (3001, 3))
    poses = np.zeros((301, 4), np.float32)
    poses[:, 0] = rng.uniform(bounds[0], bounds[3], 301)
    poses[:, 1] = rng.uniform(bounds[1], bounds[4], 301)
    poses[:, 2] = rng.uniform(bounds[2], bounds[5], 301)
    poses[:, 3] = rng.uniform(-3.2, 3.2, 301)
    poses[:8, 0] = [bounds[0], np.nextafter(np.float32(bounds[3]), np.float32(-9)), 4.12, 4.06, -3.07, 0, 0, 0]
    poses[8:12, 2] = [-0.33, 2.19, 2.13, 2.199999]
    check_batch(cuda_ctx, port, g, cells, dims, bounds, cloud, poses, np.float32(0.03), np.float32(-0.05))
    g.close()


def test_lattice_aligned_clouds_in_the_batched_kernel(cuda_ctx, port):
    """Transformed coordinates land exactly on voxel faces (and on the map's far faces): the estimate must hand
    every one of them to the exact path."""
    bounds = np.array([0.0, 0.0, 0.0, 6.4, 6.4, 3.2, 0.1])
    g, cells, dims = random_grid(cuda_ctx, bounds, 21)
    k = np.arange(-40, 41, dtype=np.float64)
    xs, ys = np.meshgrid(k * 0.1, k * 0.1, indexing="ij")
    cloud = np.zeros((xs.size, 4), np.float32)
    cloud[:, 0] = xs.ravel().astype(np.float32)
    cloud[:, 1] = ys.ravel().astype(np.float32)
    cloud[:, 2] = np.float32(0.2) * (np.arange(xs.size) % 7)
    poses = []
    for tx in (0.0, 0.1, 3.2, 3.25, 6.3, 6.399999):
        for ty in (0.0, 1.7, 3.2):
            for yaw in (0.0, np.pi / 2, np.pi, -np.pi / 2):
                poses.append((tx, ty, 0.4, yaw))
    poses = np.array(poses, np.float32)
    check_batch(cuda_ctx, port, g, cells, dims, bounds, cloud, poses, np.float32(0.0), np.float32(0.0))
    g.close()


def test_large_coordinates_fine_resolution(cuda_ctx, port):
    # 2000 x 2000 x 8 voxels @ 0.05 m: the error band of the estimate is widest here (coordinates up to 2000 voxels,
    # points up to 60 m from the sensor)
    bounds = np.array([-50.0, -50.0, 0.0, 50.0, 50.0, 0.4, 0.05])
    g, cells, dims = random_grid(cuda_ctx, bounds, 31)
    rng = np.random.default_rng(32)
    cloud = np.zeros((4099, 4), np.float32)
    r = rng.uniform(0.5, 60.0, 4099)
    th = rng.uniform(-np.pi, np.pi, 4099)
    cloud[:, 0] = r * np.cos(th)
    cloud[:, 1] = r * np.sin(th)
    cloud[:, 2] = rng.uniform(-0.2, 0.2, 4099)
    poses = np.zeros((200, 4), np.float32)
    poses[:, 0] = rng.uniform(-49.9, 49.9, 200)
    poses[:, 1] = rng.uniform(-49.9, 49.9, 200)
    poses[:, 2] = rng.uniform(0.0, 0.39, 200)
    poses[:, 3] = rng.uniform(-3.2, 3.2, 200)
    check_batch(cuda_ctx, port, g, cells, dims, bounds, cloud, poses, np.float32(0.002), np.float32(0.001), stride=3)
    g.close()


def test_large_coordinates_bricked(cuda_ctx, port):
    import amcl3d_b200
    ctx = amcl3d_b200.Context(0)
    ctx.set_option("grid_layout", 2)
    try:
        bounds = np.array([-20.0, -20.0, -1.0, 20.03, 19.98, 3.0, 0.05])
        g, cells, dims = random_grid(ctx, bounds, 41)
        rng = np.random.default_rng(42)
        cloud = np.zeros((2500, 4), np.float32)
        cloud[:, :2] = rng.normal(0, 12.0, (2500, 2))
        cloud[:, 2] = rng.normal(0, 1.5, 2500)
        poses = np.zeros((257, 4), np.float32)
        poses[:, 0] = rng.uniform(-20, 20, 257)
        poses[:, 1] = rng.uniform(-20, 19.9, 257)
        poses[:, 2] = rng.uniform(-1, 3, 257)
        poses[:, 3] = rng.uniform(-3.2, 3.2, 257)
        check_batch(ctx, port, g, cells, dims, bounds, cloud, poses, np.float32(-0.02), np.float32(0.04), stride=2)
        g.close()
    finally:
        ctx.close()


def test_non_finite_points_are_skipped_like_the_reference(cuda_ctx, port):
    bounds = np.array([-4.0, -4.0, 0.0, 4.0, 4.0, 2.0, 0.1])
    g, cells, dims = random_grid(cuda_ctx, bounds, 51)
    rng = np.random.default_rng(52)
    cloud = np.zeros((1500, 4), np.float32)
    cloud[:, :3] = rng.normal(0, 2.0, (1500, 3))
    cloud[100, 0] = np.nan
    cloud[700, 1] = np.inf
    cloud[701, 2] = -np.inf
    cloud[1400, :3] = 3e38
    poses = np.zeros((70, 4), np.float32)
    poses[:, :2] = rng.uniform(-3.9, 3.9, (70, 2))
    poses[:, 2] = rng.uniform(0.1, 1.9, 70)
    poses[:, 3] = rng.uniform(-3.2, 3.2, 70)
    check_batch(cuda_ctx, port, g, cells, dims, bounds, cloud, poses, np.float32(0.01), np.float32(0.02))
    g.close()
