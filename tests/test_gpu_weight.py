"""GPU parity: Grid3d::computeCloudWeight / isIntoMap through the C-ABI vs the CPU oracle."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def grid_S(cuda_ctx, cfg1, cfg1_cells):
    import amcl3d_b200
    cells, dims = cfg1_cells
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    assert list(g.dims) == list(dims)
    g.upload_cells(cells, cfg1["sensor_dev"])
    yield g
    g.close()


def test_unopened_grid_gives_zero(cuda_ctx, cfg1):
    import amcl3d_b200
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    w, n = g.cloud_weight(cfg1["cloud"], (0, 0, 2.5, 0, 0, 0.3))  # Grid3d.cpp:136-137
    assert float(w) == 0.0 and n == 0
    g.close()


def test_cells_round_trip(grid_S, cfg1_cells):
    cells, _ = cfg1_cells
    assert np.array_equal(bits(grid_S.download_cells()), bits(cells))
    assert np.array_equal(bits(grid_S.download_prob()), bits(cells[:, 1]))


def test_single_pose_bit_exact(grid_S, port, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    rng = np.random.default_rng(5)
    poses = [(0, 0, 2.5, 0, 0, 0.3), (0, 0, 2.5, 0.01, -0.02, 0.3), (3.3, -4.1, 1.2, 0.1, 0.2, -2.5),
             (-9.95, 9.95, 0.05, 0, 0, 3.1), (9.9, 0, 4.9, -0.3, 0.25, 1.0), (50, 50, 50, 0, 0, 0)]
    poses += [tuple(rng.uniform(-8, 8, 2)) + (rng.uniform(0.2, 4.8),) + tuple(rng.uniform(-0.2, 0.2, 2)) +
              (rng.uniform(-3.1, 3.1),) for _ in range(10)]
    for pose in poses:
        w_o, n_o, idx_o = port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], pose, want_idx=True)
        w_g, n_g, idx_g = grid_S.cloud_weight(cfg1["cloud"], pose, want_idx=True)
        assert np.array_equal(idx_g, idx_o), pose          # voxel indices bit-exact
        assert n_g == n_o
        assert bits(w_g) == bits(w_o), pose                # sequential float sum reproduced


def test_lattice_aligned_points_hit_exact_division(grid_S, port, cfg1, cfg1_cells):
    """Points that land exactly on voxel faces (k * 0.1 as floats) force the exact-division slow path."""
    cells, dims = cfg1_cells
    k = np.arange(-60, 60, dtype=np.float64)
    xs, ys = np.meshgrid(k * 0.1, k * 0.1, indexing="ij")
    cloud = np.zeros((xs.size, 4), np.float32)
    cloud[:, 0] = xs.ravel().astype(np.float32)
    cloud[:, 1] = ys.ravel().astype(np.float32)
    cloud[:, 2] = np.float32(0.3)
    for pose in [(0, 0, 1.0, 0, 0, 0), (0.1, -0.2, 2.0, 0, 0, 0), (0.5, 0.5, 0.2, 0, 0, np.pi / 2), (1, 1, 1, 0, 0, 0)]:
        w_o, n_o, idx_o = port.cloud_weight(cells, dims, cfg1["bounds"], cloud, pose, want_idx=True)
        w_g, n_g, idx_g = grid_S.cloud_weight(cloud, pose, want_idx=True)
        assert np.array_equal(idx_g, idx_o), pose
        assert bits(w_g) == bits(w_o)


def test_few_points_gives_zero(grid_S, port, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    for m in (0, 1, 10, 11, 12):
        cloud = cfg1["cloud"][:m]
        w_g, n_g = grid_S.cloud_weight(cloud, (0, 0, 2.5, 0, 0, 0.3))
        if m == 0:
            assert float(w_g) == 0.0
            continue
        w_o, n_o = port.cloud_weight(cells, dims, cfg1["bounds"], cloud, (0, 0, 2.5, 0, 0, 0.3))
        assert n_g == n_o and bits(w_g) == bits(w_o)   # n <= 10 -> 0 (Grid3d.cpp:198)


def test_batch_unsplit_bit_exact(cuda_ctx, grid_S, port, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    roll, pitch = np.float32(cfg1["roll"]), np.float32(cfg1["pitch"])
    poses = cfg1["particles"][:, :4].copy()
    poses[5, 0] = 100.0  # out of the map: still evaluated by this entry (no isIntoMap gate in computeCloudWeight)
    cuda_ctx.set_option("weight_point_splits", 1)
    try:
        w_g, n_g = grid_S.cloud_weight_batch(cfg1["cloud"], poses, roll, pitch)
    finally:
        cuda_ctx.set_option("weight_point_splits", 0)
    for i in list(range(0, 600, 7)) + [5]:
        p = poses[i]
        w_o, n_o = port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], (p[0], p[1], p[2], roll, pitch, p[3]))
        if i == 5:
            continue  # the batched kernel gates on isIntoMap like ParticleFilter::update does
        assert n_g[i] == n_o and bits(w_g[i]) == bits(w_o), i


def test_batch_auto_split_within_tolerance(cuda_ctx, grid_S, port, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    roll, pitch = np.float32(cfg1["roll"]), np.float32(cfg1["pitch"])
    poses = cfg1["particles"][:, :4]
    for splits in (0, 3, 16):
        cuda_ctx.set_option("weight_point_splits", splits)
        w_g, n_g = grid_S.cloud_weight_batch(cfg1["cloud"], poses, roll, pitch)
        cuda_ctx.set_option("weight_point_splits", 0)
        for i in range(0, 600, 13):
            p = poses[i]
            w_o, n_o = port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], (p[0], p[1], p[2], roll, pitch, p[3]))
            assert n_g[i] == n_o
            assert abs(float(w_g[i]) - float(w_o)) <= 1e-5 * abs(float(w_o))  # north_star tolerance: 1e-5 relative


def test_block_size_variants_agree(cuda_ctx, grid_S, cfg1):
    roll, pitch = np.float32(0.0), np.float32(0.0)
    poses = cfg1["particles"][:, :4]
    cuda_ctx.set_option("weight_point_splits", 1)
    outs = []
    for bt in (64, 128, 256):
        cuda_ctx.set_option("weight_block_threads", bt)
        outs.append(grid_S.cloud_weight_batch(cfg1["cloud"], poses, roll, pitch))
    cuda_ctx.set_option("weight_block_threads", 0)
    cuda_ctx.set_option("weight_point_splits", 0)
    for w, n in outs[1:]:
        assert np.array_equal(bits(w), bits(outs[0][0])) and np.array_equal(n, outs[0][1])


def test_is_into_map(grid_S, port, cfg1):
    for xyz in [(0, 0, 2.5), (-10, -10, 0), (10, 0, 1), (9.999999, 9.999999, 4.999999), (0, 0, -1e-6), (1, 1, 1),
                (-100, -100, -100)]:
        assert grid_S.is_into_map(*xyz) == port.is_into_map(cfg1["bounds"], *xyz)


def test_empty_and_ragged_batches(grid_S, cfg1):
    w, n = grid_S.cloud_weight_batch(cfg1["cloud"], np.zeros((0, 4), np.float32), 0, 0)
    assert len(w) == 0
    w, n = grid_S.cloud_weight_batch(cfg1["cloud"][:0], cfg1["particles"][:33, :4], 0, 0)
    assert np.all(w == 0) and np.all(n == 0)
    # a particle count that is not a multiple of the block, a cloud that is not a multiple of the tile/unroll
    w, n = grid_S.cloud_weight_batch(cfg1["cloud"][:1031], cfg1["particles"][:129, :4], 0, 0)
    assert len(w) == 129 and np.all(n <= 1031)
