"""GPU parity: computeGrid (PointCloudTools.cpp:84-149) through the C-ABI vs the CPU oracle and the goldens."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def assert_grid_close(got, want):
    """dist: the float minimum is unique -> bit-exact.  prob: device expf vs glibc expf, <= 1e-5 relative
    (north_star), with an absolute floor for the denormal tail just before the underflow to 0."""
    assert np.array_equal(bits(got[:, 0]), bits(want[:, 0]))
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=1e-5, atol=1e-36)


def test_compute_grid_room_vs_oracle(cuda_ctx, cfg1, cfg1_cells):
    import amcl3d_b200
    cells, dims = cfg1_cells
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.compute(cfg1["map_points"], cfg1["sensor_dev"])
    assert_grid_close(g.download_cells(), cells)
    g.close()


def test_compute_grid_matches_committed_reference_sample(cuda_ctx, cfg1, ref_cfg1):
    import amcl3d_b200
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.compute(cfg1["map_points"], cfg1["sensor_dev"])
    got = g.download_cells()[ref_cfg1["cells_sample_idx"]]
    assert_grid_close(got, ref_cfg1["cells_sample"])
    g.close()


def test_kat1_kat2_on_gpu_grid(cuda_ctx, kat, port):
    """The reference's goldens end to end on the device: computeGrid of map T, then computeCloudWeight."""
    import amcl3d_b200
    g = amcl3d_b200.Grid(cuda_ctx, kat["bounds"])
    assert list(g.dims) == [522, 384, 153]
    g.compute(kat["map_points"], float(kat["sensor_dev"]))
    w, n = g.cloud_weight(kat["sensor_cloud"], kat["kat1_pose"])
    assert n == 934
    assert abs(float(w) - float(kat["kat1_expected"])) <= float(kat["kat1_tol"])   # tests/Grid3dTest.cpp:168
    cells = g.download_cells()
    # KAT-2: probability slice (int8 scaling of Grid3d.cpp:100-118 done by the oracle's slice builder on GPU cells)
    sl = port.grid_slice(cells, g.dims, kat["bounds"], float(kat["bounds"][2]) + 1.02)
    diff = np.abs(sl.astype(np.int32) - kat["nav_slice"].astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999
    # and the oracle's own weight on the GPU-built cells agrees bit for bit with the GPU gather
    w_o, n_o = port.cloud_weight(cells, g.dims, kat["bounds"], kat["sensor_cloud"], kat["kat1_pose"])
    assert bits(w) == bits(w_o)
    # spot-check distances against exhaustive search
    rng = np.random.default_rng(2)
    for _ in range(25):
        ix, iy, iz = int(rng.integers(522)), int(rng.integers(384)), int(rng.integers(153))
        d = port.nn_dist2_bruteforce(kat["map_points"], kat["bounds"], ix, iy, iz)
        assert np.float32(d) == cells[ix + 522 * (iy + 384 * iz), 0]
    g.close()


def test_crop_of_warehouse_vs_oracle(cuda_ctx, port):
    """Config 3 parity recipe: a sub-volume of the warehouse map whose bounds are restricted while ALL map points
    (also those outside the crop) stay in the cloud."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    pts, bounds = synth.map_warehouse(size=(16.0, 12.0, 10.0), res=0.05, n_pallets=12, seed=5)
    crop = np.array([-3.0, -2.0, 0.5, 1.5, 2.05, 2.15, 0.05])   # 90 x 81 x 33 voxels, ragged (not multiples of 8)
    want, dims = port.compute_grid(pts, crop, 0.05)
    g = amcl3d_b200.Grid(cuda_ctx, crop)
    assert list(g.dims) == list(dims)
    g.compute(pts, 0.05)
    assert_grid_close(g.download_cells(), want)
    g.close()


def test_far_field_and_offgrid_points(cuda_ctx, port):
    """Sparse, off-lattice points far from most voxels: exercises many search rings and points outside the bounds."""
    import amcl3d_b200
    rng = np.random.default_rng(9)
    pts = rng.uniform(-6, 6, (40, 3)).astype(np.float32)
    bounds = np.array([-4.0, -4.0, -2.0, 4.0, 4.0, 2.0, 0.1])
    want, dims = port.compute_grid(pts, bounds, 0.05)
    g = amcl3d_b200.Grid(cuda_ctx, bounds)
    g.compute(pts, 0.05)
    assert_grid_close(g.download_cells(), want)
    g.close()


def test_empty_map_cloud(cuda_ctx):
    import amcl3d_b200
    g = amcl3d_b200.Grid(cuda_ctx, [0, 0, 0, 1, 1, 1, 0.1])
    g.compute(np.zeros((0, 3), np.float32), 0.05)
    cells = g.download_cells()
    assert np.all(cells[:, 0] == -1.0) and np.all(cells[:, 1] == 0.0)   # PointCloudTools.cpp:139-143
    g.close()


def test_cell_cap_is_enforced(cuda_ctx):
    import amcl3d_b200
    cuda_ctx.set_option("max_cells", 250000000)
    with pytest.raises(amcl3d_b200.Amcl3dCudaError):
        amcl3d_b200.Grid(cuda_ctx, [-50, -50, 0, 50, 50, 20, 0.05])   # 1.6 G cells > the reference's cap
    cuda_ctx.set_option("max_cells", 0)


def test_prob_only_grid(cuda_ctx, cfg1, cfg1_cells):
    import amcl3d_b200
    cells, _ = cfg1_cells
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.compute(cfg1["map_points"], cfg1["sensor_dev"], keep_dist=False)
    np.testing.assert_allclose(g.download_prob(), cells[:, 1], rtol=1e-5, atol=1e-36)
    g.close()


@pytest.mark.parametrize("sensor_dev", [0.05, 0.3])
def test_prob_only_far_field_cutoff(cuda_ctx, port, sensor_dev):
    """Probability-only builds stop the nearest-neighbour search where prob underflows to exactly +0
    (DfParams::d2_cut): the probability plane must not change, near or far, for narrow and wide sensor models."""
    import amcl3d_b200
    rng = np.random.default_rng(19)
    pts = rng.uniform(-6, 6, (60, 3)).astype(np.float32)
    bounds = np.array([-4.0, -4.0, -2.0, 4.0, 4.0, 2.0, 0.1])
    want, _ = port.compute_grid(pts, bounds, sensor_dev)
    g = amcl3d_b200.Grid(cuda_ctx, bounds)
    g.compute(pts, sensor_dev, keep_dist=False)
    got = g.download_prob()
    np.testing.assert_allclose(got, want[:, 1], rtol=1e-5, atol=1e-36)
    # zeros where the reference underflows (up to one denormal step at the underflow edge of expf)
    assert np.all(got[want[:, 1] == 0] < 3e-45) and np.all(want[got == 0, 1] < 3e-45)
    assert (got == 0).any() and (got > 0).any()
    g.close()
