"""GPU: the bricked physical grid layout (32^3-voxel bricks, used automatically for grids larger than L2) must be
invisible through the C-ABI: same logical indices, same cells, same weights as the linear layout and the oracle."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


@pytest.fixture()
def bricked(cuda_ctx):
    cuda_ctx.set_option("grid_layout", 2)
    yield cuda_ctx
    cuda_ctx.set_option("grid_layout", 0)


def test_cells_round_trip_and_ranges(bricked, cfg1, cfg1_cells):
    import amcl3d_b200
    cells, dims = cfg1_cells          # 200 x 200 x 50: none of the axes is a multiple of 32 -> partial bricks
    g = amcl3d_b200.Grid(bricked, cfg1["bounds"])
    g.upload_cells(cells, 0.05)
    assert np.array_equal(bits(g.download_cells()), bits(cells))
    assert np.array_equal(bits(g.download_prob()), bits(cells[:, 1]))
    for first, count in ((0, 1), (12345, 40200), (1999000, 2000), (1990000, 50000)):
        got = g.download_prob_range(first, count)
        want = np.zeros(count, np.float32)
        m = max(0, min(count, len(cells) - first))
        want[:m] = cells[first:first + m, 1]
        assert np.array_equal(bits(got), bits(want))
    g.close()


def test_compute_grid_bricked_vs_oracle(bricked, cfg1, cfg1_cells):
    import amcl3d_b200
    cells, _ = cfg1_cells
    g = amcl3d_b200.Grid(bricked, cfg1["bounds"])
    g.compute(cfg1["map_points"], 0.05)
    got = g.download_cells()
    assert np.array_equal(bits(got[:, 0]), bits(cells[:, 0]))
    np.testing.assert_allclose(got[:, 1], cells[:, 1], rtol=1e-5, atol=1e-36)
    g.close()


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_weights_bit_exact_on_bricked_grid(bricked, port, cfg1, cfg1_cells, variant):
    import amcl3d_b200
    cells, dims = cfg1_cells
    g = amcl3d_b200.Grid(bricked, cfg1["bounds"])
    g.upload_cells(cells, 0.05)
    roll, pitch = np.float32(0.01), np.float32(-0.02)
    poses = cfg1["particles"][:, :4]
    bricked.set_option("weight_point_splits", 1)
    bricked.set_option("weight_variant", variant)
    try:
        w_g, n_g = g.cloud_weight_batch(cfg1["cloud"], poses, roll, pitch)
    finally:
        bricked.set_option("weight_point_splits", 0)
        bricked.set_option("weight_variant", 0)
    for i in range(0, 600, 11):
        p = poses[i]
        w_o, n_o = port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], (p[0], p[1], p[2], roll, pitch, p[3]))
        assert n_g[i] == n_o and bits(w_g[i]) == bits(w_o), (variant, i)
    # single-pose entry: logical voxel indices are reported, whatever the storage
    w1, n1, idx1 = g.cloud_weight(cfg1["cloud"], (0.3, -0.2, 2.4, roll, pitch, 0.25), want_idx=True)
    w2, n2, idx2 = port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], (0.3, -0.2, 2.4, roll, pitch, 0.25),
                                     want_idx=True)
    assert np.array_equal(idx1, idx2) and bits(w1) == bits(w2)
    g.close()


def test_full_update_same_bits_in_both_layouts(cuda_ctx, cfg1, cfg1_cells):
    import amcl3d_b200
    cells, _ = cfg1_cells
    out = []
    for layout in (1, 2):
        cuda_ctx.set_option("grid_layout", layout)
        cuda_ctx.set_option("weight_point_splits", 1)
        cuda_ctx.set_option("sum_mode", 1)
        g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
        g.upload_cells(cells, 0.05)
        f = amcl3d_b200.Filter(cuda_ctx)
        f.upload(cfg1["particles"])
        mean = f.update(g, cfg1["cloud"], cfg1["ranges"], 0.5, 0.53, 0.01, -0.02)
        out.append((f.download(), mean))
        f.close()
        g.close()
    for k in ("grid_layout", "weight_point_splits", "sum_mode"):
        cuda_ctx.set_option(k, 0)
    assert np.array_equal(bits(out[0][0]), bits(out[1][0])) and np.array_equal(bits(out[0][1]), bits(out[1][1]))
