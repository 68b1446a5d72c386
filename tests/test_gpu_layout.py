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


@pytest.mark.parametrize("variant", [0, 4, 5])
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


def test_sequential_chunk_launches_are_bit_exact(cuda_ctx, cfg1, cfg1_cells):
    """Large particle sets walk the cloud in sequential chunk launches with carried running sums: same bits as one
    launch over the whole cloud."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    cells, _ = cfg1_cells
    n = 120000
    particles = synth.particles_tracking(n, cfg1["pose"], (0.2, 0.2, 0.2, 0.4), seed=12)
    cloud = cfg1["cloud"][:700]
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.upload_cells(cells, 0.05)
    outs = []
    for chunk in (0, 128, 300):
        cuda_ctx.set_option("weight_point_splits", 1)
        cuda_ctx.set_option("weight_chunk_points", chunk)
        cuda_ctx.set_option("sum_mode", 2)
        f = amcl3d_b200.Filter(cuda_ctx)
        f.upload(particles)
        f.update(g, cloud, None, 0.5, 0.53, 0.01, -0.02)
        outs.append((f.download(), f.last_in_map_evals()))
        f.close()
    for k in ("weight_point_splits", "weight_chunk_points", "sum_mode"):
        cuda_ctx.set_option(k, 0)
    g.close()
    for got, evals in outs[1:]:
        assert evals == outs[0][1]
        assert np.array_equal(bits(got[:, 5]), bits(outs[0][0][:, 5]))


def test_morton_reordered_cloud_within_tolerance_and_deterministic(cuda_ctx, port, cfg1, cfg1_cells):
    """cloud_order = 2 re-orders the staged cloud along a Morton curve on the device: weights move only by float
    summation order (<= 1e-5 relative, north_star) and the permutation is deterministic."""
    import amcl3d_b200
    cells, dims = cfg1_cells
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.upload_cells(cells, 0.05)
    want, mean_o = port.update(cfg1["particles"], cells, dims, cfg1["bounds"], cfg1["cloud"], cfg1["ranges"], 0.5, 0.53,
                               0.01, -0.02)
    runs = []
    for _ in range(2):
        cuda_ctx.set_option("cloud_order", 2)
        cuda_ctx.set_option("weight_point_splits", 1)
        cuda_ctx.set_option("sum_mode", 1)
        f = amcl3d_b200.Filter(cuda_ctx)
        f.upload(cfg1["particles"])
        mean = f.update(g, cfg1["cloud"], cfg1["ranges"], 0.5, 0.53, 0.01, -0.02)
        runs.append((f.download(), mean, f.last_in_map_evals()))
        f.close()
    for k in ("cloud_order", "weight_point_splits", "sum_mode"):
        cuda_ctx.set_option(k, 0)
    g.close()
    assert np.array_equal(bits(runs[0][0]), bits(runs[1][0]))                      # deterministic
    np.testing.assert_allclose(runs[0][0][:, 4:], want[:, 4:], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(runs[0][1], mean_o, atol=1e-4)
    total = sum(port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], (p[0], p[1], p[2], 0.01, -0.02, p[3]))[1]
                for p in cfg1["particles"])
    assert runs[0][2] == total                                                     # same points hit, only re-ordered


@pytest.mark.parametrize("n,mode", [(5000, 1), (40000, 3), (120001, 3)])
def test_particle_scheduling_order_does_not_change_any_bit(cuda_ctx, cfg1, cfg1_cells, n, mode):
    """particle_order = 2 lets the weighting kernel walk the particles in pose-sorted order (order.cu; single-CTA path
    up to 32768 particles, multi-kernel path above): pure scheduling, every particle's result lands in its own slot."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    cells, _ = cfg1_cells
    particles = synth.particles_tracking(n, cfg1["pose"], (0.3, 0.3, 0.2, 0.5), seed=21)
    particles[7, 0] = 500.0          # outside the map
    particles[11, 3] = np.nan        # a broken pose must not break the permutation
    cloud = cfg1["cloud"][:900]
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.upload_cells(cells, 0.05)
    outs = []
    for order in (1, 2):
        cuda_ctx.set_option("particle_order", order)
        cuda_ctx.set_option("weight_point_splits", 1)
        cuda_ctx.set_option("sum_mode", mode)
        f = amcl3d_b200.Filter(cuda_ctx)
        f.upload(particles)
        f.update(g, cloud, None, 0.5, 0.53, 0.01, -0.02)
        outs.append((f.download(), f.last_in_map_evals()))
        f.close()
    for k in ("particle_order", "weight_point_splits", "sum_mode"):
        cuda_ctx.set_option(k, 0)
    g.close()
    assert outs[0][1] == outs[1][1] and outs[0][1] > 0
    ok = np.ones(n, bool)
    ok[11] = False  # NaN pose: weights are NaN in both runs
    assert np.array_equal(bits(outs[0][0][ok][:, 4:]), bits(outs[1][0][ok][:, 4:]))


def test_split_chunk_launches_match_the_oracle(cuda_ctx, port, cfg1, cfg1_cells):
    """Point splits / sub-chunk CTAs / a Morton-ordered cloud re-associate the per-particle sum (partials accumulated in
    double): same points as the reference (counts exact), weights within 2e-6 of the exact (fp64) sum of the same
    cells and -- on this 1 000-point cloud, where the reference's own float chain is still accurate -- within 1e-5 of the
    reference.  One split in the caller's order (the default) is the reference's chain bit for bit."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    cells, dims = cfg1_cells
    n = 120000
    particles = synth.particles_tracking(n, cfg1["pose"], (0.2, 0.2, 0.2, 0.4), seed=14)
    cloud = cfg1["cloud"][:1000]
    w_o, n_o = port.cloud_weight_batch(cells, dims, cfg1["bounds"], cloud, particles[:, :4], 0.01, -0.02)
    inmap = np.array([port.is_into_map(cfg1["bounds"], *q[:3]) for q in particles])
    pick = np.nonzero(inmap)[0][::600]
    true = []
    for i in pick:
        q = particles[i]
        idx, cnt = port.cloud_indices(dims, cfg1["bounds"], cloud, (q[0], q[1], q[2], 0.01, -0.02, q[3]))
        true.append(cells[idx[idx != 0xFFFFFFFF], 1].astype(np.float64).sum() / max(cnt, 1) if cnt > 10 else 0.0)
    true = np.array(true)
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.upload_cells(cells, 0.05)
    for splits, chunk, order, ref_order in ((0, 0, 0, 1), (1, 0, 1, 1), (2, 256, 1, 1), (4, 512, 1, 1), (3, 200, 1, 1),
                                            (4, 512, 2, 1), (0, 0, 0, 0)):
        cuda_ctx.set_option("weight_point_splits", splits)
        cuda_ctx.set_option("weight_chunk_points", chunk)
        cuda_ctx.set_option("cloud_order", order)
        cuda_ctx.set_option("reference_order", ref_order)
        f = amcl3d_b200.Filter(cuda_ctx)
        f.upload(particles)
        f.update(g, cloud, None, 0.5, 0.53, 0.01, -0.02)
        w_g, n_g = f.last_cloud_weights()
        f.close()
        for k in ("weight_point_splits", "weight_chunk_points", "cloud_order"):
            cuda_ctx.set_option(k, 0)
        cuda_ctx.set_option("reference_order", 1)
        assert np.array_equal(n_g[inmap], n_o[inmap]), (splits, chunk, order)
        if (splits == 1 and order == 1) or (splits == 0 and ref_order == 1):
            assert np.array_equal(bits(w_g[inmap]), bits(w_o[inmap]))      # one float chain in the caller's order
        else:
            nz = true > 0
            assert (np.abs(w_g[pick][nz] - true[nz]) / true[nz]).max() <= 2e-6, (splits, chunk, order)
        rel = np.abs(w_g[inmap] - w_o[inmap]) / np.maximum(w_o[inmap], 1e-30)
        assert rel.max() <= 1e-5, (splits, chunk, order, float(rel.max()))
    g.close()


@pytest.mark.parametrize("n,n_pts", [(600, 2000), (5000, 3001), (40000, 700), (33, 1), (1000, 131), (4999, 513)])
def test_gather_then_replay_equals_the_direct_chain(cuda_ctx, port, cfg1, cfg1_cells, n, n_pts):
    """Reference summation order has two implementations: `direct` (one lane walks the cloud) and `replay` (the gathers
    run split over many CTAs and store their values; replay_sum_kernel adds them in the caller's order).  Same bits, and
    both are the oracle's."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    cells, dims = cfg1_cells
    particles = synth.particles_tracking(n, cfg1["pose"], (0.3, 0.3, 0.2, 0.5), seed=31)
    particles[5, 0] = 400.0
    cloud = synth.sensor_cloud(cfg1["map_points"], cfg1["pose"], n_pts, 8.0, seed=77)
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.upload_cells(cells, 0.05)
    out = {}
    for name, replay, layout_sorted in (("direct", 1, 0), ("replay", 2, 0), ("replay_morton", 2, 2), ("ordered", 1, 0)):
        cuda_ctx.set_option("replay", replay)
        cuda_ctx.set_option("ordered_mode", 2 if name == "ordered" else 1)
        if layout_sorted:
            cuda_ctx.set_option("cloud_order", 0)
        f = amcl3d_b200.Filter(cuda_ctx)
        f.upload(particles)
        f.update(g, cloud, None, 0.5, 0.53, 0.01, -0.02)
        out[name] = f.last_cloud_weights()
        f.close()
    cuda_ctx.set_option("replay", 0)
    cuda_ctx.set_option("ordered_mode", 0)
    g.close()
    pick = np.arange(0, n, max(1, n // 500))
    w_o, n_o = port.cloud_weight_batch(cells, dims, cfg1["bounds"], cloud, particles[pick, :4], 0.01, -0.02)
    inmap = np.array([port.is_into_map(cfg1["bounds"], *particles[i, :3]) for i in pick])
    for name, (w_g, n_g) in out.items():
        assert np.array_equal(n_g[pick][inmap], n_o[inmap]), name
        assert np.array_equal(bits(w_g[pick][inmap]), bits(w_o[inmap])), name
        assert np.array_equal(bits(w_g), bits(out["direct"][0])), name
