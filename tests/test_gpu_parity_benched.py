"""GPU parity on the configurations bench.py measures, through exactly the code paths it takes (every option at its
default): BASELINE.json configs[1] (10 k particles x 10 k points, map S) against the UNMODIFIED reference's update(),
and the large-map path of configs[3] / [4] (bricked grid, Morton-ordered cloud, sequential chunk launches with
sub-chunk CTAs, 32 k-point cloud, 131 072 particles, tracking and uniform poses) on a <= 250 M-cell warehouse map the
oracle can hold.  Tolerances are north_star's: counts / indices exact, weights 1e-5 relative, mean pose 1e-4 m."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu

ALPHA, SIGMA, ROLL, PITCH = 0.5, 0.53, 0.01, -0.02


def _rel(a, b):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.abs(b.astype(np.float64)), 1e-300)


def test_cfg2_exactly_as_benched_matches_the_unmodified_reference(cuda_ctx, reference, port, cfg1_cells):
    """10 000 particles x 10 000 points, three beacons, all options default: point splits chosen by the library (29 on a
    B200), pose-sorted scheduling, packed-pair kernel, segmented exact sums, host-buffer entry point -- the bench's e2e
    call -- against ParticleFilter::update of the reference sources compiled in oracle/_ref."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    w = synth.make_workload("cfg2")
    cells, dims = cfg1_cells                     # map S (same map as cfg1), oracle-built
    G = reference.grid()
    assert G.set_cells(w["map_points"], w["bounds"], w["sensor_dev"], dims, cells)
    G.set_cloud(w["cloud"])
    F = reference.filter()
    F.set_particles(w["particles"])
    F.update(G, w["ranges"], ALPHA, SIGMA, ROLL, PITCH)
    want, mean_o = F.particles(), F.mean()[:4]

    for build in ("gpu_grid", "oracle_grid"):
        grid = amcl3d_b200.Grid(cuda_ctx, w["bounds"])
        if build == "gpu_grid":
            grid.compute(w["map_points"], w["sensor_dev"], keep_dist=False)    # what bench.py does
        else:
            grid.upload_cells(cells, w["sensor_dev"])
        assert cuda_ctx.get_option("sum_mode") == 0 and cuda_ctx.get_option("weight_point_splits") == 0
        pf = amcl3d_b200.Filter(cuda_ctx)
        pf.upload(w["particles"])
        mean_g = pf.update(grid, w["cloud"], w["ranges"], ALPHA, SIGMA, ROLL, PITCH)
        got = pf.download()
        raw_w, raw_n = pf.last_cloud_weights()
        pf.close()
        grid.close()
        # contributing points per particle: exact (voxel indices are bit-exact, whatever the cell values)
        _, n_o = port.cloud_weight_batch(None, dims, w["bounds"], w["cloud"], w["particles"][:, :4], ROLL, PITCH)
        inmap = np.array([port.is_into_map(w["bounds"], *q[:3]) for q in w["particles"]])
        assert np.array_equal(raw_n[inmap], n_o[inmap])
        assert np.array_equal(bits(got[:, :4]), bits(want[:, :4]))
        for col, name in ((5, "wp"), (6, "wr"), (4, "w")):
            r = _rel(got[:, col], want[:, col])
            assert r.max() <= 1e-5, (build, name, float(r.max()))
        assert np.abs(mean_g - mean_o).max() <= 1e-4, (build, mean_g, mean_o)
        if build == "oracle_grid":
            # identical cells: the only differences left are the summation order of the point chunks (partials added in
            # double) and the device's exp() in wr
            assert _rel(got[:, 5], want[:, 5]).max() <= 2e-6


@pytest.fixture(scope="module")
def hall(cuda_ctx):
    """A 32 x 32 x 10 m warehouse hall @ 0.05 m = 640 x 640 x 200 = 81.9 M cells (328 MB probability plane: larger than
    L2, so the library bricks it), built on the GPU; the same cells are handed to the oracle."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    pts, bounds = synth.map_warehouse(size=(32.0, 32.0, 10.0), res=0.05, n_pallets=60, seed=5)
    pose = np.array([-3.0, 1.0, 1.5, 0.2])
    cloud = synth.sensor_cloud(pts, pose, 32768, 14.0, seed=7)
    grid = amcl3d_b200.Grid(cuda_ctx, bounds)
    grid.compute(pts, 0.05, keep_dist=False)
    prob = grid.download_prob()
    cells = np.zeros((len(prob), 2), np.float32)
    cells[:, 0] = -1.0
    cells[:, 1] = prob
    del prob
    yield dict(points=pts, bounds=bounds, pose=pose, cloud=cloud, grid=grid, cells=cells, dims=grid.dims.copy())
    grid.close()


@pytest.mark.parametrize("poses", ["tracking", "uniform"])
def test_large_map_path_matches_the_oracle(cuda_ctx, port, hall, poses):
    """The configs[3] / [4] code path: bricked grid, cloud re-ordered along a Morton curve on the device, sequential
    chunk launches with sub-chunk CTAs and double accumulators, pose-sorted scheduling, 131 072 particles x 32 768
    points.  (1) per-particle computeCloudWeight against the oracle on a 1 024-particle subsample: counts exact, weights
    <= 1e-5 relative; (2) the normalisations / mean over ALL particles: the reference's loops fed with the GPU's raw
    weights must give the GPU's final particles bit for bit (exact chains)."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    n = 131072
    if poses == "tracking":
        particles = synth.particles_tracking(n, hall["pose"], (0.5, 0.5, 0.5, 0.2), seed=6)
    else:
        particles = synth.particles_uniform(n, hall["bounds"], seed=8)
    assert cuda_ctx.get_option("cloud_order") == 0 and cuda_ctx.get_option("weight_chunk_points") == 0
    pf = amcl3d_b200.Filter(cuda_ctx)
    pf.upload(particles)
    mean_g = pf.update(hall["grid"], hall["cloud"], None, ALPHA, SIGMA, ROLL, PITCH)
    got = pf.download()
    raw_w, raw_n = pf.last_cloud_weights()
    launches = cuda_ctx.launch_count()
    pf.update(hall["grid"], hall["cloud"], None, ALPHA, SIGMA, ROLL, PITCH)
    launches = cuda_ctx.launch_count() - launches
    pf.close()
    assert launches >= 4, launches       # several sequential chunk launches: this IS the large-map path

    pick = np.arange(0, n, n // 1024)
    w_o, n_o = port.cloud_weight_batch(hall["cells"], hall["dims"], hall["bounds"], hall["cloud"], particles[pick, :4],
                                       ROLL, PITCH)
    inmap = np.array([port.is_into_map(hall["bounds"], *particles[i, :3]) for i in pick])
    assert inmap.sum() > 900
    assert np.array_equal(raw_n[pick][inmap], n_o[inmap])
    r = _rel(raw_w[pick][inmap], w_o[inmap])
    assert r.max() <= 1e-5, float(r.max())
    assert (n_o[inmap] > 10).mean() > 0.5       # the comparison is not vacuous: most particles see the map

    q = particles.copy()
    q[:, 5] = raw_w
    q[:, 6] = 0.0
    want, mean_o = port.update_from_weights(q, hall["bounds"], ALPHA)
    assert np.array_equal(bits(got[:, 4]), bits(want[:, 4]))                       # w
    assert np.array_equal(bits(got[:, 5]), bits(want[:, 5]))                       # wp (normalised)
    assert np.array_equal(bits(mean_g), bits(mean_o))


def test_caller_order_mode_is_bit_exact_on_the_large_map(cuda_ctx, port, hall):
    """weight_point_splits = 1 + cloud_order = 1 (AMCL3D_EXACT=1 in the drop-in classes): one float chain per particle in
    the caller's cloud order, carried through the sequential chunk launches -- Grid3d.cpp:191 bit for bit."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    n = 16384
    particles = synth.particles_tracking(n, hall["pose"], (0.5, 0.5, 0.5, 0.2), seed=16)
    cuda_ctx.set_option("weight_point_splits", 1)
    cuda_ctx.set_option("cloud_order", 1)
    cuda_ctx.set_option("weight_chunk_points", 4096)
    pf = amcl3d_b200.Filter(cuda_ctx)
    pf.upload(particles)
    pf.update(hall["grid"], hall["cloud"], None, ALPHA, SIGMA, ROLL, PITCH)
    raw_w, raw_n = pf.last_cloud_weights()
    pf.close()
    for k in ("weight_point_splits", "cloud_order", "weight_chunk_points"):
        cuda_ctx.set_option(k, 0)
    pick = np.arange(0, n, 16)
    w_o, n_o = port.cloud_weight_batch(hall["cells"], hall["dims"], hall["bounds"], hall["cloud"], particles[pick, :4],
                                       ROLL, PITCH)
    assert np.array_equal(raw_n[pick], n_o) and np.array_equal(bits(raw_w[pick]), bits(w_o))
