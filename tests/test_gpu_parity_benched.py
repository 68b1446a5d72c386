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
            # identical cells: every particle's cloud sum is the reference's own float chain (default: reference order),
            # the sums over particles are its sequential sums; only the device's exp() in wr is not bit-identical
            w_ref, _ = port.cloud_weight_batch(cells, dims, w["bounds"], w["cloud"], w["particles"][:, :4], ROLL, PITCH)
            assert np.array_equal(bits(raw_w[inmap]), bits(w_ref[inmap]))
            assert np.array_equal(bits(got[:, 5]), bits(want[:, 5]))
            assert _rel(got[:, 4], want[:, 4]).max() <= 2e-6


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
    """The configs[3] / [4] code path with every option at its default: bricked grid, sequential chunk launches in the
    caller's cloud order with carried float sums, pose-sorted scheduling, 131 072 particles x 32 768 points.
    (1) per-particle computeCloudWeight against the oracle on a 1 024-particle subsample: counts AND weights bit for
    bit; (2) the normalisations / mean over ALL particles: the reference's loops fed with the GPU's raw weights must
    give the GPU's final particles bit for bit (exact chains; a mean component that hovers around zero is returned as
    the fp64 sum and only has to agree to 1e-6 m)."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    n = 131072
    if poses == "tracking":
        particles = synth.particles_tracking(n, hall["pose"], (0.5, 0.5, 0.5, 0.2), seed=6)
    else:
        particles = synth.particles_uniform(n, hall["bounds"], seed=8)
    assert cuda_ctx.get_option("cloud_order") == 0 and cuda_ctx.get_option("weight_chunk_points") == 0
    assert cuda_ctx.get_option("reference_order") == 1
    pf = amcl3d_b200.Filter(cuda_ctx)
    pf.upload(particles)
    mean_g = pf.update(hall["grid"], hall["cloud"], None, ALPHA, SIGMA, ROLL, PITCH)
    got = pf.download()
    raw_w, raw_n = pf.last_cloud_weights()
    exact_mask = pf.mean_exact_mask()
    launches = cuda_ctx.launch_count()
    pf.update(hall["grid"], hall["cloud"], None, ALPHA, SIGMA, ROLL, PITCH)
    launches = cuda_ctx.launch_count() - launches
    pf.close()
    assert launches >= 4, launches       # several chunk launches (+ replay): this IS the large-map path

    pick = np.arange(0, n, n // 1024)
    w_o, n_o = port.cloud_weight_batch(hall["cells"], hall["dims"], hall["bounds"], hall["cloud"], particles[pick, :4],
                                       ROLL, PITCH)
    inmap = np.array([port.is_into_map(hall["bounds"], *particles[i, :3]) for i in pick])
    assert inmap.sum() > 900
    assert np.array_equal(raw_n[pick][inmap], n_o[inmap])
    assert np.array_equal(bits(raw_w[pick][inmap]), bits(w_o[inmap]))
    assert (n_o[inmap] > 10).mean() > 0.5       # the comparison is not vacuous: most particles see the map

    q = particles.copy()
    q[:, 5] = raw_w
    q[:, 6] = 0.0
    want, mean_o = port.update_from_weights(q, hall["bounds"], ALPHA)
    assert np.array_equal(bits(got[:, 4]), bits(want[:, 4]))                       # w
    assert np.array_equal(bits(got[:, 5]), bits(want[:, 5]))                       # wp (normalised)
    for k in range(4):
        if (exact_mask >> k) & 1:
            assert bits(mean_g[k:k + 1])[0] == bits(mean_o[k:k + 1])[0], (k, mean_g, mean_o)
        else:
            assert abs(float(mean_g[k]) - float(mean_o[k])) <= 1e-5, (k, mean_g, mean_o)
    if poses == "tracking":
        assert exact_mask == 0xF       # pose (-3, 1, 1.5, 0.2): no component hovers around zero


def test_fast_mode_is_the_true_sum_not_the_reference_chain(cuda_ctx, port, hall):
    """reference_order = 0 (Morton-ordered cloud, sub-chunk CTAs, double accumulators): the per-particle sums are within
    2e-6 of the EXACT (fp64) sum of the same cells -- while the reference's own sequential float chain is up to ~1e-4
    away from it on a 32 768-point cloud (values below half an ulp of the running sum are dropped).  That is why the
    default follows the reference's order: no re-associated sum can be within 1e-5 of the reference here."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    n = 131072
    particles = synth.particles_tracking(n, hall["pose"], (0.5, 0.5, 0.5, 0.2), seed=6)
    cuda_ctx.set_option("reference_order", 0)
    pf = amcl3d_b200.Filter(cuda_ctx)
    pf.upload(particles)
    pf.update(hall["grid"], hall["cloud"], None, ALPHA, SIGMA, ROLL, PITCH)
    raw_w, raw_n = pf.last_cloud_weights()
    pf.close()
    cuda_ctx.set_option("reference_order", 1)
    pick = np.arange(0, n, n // 128)
    w_ref, n_ref = port.cloud_weight_batch(hall["cells"], hall["dims"], hall["bounds"], hall["cloud"], particles[pick, :4],
                                           ROLL, PITCH)
    assert np.array_equal(raw_n[pick], n_ref)
    dev_true, dev_ref, ref_true = 0.0, 0.0, 0.0
    for k, i in enumerate(pick):
        p = particles[i]
        idx, cnt = port.cloud_indices(hall["dims"], hall["bounds"], hall["cloud"], (p[0], p[1], p[2], ROLL, PITCH, p[3]))
        if cnt <= 10:
            continue
        true = hall["cells"][idx[idx != 0xFFFFFFFF], 1].astype(np.float64).sum() / cnt
        dev_true = max(dev_true, abs(float(raw_w[i]) - true) / true)
        dev_ref = max(dev_ref, abs(float(raw_w[i]) - float(w_ref[k])) / float(w_ref[k]))
        ref_true = max(ref_true, abs(float(w_ref[k]) - true) / true)
    assert dev_true <= 2e-6, dev_true
    assert dev_ref <= 1e-3, dev_ref
    assert ref_true > 1e-5, ref_true         # the premise: the reference itself is not within 1e-5 of the exact sum


@pytest.mark.parametrize("mode", ["direct_chunks", "direct_b64", "direct_b128", "replay_morton", "ordered_fused"])
def test_reference_order_implementations_on_the_large_map(cuda_ctx, port, hall, mode):
    """The reference-order paths on the bricked map: one float chain per particle carried through sequential chunk
    launches (`direct`, any CTA width), gather + replay with a Morton-ordered cloud, and the fused gatherer / adder kernel
    (weight_ordered.cuh, forced onto the bricked layout) -- Grid3d.cpp:191 bit for bit."""
    import amcl3d_b200
    from amcl3d_b200 import synth
    n = 16384
    particles = synth.particles_tracking(n, hall["pose"], (0.5, 0.5, 0.5, 0.2), seed=16)
    opts = {"direct_chunks": {"replay": 1, "weight_chunk_points": 4096},
            "direct_b64": {"replay": 1, "weight_block_threads": 64},
            "direct_b128": {"replay": 1, "weight_block_threads": 128},
            "replay_morton": {"replay": 2},
            "ordered_fused": {"ordered_mode": 2}}[mode]
    for k, v in opts.items():
        cuda_ctx.set_option(k, v)
    pf = amcl3d_b200.Filter(cuda_ctx)
    pf.upload(particles)
    pf.update(hall["grid"], hall["cloud"], None, ALPHA, SIGMA, ROLL, PITCH)
    raw_w, raw_n = pf.last_cloud_weights()
    pf.close()
    for k in opts:
        cuda_ctx.set_option(k, 0)
    pick = np.arange(0, n, 16)
    w_o, n_o = port.cloud_weight_batch(hall["cells"], hall["dims"], hall["bounds"], hall["cloud"], particles[pick, :4],
                                       ROLL, PITCH)
    inmap = np.array([port.is_into_map(hall["bounds"], *particles[i, :3]) for i in pick])   # the reference skips the rest
    assert np.array_equal(raw_n[pick][inmap], n_o[inmap])
    assert np.array_equal(bits(raw_w[pick][inmap]), bits(w_o[inmap]))
    assert np.all(raw_n[pick][~inmap] == 0)
