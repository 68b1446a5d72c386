"""GPU parity: ParticleFilter predict / update / resample / init through the C-ABI vs the CPU oracle."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def grid_S(cuda_ctx, cfg1, cfg1_cells):
    import amcl3d_b200
    cells, dims = cfg1_cells
    g = amcl3d_b200.Grid(cuda_ctx, cfg1["bounds"])
    g.upload_cells(cells, cfg1["sensor_dev"])
    yield g
    g.close()


@pytest.fixture(params=["one_cta", "segmented"])
def exact(cuda_ctx, request):
    """Both exact implementations must return the reference's bits: the single-CTA chains (sum_mode 1 / resample_mode 1)
    and the segmented chains of filter_exact.cu (3 / 3: what runs by default from 2049 particles on and on any sharded
    set)."""
    mode = 1 if request.param == "one_cta" else 3
    cuda_ctx.set_option("weight_point_splits", 1)
    cuda_ctx.set_option("sum_mode", mode)
    cuda_ctx.set_option("resample_mode", mode)
    yield cuda_ctx
    cuda_ctx.set_option("weight_point_splits", 0)
    cuda_ctx.set_option("sum_mode", 0)
    cuda_ctx.set_option("resample_mode", 0)


def assert_mean(f, mean_g, mean_o, atol=1e-5):
    """Mean components flagged exact are the reference's float chain bit for bit; a component that hovers around zero
    comes back as the fp64 sum (amcl3d_cuda_pf_mean_exact_mask) and agrees far inside the 1e-4 m tolerance."""
    mask = f.mean_exact_mask()
    for k in range(4):
        if (mask >> k) & 1:
            assert bits(mean_g[k:k + 1])[0] == bits(mean_o[k:k + 1])[0], (k, mean_g, mean_o)
        else:
            assert abs(float(mean_g[k]) - float(mean_o[k])) <= atol, (k, mean_g, mean_o)
    return mask


def new_filter(ctx, particles):
    import amcl3d_b200
    f = amcl3d_b200.Filter(ctx)
    f.upload(particles)
    return f


def test_particles_round_trip(cuda_ctx, cfg1):
    f = new_filter(cuda_ctx, cfg1["particles"])
    assert f.size() == 600
    assert np.array_equal(bits(f.download()), bits(cfg1["particles"]))
    f.close()


def test_full_cycle_bit_exact_vs_committed_reference(exact, grid_S, cfg1, ref_cfg1):
    """predict -> update -> resample on cfg1 with the reference's mt19937 draws injected: every particle field
    and the mean equal the unmodified reference's output bit for bit."""
    f = new_filter(exact, cfg1["particles"])
    f.predict(cfg1["odom_mods"], cfg1["deltas"], noise_n4=ref_cfg1["predict_noise"])
    assert np.array_equal(bits(f.download()), bits(ref_cfg1["after_predict"]))
    mean = f.update(grid_S, cfg1["cloud"], cfg1["ranges"], cfg1["alpha"], cfg1["sigma_range"], cfg1["roll"], cfg1["pitch"])
    got = f.download()
    assert np.array_equal(bits(got[:, :4]), bits(ref_cfg1["after_update"][:, :4]))
    assert np.array_equal(bits(got[:, 5]), bits(ref_cfg1["after_update"][:, 5]))   # wp: bit-exact
    np.testing.assert_allclose(got[:, 6], ref_cfg1["after_update"][:, 6], rtol=1e-6, atol=0)   # wr: device exp()
    np.testing.assert_allclose(got[:, 4], ref_cfg1["after_update"][:, 4], rtol=1e-6, atol=0)
    np.testing.assert_allclose(mean, ref_cfg1["mean_after_update"], atol=1e-6)
    idx = f.resample(ref_cfg1["resample_u01"], want_idx=True)
    after = f.download()
    assert np.array_equal(bits(after[:, :4]), bits(ref_cfg1["after_resample"][:, :4]))  # resample indices bit-exact
    assert np.all(after[:, 4] == np.float32(1.0) / np.float32(600))
    assert np.all(np.diff(idx.astype(np.int64)) >= 0)
    f.close()


def test_update_exact_vs_port_no_beacons(exact, grid_S, port, cfg1, cfg1_cells):
    """Without the range term nothing transcendental is evaluated per particle: everything is bit-exact."""
    cells, dims = cfg1_cells
    p0 = cfg1["particles"].copy()
    p0[7, 0] = 55.0      # out of the map
    p0[8, 2] = -0.5      # below the floor
    p0[:, 5] = 0.125     # stale wp / wr values survive for skipped particles (ParticleFilter.cpp:140-141)
    p0[:, 6] = 0.25
    want, mean_o = port.update(p0, cells, dims, cfg1["bounds"], cfg1["cloud"], np.zeros((0, 4)), 0.5, 0.53, 0.01, -0.02)
    f = new_filter(exact, p0)
    mean_g = f.update(grid_S, cfg1["cloud"], None, 0.5, 0.53, 0.01, -0.02)
    got = f.download()
    assert np.array_equal(bits(got), bits(want))
    mask = assert_mean(f, mean_g, mean_o)
    assert mask & 0b1100 == 0b1100          # z (2.5 m) and yaw (0.3) do not hover: exact in every mode
    if exact.get_option("sum_mode") == 1:
        assert mask == 0xF                  # the one-CTA kernel chains everything serially
    assert got[7, 4] == 0 and got[8, 4] == 0
    f.close()


def test_update_fast_mode_within_tolerance(cuda_ctx, grid_S, port, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    want, mean_o = port.update(cfg1["particles"], cells, dims, cfg1["bounds"], cfg1["cloud"], cfg1["ranges"], 0.5, 0.53,
                               0.01, -0.02)
    cuda_ctx.set_option("sum_mode", 2)
    f = new_filter(cuda_ctx, cfg1["particles"])
    mean_g = f.update(grid_S, cfg1["cloud"], cfg1["ranges"], 0.5, 0.53, 0.01, -0.02)
    cuda_ctx.set_option("sum_mode", 0)
    got = f.download()
    np.testing.assert_allclose(got[:, 4:], want[:, 4:], rtol=1e-5, atol=1e-12)   # weights: 1e-5 relative
    np.testing.assert_allclose(mean_g[:3], mean_o[:3], atol=1e-4)               # mean pose: 1e-4 m
    np.testing.assert_allclose(mean_g[3], mean_o[3], atol=1e-4)
    f.close()


def test_update_all_outside(exact, grid_S, cfg1):
    p = cfg1["particles"][:100].copy()
    p[:, 1] -= 500.0
    f = new_filter(exact, p)
    mean = f.update(grid_S, cfg1["cloud"], cfg1["ranges"], 0.5, 0.53, 0, 0)
    assert np.all(mean == 0) and np.all(f.download()[:, 4] == 0)   # ParticleFilter.cpp:185-188
    assert f.last_in_map_evals() == 0
    f.close()


def test_in_map_eval_count(exact, grid_S, port, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    f = new_filter(exact, cfg1["particles"][:40])
    f.update(grid_S, cfg1["cloud"], None, 0.5, 0.53, 0, 0)
    total = 0
    for p in cfg1["particles"][:40]:
        total += port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], (p[0], p[1], p[2], 0, 0, p[3]))[1]
    assert f.last_in_map_evals() == total
    f.close()


@pytest.mark.parametrize("n", [1, 2, 31, 600, 4097, 50000])
def test_resample_exact_chain_bit_exact(exact, port, n):
    rng = np.random.default_rng(n)
    p = np.zeros((n, 7), np.float32)
    p[:, :4] = rng.normal(0, 1, (n, 4)).astype(np.float32)
    w = rng.gamma(0.3, 1.0, n)
    w[rng.uniform(size=n) < 0.2] = 0.0          # zero-weight (out-of-map) particles
    p[:, 4] = (w / max(w.sum(), 1e-30)).astype(np.float32)
    p[:, 5:] = rng.uniform(0, 1, (n, 2)).astype(np.float32)
    for u01 in (0.0, 0.37, 0.99999994):
        want, idx_o = port.resample(p, u01)
        f = new_filter(exact, p)
        idx_g = f.resample(u01, want_idx=True)
        assert np.array_equal(idx_g, idx_o)
        assert np.array_equal(bits(f.download()), bits(want))
        f.close()


def test_resample_runoff_clamps(exact, port):
    p = np.zeros((64, 7), np.float32)
    p[:, 0] = np.arange(64)
    p[:, 4] = 0.01           # chain ends at 0.64 < u for the last slots
    want, idx_o = port.resample(p, 0.5)
    f = new_filter(exact, p)
    idx_g = f.resample(0.5, want_idx=True)
    assert np.array_equal(idx_g, idx_o) and idx_g[-1] == 63
    f.close()


def test_default_resample_is_the_reference_walk_at_200k(cuda_ctx, port):
    """No option set: the default resample (segmented exact chain) returns the reference's indices bit for bit."""
    n = 200000
    rng = np.random.default_rng(3)
    p = np.zeros((n, 7), np.float32)
    p[:, 0] = np.arange(n) % 1000
    w = rng.gamma(0.5, 1.0, n)
    p[:, 4] = (w / w.sum()).astype(np.float32)
    p[:, 5:] = rng.uniform(0, 1, (n, 2)).astype(np.float32)
    want, idx_o = port.resample(p, 0.25)
    f = new_filter(cuda_ctx, p)
    idx = f.resample(0.25, want_idx=True)
    assert np.array_equal(idx, idx_o)
    assert np.array_equal(bits(f.download()), bits(want))
    # a second resample on the resampled set (uniform weights 1/n: systematic rounding in the chain)
    want2, idx_o2 = port.resample(want, 0.9)
    idx2 = f.resample(0.9, want_idx=True)
    assert np.array_equal(idx2, idx_o2)
    assert np.array_equal(bits(f.download()), bits(want2))
    f.close()


def test_predict_injected_noise_bit_exact(cuda_ctx, port, reference, cfg1):
    rng = reference.rng(77)
    for deltas in [cfg1["deltas"], (0.0, 0.0, 0.0, 0.0), (1.5, -2.0, 0.3, -0.7)]:
        noise = rng.predict_noise(600, cfg1["odom_mods"], deltas)
        want = port.predict(cfg1["particles"], cfg1["odom_mods"], deltas, noise)
        f = new_filter(cuda_ctx, cfg1["particles"])
        f.predict(cfg1["odom_mods"], deltas, noise_n4=noise)
        assert np.array_equal(bits(f.download()), bits(want))
        f.close()


def test_predict_philox_statistics(cuda_ctx):
    n = 400000
    p = np.zeros((n, 7), np.float32)
    f = new_filter(cuda_ctx, p)
    mods, deltas = (0.5, 0.5, 0.5, 0.5), (2.0, -1.0, 0.4, 0.2)
    f.predict(mods, deltas, seed=42, step=7)
    a = f.download()
    for k in range(4):
        sd = abs(deltas[k] * mods[k])
        assert abs(a[:, k].mean() - deltas[k]) < 5 * sd / np.sqrt(n) + 1e-6     # yaw 0: x,y unrotated
        assert abs(a[:, k].std() - sd) < 0.01 * sd
    # counter-based: same (seed, step) -> same draws; another step -> different draws
    g = new_filter(cuda_ctx, p)
    g.predict(mods, deltas, seed=42, step=7)
    assert np.array_equal(bits(g.download()), bits(a))
    h = new_filter(cuda_ctx, p)
    h.predict(mods, deltas, seed=42, step=8)
    assert not np.array_equal(bits(h.download()), bits(a))
    # normality: excess kurtosis ~ 0, channels uncorrelated
    z = (a[:, 0] - a[:, 0].mean()) / a[:, 0].std()
    assert abs((z ** 4).mean() - 3.0) < 0.05
    assert abs(np.corrcoef(a[:, 0], a[:, 1])[0, 1]) < 0.01
    for x in (f, g, h):
        x.close()


def test_init_injected_noise_vs_reference(cuda_ctx, ref_cfg1):
    import amcl3d_b200
    f = amcl3d_b200.Filter(cuda_ctx)
    mean = f.init(600, (0.0, 0.0, 2.5, 0.3), (0.05, 0.05, 0.05, 0.1), noise_n4=ref_cfg1["init_noise"])
    got = f.download()
    assert np.array_equal(bits(got[:, :4]), bits(ref_cfg1["init_particles"][:, :4]))
    np.testing.assert_allclose(got[:, 4], ref_cfg1["init_particles"][:, 4], rtol=1e-6)
    np.testing.assert_allclose(mean, ref_cfg1["init_mean"], atol=1e-6)
    f.close()


def test_empty_filter_is_harmless(cuda_ctx, grid_S, cfg1):
    import amcl3d_b200
    f = amcl3d_b200.Filter(cuda_ctx)
    assert f.size() == 0
    f.predict(cfg1["odom_mods"], cfg1["deltas"])
    assert np.all(f.update(grid_S, cfg1["cloud"], None, 0.5, 0.53, 0, 0) == 0)
    f.resample(0.5)
    f.close()


def _weights(kind, n, rng):
    if kind == "gamma":
        w = rng.gamma(0.4, 1.0, n)
        return (w / w.sum()).astype(np.float32)
    if kind == "ties":          # multiples of 2^-22 with odd factors: half-ulp ties as the chain grows
        return (rng.choice([1, 3, 5, 7], n) * 2.0 ** -22).astype(np.float32)
    if kind == "uniform":       # every particle 1/n, the state right after a resample
        return np.full(n, np.float32(1.0) / np.float32(n), np.float32)
    if kind == "range":         # 60 orders of magnitude, zeros, one dominant particle
        w = (10.0 ** rng.uniform(-38, -1, n)).astype(np.float32)
        w[rng.uniform(size=n) < 0.3] = 0
        w[n // 3] = 0.5
        return w
    if kind == "denormal":
        return (rng.uniform(0, 1, n) * 1e-40).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["gamma", "ties", "uniform", "range", "denormal"])
@pytest.mark.parametrize("n", [1, 95, 96, 97, 4192, 5000, 100003, 1048576])
def test_windowed_exact_chain_equals_sequential_sum(cuda_ctx, port, kind, n):
    """exact_scan.cuh: the parallel integer-domain scan reproduces the reference's sequential float chain
    (ParticleFilter.cpp:203-214) bit for bit -- resample indices at 1 M particles included."""
    rng = np.random.default_rng(hash((kind, n)) % (2 ** 32))
    p = np.zeros((n, 7), np.float32)
    p[:, 0] = np.arange(n, dtype=np.float32)
    p[:, 4] = _weights(kind, n, rng)
    want, idx_o = port.resample(p, 0.41)
    got = {}
    for name, mode, serial in (("windowed", 1, 0), ("serial", 1, 1), ("segmented", 3, 0)):
        if serial and n > 200000:
            continue        # the single-lane chain is only the small-n cross-check
        cuda_ctx.set_option("resample_mode", mode)
        cuda_ctx.set_option("serial_chain", serial)
        f = new_filter(cuda_ctx, p)
        got[name] = (f.resample(0.41, want_idx=True), f.download())
        f.close()
    cuda_ctx.set_option("serial_chain", 0)
    cuda_ctx.set_option("resample_mode", 0)
    for name, (idx, after) in got.items():
        assert np.array_equal(idx, idx_o), (kind, n, name, int(np.count_nonzero(idx != idx_o)))
        assert np.array_equal(bits(after), bits(want)), (kind, n, name)


def test_windowed_exact_chain_with_hostile_terms(cuda_ctx, port):
    """Negative, infinite and NaN weights are not meaningful, but the chain must still be the sequential one."""
    rng = np.random.default_rng(5)
    n = 3000
    p = np.zeros((n, 7), np.float32)
    p[:, 0] = np.arange(n, dtype=np.float32)
    w = (rng.gamma(0.5, 1.0, n) / n).astype(np.float32)
    w[500] = -0.01
    w[1500] = -w[:1500].sum() * 2     # drives the running value negative for a while
    w[2000] = 1.0
    p[:, 4] = w
    want, idx_o = port.resample(p, 0.2)
    for mode in (1, 3):
        cuda_ctx.set_option("resample_mode", mode)
        f = new_filter(cuda_ctx, p)
        idx = f.resample(0.2, want_idx=True)
        f.close()
        cuda_ctx.set_option("resample_mode", 0)
        # With a non-monotone chain the reference's forward walk and a binary search need not agree, so only the
        # contract is checked here: the call terminates and every pick is a valid particle index.
        assert idx.min() >= 0 and idx.max() < n and len(idx_o) == n


@pytest.mark.parametrize("mode", [1, 3])
def test_update_exact_weights_at_20k_particles(cuda_ctx, grid_S, port, cfg1, cfg1_cells, mode):
    """sum_mode 1 / 3 at a particle count far beyond the reference's operating point: the windowed exact scan keeps
    wtp / wtr / wt -- hence every normalised weight -- bit-identical to the reference's sequential float sums."""
    from amcl3d_b200 import synth
    cells, dims = cfg1_cells
    n = 20000
    particles = synth.particles_tracking(n, cfg1["pose"], (0.3, 0.3, 0.3, 0.5), seed=21)
    particles[::97, 0] += 40.0       # some particles outside the map
    cloud = cfg1["cloud"][:257]
    want, mean_o = port.update(particles, cells, dims, cfg1["bounds"], cloud, np.zeros((0, 4)), 0.5, 0.53, 0.01, -0.02)
    cuda_ctx.set_option("weight_point_splits", 1)
    cuda_ctx.set_option("sum_mode", mode)
    f = new_filter(cuda_ctx, particles)
    mean_g = f.update(grid_S, cloud, None, 0.5, 0.53, 0.01, -0.02)
    got = f.download()
    assert_mean(f, mean_g, mean_o)                               # exact where the sum does not hover around zero
    f.close()
    cuda_ctx.set_option("weight_point_splits", 0)
    cuda_ctx.set_option("sum_mode", 0)
    assert np.array_equal(bits(got), bits(want))                 # x,y,z,a,w,wp,wr: all bit-exact (no beacons, no exp)
    # and the resample that follows picks the same particles as the reference
    cuda_ctx.set_option("resample_mode", mode)
    f = new_filter(cuda_ctx, got)
    idx = f.resample(0.77, want_idx=True)
    f.close()
    cuda_ctx.set_option("resample_mode", 0)
    _, idx_o = port.resample(want, 0.77)
    assert np.array_equal(idx, idx_o)


@pytest.mark.parametrize("n,beacons", [(2049, True), (70000, True), (1048576, False), (1048576, True)])
def test_segmented_exact_update_is_the_reference_at_any_size(cuda_ctx, grid_S, port, cfg1, cfg1_cells, n, beacons):
    """The default update above 2048 particles (sum_mode 3, filter_exact.cu): wtp / wtr / wt, every normalised weight and
    the mean are the reference's sequential float chains bit for bit -- at 10^6 particles too, where those chains differ
    from an fp64 sum by ~1e-5 relative (weights) and ~1e-3 m (mean).  With beacons the per-particle wr goes through the
    device's exp(double), which may differ from glibc in the last bit of a few values: tolerance 5e-6 there."""
    from amcl3d_b200 import synth
    cells, dims = cfg1_cells
    particles = synth.particles_tracking(n, (3.0, -4.0, 2.5, 0.3), (0.4, 0.4, 0.3, 0.5), seed=5)
    particles[::1013, 1] += 60.0      # some particles outside the map
    particles[:, 5] = 0.125           # stale wp / wr survive for them (ParticleFilter.cpp:140-141)
    particles[:, 6] = 0.25
    cloud = cfg1["cloud"][:64]
    ranges = cfg1["ranges"] if beacons else np.zeros((0, 4), np.float32)
    want, mean_o = port.update(particles, cells, dims, cfg1["bounds"], cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    cuda_ctx.set_option("weight_point_splits", 1)
    f = new_filter(cuda_ctx, particles)
    mean_g = f.update(grid_S, cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    got = f.download()
    raw_w, raw_n = f.last_cloud_weights()
    mask = f.mean_exact_mask()
    f.close()
    assert mask == 0xF                # pose (3, -4, 2.5, 0.3): no component hovers around zero
    cuda_ctx.set_option("weight_point_splits", 0)
    # the weighting step alone, on a subsample (one split, caller's cloud order: bit-exact)
    pick = np.arange(0, n, max(1, n // 3000))
    w_o, n_o = port.cloud_weight_batch(cells, dims, cfg1["bounds"], cloud, particles[pick, :4], 0.01, -0.02)
    inmap = np.array([port.is_into_map(cfg1["bounds"], *particles[i, :3]) for i in pick])
    assert np.array_equal(raw_n[pick][inmap], n_o[inmap])
    assert np.array_equal(bits(raw_w[pick][inmap]), bits(w_o[inmap]))
    assert np.all(raw_n[pick][~inmap] == 0)
    if not beacons:
        assert np.array_equal(bits(got), bits(want))
        assert np.array_equal(bits(mean_g), bits(mean_o))
    else:
        np.testing.assert_allclose(got[:, 4:], want[:, 4:], rtol=5e-6, atol=1e-30)
        np.testing.assert_allclose(mean_g, mean_o, atol=5e-6)
        assert np.array_equal(bits(got[:, 5]), bits(want[:, 5]))       # wp does not depend on exp(): bit-exact


def test_long_hovering_mean_chains_are_evaluated_exactly(cuda_ctx, grid_S, port, cfg1, cfg1_cells):
    """Global-localisation shape: 2 M particles spread over the whole map, so the x / y / yaw mean sums hover around zero
    (|sum| << sum |term|).  Short hovering chains are returned as fp64 sums; at this length the float chain's own drift
    is no longer negligible against 1e-4 m, so components whose bound N * 2^-24 * |mean| exceeds 1e-3 m are evaluated
    exactly by the joint single-lane chain: the reference's bits."""
    cells, dims = cfg1_cells
    n = 2_000_000
    rng = np.random.default_rng(9)
    particles = np.zeros((n, 7), np.float32)
    particles[:, 0] = rng.uniform(-9.5, 9.5, n) + 0.06
    particles[:, 1] = rng.uniform(-9.5, 9.5, n)
    particles[:, 2] = rng.uniform(0.3, 4.7, n)
    particles[:, 3] = rng.uniform(-3.0, 3.0, n) + 0.03
    particles[:, 4] = 1.0 / n
    cloud = cfg1["cloud"][:16]
    ranges = np.zeros((0, 4), np.float32)
    want, mean_o = port.update(particles, cells, dims, cfg1["bounds"], cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    cuda_ctx.set_option("weight_point_splits", 1)
    f = new_filter(cuda_ctx, particles)
    mean_g = f.update(grid_S, cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    got = f.download()
    mask = f.mean_exact_mask()
    f.close()
    cuda_ctx.set_option("weight_point_splits", 0)
    assert np.array_equal(bits(got), bits(want))
    assert mask & 0b0100                                 # z: an ordinary chain
    assert mask & 0b1001, (mask, mean_g, mean_o)          # x and / or yaw: long hovering chains, evaluated exactly
    for k in range(4):
        if (mask >> k) & 1:
            assert bits(mean_g[k:k + 1])[0] == bits(mean_o[k:k + 1])[0], (k, mask, mean_g, mean_o)
        else:
            assert abs(float(mean_g[k]) - float(mean_o[k])) <= 2e-5, (k, mask, mean_g, mean_o)
