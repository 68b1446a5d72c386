"""CPU: the plain-C restatement against the unmodified reference compiled in oracle/_ref, bit for bit,
and against the committed outputs of that reference (tests/golden/ref_cfg1.npz)."""
import numpy as np
import pytest

from conftest import bits


def test_compute_grid_bit_exact(port, reference, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    g = reference.grid()
    assert g.open_from_cloud(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"])
    assert list(g.dims()) == list(dims)
    assert np.array_equal(bits(g.cells()), bits(cells))


def test_compute_grid_matches_committed_sample(cfg1_cells, ref_cfg1):
    cells, _ = cfg1_cells
    assert np.array_equal(bits(cells[ref_cfg1["cells_sample_idx"]]), bits(ref_cfg1["cells_sample"]))
    assert float(cells[:, 1].astype(np.float64).sum()) == float(ref_cfg1["cells_prob_sum64"])


def test_nn_bucketed_equals_bruteforce(port, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    rng = np.random.default_rng(11)
    for _ in range(40):
        ix, iy, iz = int(rng.integers(dims[0])), int(rng.integers(dims[1])), int(rng.integers(dims[2]))
        d = port.nn_dist2_bruteforce(cfg1["map_points"], cfg1["bounds"], ix, iy, iz)
        assert np.float32(d) == cells[ix + int(dims[0]) * (iy + int(dims[1]) * iz), 0]


def test_cloud_weight_bit_exact(port, reference, cfg1, cfg1_cells, ref_cfg1):
    cells, dims = cfg1_cells
    roll, pitch = np.float32(cfg1["roll"]), np.float32(cfg1["pitch"])
    mine = np.array([port.cloud_weight(cells, dims, cfg1["bounds"], cfg1["cloud"], (p[0], p[1], p[2], roll, pitch, p[3]))[0]
                     for p in cfg1["particles"][:64]], np.float32)
    assert np.array_equal(bits(mine), bits(ref_cfg1["wp_single"]))
    g = reference.grid()
    g.set_cells(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"], dims, cells)
    g.set_cloud(cfg1["cloud"])
    live = np.array([g.cloud_weight(p[0], p[1], p[2], roll, pitch, p[3]) for p in cfg1["particles"][:64]], np.float32)
    assert np.array_equal(bits(mine), bits(live))


def test_full_cycle_bit_exact(port, cfg1, cfg1_cells, ref_cfg1):
    """predict -> update -> resample with the reference's own mt19937 draws injected."""
    cells, dims = cfg1_cells
    p = port.predict(cfg1["particles"], cfg1["odom_mods"], cfg1["deltas"], ref_cfg1["predict_noise"])
    assert np.array_equal(bits(p), bits(ref_cfg1["after_predict"]))
    p, mean = port.update(p, cells, dims, cfg1["bounds"], cfg1["cloud"], cfg1["ranges"], cfg1["alpha"],
                          cfg1["sigma_range"], cfg1["roll"], cfg1["pitch"])
    assert np.array_equal(bits(p), bits(ref_cfg1["after_update"]))
    assert np.array_equal(bits(mean), bits(ref_cfg1["mean_after_update"]))
    p, idx = port.resample(p, ref_cfg1["resample_u01"])
    assert np.array_equal(bits(p), bits(ref_cfg1["after_resample"]))
    assert np.all(np.diff(idx.astype(np.int64)) >= 0)


def test_init_bit_exact(port, ref_cfg1):
    p, mean = port.init(600, (0.0, 0.0, 2.5, 0.3), (0.05, 0.05, 0.05, 0.1), ref_cfg1["init_noise"])
    assert np.array_equal(bits(p), bits(ref_cfg1["init_particles"]))
    assert np.array_equal(bits(mean), bits(ref_cfg1["init_mean"]))


def test_live_reference_cycle_matches_committed(reference, cfg1, ref_cfg1):
    """Guards the committed vectors against drift of the oracle build (same seed -> same bits)."""
    g = reference.grid()
    assert g.open_from_cloud(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"])
    g.set_cloud(cfg1["cloud"])
    f = reference.filter()
    f.seed(1234)
    f.set_particles(cfg1["particles"])
    f.predict(cfg1["odom_mods"], cfg1["deltas"])
    f.update(g, cfg1["ranges"], cfg1["alpha"], cfg1["sigma_range"], cfg1["roll"], cfg1["pitch"])
    assert np.array_equal(bits(f.particles()), bits(ref_cfg1["after_update"]))
    f.resample()
    assert np.array_equal(bits(f.particles()), bits(ref_cfg1["after_resample"]))
    sl, _ = g.slice(1.0)
    assert np.array_equal(sl, ref_cfg1["grid_slice_z1"])


def test_grid_slice_port_vs_reference(port, cfg1, cfg1_cells, ref_cfg1):
    cells, dims = cfg1_cells
    sl = port.grid_slice(cells, dims, cfg1["bounds"], 1.0)
    assert np.array_equal(sl, ref_cfg1["grid_slice_z1"][:len(sl)])
    assert port.grid_slice(cells, dims, cfg1["bounds"], -100.0) is None


def test_range_weight_no_beacons_is_zero(port):
    assert float(port.range_weight(1, 2, 3, np.zeros((0, 4), np.float32), 0.53)) == 0.0  # ParticleFilter.cpp:227-228


def test_update_all_outside_gives_zero_mean(port, cfg1, cfg1_cells):
    # ParticleFilter.cpp:185-188: wt == 0 -> all weights 0 and the mean is (0,0,0,0), not NaN
    cells, dims = cfg1_cells
    p = cfg1["particles"][:50].copy()
    p[:, 0] += 1000.0
    q, mean = port.update(p, cells, dims, cfg1["bounds"], cfg1["cloud"], cfg1["ranges"], 0.5, 0.53, 0, 0)
    assert np.all(q[:, 4] == 0) and np.all(mean == 0)


def test_resample_runoff_clamps(port):
    # cumulative chain falls short of u for the last slots (the reference reads past the end there)
    p = np.zeros((8, 7), np.float32)
    p[:, 0] = np.arange(8)
    p[:, 4] = 0.05
    q, idx = port.resample(p, 0.5)
    assert idx[-1] == 7 and np.all(np.diff(idx.astype(int)) >= 0)
    assert np.all(q[:, 4] == np.float32(1.0) / np.float32(8))
