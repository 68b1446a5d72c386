"""GPU: the drop-in check.  The SAME harness source (tests/harness/class_harness.cpp) is compiled against the
unmodified reference classes and against this repo's host classes; the two libraries are driven with identical
call sequences, identical mt19937 seeds, and their outputs compared."""
import os

import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from oracle.bindings import HostBuild
    h = HostBuild()
    assert h.impl == "amcl3d_b200"
    return h


@pytest.fixture()
def exact(host):
    for k, v in (("weight_point_splits", 1), ("sum_mode", 1), ("resample_mode", 1)):
        assert host.set_option(k, v) == 0
    yield host
    for k in ("weight_point_splits", "sum_mode", "resample_mode"):
        host.set_option(k, 0)


def test_grid3d_before_open(host, reference, cfg1):
    for h in (host, reference):
        g = h.grid()
        g.set_cloud(cfg1["cloud"])
        assert float(g.cloud_weight(0, 0, 2.5, 0, 0, 0.3)) == 0.0        # Grid3dTest.cpp:160-161
        assert not g.is_into_map(1, 1, 1)
        assert g.slice(0.0)[0] is None                                    # Grid3dTest.cpp:70-71
        assert g.map_cloud() is None                                      # Grid3dTest.cpp:108-109
        assert g.dims() is None


def test_grid3d_compute_and_queries(host, reference, cfg1):
    gh, gr = host.grid(), reference.grid()
    assert gh.open_from_cloud(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"])
    assert gr.open_from_cloud(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"])
    assert list(gh.dims()) == list(gr.dims())
    ch, cr = gh.cells(), gr.cells()
    assert np.array_equal(bits(ch[:, 0]), bits(cr[:, 0]))
    np.testing.assert_allclose(ch[:, 1], cr[:, 1], rtol=1e-5, atol=1e-36)
    assert np.array_equal(gh.map_cloud(), gr.map_cloud())
    nh, bh = gh.map_info()
    nr, br = gr.map_info()
    assert nh == nr and np.array_equal(bh, br)
    for xyz in [(1, 1, 1), (-100, -100, -100), (-10, -10, 0), (10, 10, 5), (9.99, -9.99, 4.99)]:
        assert gh.is_into_map(*xyz) == gr.is_into_map(*xyz)
    for z in (0.0, 1.0, 2.55, 4.99, 5.0):
        sh, ih = gh.slice(z)
        sr, ir = gr.slice(z)
        assert len(sh) == len(sr) and np.array_equal(ih, ir)              # same (quirky) payload length
        # the payload is 200 entries longer than a layer; on the top layer(s) the reference reads past the end of
        # its vector there (undefined), this build returns 0 -- compare the part that lies inside the grid
        first = int(np.float32(z) / 0.1) * 40000
        inside = max(0, min(len(sh), 2000000 - first))
        if inside == len(sh):     # (when the reference's scan runs off the end even its scale factor is undefined)
            assert np.abs(sh.astype(int) - sr.astype(int)).max() <= 1
        assert np.all(sh[inside:] == 0)
    assert gh.slice(-100.0)[0] is None and gh.slice(5.01)[0] is None


def test_grid3d_cloud_weight_bit_exact_on_same_cells(host, reference, cfg1, cfg1_cells):
    cells, dims = cfg1_cells
    gh, gr = host.grid(), reference.grid()
    assert gh.set_cells(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"], dims, cells)
    assert gr.set_cells(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"], dims, cells)
    gh.set_cloud(cfg1["cloud"])
    gr.set_cloud(cfg1["cloud"])
    for p in cfg1["particles"][:48]:
        a = gh.cloud_weight(p[0], p[1], p[2], 0.01, -0.02, p[3])
        b = gr.cloud_weight(p[0], p[1], p[2], 0.01, -0.02, p[3])
        assert bits(a) == bits(b)


@pytest.mark.parametrize("options", ["default", "one_cta_chains"])
def test_particle_filter_cycles_vs_reference(request, host, reference, cfg1, cfg1_cells, options):
    """Three predict -> update -> resample cycles with both filters seeded identically.  `default`: every library option
    at its default (reference summation order, exact sums) -- what a drop-in user gets; `one_cta_chains`: the one-CTA
    chain kernels forced (the cross-check implementation)."""
    cells, dims = cfg1_cells
    out = {}
    exact = request.getfixturevalue("exact") if options == "one_cta_chains" else host
    for name, h in (("host", exact), ("ref", reference)):
        g = h.grid()
        assert g.set_cells(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"], dims, cells)
        g.set_cloud(cfg1["cloud"])
        f = h.filter()
        assert not f.is_initialized()
        f.seed(4321)
        f.init(600, (0.0, 0.0, 2.5, 0.3), (0.05, 0.05, 0.05, 0.1))
        assert f.is_initialized() and f.size() == 600
        trace = [f.particles(), f.mean()]
        for _ in range(3):
            last_mean = trace[-1] if len(trace) == 2 else trace[-2]
            f.predict(cfg1["odom_mods"], cfg1["deltas"])
            assert np.array_equal(bits(f.mean()), bits(last_mean))   # predict leaves mean_ untouched (ParticleFilterTest.cpp:60-71)
            f.update(g, cfg1["ranges"], cfg1["alpha"], cfg1["sigma_range"], cfg1["roll"], cfg1["pitch"])
            trace += [f.particles(), f.mean()]
            f.resample()
            trace += [f.particles()]
        trace.append(f.pose_msg())
        out[name] = trace
    for a, b in zip(out["host"], out["ref"]):
        a, b = np.asarray(a), np.asarray(b)
        if a.ndim == 2 and a.shape[1] == 7 and a.dtype == np.float32:
            assert np.array_equal(bits(a[:, :4]), bits(b[:, :4]))          # poses and resample picks: bit-exact
            np.testing.assert_allclose(a[:, 4:], b[:, 4:], rtol=2e-6, atol=1e-12)
        else:
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-6)


def test_predict_keeps_mean(host, cfg1):
    # tests/ParticleFilterTest.cpp:39-72
    f = host.filter()
    f.init(600, (0.0, 0.0, 2.5, 0.3), (0.05, 0.05, 0.05, 0.1))
    before = f.mean()
    f.predict((0.1, 0.1, 0.1, 0.3), (-0.067421, -0.006161, 0.130909, 0.052421))
    np.testing.assert_allclose(f.mean(), before, atol=1e-4)


def test_update_with_unopened_grid(host, reference, cfg1):
    for h in (host, reference):
        g = h.grid()
        g.set_cloud(cfg1["cloud"])
        f = h.filter()
        f.set_particles(cfg1["particles"][:32])
        f.update(g, cfg1["ranges"], 0.5, 0.53, 0, 0)
        assert np.all(f.particles()[:, 4] == 0) and np.all(f.mean()[:4] == 0)


def test_octomap_files_and_grid_cache(host, reference, tmp_path):
    """Grid3d::open on real files: .bt and .ot, pruned coarse leaves, free leaves widening the bounds, the .grid
    cache written by one build and read back by the other (byte-compatible format)."""
    rng = np.random.default_rng(21)
    res = 0.1
    occ = np.unique(rng.integers(-12, 12, (300, 3)), axis=0)
    pts = ((occ + 0.5) * res).astype(np.float32)
    depths = np.full(len(pts), 16, np.uint8)
    depths[:6] = 15                                    # a few 2x2x2 leaves -> centres on integer lattice coordinates
    free = np.array([[-2.05, -2.05, -1.55], [2.05, 2.05, 1.55]], np.float32)
    for ext, as_ot in ((".bt", False), (".ot", True)):
        path = str(tmp_path / ("map" + ext))
        assert host.write_octomap(path, pts, res, depths=depths, free_points=free, as_ot=as_ot)
        ph, bh = host.load_octomap(path)
        pr, br = reference.load_octomap(path)
        assert np.array_equal(ph, pr) and np.array_equal(bh, br)
        assert abs(bh[6] - res) < 1e-12 and bh[0] <= -2.0 and bh[3] >= 2.0
        gr = reference.grid()
        assert gr.open(path, 0.05)                     # reference computes and writes map.grid
        cache = str(tmp_path / "map.grid")
        assert os.path.exists(cache)
        gh = host.grid()
        assert gh.open(path, 0.05)                     # host build loads the reference's cache
        assert np.array_equal(bits(gh.cells()), bits(gr.cells()))
        os.remove(cache)
        gh2 = host.grid()
        assert gh2.open(path, 0.05)                    # host build computes on the GPU and writes its own cache
        assert os.path.exists(cache)
        c2 = gh2.cells()
        assert np.array_equal(bits(c2[:, 0]), bits(gr.cells()[:, 0]))
        np.testing.assert_allclose(c2[:, 1], gr.cells()[:, 1], rtol=1e-5, atol=1e-36)
        gr2 = reference.grid()
        assert gr2.open(path, 0.05)                    # ... which the reference accepts
        assert np.array_equal(bits(gr2.cells()), bits(c2))
        assert host.grid().open(path, 0.06)            # different sensor_dev: cache rejected, grid recomputed
        os.remove(cache)


def test_open_failures_match_reference(host, reference, tmp_path):
    # tests/Grid3dTest.cpp:15-43, tests/PointCloudToolsTest.cpp:61-103
    empty_bt, empty_ot, unk = tmp_path / "mapfile_null.bt", tmp_path / "mapfile_null.ot", tmp_path / "m.unk"
    for p in (empty_bt, empty_ot, unk):
        p.write_bytes(b"")
    for h in (host, reference):
        for path in (str(tmp_path / "missing.bt"), str(empty_bt), str(empty_ot), str(unk), ".bt"):
            assert not h.grid().open(path, 0.05)
        msgs = []
        for path in (str(tmp_path / "missing.bt"), str(empty_bt), str(empty_ot), str(unk)):
            with pytest.raises(RuntimeError) as e:
                h.load_octomap(path)
            msgs.append(str(e.value).replace(str(tmp_path), ""))
        assert h.null_tree_throws()
        if h is host:
            host_msgs = msgs
    assert host_msgs == msgs


def test_node_voxel_filter_through_the_pcl_stand_in(host, port, cfg1):
    """Node.cpp:131-137 compiled verbatim against compat/pcl/filters/voxel_grid.h: the node's down-sampling runs on
    the device and hands over the cloud PCL's algorithm would (leaf order, centroids) -- bit-equal to the restatement."""
    from amcl3d_b200 import synth
    raw = synth.sensor_cloud(cfg1["map_points"], cfg1["pose"], 30000, 8.0, seed=91)
    rng = np.random.default_rng(92)
    raw = np.repeat(raw, 2, axis=0)
    raw[:, :3] += rng.normal(0, 0.02, (len(raw), 3)).astype(np.float32)
    got = host.node_voxel_filter(raw, 0.1)
    want = port.voxel_grid(raw, 0.1)
    assert got.shape == want.shape and len(got) < len(raw)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert len(host.node_voxel_filter(np.zeros((0, 4), np.float32), 0.1)) == 0
