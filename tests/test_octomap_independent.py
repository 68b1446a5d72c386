"""CPU: the self-written octomap stand-in (amcl3d_b200/host/compat/octomap/OcTree.h), through the UNMODIFIED
openOcTree / computePointCloud of the reference (PointCloudTools.cpp:26-82, compiled in oracle/_ref), against an
independent Python restatement of the octomap file formats and coordinate conventions (tests/octomap_py.py): files written
by the Python model must load to exactly the leaf centres, order and metric bounds the model predicts.  Not a pin on a real
map file (none is available offline) -- a second, independently written implementation of the published format."""
import numpy as np
import pytest

import octomap_py as om


@pytest.fixture(scope="module", params=["reference_build", "b200_host_build"])
def loader(request, reference):
    """Both builds of the free functions: the UNMODIFIED reference TU (oracle/_ref) and this repo's host classes
    (amcl3d_b200/host/PointCloudTools.cpp; pure host code, loads without a GPU)."""
    if request.param == "reference_build":
        return reference
    from oracle.bindings import HostBuild
    return HostBuild()


def random_tree(seed, n_fine=400, n_coarse=12, n_free=30, span=200):
    rng = np.random.default_rng(seed)
    leaves, taken = [], set()

    def blocks(px, py, pz, depth):       # the depth-16 voxels a leaf covers, at 2^(16-depth) granularity markers
        s = 16 - depth
        return (px << s, py << s, pz << s, s)

    def free_of_overlap(px, py, pz, depth):
        x0, y0, z0, s = blocks(px, py, pz, depth)
        for (qx, qy, qz, qs) in taken:
            m = max(s, qs)
            if (x0 >> m, y0 >> m, z0 >> m) == (qx >> m, qy >> m, qz >> m):
                return False
        return True

    def add(px, py, pz, depth, occ):
        if free_of_overlap(px, py, pz, depth):
            taken.add(blocks(px, py, pz, depth))
            leaves.append((px, py, pz, depth, occ))

    c = om.TREE_MAX_VAL
    for _ in range(n_coarse):
        depth = int(rng.integers(12, 16))
        p = (c + rng.integers(-span, span, 3)) >> (16 - depth)
        add(int(p[0]), int(p[1]), int(p[2]), depth, True)
    for _ in range(n_fine):
        p = c + rng.integers(-span, span, 3)
        add(int(p[0]), int(p[1]), int(p[2]), 16, True)
    for _ in range(n_free):
        p = c + rng.integers(-2 * span, 2 * span, 3)
        add(int(p[0]), int(p[1]), int(p[2]), 16, False)
    return leaves


@pytest.mark.parametrize("seed,res", [(1, 0.05), (2, 0.1), (3, 0.25), (4, 0.013)])
@pytest.mark.parametrize("ext", [".bt", ".ot"])
def test_python_written_octomap_loads_as_predicted(tmp_path, loader, seed, res, ext):
    root = om.build_tree(random_tree(seed))
    path = str(tmp_path / ("map" + ext))
    (om.write_bt if ext == ".bt" else om.write_ot)(path, root, res)
    want_pts, want_bounds = om.expected_point_cloud(root, res)
    pts, bounds = loader.load_octomap(path)
    assert len(pts) == len(want_pts) > 300
    assert np.array_equal(pts[:, :3].view(np.uint32), want_pts.view(np.uint32))      # centres and ORDER, bit for bit
    assert np.array_equal(bounds, want_bounds)


def test_single_coarse_leaf_and_negative_quadrant(tmp_path, loader):
    # one occupied 8x8x8-voxel leaf (depth 13) in the all-negative octant + one free voxel far away in the positive one
    leaves = [((om.TREE_MAX_VAL - 800) >> 3, (om.TREE_MAX_VAL - 160) >> 3, (om.TREE_MAX_VAL - 8) >> 3, 13, True),
              (om.TREE_MAX_VAL + 500, om.TREE_MAX_VAL + 3, om.TREE_MAX_VAL + 77, 16, False)]
    root = om.build_tree(leaves)
    for ext, writer in ((".bt", om.write_bt), (".ot", om.write_ot)):
        path = str(tmp_path / ("one" + ext))
        writer(path, root, 0.05)
        want_pts, want_bounds = om.expected_point_cloud(root, 0.05)
        pts, bounds = loader.load_octomap(path)
        assert len(pts) == 1 and np.array_equal(pts[:, :3].view(np.uint32), want_pts.view(np.uint32))
        assert np.array_equal(bounds, want_bounds)
        assert bounds[3] > 25.0 and bounds[0] < -39.9        # the free leaf widens the bounds, the coarse leaf sets the minimum


@pytest.mark.parametrize("as_ot", [False, True])
def test_files_written_by_the_stand_in_parse_as_the_python_model_predicts(tmp_path, loader, as_ot):
    """The other direction: a file written by the stand-in's writer (the one the GPU host-class tests feed to
    Grid3d::open), parsed by the independent Python reader, predicts exactly what the reference's functions load from it."""
    rng = np.random.default_rng(11)
    pts = np.zeros((600, 4), np.float32)
    pts[:, :3] = rng.uniform(-6.0, 6.0, (600, 3))
    depths = np.where(rng.uniform(size=600) < 0.03, 14, 16).astype(np.uint8)     # a few pruned 4x4x4-voxel leaves
    free = np.zeros((25, 4), np.float32)
    free[:, :3] = rng.uniform(-9.0, 9.0, (25, 3))
    path = str(tmp_path / ("w.ot" if as_ot else "w.bt"))
    assert loader.write_octomap(path, pts, 0.1, depths=depths, free_points=free, as_ot=as_ot)
    root, res = om.read_file(path)
    assert res == 0.1
    want_pts, want_bounds = om.expected_point_cloud(root, res)
    got_pts, got_bounds = loader.load_octomap(path)
    assert len(got_pts) == len(want_pts) > 400
    assert np.array_equal(got_pts[:, :3].view(np.uint32), want_pts.view(np.uint32))
    assert np.array_equal(got_bounds, want_bounds)


def test_the_reference_test_map_rebuilt_as_a_bt_file_loads_to_its_golden_cloud_and_bounds(tmp_path, loader, kat):
    """The reference's own fixtures hold what the REAL octomap library returned for its test map: the occupied leaf centres
    (tests/data/mappointcloud_msg.bin, 179 551 points) and the metric bounds / resolution asserted in
    PointCloudToolsTest.cpp:42-53.  The map file itself is not shipped, so it is rebuilt here: every golden centre becomes
    an occupied leaf at the depth its lattice position implies (179 415 voxels + 136 pruned 2x2x2 leaves), free voxels are placed in the
    corners of the golden bounds, the tree is written as a .bt by the Python model -- and openOcTree + computePointCloud
    must return the golden cloud and the golden bounds exactly."""
    res = 0.05
    gold = kat["map_points"].astype(np.float64)
    bounds = kat["bounds"]
    assert np.array_equal(bounds, [-17.35, -9.5, -1.4, 8.75, 9.7, 6.25, 0.05])
    leaves, n_coarse = [], 0
    for depth in range(16, 11, -1):
        size = res * (1 << (16 - depth))
        j = gold / size - 0.5
        on = np.all(np.abs(j - np.rint(j)) < 1e-3, axis=1)
        if depth < 16:
            on &= ~assigned
            n_coarse += int(on.sum())
        else:
            assigned = np.zeros(len(gold), bool)
        assigned |= on
        pref = np.rint(j[on]).astype(np.int64) + (om.TREE_MAX_VAL >> (16 - depth))
        leaves += [(int(a), int(b), int(c), depth, True) for a, b, c in pref]
    assert assigned.all() and n_coarse == 136
    # free voxels in the two extreme corners: they only widen the metric bounds (calcMinMax runs over ALL leaves)
    lo = np.rint(bounds[:3] / res).astype(np.int64) + om.TREE_MAX_VAL
    hi = np.rint(bounds[3:6] / res).astype(np.int64) + om.TREE_MAX_VAL - 1
    occupied = {(a, b, c) for a, b, c, d, _ in leaves if d == 16}
    for corner in (lo, hi):
        if tuple(int(v) for v in corner) not in occupied:
            leaves.append((int(corner[0]), int(corner[1]), int(corner[2]), 16, False))
    root = om.build_tree(sorted(leaves, key=lambda l: l[3]))        # coarse leaves first: nothing may lie below them
    path = str(tmp_path / "rebuilt_test_map.bt")
    om.write_bt(path, root, res)
    pts, got_bounds = loader.load_octomap(path)
    assert len(pts) == len(gold) == 179551
    np.testing.assert_allclose(got_bounds, bounds, rtol=0, atol=1e-12)
    # the golden centres went through the legacy message's shift by -octo_min and back in float: compare on the lattice
    def lattice(p):
        return np.rint(np.asarray(p, np.float64) / (res / 2)).astype(np.int64)
    a = lattice(pts[:, :3])
    b = lattice(gold)
    order_a = np.lexsort(a.T[::-1])
    order_b = np.lexsort(b.T[::-1])
    assert np.array_equal(a[order_a], b[order_b])
    assert np.abs(pts[order_a, :3] - kat["map_points"][order_b]).max() <= 2e-6
