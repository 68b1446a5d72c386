"""CPU: the C restatement of pcl::VoxelGrid (oracle_voxel_grid, Node.cpp:131-137's filter) against an independent
numpy restatement of the same published algorithm.  PCL itself is not available here: parity unpinned."""
import numpy as np
import pytest


def numpy_voxel_grid(cloud, leaf):
    p = np.ascontiguousarray(cloud, np.float32)[:, :3]
    ok = np.isfinite(p).all(1)
    src = np.nonzero(ok)[0]
    q = p[ok]
    if len(q) == 0:
        return np.zeros((0, 4), np.float32)
    inv = (np.float32(1.0) / np.asarray(leaf, np.float32)).astype(np.float32)
    lo, hi = q.min(0), q.max(0)
    d = ((hi - lo) * inv).astype(np.int64) + 1
    if int(d[0]) * int(d[1]) * int(d[2]) > 2147483647:
        return None
    min_b = np.floor(lo * inv).astype(np.int32)
    max_b = np.floor(hi * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(q * inv) - min_b.astype(np.float32)).astype(np.int32)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * (div[0] * div[1])
    order = np.lexsort((src, idx))
    idx_s, q_s = idx[order], q[order]
    starts = np.nonzero(np.r_[True, idx_s[1:] != idx_s[:-1]])[0]
    ends = np.r_[starts[1:], len(idx_s)]
    out = np.ones((len(starts), 4), np.float32)
    for k, (a, b) in enumerate(zip(starts, ends)):
        out[k, :3] = np.cumsum(q_s[a:b], axis=0, dtype=np.float32)[-1] / np.float32(b - a)
    return out


@pytest.mark.parametrize("n,leaf,spread", [(1, 0.1, 1.0), (7, 0.1, 0.05), (5000, 0.1, 3.0), (20000, 0.25, 10.0),
                                            (3000, (0.1, 0.2, 0.4), 2.0)])
def test_port_matches_numpy_restatement(port, n, leaf, spread):
    rng = np.random.default_rng(n)
    cloud = np.zeros((n, 4), np.float32)
    cloud[:, :3] = rng.normal(0, spread, (n, 3))
    leaf3 = leaf if np.ndim(leaf) else (leaf, leaf, leaf)
    got = port.voxel_grid(cloud, leaf)
    want = numpy_voxel_grid(cloud, leaf3)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert len(got) <= n and (n < 100 or len(got) < n)


def test_non_finite_points_are_dropped_and_empty_input(port):
    cloud = np.zeros((6, 4), np.float32)
    cloud[:, 0] = [0.01, 0.02, np.nan, 5.0, np.inf, 0.03]
    got = port.voxel_grid(cloud, 0.1)
    assert len(got) == 2
    np.testing.assert_allclose(got[0, 0], np.float32(np.float32(np.float32(0.01) + np.float32(0.02)) + np.float32(0.03)) / np.float32(3))
    assert got[1, 0] == 5.0
    assert len(port.voxel_grid(np.zeros((0, 4), np.float32), 0.1)) == 0
    assert len(port.voxel_grid(np.full((3, 4), np.nan, np.float32), 0.1)) == 0


def test_output_is_sorted_by_leaf_index_x_fastest(port):
    rng = np.random.default_rng(3)
    cloud = np.zeros((4000, 4), np.float32)
    cloud[:, :3] = rng.uniform(-2, 2, (4000, 3))
    got = port.voxel_grid(cloud, 0.5)
    cell = np.floor(got[:, :3] / 0.5).astype(np.int64)
    key = (cell[:, 2] * 1000 + cell[:, 1]) * 1000 + cell[:, 0]
    assert np.all(np.diff(key) > 0)


def test_leaf_too_small_is_a_pass_through(port):
    cloud = np.zeros((4, 4), np.float32)
    cloud[:, 0] = [0, 1000, 2000, 3000]
    cloud[:, 1] = [0, 1000, 2000, 3000]
    cloud[:, 2] = [0, 1000, 2000, 3000]
    assert port.voxel_grid(cloud, 0.001) is None
    assert numpy_voxel_grid(cloud, (0.001,) * 3) is None
