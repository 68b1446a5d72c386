"""CPU: the oracle against the reference's own known-answer vectors."""
import numpy as np
import pytest

from conftest import bits


@pytest.fixture(scope="module")
def kat_grid_port(port, kat):
    """computeGrid of the test map T restricted to what the KATs need would still span most layers, so the
    port computes the full 522 x 384 x 153 grid once (OpenMP; ~10 s)."""
    cells, dims = port.compute_grid(kat["map_points"], kat["bounds"], float(kat["sensor_dev"]))
    return cells, dims


def test_grid_dims_T(port, kat):
    # tests/PointCloudToolsTest.cpp:42-53 bounds -> 522 x 384 x 153 (y is 383.99999999999994 before ceil)
    assert list(port.grid_dims(kat["bounds"])) == [522, 384, 153]


def test_grid_dims_named_configs(port):
    assert list(port.grid_dims([-10, -10, 0, 10, 10, 5, 0.1])) == [200, 200, 50]
    assert list(port.grid_dims([-50, -50, 0, 50, 50, 20, 0.05])) == [2000, 2000, 400]


@pytest.mark.slow
def test_kat1_cloud_weight_port(port, kat, kat_grid_port):
    # tests/Grid3dTest.cpp:128-132,167-168
    cells, dims = kat_grid_port
    w, n = port.cloud_weight(cells, dims, kat["bounds"], kat["sensor_cloud"], kat["kat1_pose"])
    assert abs(float(w) - float(kat["kat1_expected"])) <= float(kat["kat1_tol"])
    assert float(w) == 3.8109049797058105  # the restatement reproduces the golden to the last bit
    assert n == 934


@pytest.mark.slow
def test_kat2_nav_slice_port(port, kat, kat_grid_port):
    # nav_msg.bin: golden probability slice at legacy z = 1.0 <=> grid layer 20
    cells, dims = kat_grid_port
    z = float(kat["bounds"][2]) + float(kat["nav_origin_z"]) + 0.02
    sl = port.grid_slice(cells, dims, kat["bounds"], z)
    gold = kat["nav_slice"]
    assert len(sl) == len(gold) == int(kat["nav_width"]) * int(kat["nav_height"])
    diff = np.abs(sl.astype(np.int32) - gold.astype(np.int32))
    assert diff.max() <= 1
    assert (diff == 0).mean() > 0.999


def test_kat3_is_into_map(port, kat):
    # tests/Grid3dTest.cpp:245-266
    assert port.is_into_map(kat["bounds"], 1, 1, 1)
    assert not port.is_into_map(kat["bounds"], -100, -100, -100)


def test_particle_info_fixture_is_empty(kat):
    # computeCloudWeightParticlesTest iterates over 0 poses: it pins nothing
    assert int(kat["particle_info_count"]) == 0


@pytest.mark.slow
def test_kat1_reference_build(reference, kat):
    """The unmodified reference compiled here reproduces its own golden exactly."""
    assert reference.math_overloads_are_double()
    g = reference.grid()
    g.set_cloud(kat["sensor_cloud"])
    p = kat["kat1_pose"]
    assert float(g.cloud_weight(*p)) == 0.0  # before open (Grid3dTest.cpp:160-161)
    assert g.open_from_cloud(kat["map_points"], kat["bounds"], float(kat["sensor_dev"]))
    assert float(g.cloud_weight(*p)) == 3.8109049797058105
    assert g.is_into_map(1, 1, 1) and not g.is_into_map(-100, -100, -100)
    sl, info = g.slice(float(kat["bounds"][2]) + 1.02)
    diff = np.abs(sl.astype(np.int32) - kat["nav_slice"].astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999
    assert g.slice(-100.0)[0] is None  # Grid3dTest.cpp:70-76
