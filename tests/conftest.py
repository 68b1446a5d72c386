import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement (oracle/amcl3d_oracle.c)."""
    from oracle.bindings import Port
    return Port()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind the class harness; skipped when neither the prebuilt library nor
    /root/reference is available."""
    from oracle import bindings
    if not os.path.exists(bindings.reference_path()) and not os.path.isdir("/root/reference/amcl3d/src"):
        pytest.skip("oracle/_ref/libamcl3d_ref.so not built and /root/reference absent")
    return bindings.Reference()


@pytest.fixture(scope="session")
def kat():
    return dict(np.load(os.path.join(GOLDEN, "kat_map_T.npz")))


@pytest.fixture(scope="session")
def ref_cfg1():
    return dict(np.load(os.path.join(GOLDEN, "ref_cfg1.npz")))


@pytest.fixture(scope="session")
def cfg1():
    from amcl3d_b200 import synth
    return synth.make_workload("cfg1")


@pytest.fixture(scope="session")
def cfg1_cells(port, cfg1):
    """Oracle grid of map S (port; bit-identical to the reference build, see test_oracle_port_vs_reference)."""
    cells, dims = port.compute_grid(cfg1["map_points"], cfg1["bounds"], cfg1["sensor_dev"])
    return cells, dims


@pytest.fixture(scope="session")
def cuda_ctx():
    import amcl3d_b200
    ctx = amcl3d_b200.Context(0)
    yield ctx
    ctx.close()


def bits(a):
    """float32 array -> uint32 view for bit-exact comparisons."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
