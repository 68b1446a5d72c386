"""GPU, >= 2 devices: particle-sharded predict / update / resample and z-slab computeGrid over NCCL against the
single-GPU result of the same work.  Skipped on single-GPU boxes (the CPU gloo tests cover the algebra there)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_cycle_matches_single_gpu(tmp_path, world):
    if n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "mg")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_nccl_worker.py"), out],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(world)]
    for p in procs:
        log, _ = p.communicate(timeout=600)
        assert p.returncode == 0, log.decode(errors="replace")[-3000:]
    solo = np.load(out + ".solo.npz")
    ranks = [np.load(out + ".%d.npz" % r) for r in range(world)]
    assert bool(solo["cells_equal"])                                  # z-slab build == single-GPU build, bit for bit
    assert len(set(float(r["cells_sum"]) for r in ranks)) == 1        # every rank ended up with the same grid
    cat = lambda k: np.concatenate([r[k] for r in ranks])             # noqa: E731
    # predict: Philox keyed by the global particle index -> independent of the GPU count
    assert np.array_equal(cat("after_predict").view(np.uint32), solo["after_predict"].view(np.uint32))
    # update: per-particle cloud weights identical; normalisation through the all-reduced partials
    np.testing.assert_allclose(cat("after_update")[:, 4:], solo["after_update"][:, 4:], rtol=1e-5, atol=1e-12)
    for r in ranks:
        np.testing.assert_allclose(r["mean"], solo["mean"], atol=1e-5)
    # repeated updates on already-normalised weights change nothing structural: same mean on every rank, every time,
    # whichever way the partial sums travelled (peer memory inside the kernels, or ncclAllReduce)
    assert all(bool(r["peer_active"]) for r in ranks) or not any(bool(r["peer_active"]) for r in ranks)
    for r in ranks:
        for k in ("mean_again", "mean_nccl", "mean_back"):
            np.testing.assert_allclose(r[k], solo["mean"], atol=1e-5)
            assert np.array_equal(r[k], ranks[0][k])                  # identical bits on all ranks
    np.testing.assert_allclose(cat("after_update_again")[:, 4], cat("after_update_nccl")[:, 4], rtol=1e-6, atol=1e-12)
    # global resample
    idx = cat("idx")
    assert np.all(np.diff(idx.astype(np.int64)) >= 0)
    assert np.mean(idx != solo["idx"]) < 1e-3                         # fp64 sums differ in the last bit only
    same = idx == solo["idx"]
    assert np.array_equal(cat("after_resample")[same][:, :4].view(np.uint32),
                          solo["after_resample"][same][:, :4].view(np.uint32))
