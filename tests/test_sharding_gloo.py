"""CPU, world_size 2 and 3 (gloo): the sharded update / resample algebra (amcl3d_b200/shard.py -- the specification the
CUDA library's multi-GPU path implements) against the single-process oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_update_and_resample_match_oracle(tmp_path, world, port):
    out = str(tmp_path / "shard")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(world),
               OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_gloo_worker.py"), out],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(world)]
    for p in procs:
        log, _ = p.communicate(timeout=300)
        assert p.returncode == 0, log.decode(errors="replace")[-2000:]
    oracle = np.load(out + ".oracle.npz")
    full, mean_o = oracle["full"], oracle["mean"]
    shards = [np.load(out + ".%d.npz" % r) for r in range(world)]
    wn = np.concatenate([s["wn"] for s in shards])
    assert len(wn) == len(full)
    np.testing.assert_allclose(wn, full[:, 4], rtol=1e-5, atol=1e-12)            # weights: 1e-5 relative
    for s in shards:
        np.testing.assert_allclose(s["mean"], mean_o, atol=1e-4)                 # mean pose: 1e-4 m, same on all ranks
    # global resample: the union of the ranks' slots equals the single-process scan-mode resample of the same weights
    from amcl3d_b200 import shard
    idx = np.concatenate([s["idx"] for s in shards])
    want = shard.resample_indices(shards[0]["w_all"], 0.37, 0, len(full))
    assert np.array_equal(idx, want) and np.all(np.diff(idx.astype(int)) >= 0)
    # and agrees with the reference's float chain except where the chain's rounding moves a boundary
    _, idx_chain = port.resample(np.column_stack([full[:, :4], wn, full[:, 5:]]), 0.37)
    assert np.mean(idx != idx_chain) < 0.02


def test_partition_covers_everything():
    from amcl3d_b200 import shard
    for n in (0, 1, 7, 96, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard.partition(n, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for first, count in spans:
                assert first == pos or count == 0
                pos += count
