"""GPU, >= 2 devices: the particle filter sharded over several GPUs (one process per GPU, unequal shards, an empty
shard) against the CPU ORACLE: predict, update with the exact cross-rank chains, global resample with peer-memory
gathers.  Skipped on single-GPU boxes (bench.py's `parity` record carries the multi-GPU check into the scaling runs)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def assert_mean(mean_g, mean_o, mask):
    """Components flagged exact are the reference's float chain bit for bit on every rank; a component that hovers around
    zero is returned as the fp64 sum (see amcl3d_cuda_pf_mean_exact_mask)."""
    for k in range(4):
        if (int(mask) >> k) & 1:
            assert bits(mean_g[k:k + 1])[0] == bits(mean_o[k:k + 1])[0], (k, mean_g, mean_o)
        else:
            assert abs(float(mean_g[k]) - float(mean_o[k])) <= 1e-6, (k, mean_g, mean_o)


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,scenario", [(2, "uneven"), (2, "empty"), (3, "uneven"), (4, "uneven"), (8, "uneven"),
                                            (8, "empty")])
def test_sharded_filter_matches_the_oracle(tmp_path, port, cuda_ctx, cfg1, cfg1_cells, world, scenario):
    if n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import amcl3d_b200
    from amcl3d_b200 import synth
    cells, dims = cfg1_cells
    n_total = 4096 * world + 37                      # not divisible by the rank count
    particles = synth.particles_tracking(n_total, cfg1["pose"], (0.2, 0.2, 0.2, 0.4), seed=3)
    particles[17, 1] = -300.0                        # one particle outside the map
    cloud = cfg1["cloud"][:1500]
    out = str(tmp_path / "mg")
    np.savez(out + ".input.npz", particles=particles, cloud=cloud, ranges=cfg1["ranges"], cells=cells,
             bounds=cfg1["bounds"], mods=np.asarray(cfg1["odom_mods"]), deltas=np.asarray(cfg1["deltas"]))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_nccl_worker.py"), out, scenario],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(world)]
    for p in procs:
        log, _ = p.communicate(timeout=900)
        assert p.returncode == 0, log.decode(errors="replace")[-3000:]
    ranks = [np.load(out + ".%d.npz" % r) for r in range(world)]
    cat = lambda k: np.concatenate([r[k] for r in ranks])             # noqa: E731
    assert sum(int(r["count"]) for r in ranks) == n_total
    if scenario == "empty":
        assert int(ranks[-1]["count"]) == 0
    assert all(bool(r["peer_active"]) for r in ranks)

    # ---- predict: Philox keyed by the GLOBAL particle index -> same bits as one GPU holding the whole set
    solo = amcl3d_b200.Filter(cuda_ctx)
    solo.upload(particles)
    solo.predict(cfg1["odom_mods"], cfg1["deltas"], seed=5, step=3)
    p1 = solo.download()
    assert np.array_equal(bits(cat("after_predict")), bits(p1))

    # ---- cycle 1 (beacons) against the full oracle update
    want1, mean_o1 = port.update(p1, cells, dims, cfg1["bounds"], cloud, cfg1["ranges"], 0.5, 0.53, 0.01, -0.02)
    got1 = cat("after_update1")
    _, cnt_o = port.cloud_weight_batch(None, dims, cfg1["bounds"], cloud, p1[:, :4], 0.01, -0.02)
    inmap = np.array([port.is_into_map(cfg1["bounds"], *q[:3]) for q in p1])
    assert np.array_equal(cat("cnt1")[inmap], cnt_o[inmap])
    np.testing.assert_allclose(got1[:, 4:], want1[:, 4:], rtol=1e-5, atol=1e-30)
    # pose-balanced weighting over all ranks (default) vs every rank weighing its own shard: identical bits
    assert np.array_equal(bits(cat("raw1")), bits(cat("raw1_local")))
    assert np.array_equal(cat("cnt1"), cat("cnt1_local"))
    for r in ranks:
        assert np.array_equal(bits(r["mean1_local"]), bits(r["mean1_again"]))
        np.testing.assert_allclose(r["mean1"], mean_o1, atol=1e-5)
        assert np.array_equal(bits(r["mean1"]), bits(ranks[0]["mean1"]))           # identical bits on all ranks
        assert np.array_equal(bits(r["mean1_again"]), bits(ranks[0]["mean1_again"]))
        np.testing.assert_allclose(r["mean_fast"], mean_o1, atol=1e-5)             # fp64 sums: tolerance level
        np.testing.assert_allclose(r["mean_fast_nccl"], mean_o1, atol=1e-5)
        assert np.array_equal(bits(r["mean_fast"]), bits(ranks[0]["mean_fast"]))

    # ---- cycle 2 (no beacons): the reference's loops fed with the GPUs' raw weights give the GPUs' bits
    q = p1.copy()
    q[:, 4:] = cat("after_update_fast")[:, 4:]      # the state the update started from (stale wp / wr of skipped ones)
    q[inmap, 5] = cat("raw2")[inmap]
    q[inmap, 6] = 0.0
    want2, mean_o2 = port.update_from_weights(q, cfg1["bounds"], 0.5)
    got2 = cat("after_update2")
    assert np.array_equal(bits(got2[inmap]), bits(want2[inmap]))
    assert np.array_equal(bits(got2[:, 4]), bits(want2[:, 4]))
    for r in ranks:
        assert_mean(r["mean2"], mean_o2, r["mask2"])
        assert int(r["mask2"]) & 0b1100 == 0b1100                     # z and yaw do not hover
        assert np.array_equal(bits(r["mean2"]), bits(ranks[0]["mean2"])) and int(r["mask2"]) == int(ranks[0]["mask2"])
    # and the raw weights themselves are within tolerance of the oracle's
    w_o, _ = port.cloud_weight_batch(cells, dims, cfg1["bounds"], cloud, p1[:, :4], 0.01, -0.02)
    rel = np.abs(cat("raw2")[inmap] - w_o[inmap]) / np.maximum(w_o[inmap], 1e-30)
    assert rel.max() <= 1e-5

    # ---- global resample: the reference's walk over the concatenated set, bit for bit
    want_r1, idx_o1 = port.resample(got2, 0.61)
    assert np.array_equal(cat("idx1"), idx_o1)
    assert np.array_equal(bits(cat("after_resample1")), bits(want_r1))

    # ---- cycle 3 on the resampled set
    solo.upload(want_r1)
    solo.predict(cfg1["odom_mods"], cfg1["deltas"], seed=5, step=4)
    p2 = solo.download()
    solo.close()
    assert np.array_equal(bits(cat("after_predict2")), bits(p2))
    inmap2 = np.array([port.is_into_map(cfg1["bounds"], *z[:3]) for z in p2])
    q = p2.copy()
    q[inmap2, 5] = cat("raw3")[inmap2]
    q[inmap2, 6] = 0.0
    want3, mean_o3 = port.update_from_weights(q, cfg1["bounds"], 0.5)
    got3 = cat("after_update3")
    assert np.array_equal(bits(got3[:, 4]), bits(want3[:, 4]))
    assert np.array_equal(bits(got3[inmap2]), bits(want3[inmap2]))
    for r in ranks:
        assert_mean(r["mean3"], mean_o3, r["mask3"])
        assert np.array_equal(bits(r["mean3"]), bits(ranks[0]["mean3"]))
    want_r2, idx_o2 = port.resample(got3, 0.07)
    assert np.array_equal(cat("idx2"), idx_o2)
    assert np.array_equal(bits(cat("after_resample2")), bits(want_r2))
    want_r3, idx_o3 = port.resample(want_r2, 0.93)
    assert np.array_equal(cat("idx3"), idx_o3)
    assert np.array_equal(bits(cat("after_resample3")), bits(want_r3))
