"""Worker of tests/test_gpu_multigpu.py: one rank (one GPU) of a sharded predict / update / resample run with UNEQUAL
shards (and, in the "empty" scenario, one rank without particles).  It only runs the library and records what it got;
the parent test compares the concatenated shards with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import amcl3d_b200  # noqa: E402
from amcl3d_b200 import shard, synth  # noqa: E402


def main():
    out_path, scenario = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    ctx = amcl3d_b200.Context(rank)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(ctx.unique_id().copy())
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    ctx.comm_init(uid.cpu().numpy(), rank, world)
    assert ctx.comm_rank() == (rank, world)
    peer_active = ctx.comm_peer_active()

    inp = np.load(out_path + ".input.npz")
    particles, cloud, ranges, cells, bounds = inp["particles"], inp["cloud"], inp["ranges"], inp["cells"], inp["bounds"]
    n_total = len(particles)
    if scenario == "empty":
        # the last rank holds nothing, the others share the set unevenly
        first, count = shard.partition(n_total, rank, world - 1) if rank < world - 1 else (n_total, 0)
    else:
        first, count = shard.partition(n_total, rank, world)
    grid = amcl3d_b200.Grid(ctx, bounds)
    grid.upload_cells(cells, 0.05)
    rec = dict(first=first, count=count, peer_active=peer_active)

    pf = amcl3d_b200.Filter(ctx)
    pf.upload(particles[first:first + count])
    pf.predict(inp["mods"], inp["deltas"], seed=5, step=3)       # Philox counters = global particle index
    rec["after_predict"] = pf.download()
    # cycle 1: beacons
    rec["mean1"] = pf.update(grid, cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    rec["after_update1"] = pf.download()
    rec["raw1"], rec["cnt1"] = pf.last_cloud_weights()
    # a few more updates exercise the parity double-buffering of the mailboxes
    for _ in range(3):
        rec["mean1_again"] = pf.update(grid, cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    # the same update with every rank weighing its OWN shard (the default deals the weighting work out by pose over all
    # ranks): scheduling only, the cloud weights must not change by a bit
    ctx.set_option("global_schedule", 1)
    rec["mean1_local"] = pf.update(grid, cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    rec["raw1_local"], rec["cnt1_local"] = pf.last_cloud_weights()
    ctx.set_option("global_schedule", 0)
    # the fast (fp64) sums: peer memory inside the kernels, then the ncclAllReduce route
    ctx.set_option("sum_mode", 2)
    rec["mean_fast"] = pf.update(grid, cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    rec["after_update_fast"] = pf.download()
    ctx.set_option("peer_reduce", 1)
    rec["mean_fast_nccl"] = pf.update(grid, cloud, ranges, 0.5, 0.53, 0.01, -0.02)
    ctx.set_option("peer_reduce", 0)
    ctx.set_option("sum_mode", 0)
    # cycle 2: no beacons (nothing transcendental: the cross-rank chains can be checked bit for bit)
    rec["mean2"] = pf.update(grid, cloud, None, 0.5, 0.53, 0.01, -0.02)
    rec["mask2"] = pf.mean_exact_mask()
    rec["after_update2"] = pf.download()
    rec["raw2"], _ = pf.last_cloud_weights()
    rec["idx1"] = pf.resample(0.61, want_idx=True)
    rec["after_resample1"] = pf.download()
    # cycle 3 on the resampled set (state / cumulative buffers flipped on every rank)
    pf.predict(inp["mods"], inp["deltas"], seed=5, step=4)
    rec["after_predict2"] = pf.download()
    rec["mean3"] = pf.update(grid, cloud, None, 0.5, 0.53, 0.01, -0.02)
    rec["mask3"] = pf.mean_exact_mask()
    rec["after_update3"] = pf.download()
    rec["raw3"], _ = pf.last_cloud_weights()
    rec["idx2"] = pf.resample(0.07, want_idx=True)
    rec["after_resample2"] = pf.download()
    # two resamples back to back (no update in between)
    rec["idx3"] = pf.resample(0.93, want_idx=True)
    rec["after_resample3"] = pf.download()
    np.savez(out_path + ".%d.npz" % rank, **rec)

    dist.barrier()
    pf.close()
    grid.close()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
