"""Worker of tests/test_gpu_multigpu.py: one rank (one GPU) of a sharded predict / update / resample / computeGrid run.
Rank 0 also runs the same work on a single-GPU context and compares."""
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
    out_path = sys.argv[1]
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

    w = synth.make_workload("cfg1", n_particles=4096 * world, n_points=1500)
    w["particles"][17, 1] = -300.0
    # ---- computeGrid: z-slabs per rank + broadcast => every rank holds the full grid
    grid = amcl3d_b200.Grid(ctx, w["bounds"])
    grid.compute(w["map_points"], w["sensor_dev"])
    cells = grid.download_cells()

    first, count = shard.partition(len(w["particles"]), rank, world)
    pf = amcl3d_b200.Filter(ctx)
    pf.upload(w["particles"][first:first + count])
    pf.predict(w["odom_mods"], w["deltas"], seed=5, step=3)       # Philox counters = global particle index
    after_predict = pf.download()
    mean = pf.update(grid, w["cloud"], w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
    after_update = pf.download()
    # the ten partial sums travel through peer memory inside the kernels when CUDA IPC is available; a few more
    # updates exercise the parity double-buffering, then the NCCL route must give the same numbers
    peer_active = ctx.comm_peer_active()
    for _ in range(5):
        mean_again = pf.update(grid, w["cloud"], w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
    after_update_again = pf.download()
    ctx.set_option("peer_reduce", 1)
    assert not ctx.comm_peer_active()
    mean_nccl = pf.update(grid, w["cloud"], w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
    after_update_nccl = pf.download()
    ctx.set_option("peer_reduce", 0)
    mean_back = pf.update(grid, w["cloud"], w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
    idx = pf.resample(0.61, want_idx=True)
    after_resample = pf.download()
    np.savez(out_path + ".%d.npz" % rank, first=first, count=count, after_predict=after_predict,
             after_update=after_update, mean=mean, idx=idx, after_resample=after_resample,
             peer_active=peer_active, mean_again=mean_again, after_update_again=after_update_again,
             mean_nccl=mean_nccl, after_update_nccl=after_update_nccl, mean_back=mean_back,
             cells_sum=np.float64(cells.astype(np.float64).sum()))

    if rank == 0:
        solo = amcl3d_b200.Context(0)
        solo.set_option("sum_mode", 2)
        solo.set_option("resample_mode", 2)
        g1 = amcl3d_b200.Grid(solo, w["bounds"])
        g1.compute(w["map_points"], w["sensor_dev"])
        c1 = g1.download_cells()
        f1 = amcl3d_b200.Filter(solo)
        f1.upload(w["particles"])
        f1.predict(w["odom_mods"], w["deltas"], seed=5, step=3)
        p1 = f1.download()
        m1 = f1.update(g1, w["cloud"], w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
        u1 = f1.download()
        i1 = f1.resample(0.61, want_idx=True)
        r1 = f1.download()
        np.savez(out_path + ".solo.npz", after_predict=p1, after_update=u1, mean=m1, idx=i1, after_resample=r1,
                 cells_equal=np.array_equal(c1.view(np.uint32), cells.view(np.uint32)))
        f1.close()
        g1.close()
        solo.close()
    dist.barrier()
    pf.close()
    grid.close()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
