#!/usr/bin/env python
"""BASELINE.json configs[2]: Grid3d::computeGrid (PointCloudTools.cpp:84-149) on the synthetic 100 x 100 x 20 m
warehouse map @ 0.05 m (1.6 G voxels, ~29 M occupied-leaf points), z-slab sharded over the ranks of one node.

    python tools/grid_bench.py [--reps 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/grid_bench.py

Prints one JSON line on rank 0: voxels/s of the whole build (host point upload + bucket sort + distance field +
slab broadcast; wall clock bracketed by barriers, max over ranks) and of the distance-field kernels alone
(CUDA events inside the library, option kernel_timing)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--map", default="warehouse", choices=["warehouse", "room"])
    args = ap.parse_args()
    import torch
    import amcl3d_b200
    from amcl3d_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = amcl3d_b200.Context(local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(ctx.unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(uid.cpu().numpy(), rank, world)
    pts, bounds = synth.make_map(args.map)
    ctx.set_option("kernel_timing", 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wall, kern = [], []
    cells_total = None
    for rep in range(args.reps + 1):
        grid = amcl3d_b200.Grid(ctx, bounds)
        barrier()
        t0 = time.perf_counter()
        grid.compute(pts, 0.05, keep_dist=False)
        ctx.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        if rep > 0:  # rep 0 warms up allocations / NCCL
            wall.append(dt)
            try:
                kern.append(ctx.last_kernel_ms())
            except Exception:
                kern.append(0.0)  # this rank's z-slab is empty (more ranks than brick rows): no kernel was timed
        cells_total = int(np.prod([int(d) for d in grid.dims]))
        if rep == args.reps:
            probe = grid.download_prob_range(cells_total // 2, 4096)
            checksum = float(np.asarray(probe, np.float64).sum())
        grid.close()
    t = torch.tensor([min(wall), float(np.median(kern))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "metric": "computeGrid_voxels_per_s", "value": cells_total / float(t[0]), "unit": "voxels/s",
            "n_gpus": world, "wall_s": float(t[0]), "df_kernels_ms_per_rank": float(t[1]),
            "voxels": cells_total, "map_points": int(len(pts)), "scaling": "strong", "checksum_mid_4096": checksum,
            "config": {"workload": "cfg3: %s map, z-slabs over %d rank(s), slab broadcast so that every rank "
                                   "ends with the full grid" % (args.map, world)}}), flush=True)
    if world > 1:
        ctx.comm_destroy()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
