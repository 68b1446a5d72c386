"""Worker of tests/test_sharding_gloo.py: one rank of a world_size-N CPU (gloo) run of the sharded update/resample."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from amcl3d_b200 import shard, synth  # noqa: E402
from oracle.bindings import Port  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    port = Port()
    pts, bounds = synth.map_room(size=(6.0, 6.0, 3.0), res=0.1, n_boxes=5, seed=1)
    pose = (0.0, 0.0, 1.5, 0.3)
    cloud = synth.sensor_cloud(pts, pose, 400, 4.0, seed=2)
    P = synth.particles_tracking(96, pose, (0.1, 0.1, 0.1, 0.2), seed=3)
    P[5, 0] = 77.0   # one particle outside the map
    ranges = synth.beacons(pose, positions=((-2.5, -2.5, 2.5), (2.5, -2.5, 2.5), (0.0, 2.5, 2.5)))
    alpha, sigma, roll, pitch = 0.5, 0.53, np.float32(0.01), np.float32(-0.02)
    cells, dims = port.compute_grid(pts, bounds, 0.05)

    first, count = shard.partition(len(P), rank, world)
    mine = P[first:first + count]
    inside = np.array([port.is_into_map(bounds, *p[:3]) for p in mine])
    wp = np.array([port.cloud_weight(cells, dims, bounds, cloud, (p[0], p[1], p[2], roll, pitch, p[3]))[0] if ok else 0.0
                   for p, ok in zip(mine, inside)], np.float32)
    wr = np.array([port.range_weight(p[0], p[1], p[2], ranges, sigma) if ok else 0.0 for p, ok in zip(mine, inside)],
                  np.float32)
    partials = torch.from_numpy(shard.update_partials(mine[:, :4], wp, wr, inside))
    dist.all_reduce(partials)                                    # the ONE collective of a sharded update
    wpn, wrn, wn, mean = shard.finish_update(partials.numpy(), wp, wr, inside, alpha)
    # global resample: all-gather the weights, each rank fills its own output slots
    gathered = [torch.zeros(count, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(wn))
    w_all = torch.cat(gathered).numpy()
    idx = shard.resample_indices(w_all, 0.37, first, count)
    np.savez(out_path + ".%d.npz" % rank, first=first, count=count, wn=wn, mean=mean, idx=idx, w_all=w_all)
    if rank == 0:
        full, mean_o = port.update(P, cells, dims, bounds, cloud, ranges, alpha, sigma, roll, pitch)
        np.savez(out_path + ".oracle.npz", full=full, mean=mean_o)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
