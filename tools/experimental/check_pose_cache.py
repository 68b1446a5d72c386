"""Quick hardware check of the pose-parameter cache: bit-equality with the cache off, both layouts, + kernel time."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amcl3d_b200
from amcl3d_b200 import synth

def run(ctx, bounds, n, n_pts, opts, seed):
    rng = np.random.default_rng(seed)
    for k, v in opts.items():
        ctx.set_option(k, v)
    g = amcl3d_b200.Grid(ctx, bounds)
    dims = [int(d) for d in g.dims]
    cells = np.empty((dims[0] * dims[1] * dims[2], 2), np.float32)
    cells[:, 0] = 0.1
    cells[:, 1] = rng.random(len(cells), dtype=np.float32) + np.float32(0.01)
    g.upload_cells(cells, 0.05)
    mid = (bounds[:3] + bounds[3:6]) / 2
    P = synth.particles_tracking(n, (mid[0], mid[1], mid[2], 0.3), (0.4, 0.4, 0.2, 0.5), seed=seed)
    P[3, 0] = 1e4
    cloud = np.zeros((n_pts, 4), np.float32)
    cloud[:, :3] = rng.normal(0, 3.0, (n_pts, 3))
    outs = []
    for cache in (1, 0):
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_option("pose_cache", cache)
        ctx.set_option("kernel_timing", 1)
        f = amcl3d_b200.Filter(ctx)
        f.upload(P)
        f.update(g, cloud, None, 0.5, 0.53, 0.01, -0.02)
        f.update(g, cloud, None, 0.5, 0.53, 0.01, -0.02)
        ms = ctx.last_kernel_ms()
        outs.append((f.download(), f.last_in_map_evals(), ms))
        f.close()
    for k in opts:
        ctx.set_option(k, 0)
    ctx.set_option("pose_cache", 0)
    g.close()
    same = np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32)) and outs[0][1] == outs[1][1]
    return same, outs[0][2], outs[1][2], outs[0][1]

ctx = amcl3d_b200.Context(0)
ok = True
cases = [
    ("linear splits=4 exact sums", np.array([-6.0, -6.0, 0.0, 6.0, 6.0, 3.0, 0.1]), 6000, 3000, {"weight_point_splits": 4, "sum_mode": 3, "grid_layout": 1}),
    ("linear auto 10k x 10k", np.array([-10.0, -10.0, 0.0, 10.0, 10.0, 5.0, 0.1]), 10000, 10000, {"sum_mode": 3, "grid_layout": 1}),
    ("partial bounds order off", np.array([-6.03, -5.98, 0.0, 6.01, 6.0, 3.04, 0.07]), 5000, 2000, {"weight_point_splits": 3, "sum_mode": 3, "particle_order": 1, "grid_layout": 1}),
    ("bricked chunk launches", np.array([-6.0, -6.0, 0.0, 6.0, 6.0, 3.0, 0.05]), 160000, 1500, {"weight_point_splits": 1, "weight_chunk_points": 512, "sum_mode": 3, "grid_layout": 2}),
    ("bricked sub-chunks", np.array([-6.0, -6.0, 0.0, 6.0, 6.0, 3.0, 0.05]), 90000, 2100, {"weight_point_splits": 2, "weight_chunk_points": 1024, "sum_mode": 3, "grid_layout": 2}),
]
for name, b, n, m, o in cases:
    same, ms_off, ms_on, ev = run(ctx, b, n, m, o, 7)
    ok &= same and ev > 0
    print("%-28s same_bits=%s evals=%d kernel ms cache off %.4f on %.4f" % (name, same, ev, ms_off, ms_on), flush=True)
print("POSE_CACHE_OK" if ok else "POSE_CACHE_MISMATCH")
ctx.close()
