"""Debug: exact-order weighting on the bricked hall map, variants 0 / 4 / 5 vs the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import amcl3d_b200
from amcl3d_b200 import synth
from oracle.bindings import Port

port = Port()
ctx = amcl3d_b200.Context(0)
pts, bounds = synth.map_warehouse(size=(32.0, 32.0, 10.0), res=0.05, n_pallets=60, seed=5)
pose = np.array([-3.0, 1.0, 1.5, 0.2])
cloud = synth.sensor_cloud(pts, pose, 32768, 14.0, seed=7)
grid = amcl3d_b200.Grid(ctx, bounds)
grid.compute(pts, 0.05, keep_dist=False)
prob = grid.download_prob()
cells = np.zeros((len(prob), 2), np.float32); cells[:, 1] = prob
dims = grid.dims.copy()
for n in (16384, 131072):
    particles = synth.particles_tracking(n, pose, (0.5, 0.5, 0.5, 0.2), seed=16)
    pick = np.arange(0, n, n // 512)
    w_o, n_o = port.cloud_weight_batch(cells, dims, bounds, cloud, particles[pick, :4], 0.01, -0.02)
    # "exact" sums of the same gathered values in double, per particle
    w_d = []
    for i in pick[:64]:
        p = particles[i]
        idx, cnt = port.cloud_indices(dims, bounds, cloud, (p[0], p[1], p[2], 0.01, -0.02, p[3]))
        vals = cells[idx[idx != 0xFFFFFFFF], 1].astype(np.float64)
        w_d.append(vals.sum() / max(cnt, 1))
    w_d = np.array(w_d)
    print("n", n, "reference float chain vs fp64 sum: max rel", float(np.max(np.abs(w_o[:64] - w_d) / w_d)))
    for variant, splits, corder, chunk in ((0, 1, 1, 0), (0, 1, 1, 4096), (4, 1, 1, 4096), (0, 0, 0, 0), (0, 0, 1, 0)):
        for k, v in (("weight_variant", variant), ("weight_point_splits", splits), ("cloud_order", corder), ("weight_chunk_points", chunk)):
            ctx.set_option(k, v)
        pf = amcl3d_b200.Filter(ctx)
        pf.upload(particles)
        l0 = ctx.launch_count()
        pf.update(grid, cloud, None, 0.5, 0.53, 0.01, -0.02)
        l1 = ctx.launch_count()
        rw, rn = pf.last_cloud_weights()
        pf.close()
        bad = np.nonzero(rw[pick].view(np.uint32) != w_o.view(np.uint32))[0]
        print("n", n, "variant", variant, "splits", splits, "cloud_order", corder, "chunk", chunk, "launches", l1 - l0,
              "counts equal", np.array_equal(rn[pick], n_o), "n bad", len(bad),
              "max rel vs reference", float(np.max(np.abs(rw[pick] - w_o) / np.maximum(w_o, 1e-30))),
              "max rel vs fp64", float(np.max(np.abs(rw[pick][:64] - w_d) / w_d)))
