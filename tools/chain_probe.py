"""Probe: device time of the exact-chain kernels vs particle count (run under `ncu --metrics gpu__time_duration.sum`)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amcl3d_b200
from amcl3d_b200 import synth

ctx = amcl3d_b200.Context(0)
w = synth.make_workload("cfg1", n_points=256)
grid = amcl3d_b200.Grid(ctx, w["bounds"])
grid.compute(w["map_points"], 0.05, keep_dist=False)
ctx.set_option("sum_mode", 1)
ctx.set_option("resample_mode", 1)
for n in (1024, 4096, 16384, 65536):
    p = synth.particles_tracking(n, w["pose"], (0.05, 0.05, 0.05, 0.1))
    f = amcl3d_b200.Filter(ctx)
    f.upload(p)
    for _ in range(2):
        f.update(grid, w["cloud"], w["ranges"], 0.5, 0.53, 0.0, 0.0)
        f.resample(0.3)
    f.close()
print("done")
