"""Minimal driver for ncu captures: builds the workload's grid, runs a few updates with the given options.
    python tools/prof_run.py cfg4 --particles 131072 --opt cloud_order=1 --opt weight_point_splits=1 --updates 2"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amcl3d_b200
from amcl3d_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--particles", type=int, default=None)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--updates", type=int, default=2)
ap.add_argument("--grid-only", action="store_true")
args = ap.parse_args()
os.environ.setdefault("AMCL3D_SYNTH_CACHE", "/tmp/amcl3d_synth_%d" % os.getuid())
w = synth.make_workload(args.workload, n_particles=args.particles)
ctx = amcl3d_b200.Context(0)
for o in args.opt:
    k, v = o.split("=")
    ctx.set_option(k, int(v))
grid = amcl3d_b200.Grid(ctx, w["bounds"])
grid.compute(w["map_points"], w["sensor_dev"], keep_dist=False)
if not args.grid_only:
    pf = amcl3d_b200.Filter(ctx)
    pf.upload(w["particles"])
    for _ in range(args.updates):
        m = pf.update(grid, w["cloud"], w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
    print("mean", m, "launches", ctx.launch_count())
    pf.close()
grid.close()
ctx.close()
