#!/bin/bash
# usage: tools/debug/mg_debug.sh <world> <scenario> <outdir>  -- runs tests/mp_nccl_worker.py by hand with per-rank logs
W=$1; SC=$2; OUT=$3
mkdir -p $OUT
python - <<PY
import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from amcl3d_b200 import synth
import amcl3d_b200
w = synth.make_workload("cfg1")
ctx = amcl3d_b200.Context(0)
g = amcl3d_b200.Grid(ctx, w["bounds"]); g.compute(w["map_points"], w["sensor_dev"], keep_dist=True)
cells = g.download_cells()
n_total = 4096 * $W + 37
particles = synth.particles_tracking(n_total, w["pose"], (0.2, 0.2, 0.2, 0.4), seed=3)
particles[17, 1] = -300.0
np.savez("$OUT/mg.input.npz", particles=particles, cloud=w["cloud"][:1500], ranges=w["ranges"], cells=cells, bounds=w["bounds"], mods=np.asarray(w["odom_mods"]), deltas=np.asarray(w["deltas"]))
PY
for r in $(seq 0 $((W-1))); do
  MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 WORLD_SIZE=$W RANK=$r LOCAL_RANK=$r timeout -s ABRT 90 python -X faulthandler tests/mp_nccl_worker.py $OUT/mg $SC > $OUT/rank$r.log 2>&1 &
done
wait
echo done
