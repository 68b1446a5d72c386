"""Debug: segmented resample vs the oracle at several sizes, repeated (flakiness), with the first mismatch located."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import amcl3d_b200
from oracle.bindings import Port

port = Port()
ctx = amcl3d_b200.Context(0)
sizes = [int(a) for a in sys.argv[1:]] or [20000, 100003, 300000, 600000, 1048576]
for n in sizes:
    rng = np.random.default_rng(n)
    p = np.zeros((n, 7), np.float32)
    p[:, 0] = np.arange(n, dtype=np.float32)
    w = rng.gamma(0.4, 1.0, n)
    w[::97] = 0
    p[:, 4] = (w / w.sum()).astype(np.float32)
    want, idx_o = port.resample(p, 0.41)
    for rep in range(3):
        for mode in (3, 1):
            ctx.set_option("resample_mode", mode)
            f = amcl3d_b200.Filter(ctx)
            f.upload(p)
            idx = f.resample(0.41, want_idx=True)
            f.close()
            bad = np.nonzero(idx != idx_o)[0]
            print("n", n, "rep", rep, "mode", mode, "mismatches", len(bad), "first", (int(bad[0]), int(idx[bad[0]]), int(idx_o[bad[0]])) if len(bad) else None, flush=True)
ctx.close()
