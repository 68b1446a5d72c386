import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import amcl3d_b200
from oracle.bindings import Port
port = Port()
ctx = amcl3d_b200.Context(0)
lib = ctx.lib
lib.amcl3d_cuda_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
for n, zeros in ((20000, True), (20000, False), (1048576, False)):
    rng = np.random.default_rng(n)
    p = np.zeros((n, 7), np.float32)
    w = rng.gamma(0.4, 1.0, n)
    if zeros:
        w[::97] = 0
    p[:, 4] = (w / w.sum()).astype(np.float32)
    ctx.set_option("resample_mode", 3)
    f = amcl3d_b200.Filter(ctx)
    f.upload(p)
    idx = f.resample(0.41, want_idx=True)
    cum = np.zeros(n, np.float32)
    lib.amcl3d_cuda_debug_read(f.h, 0, cum.ctypes.data_as(C.c_void_p), n * 4)
    n_seg = (n + 2047) // 2048
    carry = np.zeros(n_seg, np.float32); slow = np.zeros(n_seg, np.uint32)
    lib.amcl3d_cuda_debug_read(f.h, 1, carry.ctypes.data_as(C.c_void_p), n_seg * 4)
    lib.amcl3d_cuda_debug_read(f.h, 2, slow.ctypes.data_as(C.c_void_p), n_seg * 4)
    f.close()
    _, idx_o = port.resample(p, 0.41)
    badi = np.nonzero(idx != idx_o)[0]
    print("n", n, "idx mismatches", len(badi), "first", [(int(b), int(idx[b]), int(idx_o[b])) for b in badi[:3]])
    factor = np.float32(1.0) / np.float32(n)
    u = (factor * np.float32(0.41) + factor * np.arange(n, dtype=np.uint32).astype(np.float32)).astype(np.float32)
    mine = np.minimum(np.searchsorted(cum, u, side="left"), n - 1)
    print("   numpy search on the GPU's cum == oracle idx:", np.array_equal(mine.astype(np.uint32), idx_o))
    want = np.cumsum(p[:, 4], dtype=np.float32)
    bad = np.nonzero(cum.view(np.uint32) != want.view(np.uint32))[0]
    print("n", n, "zeros", zeros, "cum mismatches", len(bad), "first", bad[:5], "got", cum[bad[:5]], "want", want[bad[:5]])
    print("  slow flags", slow[:12], "carry", carry[:6], "want carry", np.concatenate([[0], want[2047::2048][:5]]))
    if len(bad):
        b = bad[0]
        print("  around first bad:", cum[max(0, b - 3):b + 4], want[max(0, b - 3):b + 4], "w", p[max(0, b - 3):b + 4, 4])
        runs = np.nonzero(np.diff(np.concatenate([[0], (cum.view(np.uint32) != want.view(np.uint32)).astype(np.int8), [0]])))[0]
        print("  bad runs (start,end) first 6:", runs[:12].reshape(-1, 2)[:6].tolist())
