"""CPU: the error analysis behind weight_v4_kernel's fused-scale voxel estimate (amcl3d_b200/csrc/weight.cu header).

The kernel estimates the voxel coordinate with three FFMAs per axis, q = fma(px,R0, fma(py,R1, fma(pz,R2, F))), and
trusts rint(q) whenever q is farther from a half-integer than  u (8 |p|_2/res + K + 3.6)  (x, y) resp.
u (3 |s_z|/res + K_z + 2.6)  (z), times 1.25.  Here the same arithmetic is emulated in numpy (float32 operations,
FMA = one rounding of the exact double product-sum) against the reference's own arithmetic
(Grid3d.cpp:146-149,174-183: float product/sum chain, double offset add, float rounding, double division) on random
poses and points at the scales of map S and map L; the observed error must stay inside the analytic bound (without
the safety factor) and every coordinate the kernel would trust must floor to the reference's voxel."""
import numpy as np
import pytest

U = np.float32(2.0 ** -24)
F32 = np.float32


def fma32(a, b, c):
    # a*b is exact in float64 (24 + 24 bits); the sum is rounded to float64 and then to float32.  That double
    # rounding differs from a true FMA only in ~2^-29 of the cases and by one float32 ulp -- irrelevant for a bound
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def rotation_rows(roll, pitch, yaw):
    # Grid3d.cpp:139-149: double sin/cos of the float angles, entries rounded to float
    sr, cr, sp, cp = (np.sin(np.float64(F32(roll))), np.cos(np.float64(F32(roll))), np.sin(np.float64(F32(pitch))),
                      np.cos(np.float64(F32(pitch))))
    sy, cy = np.sin(yaw.astype(np.float32).astype(np.float64)), np.cos(yaw.astype(np.float32).astype(np.float64))
    r0 = [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr]
    r1 = [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr]
    r2 = [np.full_like(sy, -sp), np.full_like(sy, cp * sr), np.full_like(sy, cp * cr)]
    return [[c.astype(np.float32) for c in r] for r in (r0, r1, r2)]


@pytest.mark.parametrize("res,extent,reach,n", [(0.1, (20.0, 20.0, 5.0), 12.0, 400000), (0.05, (100.0, 100.0, 20.0), 30.0, 400000),
                                                (0.05, (100.0, 100.0, 20.0), 3.0, 200000), (0.02, (300.0, 200.0, 30.0), 60.0, 200000)])
def test_estimate_error_is_inside_the_analytic_bound(res, extent, reach, n):
    rng = np.random.default_rng(int(res * 1000) + int(reach))
    mn = -np.array(extent) / 2
    t = (mn + rng.uniform(0, 1, (n, 3)) * np.array(extent)).astype(np.float32)       # in-map particle positions
    yaw = rng.uniform(-3.2, 3.2, n)
    p = (rng.normal(0, 1, (n, 3)) * rng.uniform(0.05, reach, (n, 1))).astype(np.float32)
    rows = rotation_rows(0.01, -0.02, yaw)
    inv = 1.0 / res
    inv_f = F32(inv)
    worst = 0.0
    for axis in range(3):
        r = rows[axis]
        # reference: ((px*r0 + py*r1) + pz*r2) in float, + offset in double, to float, / res in double
        s = ((p[:, 0] * r[0]).astype(np.float32) + (p[:, 1] * r[1]).astype(np.float32)).astype(np.float32)
        s = (s + (p[:, 2] * r[2]).astype(np.float32)).astype(np.float32)
        off = t[:, axis].astype(np.float64) - mn[axis]
        v = (s.astype(np.float64) + off).astype(np.float32)
        T = v.astype(np.float64) / res
        # estimate, exactly as make/estimate in weight_v4_kernel
        dx = off * inv
        K = np.floor(dx)
        F = ((dx - K) - 0.5).astype(np.float32)
        if axis < 2:
            R = [(c.astype(np.float64) * inv).astype(np.float32) for c in r]
            q = fma32(p[:, 0], R[0], fma32(p[:, 1], R[1], fma32(p[:, 2], R[2], F)))
            A = np.sqrt((p.astype(np.float64) ** 2).sum(1))
            bound = float(U) * (8 * A / res + K + 3.6)
        else:
            q = fma32(s, np.full(n, inv_f, np.float32), F)
            bound = float(U) * (3 * np.abs(s.astype(np.float64)) / res + K + 2.6)
        err = np.abs(q.astype(np.float64) - (T - K - 0.5))
        assert np.all(err <= bound), (axis, float((err / bound).max()))
        worst = max(worst, float((err / bound).max()))
        # what the kernel concludes: rint(q) + K is the voxel whenever |q - rint(q)| < 0.5 - 1.25 * bound
        d = np.abs(q.astype(np.float64) - np.rint(q.astype(np.float64)))
        trusted = d < 0.5 - 1.25 * bound
        assert np.array_equal((np.rint(q.astype(np.float64)) + K)[trusted], np.floor(T)[trusted])
        assert trusted.mean() > 0.98
    assert worst > 0.02   # the bound is not vacuous: observed errors reach a visible fraction of it
