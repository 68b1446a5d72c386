"""CPU tests of the exact-chain algebra (amcl3d_b200/csrc/chain_fn.h): the integer model of sequential float
accumulation that the GPU kernels use (windowed scan, segment summaries passed from CTA to CTA and GPU to GPU) must
return the bits of the reference's plain loop `c += t[i]` (ParticleFilter.cpp:151-152,179,190-193,214) for ANY input."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "harness", "chain_host.cpp")
OUT = os.path.join(ROOT, "tests", "harness", "_build", "libchain_host.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(ROOT, "amcl3d_b200", "csrc", "chain_fn.h")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++14", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               SRC, "-o", OUT])
    l = C.CDLL(OUT)
    fp = C.POINTER(C.c_float)
    l.chain_seq_sum.restype = C.c_float
    l.chain_seq_sum.argtypes = [fp, C.c_uint64, C.c_float, fp]
    l.chain_window_sum.restype = C.c_float
    l.chain_window_sum.argtypes = [fp, C.c_uint64, C.c_float, C.c_uint32, C.c_uint32, fp, C.POINTER(C.c_uint64)]
    l.chain_segmented_sum.restype = C.c_float
    l.chain_segmented_sum.argtypes = [fp, C.c_uint64, C.c_float, C.c_uint32, C.c_uint32, C.c_double, C.POINTER(C.c_uint64), fp]
    return l


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def cases():
    rng = np.random.default_rng(7)
    out = {}
    n = 200_000
    out["weights_like"] = rng.gamma(4.0, 0.5, n).astype(np.float32)                      # positive, similar size
    out["normalised"] = (out["weights_like"] / out["weights_like"].sum()).astype(np.float32)
    out["equal"] = np.full(n, np.float32(1.0 / n), np.float32)                             # systematic rounding, ties
    out["pow2_ties"] = (rng.integers(1, 8, n) * np.float32(2.0 ** -20)).astype(np.float32)  # many exact ties
    out["with_zeros"] = np.where(rng.random(n) < 0.3, 0, out["weights_like"]).astype(np.float32)
    w = out["normalised"]
    out["mean_x_negative"] = (w * (-20 + 0.5 * rng.standard_normal(n)).astype(np.float32)).astype(np.float32)
    out["mean_a_mixed"] = (w * (0.2 + 0.2 * rng.standard_normal(n)).astype(np.float32)).astype(np.float32)
    out["mean_y_hover"] = (w * (0.5 * rng.standard_normal(n)).astype(np.float32)).astype(np.float32)
    out["wild"] = (rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)).astype(np.float32)
    out["denormal"] = (rng.random(n) * 1e-41).astype(np.float32)
    sp = out["weights_like"].copy()
    sp[1000] = np.inf
    out["with_inf"] = sp
    sp = out["weights_like"].copy()
    sp[5000] = np.nan
    out["with_nan"] = sp
    out["minus_zero"] = np.where(rng.random(n) < 0.5, np.float32(-0.0), out["normalised"]).astype(np.float32)
    out["all_zero"] = np.zeros(n, np.float32)
    z = out["weights_like"].copy()
    z[:50_000] = 0
    z[120_000:160_000] = 0
    out["zero_runs"] = z
    out["short"] = out["weights_like"][:37].copy()
    out["empty"] = np.zeros(0, np.float32)
    return out


@pytest.mark.parametrize("name", list(cases().keys()))
def test_windowed_chain_equals_sequential_sum(lib, name):
    t = cases()[name]
    n = len(t)
    for c0 in (0.0, 1.5, -3.25e-3):
        ref_prefix = np.zeros(max(n, 1), np.float32)
        want = lib.chain_seq_sum(_p(t), n, c0, _p(ref_prefix))
        for window, head in ((4096, 96), (1024, 0), (64, 7)):
            got_prefix = np.zeros(max(n, 1), np.float32)
            wins = C.c_uint64()
            got = lib.chain_window_sum(_p(t), n, c0, window, head, _p(got_prefix), C.byref(wins))
            assert _bits(got) == _bits(want), (name, c0, window)
            assert np.array_equal(_bits(got_prefix[:n]), _bits(ref_prefix[:n])), (name, c0, window)


@pytest.mark.parametrize("name", list(cases().keys()))
def test_segment_summaries_equal_sequential_sum(lib, name):
    t = cases()[name]
    n = len(t)
    for c0 in (0.0, 0.75):
        ref_prefix = np.zeros(max(n, 1), np.float32)
        want = lib.chain_seq_sum(_p(t), n, c0, _p(ref_prefix))
        # est_bias: a deliberately wrong hypothesis must cost speed only, never correctness
        for seg, leaf, bias in ((2048, 4, 0.0), (4096, 8, 0.0), (512, 1, 0.0), (2048, 4, 0.3), (2048, 4, -0.6)):
            fb = C.c_uint64()
            got_prefix = np.zeros(max(n, 1), np.float32)
            got = lib.chain_segmented_sum(_p(t), n, c0, seg, leaf, bias, C.byref(fb), _p(got_prefix))
            assert _bits(got) == _bits(want), (name, c0, seg, bias)
            assert np.array_equal(_bits(got_prefix[:n]), _bits(ref_prefix[:n])), (name, c0, seg, bias)


def test_fast_path_is_taken_for_weight_like_data(lib):
    """The scheme is only useful if realistic chains rarely fall back: non-negative weights and same-sign mean terms."""
    c = cases()
    for name, limit in (("weights_like", 0.25), ("normalised", 0.25), ("mean_x_negative", 0.25), ("mean_a_mixed", 0.3)):
        t = c[name]
        fb = C.c_uint64()
        lib.chain_segmented_sum(_p(t), len(t), 0.0, 2048, 4, 0.0, C.byref(fb), None)
        n_seg = (len(t) + 2047) // 2048
        assert fb.value <= max(12, limit * n_seg), (name, fb.value, n_seg)
