"""ctypes binding of include/amcl3d_cuda.h.

Plain pointers and sizes only -- numpy arrays are handed over as host buffers; nothing here computes.
The library is looked up in-tree (amcl3d_b200/lib/libamcl3d_cuda.so); if it is missing it is built with
nvcc (amcl3d_b200/build.py); if that fails the import error is raised -- there is no fallback path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f, c_d, c_u32, c_u64, c_i64, c_vp, c_int = C.c_float, C.c_double, C.c_uint32, C.c_uint64, C.c_int64, C.c_void_p, C.c_int
_P = C.POINTER

# name -> (restype, argtypes); mirrors include/amcl3d_cuda.h one to one
SIGNATURES = {
    "amcl3d_cuda_abi_version": (c_int, []),
    "amcl3d_cuda_last_error": (C.c_char_p, []),
    "amcl3d_cuda_ctx_create": (c_int, [c_int, c_vp, _P(c_vp)]),
    "amcl3d_cuda_ctx_destroy": (c_int, [c_vp]),
    "amcl3d_cuda_ctx_set_stream": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_ctx_synchronize": (c_int, [c_vp]),
    "amcl3d_cuda_ctx_device_info": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_ctx_set_option": (c_int, [c_vp, C.c_char_p, c_i64]),
    "amcl3d_cuda_ctx_get_option": (c_int, [c_vp, C.c_char_p, _P(c_i64)]),
    "amcl3d_cuda_ctx_last_kernel_ms": (c_int, [c_vp, _P(c_f)]),
    "amcl3d_cuda_ctx_last_update_phases_ms": (c_int, [c_vp, _P(c_f)]),
    "amcl3d_cuda_ctx_launch_count": (c_int, [c_vp, _P(c_u64)]),
    "amcl3d_cuda_voxel_grid": (c_int, [c_vp, c_vp, c_u64, c_f, c_f, c_f, c_vp, c_u64, _P(c_u64)]),
    "amcl3d_cuda_probe_gather": (c_int, [c_vp, c_u64, C.c_uint32, _P(C.c_double), _P(C.c_double)]),
    "amcl3d_cuda_grid_create": (c_int, [c_vp, c_vp, _P(c_vp)]),
    "amcl3d_cuda_grid_destroy": (c_int, [c_vp]),
    "amcl3d_cuda_grid_dims": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_grid_bounds": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_grid_upload_cells": (c_int, [c_vp, c_vp, c_d]),
    "amcl3d_cuda_grid_download_cells": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_grid_download_prob": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_grid_download_prob_range": (c_int, [c_vp, c_u64, c_u64, c_vp]),
    "amcl3d_cuda_grid_gather_prob": (c_int, [c_vp, c_vp, c_u64, c_vp]),
    "amcl3d_cuda_grid_has_cells": (c_int, [c_vp, _P(c_int)]),
    "amcl3d_cuda_grid_compute": (c_int, [c_vp, c_vp, c_u64, c_d, c_int]),
    "amcl3d_cuda_cloud_weight": (c_int, [c_vp, c_vp, c_u64, c_f, c_f, c_f, c_f, c_f, c_f, _P(c_f), _P(c_u32), c_vp]),
    "amcl3d_cuda_cloud_weight_batch": (c_int, [c_vp, c_vp, c_u64, c_vp, c_u64, c_f, c_f, c_vp, c_vp]),
    "amcl3d_cuda_is_into_map": (c_int, [c_vp, c_f, c_f, c_f, _P(c_int)]),
    "amcl3d_cuda_pf_create": (c_int, [c_vp, _P(c_vp)]),
    "amcl3d_cuda_pf_destroy": (c_int, [c_vp]),
    "amcl3d_cuda_pf_upload_particles": (c_int, [c_vp, c_vp, c_u64]),
    "amcl3d_cuda_pf_download_particles": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_pf_size": (c_int, [c_vp, _P(c_u64)]),
    "amcl3d_cuda_pf_init": (c_int, [c_vp, c_u64, c_vp, c_vp, c_vp, c_u64, c_vp]),
    "amcl3d_cuda_pf_predict": (c_int, [c_vp, c_vp, c_vp, c_vp, c_u64, c_u64]),
    "amcl3d_cuda_pf_stage_cloud": (c_int, [c_vp, c_vp, c_u64]),
    "amcl3d_cuda_pf_update_staged": (c_int, [c_vp, c_vp, c_vp, c_u32, c_d, c_d, c_d, c_d, c_vp]),
    "amcl3d_cuda_pf_update": (c_int, [c_vp, c_vp, c_vp, c_u64, c_vp, c_u32, c_d, c_d, c_d, c_d, c_vp]),
    "amcl3d_cuda_pf_get_mean": (c_int, [c_vp, c_vp]),
    "amcl3d_cuda_pf_mean_exact_mask": (c_int, [c_vp, _P(c_u32)]),
    "amcl3d_cuda_pf_last_cloud_weights": (c_int, [c_vp, c_vp, c_vp]),
    "amcl3d_cuda_pf_last_in_map_evals": (c_int, [c_vp, _P(c_u64)]),
    "amcl3d_cuda_pf_resample": (c_int, [c_vp, c_f, c_vp]),
    "amcl3d_cuda_comm_unique_id": (c_int, [c_vp]),
    "amcl3d_cuda_comm_init": (c_int, [c_vp, c_vp, c_int, c_int]),
    "amcl3d_cuda_comm_destroy": (c_int, [c_vp]),
    "amcl3d_cuda_comm_rank": (c_int, [c_vp, _P(c_int), _P(c_int)]),
    "amcl3d_cuda_comm_peer_active": (c_int, [c_vp, _P(c_int)]),
}


class Amcl3dCudaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("amcl3d_cuda error %d: %s" % (code, message))
        self.code = code


def library_path():
    return os.path.join(_HERE, "lib", "libamcl3d_cuda.so")


def load_library(build_if_missing=True):
    """Loads the C-ABI shared library and declares every prototype.  Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(path)
        from . import build
        build.build_cuda()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _check(rc):
    if rc != 0:
        raise Amcl3dCudaError(rc, load_library().amcl3d_cuda_last_error().decode(errors="replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(-1, cols) if cols else a


def as_xyzw(points):
    """n x 3 or n x 4 -> contiguous n x 4 float32 (pcl::PointXYZ layout)."""
    p = np.asarray(points, dtype=np.float32)
    if p.ndim == 2 and p.shape[1] == 4:
        return np.ascontiguousarray(p)
    p = p.reshape(-1, 3) if p.ndim != 2 else p
    out = np.zeros((p.shape[0], 4), np.float32)
    out[:, :3] = p[:, :3]
    out[:, 3] = 1.0
    return out


class Context:
    """amcl3d_cuda_ctx: one device, one stream.  `stream` may be a raw cudaStream_t (int), e.g.
    torch.cuda.current_stream().cuda_stream, so the caller can bracket calls with its own events."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = c_vp()
        _check(self.lib.amcl3d_cuda_ctx_create(int(device), c_vp(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.amcl3d_cuda_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        _check(self.lib.amcl3d_cuda_ctx_set_stream(self.h, c_vp(stream) if stream else None))

    def synchronize(self):
        _check(self.lib.amcl3d_cuda_ctx_synchronize(self.h))

    def device_info(self):
        info = np.zeros(4, np.int64)
        _check(self.lib.amcl3d_cuda_ctx_device_info(self.h, _ptr(info)))
        return {"sm_count": int(info[0]), "l2_bytes": int(info[1]), "l2_persist_max": int(info[2]), "cc": int(info[3])}

    def set_option(self, name, value):
        _check(self.lib.amcl3d_cuda_ctx_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name):
        v = c_i64()
        _check(self.lib.amcl3d_cuda_ctx_get_option(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def last_kernel_ms(self):
        ms = c_f()
        _check(self.lib.amcl3d_cuda_ctx_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def last_update_phases_ms(self):
        """(weighting, exchange / wait, sums over particles) device times of the last update, option kernel_timing = 1."""
        ms = (c_f * 3)()
        _check(self.lib.amcl3d_cuda_ctx_last_update_phases_ms(self.h, ms))
        return [float(v) for v in ms]

    def probe_gather(self, footprint_bytes, lanes_per_sector=1):
        """Random 4-byte gather roofline of the device: (sector GB/s, warp requests/s)."""
        gbs, req = C.c_double(0), C.c_double(0)
        _check(self.lib.amcl3d_cuda_probe_gather(self.h, int(footprint_bytes), int(lanes_per_sector), C.byref(gbs),
                                                 C.byref(req)))
        return gbs.value, req.value

    def voxel_grid(self, cloud, leaf):
        """pcl::VoxelGrid down-sampling of a sensor cloud on the device (Node.cpp:131-137): m x 4 float32."""
        cl = as_xyzw(cloud)
        leaf3 = [float(np.float32(v)) for v in (leaf if np.ndim(leaf) else (leaf, leaf, leaf))]
        out = np.zeros((max(len(cl), 1), 4), np.float32)
        m = c_u64(0)
        _check(self.lib.amcl3d_cuda_voxel_grid(self.h, _ptr(cl), len(cl), *leaf3, _ptr(out), len(out), C.byref(m)))
        return out[:int(m.value)].copy()

    def launch_count(self):
        n = c_u64()
        _check(self.lib.amcl3d_cuda_ctx_launch_count(self.h, C.byref(n)))
        return int(n.value)

    # ---- multi-GPU
    def unique_id(self):
        buf = np.zeros(128, np.uint8)
        _check(self.lib.amcl3d_cuda_comm_unique_id(_ptr(buf)))
        return buf

    def comm_init(self, unique_id, rank, n_ranks):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        _check(self.lib.amcl3d_cuda_comm_init(self.h, _ptr(uid), int(rank), int(n_ranks)))

    def comm_destroy(self):
        _check(self.lib.amcl3d_cuda_comm_destroy(self.h))

    def comm_rank(self):
        r, n = c_int(), c_int()
        _check(self.lib.amcl3d_cuda_comm_rank(self.h, C.byref(r), C.byref(n)))
        return int(r.value), int(n.value)

    def comm_peer_active(self):
        v = c_int(0)
        _check(self.lib.amcl3d_cuda_comm_peer_active(self.h, C.byref(v)))
        return bool(v.value)


class Grid:
    """amcl3d_cuda_grid: the probability (and optionally distance) field in HBM."""

    def __init__(self, ctx, bounds7):
        self.ctx, self.lib = ctx, ctx.lib
        self.bounds = np.ascontiguousarray(bounds7, dtype=np.float64).reshape(7)
        h = c_vp()
        _check(self.lib.amcl3d_cuda_grid_create(ctx.h, _ptr(self.bounds), C.byref(h)))
        self.h = h
        d = np.zeros(3, np.uint32)
        _check(self.lib.amcl3d_cuda_grid_dims(self.h, _ptr(d)))
        self.dims = d
        self.n_cells = int(d[0]) * int(d[1]) * int(d[2])

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.amcl3d_cuda_grid_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_cells(self, cells, sensor_dev):
        c = _f32(cells)
        assert c.size == 2 * self.n_cells
        _check(self.lib.amcl3d_cuda_grid_upload_cells(self.h, _ptr(c), float(sensor_dev)))

    def download_cells(self):
        out = np.zeros((self.n_cells, 2), np.float32)
        _check(self.lib.amcl3d_cuda_grid_download_cells(self.h, _ptr(out)))
        return out

    def download_prob(self):
        out = np.zeros(self.n_cells, np.float32)
        _check(self.lib.amcl3d_cuda_grid_download_prob(self.h, _ptr(out)))
        return out

    def download_prob_range(self, first, count):
        out = np.zeros(int(count), np.float32)
        _check(self.lib.amcl3d_cuda_grid_download_prob_range(self.h, int(first), int(count), _ptr(out)))
        return out

    def gather_prob(self, idx):
        """Probabilities at logical voxel indices (0xFFFFFFFF -> 0)."""
        i = np.ascontiguousarray(idx, dtype=np.uint32)
        out = np.zeros(len(i), np.float32)
        _check(self.lib.amcl3d_cuda_grid_gather_prob(self.h, _ptr(i), len(i), _ptr(out)))
        return out

    def has_cells(self):
        r = c_int()
        _check(self.lib.amcl3d_cuda_grid_has_cells(self.h, C.byref(r)))
        return bool(r.value)

    def compute(self, points, sensor_dev, keep_dist=True):
        pts = as_xyzw(points)
        _check(self.lib.amcl3d_cuda_grid_compute(self.h, _ptr(pts), len(pts), float(sensor_dev), 1 if keep_dist else 0))

    def cloud_weight(self, cloud, pose6, want_idx=False):
        cl = as_xyzw(cloud)
        w, n = c_f(), c_u32()
        idx = np.zeros(len(cl), np.uint32) if want_idx else None
        a = [float(np.float32(v)) for v in pose6]
        _check(self.lib.amcl3d_cuda_cloud_weight(self.h, _ptr(cl), len(cl), *a, C.byref(w), C.byref(n), _ptr(idx)))
        return (np.float32(w.value), int(n.value), idx) if want_idx else (np.float32(w.value), int(n.value))

    def cloud_weight_batch(self, cloud, poses_xyza, roll, pitch):
        cl = as_xyzw(cloud)
        poses = _f32(poses_xyza, 4)
        w = np.zeros(len(poses), np.float32)
        n = np.zeros(len(poses), np.uint32)
        _check(self.lib.amcl3d_cuda_cloud_weight_batch(self.h, _ptr(cl), len(cl), _ptr(poses), len(poses),
                                                       float(np.float32(roll)), float(np.float32(pitch)), _ptr(w), _ptr(n)))
        return w, n

    def is_into_map(self, x, y, z):
        r = c_int()
        _check(self.lib.amcl3d_cuda_is_into_map(self.h, float(np.float32(x)), float(np.float32(y)), float(np.float32(z)),
                                                C.byref(r)))
        return bool(r.value)


class Filter:
    """amcl3d_cuda_pf: device-resident particle set + predict / update / resample."""

    def __init__(self, ctx):
        self.ctx, self.lib = ctx, ctx.lib
        h = c_vp()
        _check(self.lib.amcl3d_cuda_pf_create(ctx.h, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.amcl3d_cuda_pf_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        n = c_u64()
        _check(self.lib.amcl3d_cuda_pf_size(self.h, C.byref(n)))
        return int(n.value)

    def upload(self, particles7):
        p = _f32(particles7, 7)
        _check(self.lib.amcl3d_cuda_pf_upload_particles(self.h, _ptr(p), len(p)))

    def download(self):
        out = np.zeros((self.size(), 7), np.float32)
        _check(self.lib.amcl3d_cuda_pf_download_particles(self.h, _ptr(out)))
        return out

    def init(self, n, pose4, devs4, noise_n4=None, seed=0):
        pose = _f32(pose4)
        devs = _f32(devs4)
        nz = None if noise_n4 is None else _f32(noise_n4, 4)
        mean = np.zeros(4, np.float32)
        _check(self.lib.amcl3d_cuda_pf_init(self.h, int(n), _ptr(pose), _ptr(devs), _ptr(nz), int(seed), _ptr(mean)))
        return mean

    def predict(self, mods4, deltas4, noise_n4=None, seed=0, step=0):
        m = np.ascontiguousarray(mods4, dtype=np.float64)
        d = np.ascontiguousarray(deltas4, dtype=np.float64)
        nz = None if noise_n4 is None else _f32(noise_n4, 4)
        _check(self.lib.amcl3d_cuda_pf_predict(self.h, _ptr(m), _ptr(d), _ptr(nz), int(seed), int(step)))

    def stage_cloud(self, cloud):
        cl = as_xyzw(cloud)
        self._cloud_keepalive = cl
        _check(self.lib.amcl3d_cuda_pf_stage_cloud(self.h, _ptr(cl), len(cl)))

    def update_staged(self, grid, ranges, alpha, sigma, roll, pitch, want_mean=True):
        r = _f32(ranges, 4) if ranges is not None and len(ranges) else None
        mean = np.zeros(4, np.float32) if want_mean else None
        _check(self.lib.amcl3d_cuda_pf_update_staged(self.h, grid.h, _ptr(r), 0 if r is None else len(r), float(alpha),
                                                     float(sigma), float(roll), float(pitch), _ptr(mean)))
        return mean

    def update(self, grid, cloud, ranges, alpha, sigma, roll, pitch):
        cl = as_xyzw(cloud)
        r = _f32(ranges, 4) if ranges is not None and len(ranges) else None
        mean = np.zeros(4, np.float32)
        _check(self.lib.amcl3d_cuda_pf_update(self.h, grid.h, _ptr(cl), len(cl), _ptr(r), 0 if r is None else len(r),
                                              float(alpha), float(sigma), float(roll), float(pitch), _ptr(mean)))
        return mean

    def mean(self):
        m = np.zeros(4, np.float32)
        _check(self.lib.amcl3d_cuda_pf_get_mean(self.h, _ptr(m)))
        return m

    def mean_exact_mask(self):
        """Bit k set: component k (x, y, z, yaw) of the last mean is the reference's sequential float sum bit for bit."""
        m = c_u32()
        _check(self.lib.amcl3d_cuda_pf_mean_exact_mask(self.h, C.byref(m)))
        return int(m.value)

    def last_cloud_weights(self):
        """Raw computeCloudWeight result and contributing-point count per particle of the last update."""
        n = self.size()
        w = np.zeros(n, np.float32)
        c = np.zeros(n, np.uint32)
        _check(self.lib.amcl3d_cuda_pf_last_cloud_weights(self.h, _ptr(w), _ptr(c)))
        return w, c

    def last_in_map_evals(self):
        n = c_u64()
        _check(self.lib.amcl3d_cuda_pf_last_in_map_evals(self.h, C.byref(n)))
        return int(n.value)

    def resample(self, u01, want_idx=False):
        idx = np.zeros(self.size(), np.uint32) if want_idx else None
        _check(self.lib.amcl3d_cuda_pf_resample(self.h, float(np.float32(u01)), _ptr(idx)))
        return idx
