"""ctypes bindings for the two CPU checkers built by oracle/Makefile.

* ``Port``      -> oracle/_ref/libamcl3d_oracle.so : plain-C restatement (oracle/amcl3d_oracle.c)
* ``Reference`` -> oracle/_ref/libamcl3d_ref.so    : the UNMODIFIED reference C++ sources compiled against
  stand-in headers, driven through tests/harness/class_harness.cpp

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "_ref")

c_f, c_d, c_u32, c_u64, c_i64, c_vp = C.c_float, C.c_double, C.c_uint32, C.c_uint64, C.c_int64, C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def as_xyzw(points):
    """n x 3 or n x 4 -> contiguous n x 4 float32 (pcl::PointXYZ layout)."""
    p = np.asarray(points, dtype=np.float32)
    if p.ndim != 2:
        p = p.reshape(-1, 4)
    if p.shape[1] == 4:
        return np.ascontiguousarray(p)
    out = np.zeros((p.shape[0], 4), np.float32)
    out[:, :3] = p[:, :3]
    out[:, 3] = 1.0
    return out


def build(target="all", quiet=True):
    """Runs oracle/Makefile.  `ref` needs /root/reference (only present in the authoring container)."""
    cmd = ["make", "-C", _HERE, target]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and not quiet:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return r.returncode == 0


def port_path():
    return os.path.join(_OUT, "libamcl3d_oracle.so")


def reference_path():
    return os.path.join(_OUT, "libamcl3d_ref.so")


class Port:
    """The plain-C restatement."""

    def __init__(self, path=None):
        path = path or port_path()
        if not os.path.exists(path):
            build("port", quiet=False)
        self.lib = L = C.CDLL(path)
        L.oracle_grid_dims.argtypes = [c_vp, c_vp]
        L.oracle_compute_grid.argtypes = [c_vp, c_u64, c_vp, c_d, c_vp, c_u32, c_u32, c_u64]
        L.oracle_compute_grid.restype = C.c_int
        L.oracle_nn_dist2_bruteforce.argtypes = [c_vp, c_u64, c_vp, c_u32, c_u32, c_u32]
        L.oracle_nn_dist2_bruteforce.restype = c_f
        L.oracle_cloud_weight.argtypes = [c_vp, c_vp, c_vp, c_vp, c_u64] + [c_f] * 6 + [c_vp, c_vp]
        L.oracle_cloud_weight.restype = c_f
        L.oracle_is_into_map.argtypes = [c_vp, c_f, c_f, c_f]
        L.oracle_is_into_map.restype = C.c_int
        L.oracle_range_weight.argtypes = [c_f, c_f, c_f, c_vp, c_u32, c_d]
        L.oracle_range_weight.restype = c_f
        L.oracle_update.argtypes = [c_vp, c_u64, c_vp, c_vp, c_vp, c_vp, c_u64, c_vp, c_u32, c_d, c_d, c_d, c_d, c_vp]
        L.oracle_update_from_weights.argtypes = [c_vp, c_u64, c_vp, c_d, c_vp]
        L.oracle_cloud_weight_batch.argtypes = [c_vp, c_vp, c_vp, c_vp, c_u64, c_vp, c_u64, c_f, c_f, c_vp, c_vp]
        L.oracle_resample.argtypes = [c_vp, c_u64, c_f, c_vp]
        L.oracle_predict.argtypes = [c_vp, c_u64, c_vp, c_vp, c_vp]
        L.oracle_init.argtypes = [c_vp, c_u64] + [c_f] * 8 + [c_vp, c_vp]
        L.oracle_grid_slice.argtypes = [c_vp, c_vp, c_vp, c_d, c_vp, c_u64]
        L.oracle_grid_slice.restype = c_i64
        L.oracle_voxel_grid.argtypes = [c_vp, c_u64, c_f, c_f, c_f, c_vp]
        L.oracle_voxel_grid.restype = c_i64

    @staticmethod
    def _bounds(b):
        return np.ascontiguousarray(b, dtype=np.float64).reshape(7)

    def grid_dims(self, bounds7):
        b = self._bounds(bounds7)
        d = np.zeros(3, np.uint32)
        self.lib.oracle_grid_dims(_ptr(b), _ptr(d))
        return d

    def compute_grid(self, points, bounds7, sensor_dev, z_range=None, max_cells=250000000, cells=None):
        pts = as_xyzw(points)
        b = self._bounds(bounds7)
        dims = self.grid_dims(b)
        total = int(dims[0]) * int(dims[1]) * int(dims[2])
        if cells is None:
            cells = np.zeros((total, 2), np.float32)
            cells[:, 0] = -1.0
        z0, z1 = (0, int(dims[2])) if z_range is None else z_range
        rc = self.lib.oracle_compute_grid(_ptr(pts), len(pts), _ptr(b), float(sensor_dev), _ptr(cells), z0, z1,
                                          int(max_cells))
        if rc != 0:
            raise RuntimeError("Octomap size is too big. Grid size over 2Gb." if rc == -1 else "oracle alloc failure")
        return cells, dims

    def nn_dist2_bruteforce(self, points, bounds7, ix, iy, iz):
        pts = as_xyzw(points)
        b = self._bounds(bounds7)
        return float(self.lib.oracle_nn_dist2_bruteforce(_ptr(pts), len(pts), _ptr(b), ix, iy, iz))

    def cloud_weight(self, cells, dims, bounds7, cloud, pose6, want_idx=False):
        cl = as_xyzw(cloud)
        b = self._bounds(bounds7)
        d = np.ascontiguousarray(dims, dtype=np.uint32)
        cells = _f32(cells)
        idx = np.zeros(len(cl), np.uint32) if want_idx else None
        n = np.zeros(1, np.uint32)
        tx, ty, tz, roll, pitch, yaw = [float(np.float32(v)) for v in pose6]
        w = self.lib.oracle_cloud_weight(_ptr(cells), _ptr(d), _ptr(b), _ptr(cl), len(cl), tx, ty, tz, roll, pitch, yaw,
                                         _ptr(idx), _ptr(n))
        return (np.float32(w), int(n[0]), idx) if want_idx else (np.float32(w), int(n[0]))

    def is_into_map(self, bounds7, x, y, z):
        return bool(self.lib.oracle_is_into_map(_ptr(self._bounds(bounds7)), float(np.float32(x)), float(np.float32(y)),
                                                float(np.float32(z))))

    def range_weight(self, x, y, z, ranges, sigma):
        r = _f32(ranges, 4)
        return np.float32(self.lib.oracle_range_weight(float(np.float32(x)), float(np.float32(y)), float(np.float32(z)),
                                                       _ptr(r), len(r), float(sigma)))

    def update(self, particles, cells, dims, bounds7, cloud, ranges, alpha, sigma, roll, pitch):
        p = _f32(particles, 7).copy()
        cl = as_xyzw(cloud)
        r = _f32(ranges, 4)
        mean = np.zeros(4, np.float32)
        self.lib.oracle_update(_ptr(p), len(p), _ptr(_f32(cells)), _ptr(np.ascontiguousarray(dims, dtype=np.uint32)),
                               _ptr(self._bounds(bounds7)), _ptr(cl), len(cl), _ptr(r), len(r), float(alpha),
                               float(sigma), float(roll), float(pitch), _ptr(mean))
        return p, mean

    def update_from_weights(self, particles, bounds7, alpha):
        """Loops 2 and 3 (+ the chain totals of loop 1) of update() on particles whose wp / wr hold raw weights."""
        p = _f32(particles, 7).copy()
        mean = np.zeros(4, np.float32)
        self.lib.oracle_update_from_weights(_ptr(p), len(p), _ptr(self._bounds(bounds7)), float(alpha), _ptr(mean))
        return p, mean

    def cloud_weight_batch(self, cells, dims, bounds7, cloud, poses_xyza, roll, pitch):
        """computeCloudWeight for many poses (OpenMP over poses): (weights, contributing-point counts)."""
        cl = as_xyzw(cloud)
        poses = _f32(poses_xyza, 4)
        w = np.zeros(len(poses), np.float32)
        n = np.zeros(len(poses), np.uint32)
        self.lib.oracle_cloud_weight_batch(_ptr(None if cells is None else _f32(cells)),
                                           _ptr(np.ascontiguousarray(dims, dtype=np.uint32)), _ptr(self._bounds(bounds7)),
                                           _ptr(cl), len(cl), _ptr(poses), len(poses), float(np.float32(roll)),
                                           float(np.float32(pitch)), _ptr(w), _ptr(n))
        return w, n

    def cloud_indices(self, dims, bounds7, cloud, pose6):
        """Voxel index per cloud point (0xFFFFFFFF where the reference skips it) without touching any cell."""
        cl = as_xyzw(cloud)
        idx = np.zeros(len(cl), np.uint32)
        n = np.zeros(1, np.uint32)
        tx, ty, tz, roll, pitch, yaw = [float(np.float32(v)) for v in pose6]
        self.lib.oracle_cloud_weight(None, _ptr(np.ascontiguousarray(dims, dtype=np.uint32)), _ptr(self._bounds(bounds7)),
                                     _ptr(cl), len(cl), tx, ty, tz, roll, pitch, yaw, _ptr(idx), _ptr(n))
        return idx, int(n[0])

    def resample(self, particles, u01):
        p = _f32(particles, 7).copy()
        idx = np.zeros(len(p), np.uint32)
        self.lib.oracle_resample(_ptr(p), len(p), float(np.float32(u01)), _ptr(idx))
        return p, idx

    def predict(self, particles, mods4, deltas4, noise_n4):
        p = _f32(particles, 7).copy()
        m = np.ascontiguousarray(mods4, dtype=np.float64)
        d = np.ascontiguousarray(deltas4, dtype=np.float64)
        nz = _f32(noise_n4, 4)
        assert len(nz) == len(p)
        self.lib.oracle_predict(_ptr(p), len(p), _ptr(m), _ptr(d), _ptr(nz))
        return p

    def init(self, n, pose4, devs4, noise_n4):
        p = np.zeros((n, 7), np.float32)
        nz = _f32(noise_n4, 4)
        mean = np.zeros(4, np.float32)
        a = [float(np.float32(v)) for v in list(pose4) + list(devs4)]
        self.lib.oracle_init(_ptr(p), n, *a, _ptr(nz), _ptr(mean))
        return p, mean

    def grid_slice(self, cells, dims, bounds7, z):
        d = np.ascontiguousarray(dims, dtype=np.uint32)
        out = np.zeros(int(d[0]) * int(d[1]) + 16, np.int8)
        n = self.lib.oracle_grid_slice(_ptr(_f32(cells)), _ptr(d), _ptr(self._bounds(bounds7)), float(z), _ptr(out),
                                       len(out))
        return None if n < 0 else out[:min(n, len(out))]


    def voxel_grid(self, cloud, leaf):
        """pcl::VoxelGrid restatement: returns the down-sampled cloud (m x 4) or None for PCL's pass-through case."""
        cl = as_xyzw(cloud)
        leaf3 = [float(np.float32(v)) for v in (leaf if np.ndim(leaf) else (leaf, leaf, leaf))]
        out = np.zeros((max(len(cl), 1), 4), np.float32)
        m = self.lib.oracle_voxel_grid(_ptr(cl), len(cl), *leaf3, _ptr(out))
        return None if m < 0 else out[:m].copy()


class ClassHarness:
    """ctypes view of tests/harness/class_harness.cpp -- identical for the reference build and the B200 host build."""

    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        L.h_impl_name.restype = C.c_char_p
        L.h_grid_new.restype = c_vp
        L.h_grid_free.argtypes = [c_vp]
        L.h_grid_open.argtypes = [c_vp, C.c_char_p, c_d]
        L.h_grid_open_from_cloud.argtypes = [c_vp, c_vp, c_u64, c_vp, c_d]
        L.h_grid_set_cells.argtypes = [c_vp, c_vp, c_u64, c_vp, c_d, c_vp, c_vp]
        L.h_grid_dims.argtypes = [c_vp, c_vp]
        L.h_grid_get_cells.argtypes = [c_vp, c_vp]
        L.h_grid_set_cloud.argtypes = [c_vp, c_vp, c_u64]
        L.h_grid_cloud_weight.argtypes = [c_vp] + [c_f] * 6
        L.h_grid_cloud_weight.restype = c_f
        L.h_grid_is_into_map.argtypes = [c_vp, c_f, c_f, c_f]
        L.h_grid_slice.argtypes = [c_vp, c_d, c_vp, c_u64, c_vp]
        L.h_grid_slice.restype = c_i64
        L.h_grid_map_cloud.argtypes = [c_vp, c_vp, c_u64]
        L.h_grid_map_cloud.restype = c_i64
        L.h_grid_map_info.argtypes = [c_vp, c_vp]
        L.h_grid_map_info.restype = c_i64
        L.h_tools_load_octomap.argtypes = [C.c_char_p, c_vp, c_vp, c_u64, C.c_char_p, c_u64]
        L.h_tools_load_octomap.restype = c_i64
        L.h_tools_write_octomap.argtypes = [C.c_char_p, c_vp, c_u64, c_vp, c_vp, c_u64, c_d, C.c_int]
        L.h_set_option.argtypes = [C.c_char_p, c_i64]
        L.h_pf_new.restype = c_vp
        L.h_pf_free.argtypes = [c_vp]
        L.h_pf_seed.argtypes = [c_vp, c_u32]
        L.h_pf_is_initialized.argtypes = [c_vp]
        L.h_pf_init.argtypes = [c_vp, C.c_int] + [c_f] * 8
        L.h_pf_size.argtypes = [c_vp]
        L.h_pf_size.restype = c_u64
        L.h_pf_set_particles.argtypes = [c_vp, c_vp, c_u64]
        L.h_pf_get_particles.argtypes = [c_vp, c_vp]
        L.h_pf_get_mean.argtypes = [c_vp, c_vp]
        L.h_pf_predict.argtypes = [c_vp, c_vp, c_vp]
        L.h_pf_update.argtypes = [c_vp, c_vp, c_vp, c_u32, c_d, c_d, c_d, c_d]
        L.h_pf_resample.argtypes = [c_vp]
        L.h_pf_pose_msg.argtypes = [c_vp, c_vp, c_u64]
        L.h_pf_pose_msg.restype = c_u64
        L.h_rng_new.argtypes = [c_u32]
        L.h_rng_new.restype = c_vp
        L.h_rng_free.argtypes = [c_vp]
        L.h_rng_gaussian.argtypes = [c_vp, c_d, c_d]
        L.h_rng_gaussian.restype = c_f
        L.h_rng_uniform01.argtypes = [c_vp]
        L.h_rng_uniform01.restype = c_f
        L.h_rng_predict_noise.argtypes = [c_vp, c_u64, c_vp, c_vp, c_vp]
        L.h_time_update.argtypes = [c_vp, c_vp, c_vp, c_u32, c_d, c_d, c_d, c_d, C.c_int]
        L.h_time_update.restype = c_d

    @property
    def impl(self):
        return self.lib.h_impl_name().decode()

    def math_overloads_are_double(self):
        return bool(self.lib.h_math_overloads_are_double())

    def grid(self):
        return HGrid(self)

    def filter(self):
        return HFilter(self)

    def rng(self, seed):
        return HRng(self, seed)

    def load_octomap(self, path, cap=1 << 22):
        b = np.zeros(7)
        pts = np.zeros((cap, 4), np.float32)
        err = C.create_string_buffer(512)
        n = self.lib.h_tools_load_octomap(path.encode(), _ptr(b), _ptr(pts), cap, err, 512)
        if n < 0:
            raise RuntimeError(err.value.decode())
        return pts[:min(n, cap)].copy(), b

    def write_octomap(self, path, points, res, depths=None, free_points=None, as_ot=False):
        pts = as_xyzw(points)
        d = None if depths is None else np.ascontiguousarray(depths, dtype=np.uint8)
        fp = as_xyzw(free_points) if free_points is not None else np.zeros((0, 4), np.float32)
        return bool(self.lib.h_tools_write_octomap(path.encode(), _ptr(pts), len(pts), _ptr(d), _ptr(fp), len(fp),
                                                   float(res), 1 if as_ot else 0))

    def set_option(self, name, value):
        return int(self.lib.h_set_option(name.encode(), int(value)))

    def node_voxel_filter(self, cloud, voxel_size):
        """Node.cpp:131-137 through the harness' pcl::VoxelGrid (B200 build: device-backed stand-in)."""
        cl = as_xyzw(cloud)
        out = np.zeros((max(len(cl), 1), 4), np.float32)
        self.lib.h_node_voxel_filter.restype = c_i64
        self.lib.h_node_voxel_filter.argtypes = [c_vp, c_u64, c_d, c_vp]
        m = self.lib.h_node_voxel_filter(_ptr(cl), len(cl), float(voxel_size), _ptr(out))
        if m < 0:
            raise RuntimeError("h_node_voxel_filter returned %d" % m)
        return out[:m].copy()

    def null_tree_throws(self):
        return bool(self.lib.h_tools_null_tree_throws())


class HGrid:
    def __init__(self, h):
        self.h, self.L = h, h.lib
        self.g = c_vp(self.L.h_grid_new())

    def __del__(self):
        try:
            self.L.h_grid_free(self.g)
        except Exception:
            pass

    def open(self, path, sensor_dev):
        return bool(self.L.h_grid_open(self.g, path.encode(), float(sensor_dev)))

    def open_from_cloud(self, points, bounds7, sensor_dev):
        pts = as_xyzw(points)
        b = np.ascontiguousarray(bounds7, dtype=np.float64)
        return bool(self.L.h_grid_open_from_cloud(self.g, _ptr(pts), len(pts), _ptr(b), float(sensor_dev)))

    def set_cells(self, points, bounds7, sensor_dev, dims, cells):
        pts = as_xyzw(points)
        b = np.ascontiguousarray(bounds7, dtype=np.float64)
        d = np.ascontiguousarray(dims, dtype=np.uint32)
        c = _f32(cells)
        return bool(self.L.h_grid_set_cells(self.g, _ptr(pts), len(pts), _ptr(b), float(sensor_dev), _ptr(d), _ptr(c)))

    def dims(self):
        d = np.zeros(3, np.uint32)
        return d if self.L.h_grid_dims(self.g, _ptr(d)) else None

    def cells(self):
        d = self.dims()
        if d is None:
            return None
        out = np.zeros((int(d[0]) * int(d[1]) * int(d[2]), 2), np.float32)
        self.L.h_grid_get_cells(self.g, _ptr(out))
        return out

    def set_cloud(self, cloud):
        cl = as_xyzw(cloud)
        self.L.h_grid_set_cloud(self.g, _ptr(cl), len(cl))

    def cloud_weight(self, tx, ty, tz, roll, pitch, yaw):
        a = [float(np.float32(v)) for v in (tx, ty, tz, roll, pitch, yaw)]
        return np.float32(self.L.h_grid_cloud_weight(self.g, *a))

    def is_into_map(self, x, y, z):
        return bool(self.L.h_grid_is_into_map(self.g, float(np.float32(x)), float(np.float32(y)), float(np.float32(z))))

    def slice(self, z, cap=1 << 24):
        out = np.zeros(cap, np.int8)
        info = np.zeros(4)
        n = self.L.h_grid_slice(self.g, float(z), _ptr(out), cap, _ptr(info))
        return (None, None) if n < 0 else (out[:min(n, cap)].copy(), info)

    def map_cloud(self, cap=1 << 22):
        out = np.zeros((cap, 4), np.float32)
        n = self.L.h_grid_map_cloud(self.g, _ptr(out), cap)
        return None if n < 0 else out[:min(n, cap)].copy()

    def map_info(self):
        b = np.zeros(7)
        n = self.L.h_grid_map_info(self.g, _ptr(b))
        return (None, None) if n < 0 else (int(n), b)


class HFilter:
    def __init__(self, h):
        self.h, self.L = h, h.lib
        self.p = c_vp(self.L.h_pf_new())

    def __del__(self):
        try:
            self.L.h_pf_free(self.p)
        except Exception:
            pass

    def seed(self, s):
        self.L.h_pf_seed(self.p, int(s))

    def is_initialized(self):
        return bool(self.L.h_pf_is_initialized(self.p))

    def init(self, n, pose4, devs4):
        a = [float(np.float32(v)) for v in list(pose4) + list(devs4)]
        self.L.h_pf_init(self.p, int(n), *a)

    def size(self):
        return int(self.L.h_pf_size(self.p))

    def set_particles(self, particles):
        p = _f32(particles, 7)
        self.L.h_pf_set_particles(self.p, _ptr(p), len(p))

    def particles(self):
        out = np.zeros((self.size(), 7), np.float32)
        self.L.h_pf_get_particles(self.p, _ptr(out))
        return out

    def mean(self):
        out = np.zeros(7, np.float32)
        self.L.h_pf_get_mean(self.p, _ptr(out))
        return out

    def predict(self, mods4, deltas4):
        m = np.ascontiguousarray(mods4, dtype=np.float64)
        d = np.ascontiguousarray(deltas4, dtype=np.float64)
        self.L.h_pf_predict(self.p, _ptr(m), _ptr(d))

    def update(self, grid, ranges, alpha, sigma, roll, pitch):
        r = _f32(ranges, 4)
        self.L.h_pf_update(self.p, grid.g, _ptr(r), len(r), float(alpha), float(sigma), float(roll), float(pitch))

    def time_update(self, grid, ranges, alpha, sigma, roll, pitch, reps=1):
        r = _f32(ranges, 4)
        return float(self.L.h_time_update(self.p, grid.g, _ptr(r), len(r), float(alpha), float(sigma), float(roll),
                                          float(pitch), int(reps)))

    def resample(self):
        self.L.h_pf_resample(self.p)

    def pose_msg(self):
        n = self.size()
        out = np.zeros((n, 7))
        self.L.h_pf_pose_msg(self.p, _ptr(out), n)
        return out


class HRng:
    """std::mt19937 + the reference's per-call distributions (ParticleFilter.cpp:246-256)."""

    def __init__(self, h, seed):
        self.L = h.lib
        self.r = c_vp(self.L.h_rng_new(int(seed)))

    def __del__(self):
        try:
            self.L.h_rng_free(self.r)
        except Exception:
            pass

    def gaussian(self, mean, sigma):
        return np.float32(self.L.h_rng_gaussian(self.r, float(mean), float(sigma)))

    def uniform01(self):
        return np.float32(self.L.h_rng_uniform01(self.r))

    def predict_noise(self, n, mods4, deltas4):
        m = np.ascontiguousarray(mods4, dtype=np.float64)
        d = np.ascontiguousarray(deltas4, dtype=np.float64)
        out = np.zeros((n, 4), np.float32)
        self.L.h_rng_predict_noise(self.r, n, _ptr(m), _ptr(d), _ptr(out))
        return out

    def init_noise(self, n, devs4):
        """init()'s draw order: particles 1..n-1, x, y, z, a each (ParticleFilter.cpp:69-72); row 0 unused."""
        out = np.zeros((n, 4), np.float32)
        for i in range(1, n):
            for k in range(4):
                out[i, k] = self.gaussian(0, float(np.float32(devs4[k])))
        return out


def HostBuild(path=None):
    """This repo's B200 host classes behind the same harness (amcl3d_b200/lib/libamcl3d_host.so)."""
    if path is None:
        path = os.path.join(os.path.dirname(_HERE), "amcl3d_b200", "lib", "libamcl3d_host.so")
    return ClassHarness(path)


def Reference(path=None):
    """The unmodified reference sources behind the class harness (oracle/_ref/libamcl3d_ref.so)."""
    path = path or reference_path()
    if not os.path.exists(path):
        if not os.path.isdir("/root/reference/amcl3d/src"):
            raise FileNotFoundError(path + " missing and /root/reference not available to build it")
        build("ref", quiet=False)
    return ClassHarness(path)
