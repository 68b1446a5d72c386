"""Seeded synthetic inputs for the BASELINE.json configurations.

Everything is defined at the PointCloudInfo level (occupied-leaf centres + metric bounds + resolution), which
is exactly what computePointCloud hands to computeGrid (PointCloudTools.cpp:51-82), so the CPU oracle and the
CUDA path consume identical inputs.  Pure numpy; deterministic for a given seed.
"""
import numpy as np


# ----------------------------------------------------------------------------------------------- maps
def _box_shell(lo, hi):
    """Voxel indices (n x 3, int64) of the 1-voxel-thick shell of the integer box [lo, hi) ."""
    lo = np.asarray(lo, np.int64)
    hi = np.asarray(hi, np.int64)
    parts = []
    for axis in range(3):
        others = [a for a in range(3) if a != axis]
        r0 = np.arange(lo[others[0]], hi[others[0]])
        r1 = np.arange(lo[others[1]], hi[others[1]])
        g0, g1 = np.meshgrid(r0, r1, indexing="ij")
        for face in (lo[axis], hi[axis] - 1):
            idx = np.empty((g0.size, 3), np.int64)
            idx[:, axis] = face
            idx[:, others[0]] = g0.ravel()
            idx[:, others[1]] = g1.ravel()
            parts.append(idx)
    return np.concatenate(parts, axis=0)


def _plane(axis, value, dims):
    others = [a for a in range(3) if a != axis]
    g0, g1 = np.meshgrid(np.arange(dims[others[0]]), np.arange(dims[others[1]]), indexing="ij")
    idx = np.empty((g0.size, 3), np.int64)
    idx[:, axis] = value
    idx[:, others[0]] = g0.ravel()
    idx[:, others[1]] = g1.ravel()
    return idx


def _finish_map(voxels, dims, vmin, res):
    """unique voxels -> leaf-centre points (float32, as static_cast<float>(it.getX()) would give) + bounds7."""
    dims = np.asarray(dims, np.int64)
    v = voxels[(voxels >= 0).all(1) & (voxels < dims).all(1)]
    lin = v[:, 0] + dims[0] * (v[:, 1] + dims[1] * v[:, 2])
    lin = np.unique(lin)
    ix = lin % dims[0]
    iy = (lin // dims[0]) % dims[1]
    iz = lin // (dims[0] * dims[1])
    pts = np.empty((lin.size, 3), np.float32)
    pts[:, 0] = (vmin[0] + (ix + 0.5) * res).astype(np.float32)
    pts[:, 1] = (vmin[1] + (iy + 0.5) * res).astype(np.float32)
    pts[:, 2] = (vmin[2] + (iz + 0.5) * res).astype(np.float32)
    bounds = np.array([vmin[0], vmin[1], vmin[2], vmin[0] + dims[0] * res, vmin[1] + dims[1] * res,
                       vmin[2] + dims[2] * res, res], np.float64)
    return pts, bounds


def map_room(size=(20.0, 20.0, 5.0), res=0.1, n_boxes=40, seed=1):
    """Map S: a closed room (floor, ceiling, four walls, 1 voxel thick) with `n_boxes` random hollow boxes
    (edge 0.3-2 m).  Bounds: x,y centred on 0, z from 0.  Default = 200 x 200 x 50 voxels."""
    rng = np.random.default_rng(seed)
    dims = np.array([round(size[0] / res), round(size[1] / res), round(size[2] / res)], np.int64)
    vmin = np.array([-size[0] / 2, -size[1] / 2, 0.0])
    parts = [_plane(2, 0, dims), _plane(2, dims[2] - 1, dims), _plane(0, 0, dims), _plane(0, dims[0] - 1, dims),
             _plane(1, 0, dims), _plane(1, dims[1] - 1, dims)]
    for _ in range(n_boxes):
        edge = np.maximum(1, np.round(rng.uniform(0.3, 2.0, 3) / res)).astype(np.int64)
        lo = (rng.uniform(0, 1, 3) * (dims - edge)).astype(np.int64)
        lo[2] = 1 if rng.uniform() < 0.7 else lo[2]  # most boxes stand on the floor
        parts.append(_box_shell(lo, lo + edge))
    return _finish_map(np.concatenate(parts, 0), dims, vmin, res)


def map_warehouse(size=(100.0, 100.0, 20.0), res=0.05, n_pallets=1000, seed=5):
    """Map L: warehouse hall -- floor/ceiling/walls, shelving rows (1 m deep, 8 m tall, every 4 m, two runs along x
    separated by a central aisle, boards every 2 m) and random pallets.  Default = 2000 x 2000 x 400 voxels."""
    rng = np.random.default_rng(seed)
    dims = np.array([round(size[0] / res), round(size[1] / res), round(size[2] / res)], np.int64)
    vmin = np.array([-size[0] / 2, -size[1] / 2, 0.0])
    m = lambda metres: int(round(metres / res))  # noqa: E731
    parts = [_plane(2, 0, dims), _plane(2, dims[2] - 1, dims), _plane(0, 0, dims), _plane(0, dims[0] - 1, dims),
             _plane(1, 0, dims), _plane(1, dims[1] - 1, dims)]
    shelf_h = min(m(8.0), dims[2] - 2)
    run_len = int(dims[0] * 0.4)
    runs = [(int(dims[0] * 0.05), int(dims[0] * 0.05) + run_len), (int(dims[0] * 0.55), int(dims[0] * 0.55) + run_len)]
    y = m(2.0)
    while y + m(1.0) < dims[1] - m(1.0):
        for (x0, x1) in runs:
            parts.append(_box_shell((x0, y, 1), (x1, y + m(1.0), 1 + shelf_h)))
            for h in range(m(2.0), shelf_h, m(2.0)):
                gx, gy = np.meshgrid(np.arange(x0, x1), np.arange(y, y + m(1.0)), indexing="ij")
                board = np.empty((gx.size, 3), np.int64)
                board[:, 0] = gx.ravel()
                board[:, 1] = gy.ravel()
                board[:, 2] = 1 + h
                parts.append(board)
        y += m(4.0)
    for _ in range(n_pallets):
        edge = np.maximum(2, np.round(rng.uniform(0.8, 1.2, 3) / res)).astype(np.int64)
        lo = (rng.uniform(0, 1, 3) * (dims - edge)).astype(np.int64)
        lo[2] = 1
        parts.append(_box_shell(lo, lo + edge))
    return _finish_map(np.concatenate(parts, 0), dims, vmin, res)


# ----------------------------------------------------------------------------------------------- sensor data
def sensor_cloud(map_points, pose4, n_points, radius, seed=2, noise=0.02, leaf=0.1):
    """A body-frame cloud of `n_points` map-surface samples within `radius` of the pose, with N(0, noise) jitter,
    ordered like pcl::VoxelGrid output (ascending voxel index, x fastest)."""
    rng = np.random.default_rng(seed)
    t = np.asarray(pose4[:3], np.float64)
    yaw = float(pose4[3])
    d2 = ((map_points.astype(np.float64) - t) ** 2).sum(1)
    cand = np.nonzero(d2 <= radius * radius)[0]
    if cand.size == 0:
        raise ValueError("no map points within the sensor radius")
    pick = rng.choice(cand, size=n_points, replace=cand.size < n_points)
    world = map_points[pick].astype(np.float64) + rng.normal(0.0, noise, (n_points, 3))
    c, s = np.cos(yaw), np.sin(yaw)
    rel = world - t
    body = np.empty_like(rel)
    body[:, 0] = c * rel[:, 0] + s * rel[:, 1]
    body[:, 1] = -s * rel[:, 0] + c * rel[:, 1]
    body[:, 2] = rel[:, 2]
    v = np.floor(body / leaf).astype(np.int64)
    v -= v.min(0)
    span = v.max(0) + 1
    order = np.argsort(v[:, 0] + span[0] * (v[:, 1] + span[1] * v[:, 2]), kind="stable")
    out = np.zeros((n_points, 4), np.float32)
    out[:, :3] = body[order].astype(np.float32)
    out[:, 3] = 1.0
    return out


def beacons(pose4, positions=((-9.0, -9.0, 4.0), (9.0, -9.0, 4.0), (0.0, 9.0, 4.0)), seed=4, noise=0.1):
    """Range measurements (r, ax, ay, az) to fixed anchors: true distance + N(0, noise)."""
    rng = np.random.default_rng(seed)
    t = np.asarray(pose4[:3], np.float64)
    out = np.zeros((len(positions), 4), np.float32)
    for i, a in enumerate(positions):
        a = np.asarray(a, np.float64)
        out[i, 0] = np.float32(np.linalg.norm(t - a) + rng.normal(0.0, noise))
        out[i, 1:] = a.astype(np.float32)
    return out


def morton_order(cloud, cell=0.5):
    """Permutation that orders a cloud along a 3-D Morton (Z-order) curve of `cell`-sized boxes: consecutive points
    are spatial neighbours, which is what keeps the gather footprint of a point chunk inside L2 on large maps."""
    q = np.floor(np.asarray(cloud)[:, :3].astype(np.float64) / cell).astype(np.int64)
    q -= q.min(0)
    key = np.zeros(len(q), np.uint64)
    for bit in range(21):
        for a in range(3):
            key |= ((q[:, a].astype(np.uint64) >> np.uint64(bit)) & np.uint64(1)) << np.uint64(3 * bit + a)
    return np.argsort(key, kind="stable")


# ----------------------------------------------------------------------------------------------- particle sets
def particles_tracking(n, pose4, devs4, seed=3):
    """`init`-like set (ParticleFilter.cpp:58-72 semantics): particle 0 is the pose, the others pose + N(0, dev);
    uniform weights 1/n.  Returns n x 7 float32 (x, y, z, a, w, wp, wr)."""
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 7), np.float32)
    pose = np.asarray(pose4, np.float32)
    p[:, :4] = pose
    if n > 1:
        for k in range(4):
            p[1:, k] = pose[k] + rng.normal(0.0, devs4[k], n - 1).astype(np.float32)
    p[:, 4] = np.float32(1.0) / np.float32(n)
    return p


def particles_uniform(n, bounds7, seed=8):
    """Global-localisation set: uniform over the map volume and yaw in [-pi, pi)."""
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 7), np.float32)
    for k in range(3):
        p[:, k] = rng.uniform(bounds7[k], bounds7[3 + k], n).astype(np.float32)
    p[:, 3] = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    p[:, 4] = np.float32(1.0) / np.float32(n)
    return p


# ----------------------------------------------------------------------------------------------- named workloads
# canonical parameter values: amcl3d/launch/amcl3d.launch:79-114, amcl3d_rosin.launch:10-17, tests/ParticleFilterTest.cpp:41-58
DEFAULTS = dict(sensor_dev=0.05, sigma_range=0.53, alpha=0.5, voxel_size=0.1,
                odom_mods=(0.1, 0.1, 0.1, 0.3), deltas=(-0.067421, -0.006161, 0.130909, 0.052421))

WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own operating point
    "cfg1": dict(map="room", n_particles=600, n_points=2000, radius=8.0, pose=(0.0, 0.0, 2.5, 0.3),
                 devs=(0.05, 0.05, 0.05, 0.1), beacons=True),
    # configs[1]: same map, 10k x 10k
    "cfg2": dict(map="room", n_particles=10000, n_points=10000, radius=12.0, pose=(0.0, 0.0, 2.5, 0.3),
                 devs=(0.2, 0.2, 0.2, 0.4), beacons=True),
    # configs[3]: 1M x 32k on the warehouse map
    "cfg4": dict(map="warehouse", n_particles=1048576, n_points=32768, radius=30.0, pose=(-20.0, 0.0, 1.5, 0.2),
                 devs=(0.5, 0.5, 0.5, 0.2), beacons=False),
    # configs[4]: global localisation, 8M uniform x 64k
    "cfg5": dict(map="warehouse", n_particles=8388608, n_points=65536, radius=40.0, pose=(-20.0, 0.0, 1.5, 0.2),
                 devs=None, beacons=False),
}


def make_map(kind, **kw):
    """Builds (or, when AMCL3D_SYNTH_CACHE names a directory, re-loads) one of the named maps."""
    import os
    cache = os.environ.get("AMCL3D_SYNTH_CACHE")
    path = None
    if cache and not kw:
        path = os.path.join(cache, "amcl3d_map_%s.npz" % kind)
        if os.path.exists(path):
            d = np.load(path)
            return d["points"], d["bounds"]
    if kind == "room":
        out = map_room(**kw)
    elif kind == "warehouse":
        out = map_warehouse(**kw)
    else:
        raise ValueError(kind)
    if path:
        # several ranks may build the same map at once: write privately, publish atomically
        os.makedirs(cache, exist_ok=True)
        tmp = "%s.%d.tmp.npz" % (path, os.getpid())
        np.savez(tmp, points=out[0], bounds=out[1])
        os.replace(tmp, path)
    return out


def make_workload(name, n_particles=None, n_points=None, map_kwargs=None):
    """Returns a dict with map points/bounds, the sensor cloud, particle set, beacon ranges and scalar parameters."""
    w = dict(WORKLOADS[name])
    if n_particles is not None:
        w["n_particles"] = int(n_particles)
    if n_points is not None:
        w["n_points"] = int(n_points)
    pts, bounds = make_map(w["map"], **(map_kwargs or {}))
    pose = np.asarray(w["pose"], np.float64)
    cloud = sensor_cloud(pts, pose, w["n_points"], w["radius"], seed=2 if w["map"] == "room" else 7)
    if w["devs"] is None:
        particles = particles_uniform(w["n_particles"], bounds, seed=8)
    else:
        particles = particles_tracking(w["n_particles"], pose, w["devs"], seed=3 if w["map"] == "room" else 6)
    ranges = beacons(pose) if w["beacons"] else np.zeros((0, 4), np.float32)
    out = dict(name=name, map_points=pts, bounds=bounds, cloud=cloud, particles=particles, ranges=ranges, pose=pose)
    out.update(DEFAULTS)
    out["roll"], out["pitch"] = 0.01, -0.02
    return out
