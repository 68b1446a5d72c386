"""Generates the committed golden fixtures under tests/golden/ (run in the authoring container only).

Sources
  * the reference's own serialized test messages in /root/reference/amcl3d/tests/data (decoded, not copied
    verbatim: only the numeric payloads are kept) and the known-answer constants of tests/Grid3dTest.cpp:128-132
    and tests/PointCloudToolsTest.cpp:42-53;
  * outputs of the UNMODIFIED reference compiled into oracle/_ref/libamcl3d_ref.so, on the seeded synthetic
    configuration cfg1 (amcl3d_b200.synth), with its mt19937 seeded through the white-box harness.

/root/reference does not exist on the GPU box; the tests read only the files written here.
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_DATA = "/root/reference/amcl3d/tests/data"


def _skip_header(buf, off=0):
    off += 12  # seq, stamp.sec, stamp.nsec
    (ln,) = struct.unpack_from("<I", buf, off)
    return off + 4 + ln


def decode_pointcloud2(path):
    d = open(path, "rb").read()
    off = _skip_header(d)
    h, w = struct.unpack_from("<II", d, off)
    off += 8
    (nf,) = struct.unpack_from("<I", d, off)
    off += 4
    for _ in range(nf):
        (ln,) = struct.unpack_from("<I", d, off)
        off += 4 + ln + 4 + 1 + 4
    _, point_step, _ = struct.unpack_from("<BII", d, off)
    off += 9
    (dl,) = struct.unpack_from("<I", d, off)
    off += 4
    pts = np.frombuffer(d, dtype=np.float32, count=dl // 4, offset=off).reshape(-1, point_step // 4)
    return pts[:, :3].copy()


def decode_posearray_positions(path):
    d = open(path, "rb").read()
    off = _skip_header(d)
    (n,) = struct.unpack_from("<I", d, off)
    off += 4
    poses = np.frombuffer(d, dtype=np.float64, count=n * 7, offset=off).reshape(-1, 7)
    return poses[:, :3].copy()


def decode_occupancy_grid(path):
    d = open(path, "rb").read()
    off = _skip_header(d)
    off += 8  # map_load_time
    (res,) = struct.unpack_from("<f", d, off)
    off += 4
    w, h = struct.unpack_from("<II", d, off)
    off += 8
    origin = struct.unpack_from("<7d", d, off)
    off += 56
    (n,) = struct.unpack_from("<I", d, off)
    off += 4
    data = np.frombuffer(d, dtype=np.int8, count=n, offset=off).copy()
    return dict(res=res, width=w, height=h, origin_z=origin[2], data=data)


def make_kat():
    """KAT-1/2/3 inputs."""
    map_shifted = decode_pointcloud2(os.path.join(REF_DATA, "mappointcloud_msg.bin"))  # stored shifted by -octo_min
    sensor = decode_posearray_positions(os.path.join(REF_DATA, "grid_info.bin")).astype(np.float32)
    nav = decode_occupancy_grid(os.path.join(REF_DATA, "nav_msg.bin"))
    n_particles = len(decode_posearray_positions(os.path.join(REF_DATA, "particle_info.bin")))
    bounds = np.array([-17.35, -9.5, -1.4, 8.75, 9.7, 6.25, 0.05])  # tests/PointCloudToolsTest.cpp:42-53
    # back to the map frame: the reference's computePointCloud stores float(centre)
    map_pts = (map_shifted.astype(np.float64) + bounds[:3]).astype(np.float32)
    np.savez_compressed(
        os.path.join(HERE, "kat_map_T.npz"), map_points=map_pts, bounds=bounds, sensor_dev=np.float64(0.05),
        sensor_cloud=sensor, kat1_pose=np.array([20.017967, 10.140815, 3.372801, 0.0, 0.0, 0.166781], np.float32),
        kat1_expected=np.float64(3.8109049797058105), kat1_tol=np.float64(1e-4),  # tests/Grid3dTest.cpp:128-132,168
        nav_slice=nav["data"], nav_width=np.int64(nav["width"]), nav_height=np.int64(nav["height"]),
        nav_origin_z=np.float64(nav["origin_z"]), nav_res=np.float64(nav["res"]),
        particle_info_count=np.int64(n_particles))
    print("kat_map_T.npz: map", map_pts.shape, "sensor", sensor.shape, "nav", nav["width"], nav["height"],
          "origin_z", nav["origin_z"], "particle_info poses", n_particles)


def make_reference_vectors():
    """Outputs of the unmodified reference on cfg1 with seeded RNG."""
    from amcl3d_b200 import synth
    from oracle.bindings import Reference
    R = Reference()
    assert R.math_overloads_are_double()
    w = synth.make_workload("cfg1")
    G = R.grid()
    assert G.open_from_cloud(w["map_points"], w["bounds"], w["sensor_dev"])
    cells = G.cells()
    G.set_cloud(w["cloud"])
    out = {}
    # per-particle cloud weights through the public single-pose method
    parts = w["particles"]
    roll, pitch = np.float32(w["roll"]), np.float32(w["pitch"])
    out["wp_single"] = np.array([G.cloud_weight(p[0], p[1], p[2], roll, pitch, p[3]) for p in parts[:64]], np.float32)
    # one full predict -> update -> resample cycle, mt19937 seeded with 1234
    F = R.filter()
    F.seed(1234)
    F.set_particles(parts)
    F.predict(w["odom_mods"], w["deltas"])
    out["after_predict"] = F.particles()
    F.update(G, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
    out["after_update"] = F.particles()
    out["mean_after_update"] = F.mean()[:4].copy()
    F.resample()
    out["after_resample"] = F.particles()
    # the same RNG stream, replayed for injection into the CUDA path
    rng = R.rng(1234)
    out["predict_noise"] = rng.predict_noise(len(parts), w["odom_mods"], w["deltas"])
    out["resample_u01"] = np.float32(rng.uniform01())
    # init() with seed 99
    F2 = R.filter()
    F2.seed(99)
    F2.init(600, (0.0, 0.0, 2.5, 0.3), (0.05, 0.05, 0.05, 0.1))
    out["init_particles"] = F2.particles()
    out["init_mean"] = F2.mean()[:4].copy()
    out["init_noise"] = R.rng(99).init_noise(600, (0.05, 0.05, 0.05, 0.1))
    # probability-grid digest (the cells themselves are 16 MB; keep a strided sample + checksum)
    out["cells_sample_idx"] = np.arange(0, len(cells), 97, dtype=np.int64)
    out["cells_sample"] = cells[out["cells_sample_idx"]]
    out["cells_prob_sum64"] = np.float64(cells[:, 1].astype(np.float64).sum())
    out["grid_slice_z1"] = G.slice(1.0)[0]
    np.savez_compressed(os.path.join(HERE, "ref_cfg1.npz"), **out)
    print("ref_cfg1.npz:", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    make_kat()
    make_reference_vectors()
