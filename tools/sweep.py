"""Option sweep on one GPU: device time of update / weighting kernel for a list of (workload, particles, options).

    python tools/sweep.py cfg4 [--out gpurun_out/sweep_cfg4.jsonl] [--spec tools/sweep_spec.json]

One process per workload (the map is built once); every variant re-creates the filter so that options apply cleanly.
Prints one JSON line per variant.  Never a bench value: exploration only.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT_SPECS = {
    "cfg4": [
        {"name": "ref_1M", "particles": 1048576, "opts": {}},
        {"name": "ref_524k", "particles": 524288, "opts": {}},
        {"name": "ref_524k_replay", "particles": 524288, "opts": {"replay": 2, "replay_max_mb": 90000}},
        {"name": "ref_262k", "particles": 262144, "opts": {}},
        {"name": "ref_262k_direct", "particles": 262144, "opts": {"replay": 1}},
        {"name": "ref_131k", "particles": 131072, "opts": {}},
        {"name": "ref_131k_direct", "particles": 131072, "opts": {"replay": 1}},
        {"name": "fast_131k", "particles": 131072, "opts": {"reference_order": 0}},
        {"name": "fast_1M", "particles": 1048576, "opts": {"reference_order": 0}},
    ],
    "cfg2": [
        {"name": "ref", "particles": 10000, "opts": {}},
        {"name": "ref_direct", "particles": 10000, "opts": {"replay": 1}},
        {"name": "fast", "particles": 10000, "opts": {"reference_order": 0}},
        {"name": "fast_fp64sums", "particles": 10000, "opts": {"reference_order": 0, "sum_mode": 2}},
    ],
    "cfg1": [
        {"name": "ref", "particles": 600, "opts": {}},
        {"name": "ref_direct", "particles": 600, "opts": {"replay": 1}},
        {"name": "fast", "particles": 600, "opts": {"reference_order": 0}},
        {"name": "fast_fp64sums", "particles": 600, "opts": {"reference_order": 0, "sum_mode": 2}},
    ],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--out", default=None)
    ap.add_argument("--spec", default=None, help="JSON file: {workload: [ {name, particles, opts, [points]} ]}")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="", help="comma-separated variant names")
    args = ap.parse_args()
    import torch
    import amcl3d_b200
    from amcl3d_b200 import synth

    specs = DEFAULT_SPECS
    if args.spec:
        specs = json.load(open(args.spec))
    variants = specs[args.workload]
    if args.only:
        keep = set(args.only.split(","))
        variants = [v for v in variants if v["name"] in keep]
    n_max = max(v["particles"] for v in variants)
    t0 = time.time()
    w = synth.make_workload(args.workload, n_particles=n_max)
    t_synth = time.time() - t0
    stream = torch.cuda.Stream()
    out = open(args.out, "a") if args.out else None
    flush = torch.empty((512 << 20) // 4, dtype=torch.float32, device="cuda")

    def emit(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    emit({"workload": args.workload, "synth_s": t_synth, "map_points": int(len(w["map_points"]))})
    cur_layout = None
    ctx = grid = None
    for v in variants:
        layout = v["opts"].get("grid_layout", 0)
        if ctx is None or layout != cur_layout:
            if grid is not None:
                grid.close()
                ctx.close()
            ctx = amcl3d_b200.Context(0, stream=stream.cuda_stream)
            ctx.set_option("grid_layout", layout)
            grid = amcl3d_b200.Grid(ctx, w["bounds"])
            t0 = time.time()
            grid.compute(w["map_points"], w["sensor_dev"], keep_dist=False)
            ctx.synchronize()
            emit({"grid_build_s": time.time() - t0, "layout": layout})
            cur_layout = layout
        # reset every option this sweep may touch
        for k in ("weight_point_splits", "sum_mode", "resample_mode", "cloud_order", "particle_order",
                  "weight_chunk_points", "weight_block_threads", "weight_variant"):
            try:
                ctx.set_option(k, 0)
            except Exception:
                pass
        ctx.set_option("reference_order", 1)
        ctx.set_option("order_clip_sigma_x10", 0)
        for ax, wt in (("x", 50), ("y", 400), ("z", 3200), ("yaw", 100)):
            ctx.set_option("order_weight_" + ax, wt)
        ctx.set_option("order_key_bits", 0)
        ctx.set_option("replay", 0)
        ctx.set_option("replay_max_mb", 40960)
        for k, val in v["opts"].items():
            if k != "grid_layout":
                ctx.set_option(k, val)
        ctx.set_option("kernel_timing", 1)
        n = v["particles"]
        pf = amcl3d_b200.Filter(ctx)
        part_all = w["particles"]
        if v.get("subset") == "yaw_slice":      # a pose-coherent shard: contiguous slice of the yaw-sorted set
            part_all = part_all[np.argsort(part_all[:, 3], kind="stable")]
            k0 = int(v.get("slice", 3)) * n
            part_all = np.ascontiguousarray(part_all[k0:k0 + n])
        pf.upload(part_all[:n])
        cloud = w["cloud"] if "points" not in v else w["cloud"][: v["points"]]
        rec = {"name": v["name"], "particles": n, "points": int(len(cloud)), "opts": v["opts"]}
        try:
            with torch.cuda.stream(stream):
                step_ms, kern_ms = [], []
                for k in range(3 + args.steps):
                    pf.stage_cloud(cloud)
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    pf.update_staged(grid, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], want_mean=False)
                    e1.record(stream)
                    e1.synchronize()
                    if k >= 3:
                        step_ms.append(e0.elapsed_time(e1))
                        kern_ms.append(ctx.last_kernel_ms())
                mean = pf.mean()
                # end-to-end wall clock through the host-buffer entry point
                e2e = []
                for k in range(args.steps):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    pf.update(grid, cloud, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
                    e2e.append(1e3 * (time.perf_counter() - t0))
                # resample + predict
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                rs, pr = [], []
                for k in range(3):
                    pf.upload(part_all[:n])
                    pf.update_staged(grid, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], want_mean=False)
                    ev[0].record(stream)
                    pf.resample(0.37)
                    ev[1].record(stream)
                    pf.predict(w["odom_mods"], w["deltas"], seed=1, step=k)
                    ev[2].record(stream)
                    ev[2].synchronize()
                    rs.append(ev[0].elapsed_time(ev[1]))
                    pr.append(ev[1].elapsed_time(ev[2]))
            rec.update({"step_ms": float(np.mean(step_ms)), "step_min_ms": float(np.min(step_ms)),
                        "kernel_ms": float(np.mean(kern_ms)), "e2e_p50_ms": float(np.median(e2e)),
                        "resample_ms": float(np.min(rs)), "predict_ms": float(np.min(pr)),
                        "evals_per_s": n * len(cloud) / (np.mean(step_ms) * 1e-3),
                        "mean": [float(x) for x in mean], "in_map": pf.last_in_map_evals()})
        except Exception as e:
            rec["error"] = str(e)
        emit(rec)
        pf.close()
    if grid is not None:
        grid.close()
        ctx.close()


if __name__ == "__main__":
    main()
