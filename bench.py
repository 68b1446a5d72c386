#!/usr/bin/env python
"""bench.py -- particle*point evaluations/s of the amcl3d measurement update (ParticleFilter::update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg1|cfg5|cfg3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one ParticleFilter::update over one synthetic sensor cloud (weighting of every particle against the
cloud, beacon likelihood, both normalisations, mean pose).  Default workload = the configuration BASELINE.json's
north_star quotes its scaling target on, configs[3]: map L (100 x 100 x 20 m @ 0.05 m = 1.6 G voxels), ONE set of
1 048 576 particles x 32 768 points, STRONG scaling: with N GPUs the particle set is block-partitioned over the ranks
(grid replicated); the ranks exchange the ten fp64 partial sums and the exact float carries of the reference's
sequential sums through peer-memory mailboxes inside the update kernels.  `--weak` gives every rank a full-size set.

Printed JSON (one line, rank 0):
  value        evals/s with the cloud already resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e          the same metric through the host-buffer C-ABI call amcl3d_cuda_pf_update (pinned cloud H2D + mean D2H
               inside the wall-clock timed region) + update latency percentiles
  cycle        predict -> update -> resample on the same particle set (device time per cycle)
  parity       max deviation from the CPU oracle, checked OUTSIDE the timed region at this N: per-particle cloud weights
               and counts on a subsample, then normalised weights / mean / resample indices over ALL particles
  latency      (N = 1) configs[1] (10 k x 10 k, map S) and configs[0] (600 x 2 k) end-to-end update latency
  roofline     the weighting kernel: sector-granular and algorithmic gather rates against the measured peaks
  cpu_baseline the UNMODIFIED reference (oracle/_ref) timed on this box's host, bounded sample (N = 1)

--impl reference times that CPU reference as its own arm (rank 0 only).  --workload cfg3 measures computeGrid.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_FLUSH_BYTES = 512 << 20
SECTOR_BYTES = 32
ALGO_BYTES_PER_EVAL = 4
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_summary.json")   # written by tools/ncu_to_json.py from ncu reports


_RESULT_FD = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torch.distributed notices) write to stdout; the contract is ONE JSON line there.
    Everything that is not the result goes to stderr: fd 1 is pointed at fd 2 and the result is written to the saved fd."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append((time.perf_counter(), int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if bit and (mask & bit):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self, t0=None, t1=None):
        self._stop_evt.set()
        self.join(timeout=1.0)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        inside = [c for (t, c) in self.samples if t0 is None or (t0 <= t <= t1)]
        use = inside if len(inside) >= 3 else [c for _, c in self.samples]
        return {"sm_mhz": float(np.median(use)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(use)}


def grid_dims(bounds):
    return [int(round((bounds[3 + a] - bounds[a]) / bounds[6])) for a in range(3)]


def workload_description(name, bounds, n_part_total, n_pts, n_ranges, n_gpus, strong):
    d = grid_dims(bounds)
    return ("%s: map %dx%dx%d voxels @%.2f m, %d particles in total (%s over %d GPU(s)) x %d points, %d beacons"
            % (name, d[0], d[1], d[2], bounds[6], n_part_total, "partitioned" if strong else "one full set per GPU",
               n_gpus, n_pts, n_ranges))


# ----------------------------------------------------------------------------------------------- CPU reference legs
LARGE_MAP = ("cfg4", "cfg5")


def reference_grid(R, w, workload):
    """A reference Grid3d ready for update() on this workload.  Map S: the reference's own computeGrid.  Map L exceeds
    the reference's 250 M-cell cap (PointCloudTools.cpp:103-105) and a kd-tree build of it would take hours, so the cells
    are INSTALLED instead (they only have to be there: update()'s cost depends on the grid's size and on which points
    land inside it, not on the values): the 64 x 64 x 20 m region around the tracked pose for cfg4 (every cloud point of
    every sampled particle stays inside), the full map for cfg5 (particles uniform over it)."""
    G = R.grid()
    if workload not in LARGE_MAP:
        if not G.open_from_cloud(w["map_points"], w["bounds"], w["sensor_dev"]):
            raise RuntimeError("reference computeGrid failed")
        return G, "the reference's own computeGrid"
    b = np.array(w["bounds"], np.float64)
    if workload == "cfg4":
        cx, cy = float(w["pose"][0]), float(w["pose"][1])
        b[0], b[3] = max(b[0], cx - 32.0), min(b[3], cx + 32.0)
        b[1], b[4] = max(b[1], cy - 32.0), min(b[4], cy + 32.0)
    d = grid_dims(b)
    rng = np.random.default_rng(1)
    plane = rng.uniform(0.0, 7.9, (d[1], d[0])).astype(np.float32)
    cells = np.empty((d[2], d[1], d[0], 2), np.float32)
    cells[..., 0] = 0.01
    cells[..., 1] = plane[None]
    ok = G.set_cells(np.zeros((1, 4), np.float32), b, w["sensor_dev"], np.array(d, np.uint32), cells.reshape(-1, 2))
    del cells
    if not ok:
        raise RuntimeError("reference set_cells failed")
    return G, "cells installed with set_cells (%dx%dx%d voxels around the pose; map L is over the reference's 250 M-cell cap)" % tuple(d)


def reference_prepare(w, workload, n_sample):
    from oracle.bindings import Reference
    R = Reference()
    G, how = reference_grid(R, w, workload)
    G.set_cloud(w["cloud"])
    return R, G, how


def _reference_worker(R, G, w, first, n_sample, reps, barrier, out, shard):
    try:
        F = R.filter()
        F.set_particles(w["particles"][first:first + n_sample])
        F.time_update(G, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=1)
        barrier.wait(timeout=1200)
        t0 = time.perf_counter()
        for _ in range(reps):
            F.time_update(G, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=1)
        out.put((shard, n_sample, time.perf_counter() - t0))
    except Exception as e:  # report, never hang the parent
        out.put((shard, 0, float("inf")))
        try:
            barrier.abort()
        except Exception:
            pass
        print("reference shard %d failed: %s" % (shard, e), file=sys.stderr)


def reference_all_cores(R, G, w, n_sample, n_pts, procs_n, reps):
    """Aggregate rate of one unmodified single-threaded reference process per host core, each on its own particle
    shard, sharing the (read-only) grid through fork.  NOT the reference as shipped (which is one thread)."""
    import multiprocessing as mp
    if procs_n <= 1:
        return None
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(procs_n), ctx.Queue()
    n_avail = len(w["particles"])
    procs = []
    for k in range(procs_n):
        first = (k * n_sample) % max(1, n_avail - n_sample + 1)
        procs.append(ctx.Process(target=_reference_worker, args=(R, G, w, first, n_sample, reps, barrier, out, k)))
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(out.get(timeout=1800))
    except Exception:
        pass
    for p in procs:
        p.join(timeout=5)
        if p.is_alive():
            p.terminate()
    if len(res) != procs_n or any(not np.isfinite(r[2]) for r in res):
        return None
    total = sum(r[1] for r in res) * n_pts * reps
    return {"value": total / max(r[2] for r in res), "unit": "evals/s", "cores": procs_n}


def ref_sample_size(args, n_pts):
    if args.ref_particles > 0:
        return args.ref_particles
    # about 2 s of single-thread work per update() at ~120 ns per evaluation
    return int(max(64, min(4096, 2.0 / (120e-9 * max(1, n_pts)))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from amcl3d_b200 import synth
    w = synth.make_workload(args.workload)
    n_pts = len(w["cloud"])
    n_sample = min(ref_sample_size(args, n_pts), len(w["particles"]))
    R, G, how = reference_prepare(w, args.workload, n_sample)
    F = R.filter()
    F.set_particles(w["particles"][:n_sample])
    times = []
    for i in range(args.warmup + args.steps):
        t = F.time_update(G, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=1)
        if i >= args.warmup:
            times.append(t)
    single = {"value": n_sample * n_pts / float(np.mean(times)), "ms_per_step": 1e3 * float(np.mean(times)), "cores": 1}
    sample = "%d of %d particles x %d points per step and process (update() is linear in particles, " \
             "ParticleFilter.cpp:129); grid: %s" % (n_sample, len(w["particles"]), n_pts, how)
    try:
        usable = len(os.sched_getaffinity(0))
    except Exception:
        usable = os.cpu_count() or 1
    procs_n = args.ref_procs if args.ref_procs > 0 else min(usable, 256)
    all_cores = None
    try:
        all_cores = reference_all_cores(R, G, w, n_sample, n_pts, procs_n, max(1, min(3, args.steps)))
    except Exception as e:
        print("all-cores reference figure unavailable: %s" % e, file=sys.stderr)
    value, cores, ms = single["value"], 1, single["ms_per_step"]
    if all_cores:
        # headline of this arm: every host core busy with the unmodified reference; the single-thread figure -- what the
        # reference as shipped delivers -- stays beside it
        value, cores = all_cores["value"], all_cores["cores"]
        ms = 1e3 * cores * n_sample * n_pts / value
        sample = "%d processes x (%s)" % (cores, sample)
    n_total = len(w["particles"])
    line = {
        "impl": "reference", "metric": "particle_point_evals_per_s", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        # the same workload string as the product arm prints (what was sampled of it is in cpu_baseline.sample)
        "config": {"workload": workload_description(args.workload, w["bounds"], n_total, n_pts, len(w["ranges"]),
                                                    args.gpus, not args.weak)},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "reference", "sample": sample,
                         "threads_note": "the reference is single-threaded by construction (Node.cpp:71-75): all host "
                                         "threads = one unmodified reference process per core, each on its own particle "
                                         "shard; host has %d cores" % (os.cpu_count() or 0),
                         "single_thread": single},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------- parity (outside timing)
def parity_check(ctx, grid, pf, w, particles_local, first, world, rank, dist, n_sub=256, weights_only=False):
    """Compares this run's update / resample with the CPU oracle (oracle/: test infrastructure, only the checker).
    Every rank checks the weighting step on a subsample of its own shard (voxel indices from the oracle's arithmetic,
    probabilities gathered from the device grid, summed in the reference's order); rank 0 then feeds the raw weights of
    ALL particles to the reference's normalisation loops and its resample walk and compares bit patterns."""
    import torch
    from oracle.bindings import Port
    port = Port()
    dims = np.array(grid.dims, np.uint32)
    bounds = w["bounds"]
    n = len(particles_local)
    pf.upload(particles_local)
    mean_g = pf.update(grid, w["cloud"], w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
    after = pf.download()
    raw_w, raw_n = pf.last_cloud_weights()
    mean_mask = int(pf.mean_exact_mask())   # bit k: component k is the reference's float chain (else its fp64 sum: a hovering chain)
    idx_g = pf.resample(0.37, want_idx=True)
    after_rs = pf.download()
    # ---- weighting step, subsample
    worst, counts_ok, checked = 0.0, True, 0
    roll, pitch = np.float32(w["roll"]), np.float32(w["pitch"])
    for i in (np.linspace(0, n - 1, min(n_sub, n)).astype(np.int64) if n else []):
        p = particles_local[i]
        if not port.is_into_map(bounds, p[0], p[1], p[2]):
            counts_ok &= raw_n[i] == 0
            continue
        idx, cnt = port.cloud_indices(dims, bounds, w["cloud"], (p[0], p[1], p[2], roll, pitch, p[3]))
        vals = grid.gather_prob(idx[idx != 0xFFFFFFFF])
        s = np.cumsum(vals, dtype=np.float32)[-1] if len(vals) else np.float32(0)      # sequential float sum (Grid3d.cpp:191)
        want = np.float32(0) if cnt <= 10 else np.float32(s) / np.float32(cnt)
        counts_ok &= int(raw_n[i]) == cnt
        if want > 0:
            worst = max(worst, abs(float(raw_w[i]) - float(want)) / float(want))
        elif raw_w[i] != 0:
            worst = float("inf")
        checked += 1
    rec = {"cloud_weight_max_rel_err": worst, "counts_exact": bool(counts_ok), "subsample_particles": int(checked)}
    if world > 1:
        t = torch.tensor([worst, 0.0 if counts_ok else 1.0, float(checked)], dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        rec = {"cloud_weight_max_rel_err": float(tm[0]), "counts_exact": bool(tm[1] == 0), "subsample_particles": int(t[2])}
    if weights_only:
        return rec
    # ---- sums over particles: gather everything on rank 0

    def gather_rows(a):
        if world == 1:
            return a
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([a.shape[0]], dtype=torch.int64, device="cuda"))
        sizes = [int(s) for s in sizes]
        cols = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
        m = max(max(sizes), 1)
        buf = torch.zeros((m, cols), dtype=torch.float64, device="cuda")
        if a.shape[0]:
            buf[:a.shape[0]] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(a.shape[0], cols)).cuda()
        outs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)]).reshape((-1,) + a.shape[1:])

    # float32 / uint32 values are exactly representable in float64, so the detour keeps every bit
    g_part = gather_rows(particles_local).astype(np.float32)
    g_after = gather_rows(after).astype(np.float32)
    g_raw = gather_rows(raw_w).astype(np.float32)
    g_idx = gather_rows(idx_g.astype(np.float64)).astype(np.uint32)
    g_rs = gather_rows(after_rs).astype(np.float32)
    if rank == 0:
        q = g_part.copy()
        q[:, 5] = g_raw
        if len(w["ranges"]):
            q[:, 6] = [port.range_weight(x, y, z, w["ranges"], w["sigma_range"]) for x, y, z in q[:, :3]]
        else:
            q[:, 6] = 0.0
        want, mean_o = port.update_from_weights(q, bounds, w["alpha"])
        exact = len(w["ranges"]) == 0      # with beacons wr goes through the device's exp(): tolerance instead of bits
        denom = np.maximum(np.abs(want[:, 4].astype(np.float64)), 1e-300)
        rec.update({
            "normalised_w_max_rel_err": float((np.abs(g_after[:, 4].astype(np.float64) - want[:, 4]) / denom)[want[:, 4] > 0].max()) if np.any(want[:, 4] > 0) else 0.0,
            "normalised_w_bit_exact": bool(np.array_equal(g_after[:, 4].view(np.uint32), want[:, 4].view(np.uint32))),
            "mean_abs_err": float(np.abs(mean_g.astype(np.float64) - mean_o.astype(np.float64)).max()),
            "mean_exact_mask": mean_mask,
            "mean_bit_exact_where_claimed": bool(all(mean_g.view(np.uint32)[k] == mean_o.view(np.uint32)[k]
                                                     for k in range(4) if (mean_mask >> k) & 1)),
            "mean_note": "components outside mean_exact_mask hover around zero (|sum| << sum |term|): returned as the fp64 "
                         "sum, inside the 1e-4 m tolerance by orders of magnitude",
            "bit_exact_expected": bool(exact),
        })
        _, idx_o = port.resample(g_after, 0.37)
        rec["resample_indices_equal"] = bool(np.array_equal(g_idx, idx_o))
        rec["resample_poses_equal"] = bool(np.array_equal(g_rs[:, :4].view(np.uint32), g_after[idx_o][:, :4].view(np.uint32)))
        rec["particles_checked"] = int(len(q))
        rec["oracle"] = ("oracle/ port (bit-identical to the reference build, tests/test_oracle_port_vs_reference.py): "
                         "indices + sequential sums for the subsample, ParticleFilter.cpp:151-218 loops for all particles")
    return rec


# ----------------------------------------------------------------------------------------------- product arm
def latency_record(amcl3d_b200, synth, torch, stream, local_rank, steps=200):
    """configs[1] and configs[0] end to end on one GPU: host wall clock around amcl3d_cuda_pf_update with a fresh
    pinned cloud every call (20 warm-ups, >= 200 timed iterations, p50 / p90 / p99)."""
    out = {}
    for name in ("cfg2", "cfg1"):
        w = synth.make_workload(name)
        ctx = amcl3d_b200.Context(local_rank, stream=stream.cuda_stream)
        grid = amcl3d_b200.Grid(ctx, w["bounds"])
        grid.compute(w["map_points"], w["sensor_dev"], keep_dist=False)
        pf = amcl3d_b200.Filter(ctx)
        pf.upload(w["particles"])
        n_pts = len(w["cloud"])
        clouds = [torch.from_numpy(synth.sensor_cloud(w["map_points"], w["pose"], n_pts, synth.WORKLOADS[name]["radius"],
                                                      seed=100 + k)).pin_memory() for k in range(4)]
        rec = {"workload": "%d particles x %d points, %d beacons, map S" % (len(w["particles"]), n_pts, len(w["ranges"])),
               "iterations": steps}
        for mode, opts in (("reference_order", {"reference_order": 1, "sum_mode": 0}),
                           ("fast", {"reference_order": 0, "sum_mode": 2})):
            for k, v in opts.items():
                ctx.set_option(k, v)
            ms = []
            with torch.cuda.stream(stream):
                for k in range(20 + steps):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    pf.update(grid, clouds[k % 4].numpy(), w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
                    if k >= 20:
                        ms.append(1e3 * (time.perf_counter() - t0))
            rec[mode] = {"update_p50_ms": float(np.percentile(ms, 50)), "update_p90_ms": float(np.percentile(ms, 90)),
                         "update_p99_ms": float(np.percentile(ms, 99)),
                         "evals_per_s_e2e": len(w["particles"]) * n_pts / (1e-3 * float(np.mean(ms)))}
        rec["note"] = ("reference_order (default): every sum is the reference's sequential float sum, results bit-identical "
                       "to the reference given the same grid; fast: re-associated cloud sums + fp64 sums over particles "
                       "(the round-1 numerics: closer to the exact sums, ~1e-5..1e-4 away from the reference's)")
        out[name] = rec
        pf.close()
        grid.close()
        ctx.close()
    return out


def run_ours(args):
    import torch
    import amcl3d_b200
    from amcl3d_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    strong = not args.weak

    stream = torch.cuda.Stream()
    ctx = amcl3d_b200.Context(local_rank, stream=stream.cuda_stream)
    info = ctx.device_info()

    # ---- workload: identical map and clouds on every rank (rank 0 builds the map, the others read its cache)
    os.environ.setdefault("AMCL3D_SYNTH_CACHE", "/tmp/amcl3d_synth_%d" % os.getuid())
    if world > 1 and rank != 0:
        dist.barrier()
    t_synth = time.perf_counter()
    w = synth.make_workload(args.workload, n_particles=args.particles, n_points=args.points)
    t_synth = time.perf_counter() - t_synth
    if world > 1 and rank == 0:
        dist.barrier()
    n_total = len(w["particles"]) if strong else len(w["particles"]) * world
    if strong:
        from amcl3d_b200 import shard
        first, count = shard.partition(len(w["particles"]), rank, world)
        particles = np.ascontiguousarray(w["particles"][first:first + count])
    else:
        first = rank * len(w["particles"])
        particles = w["particles"].copy()
        if rank:
            rng = np.random.default_rng(1000 + rank)
            particles[1:, :4] += rng.normal(0, 1e-3, (len(particles) - 1, 4)).astype(np.float32)
    n_part, n_pts = len(particles), len(w["cloud"])
    grid = amcl3d_b200.Grid(ctx, w["bounds"])
    t_grid = time.perf_counter()
    grid.compute(w["map_points"], w["sensor_dev"], keep_dist=False)   # every rank builds its replica (cfg3 measures this step)
    ctx.synchronize()
    t_grid = time.perf_counter() - t_grid
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(ctx.unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(uid.cpu().numpy(), rank, world)
    for name, val in (("sum_mode", args.exact), ("weight_point_splits", args.splits), ("weight_block_threads", args.block),
                      ("weight_variant", args.variant), ("particle_order", args.particle_order),
                      ("weight_chunk_points", args.chunk), ("peer_reduce", args.peer_reduce),
                      ("cloud_order", args.cloud_order), ("global_schedule", args.global_schedule),
                      ("global_schedule_chunk", args.deal_chunk)):
        if val is not None and val >= 0:
            ctx.set_option(name, val)
    ctx.set_option("kernel_timing", 1)
    pf = amcl3d_b200.Filter(ctx)
    pf.upload(particles)

    # a handful of distinct clouds (fresh measurement every step), in pinned host memory
    n_clouds = 4
    clouds = [torch.from_numpy(synth.sensor_cloud(w["map_points"], w["pose"], n_pts,
                                                  synth.WORKLOADS[args.workload]["radius"], seed=100 + k)).pin_memory()
              for k in range(n_clouds)]
    ranges = w["ranges"]
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")
    upd = (w["alpha"], w["sigma_range"], w["roll"], w["pitch"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    with torch.cuda.stream(stream):
        # ---- device-timed, cloud resident in HBM
        for k in range(args.warmup):
            pf.stage_cloud(clouds[k % n_clouds].numpy())
            pf.update_staged(grid, ranges, *upd, want_mean=False)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        kernel_ms, phase_ms = [], []
        launches0 = ctx.launch_count()
        t_region0 = time.perf_counter()
        align = torch.zeros(1, device="cuda")
        for k in range(args.steps):
            pf.stage_cloud(clouds[k % n_clouds].numpy())    # untimed: `value` is quoted with inputs resident
            flush.zero_()                                   # evict the grid / cloud from L2 between timed steps
            if world > 1:
                dist.all_reduce(align)                      # untimed: the ranks' streams start the step together
            ev[k][0].record(stream)
            pf.update_staged(grid, ranges, *upd, want_mean=False)
            ev[k][1].record(stream)
            ev[k][1].synchronize()
            kernel_ms.append(ctx.last_kernel_ms())
            try:
                phase_ms.append(ctx.last_update_phases_ms())
            except Exception:        # an empty shard launches no weighting kernel
                phase_ms.append([0.0, 0.0, 0.0])
        barrier()
        t_region1 = time.perf_counter()
        launches = ctx.launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in ev]
        in_map = pf.last_in_map_evals()

        # ---- end to end through the host-buffer entry point: pinned H2D + update + mean D2H, wall clock
        e2e_ms = []
        for k in range(3):
            pf.update(grid, clouds[k % n_clouds].numpy(), ranges, *upd)
        barrier()
        for k in range(args.steps):
            flush.zero_()
            barrier()                                       # untimed; includes torch.cuda.synchronize()
            t0 = time.perf_counter()
            mean = pf.update(grid, clouds[k % n_clouds].numpy(), ranges, *upd)
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
        barrier()
        clocks = sampler.stop(t_region0, t_region1)

        # ---- full filter cycle: predict -> update -> resample (device time)
        cyc = []
        n_cyc = max(3, min(args.steps, 10))
        for k in range(2 + n_cyc):
            pf.upload(particles)
            pf.stage_cloud(clouds[k % n_clouds].numpy())
            barrier()
            e0, e1, e2, e3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e0.record(stream)
            pf.predict(w["odom_mods"], w["deltas"], seed=11, step=k)
            e1.record(stream)
            pf.update_staged(grid, ranges, *upd, want_mean=False)
            e2.record(stream)
            pf.resample(0.5)
            e3.record(stream)
            e3.synchronize()
            if k >= 2:
                cyc.append((e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)))
        barrier()

    # ---- the same step with re-associated cloud sums (option reference_order = 0), for comparison
    fast = None
    try:
        ctx.set_option("reference_order", 0)
        pf.upload(particles)
        with torch.cuda.stream(stream):
            for k in range(3):
                pf.stage_cloud(clouds[k % n_clouds].numpy())
                pf.update_staged(grid, ranges, *upd, want_mean=False)
            barrier()
            n_fast = max(3, min(args.steps, 10))
            fms = []
            for k in range(n_fast):
                pf.stage_cloud(clouds[k % n_clouds].numpy())
                flush.zero_()
                if world > 1:
                    dist.all_reduce(align)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                pf.update_staged(grid, ranges, *upd, want_mean=False)
                e1.record(stream)
                e1.synchronize()
                fms.append(e0.elapsed_time(e1))
            barrier()
        fast_ms = float(np.mean(fms))
        if world > 1:
            t = torch.tensor([fast_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fast_ms = float(t[0])
        fast = {"ms_per_step": fast_ms, "evals_per_s": float(n_total) * n_pts / (fast_ms * 1e-3),
                "note": "reference_order = 0: Morton-ordered cloud, point splits, partials accumulated in double -- within "
                        "~1e-6 of the exact sums but NOT of the reference's float chain (see parity_fast)"}
    except Exception as e:
        fast = {"error": str(e)}
    finally:
        ctx.set_option("reference_order", 1)

    # ---- parity against the oracle at this N (outside every timed region)
    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(ctx, grid, pf, w, particles, first, world, rank, dist)
        except Exception as e:
            parity = {"error": str(e)}
        try:
            ctx.set_option("reference_order", 0)
            pfast = parity_check(ctx, grid, pf, w, particles, first, world, rank, dist, n_sub=64, weights_only=True)
            if isinstance(fast, dict):
                fast["parity_fast"] = pfast
        except Exception as e:
            if isinstance(fast, dict):
                fast["parity_fast"] = {"error": str(e)}
        finally:
            ctx.set_option("reference_order", 1)

    per_rank = None
    if world > 1:
        ph = np.mean(np.array(phase_ms, np.float64), axis=0)
        t = torch.tensor([float(np.mean(kernel_ms)), float(np.mean(step_ms)), ph[1], ph[2]], dtype=torch.float64, device="cuda")
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        per_rank = {"weighting_kernel_ms": [float(o[0]) for o in outs], "update_ms": [float(o[1]) for o in outs],
                    "exchange_and_wait_ms": [float(o[2]) for o in outs], "sums_over_particles_ms": [float(o[3]) for o in outs],
                    "note": "device time per rank; exchange_and_wait = from the end of this rank's weighting kernels until "
                            "the cloud sums of all ranks are back (absorbs the slowest rank); sums = update_seg_kernel"}
    else:
        ph = np.mean(np.array(phase_ms, np.float64), axis=0)
        per_rank = {"weighting_kernel_ms": [float(np.mean(kernel_ms))], "update_ms": [float(np.mean(step_ms))],
                    "exchange_and_wait_ms": [float(ph[1])], "sums_over_particles_ms": [float(ph[2])]}
    total_ms = float(np.sum(step_ms))
    total_e2e_ms = float(np.sum(e2e_ms))
    cyc_ms = np.array(cyc, np.float64).mean(0)
    if world > 1:
        t = torch.tensor([total_ms, total_e2e_ms, cyc_ms.sum()] + list(cyc_ms), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, total_e2e_ms = float(t[0]), float(t[1])
        cyc_total, cyc_ms = float(t[2]), t[3:].cpu().numpy()
    else:
        cyc_total = float(cyc_ms.sum())
    evals_per_step = float(n_total) * n_pts
    value = evals_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = evals_per_step * args.steps / (total_e2e_ms * 1e-3)

    line = None
    if rank == 0:
        peak, peak_kind = measured_peaks()
        k_ms = float(np.mean(kernel_ms))
        rate_in_map = in_map / (k_ms * 1e-3)
        plane_bytes = 4
        for a in grid_dims(w["bounds"]):
            plane_bytes *= a
        in_l2 = plane_bytes <= info["l2_bytes"] // 2
        uniform = synth.WORKLOADS[args.workload]["devs"] is None
        # regime: a plane that fits L2 is served from L2; the bricked 6.4 GB plane with a tracking particle cloud is served
        # from L2 as well (one chunk's footprint at a time, ncu: L2 hit > 90 %); particles uniform over it gather from HBM
        regime = "l2" if (in_l2 or not uniform) else "hbm"
        roofline = {"kernel": "weight_v5_kernel", "bound": regime, "unit": "GB/s",
                    "in_map_evals_per_launch_set": in_map, "evals_per_step_this_rank": float(n_part) * n_pts,
                    "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (float(np.mean(step_ms))),
                    "algorithmic_gbs": rate_in_map * ALGO_BYTES_PER_EVAL / 1e9,
                    "algorithmic_frac_of_hbm_copy": rate_in_map * ALGO_BYTES_PER_EVAL / 1e9 / peak,
                    "hbm_copy_peak_gbs": peak, "hbm_copy_peak_kind": peak_kind + " (MEASURED_PEAKS.json)"}
        try:
            l2_gbs, _ = ctx.probe_gather(min(max(plane_bytes, 1 << 20), 48 << 20), 1)
            hbm_gbs, _ = ctx.probe_gather(4 << 30, 1)
            gpeak = l2_gbs if regime == "l2" else hbm_gbs
            roofline.update({
                "achieved": rate_in_map * SECTOR_BYTES / 1e9, "peak": gpeak, "frac": rate_in_map * SECTOR_BYTES / 1e9 / gpeak,
                "definition": "achieved = 32 B x in-map evaluations / kernel time (SURVEY.md 8d: one sector per gather); "
                              "peak = this device's random 4-byte gather rate in the regime the kernel runs in, measured "
                              "in this run (amcl3d_cuda_probe_gather; L2-resident footprint / 4 GiB footprint).  Lanes "
                              "of a warp share sectors, so the hardware moves fewer sectors than `achieved` counts: "
                              "`ncu` below holds the counter-based figures",
                "gather_peak_l2_gbs": l2_gbs, "gather_peak_hbm_gbs": hbm_gbs})
        except Exception as e:
            roofline.update({"achieved": rate_in_map * SECTOR_BYTES / 1e9, "peak": peak,
                             "frac": rate_in_map * SECTOR_BYTES / 1e9 / peak, "gather_peak_error": str(e)})
        roofline["traffic"] = None
        try:
            ncu = json.load(open(NCU_SUMMARY)).get(args.workload)
            if ncu:
                # counter-based roofline: bytes the memory system really moved per evaluation (ncu --set full capture of
                # this kernel on this workload, committed under profiles/ and named in ncu["file"]) x the evaluation rate
                # measured live in this run, against the peak of the level that bounds the regime
                roofline["sector_gather"] = {"achieved": roofline.get("achieved"), "peak": roofline.get("peak"),
                                             "frac": roofline.get("frac"), "definition": roofline.pop("definition", None)}
                roofline["kernel"] = str(ncu.get("kernel", "weight_v5_kernel")).split("(")[0].replace("void ", "")
                roofline["traffic"] = ncu.get("dram_bytes_per_launch")
                roofline["ncu"] = ncu
                evals_rate = float(n_part) * n_pts / (k_ms * 1e-3)
                per_eval_l2 = ncu["l2_sectors_read_per_launch"] * SECTOR_BYTES / ncu["evals_per_launch"]
                per_eval_dram = ncu["dram_bytes_per_launch"] / ncu["evals_per_launch"]
                sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
                lts_cap = 6300.0 * sm_mhz * 1e6 / 1e9      # GB/s
                roofline.update({"l2_to_sm_gbs": per_eval_l2 * evals_rate / 1e9, "l2_to_sm_peak_gbs": lts_cap,
                                 "l2_to_sm_frac": per_eval_l2 * evals_rate / 1e9 / lts_cap,
                                 "l2_to_sm_peak_kind": "LTS throughput cap ~6300 B/clk full chip (B300_MICROARCH.md) x the SM "
                                                       "clock sampled in this run",
                                 "dram_gbs": per_eval_dram * evals_rate / 1e9, "dram_frac_of_hbm_copy": per_eval_dram * evals_rate / 1e9 / peak,
                                 "bytes_per_eval": {"algorithmic": ALGO_BYTES_PER_EVAL, "l2_to_sm": per_eval_l2, "dram": per_eval_dram}})
                if regime == "hbm":
                    roofline.update({"achieved": roofline["dram_gbs"], "peak": peak, "frac": roofline["dram_frac_of_hbm_copy"]})
                else:
                    roofline.update({"achieved": roofline["l2_to_sm_gbs"], "peak": lts_cap, "frac": roofline["l2_to_sm_frac"]})
                roofline["definition"] = ("achieved = bytes moved per evaluation at the bounding level (l2: L2->SM sectors x 32 B, "
                                          "hbm: DRAM bytes; ncu counters of the committed capture) x evaluations/s of this "
                                          "run's kernel; peak = that level's peak (l2: LTS cap, hbm: measured copy bandwidth). "
                                          "`sector_gather` keeps SURVEY 8d's figure: 32 B x in-map evaluations against the "
                                          "random-gather rate measured in this run")
                if ncu.get("thread_inst_per_eval") and clocks.get("sm_mhz"):
                    issue_peak = info["sm_count"] * 4 * 32 * clocks["sm_mhz"] * 1e6      # thread-instructions / s
                    roofline["issue_frac"] = ncu["thread_inst_per_eval"] * evals_rate / issue_peak
        except Exception as e:
            roofline["ncu_error"] = str(e)
        line = {
            "metric": "particle_point_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_description(args.workload, w["bounds"], n_total, n_pts, len(ranges), world, strong),
                       "particles_total": int(n_total), "particles_this_rank": int(n_part),
                       "l2": "flushed between timed steps by a %d MiB device write" % (L2_FLUSH_BYTES >> 20),
                       "sums": {0: "exact (auto)", 1: "exact, one CTA", 2: "fast fp64", 3: "exact, segmented"}.get(ctx.get_option("sum_mode")),
                       "point_splits": ctx.get_option("weight_point_splits"), "weight_variant": ctx.get_option("weight_variant"),
                       "exchange": ("peer-memory mailboxes inside the update kernels (fp64 partials + exact float carries)"
                                    if ctx.comm_peer_active() else "ncclAllReduce") if world > 1 else "none (one GPU)",
                       "weighting_schedule": ("pose-balanced over all ranks (poses all-gathered once per pose change, "
                                              "cloud sums returned to the owners by one exact uint32 all-reduce)"
                                              if ctx.get_option("global_schedule") != 1 else "every rank weighs its own shard")
                       if world > 1 else "one GPU",
                       "grid_build_s": t_grid, "synth_s": t_synth, "sm_count": info["sm_count"], "l2_bytes": info["l2_bytes"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": n_pts * 16 + len(ranges) * 16,
                    "d2h_bytes_per_step": 48, "update_p50_ms": float(np.percentile(e2e_ms, 50)),
                    "update_p90_ms": float(np.percentile(e2e_ms, 90)), "update_p99_ms": float(np.percentile(e2e_ms, 99)),
                    "mean_pose": [float(v) for v in mean]},
            "cycle": {"predict_ms": float(cyc_ms[0]), "update_ms": float(cyc_ms[1]), "resample_ms": float(cyc_ms[2]),
                      "total_ms": cyc_total, "evals_per_s": evals_per_step / (cyc_total * 1e-3),
                      "note": "predict (Philox) -> update -> global low-variance resample, device time, max over ranks"},
            "gpu_launches": int(launches),
            "per_rank": per_rank,
            "roofline": roofline,
            "parity": parity,
            "fast_mode": fast,
        }
    pf.close()
    grid.close()
    if world > 1:
        ctx.comm_destroy()
    ctx.close()

    if rank == 0:
        if world == 1 and not args.no_latency:
            try:
                line["latency"] = latency_record(amcl3d_b200, synth, torch, stream, local_rank)
            except Exception as e:
                line["latency"] = {"error": str(e)}
        cpu = None
        try:
            if world > 1:
                raise RuntimeError("measured at N = 1 only")
            n_ref = min(ref_sample_size(args, n_pts), len(w["particles"]))
            R, G, how = reference_prepare(w, args.workload, n_ref)
            F = R.filter()
            F.set_particles(w["particles"][:n_ref])
            F.time_update(G, ranges, *upd, reps=1)
            reps = 5
            t_best = F.time_update(G, ranges, *upd, reps=reps)
            cpu = {"value": n_ref * n_pts / t_best, "unit": "evals/s", "cores": 1, "kind": "reference",
                   "sample": "%d of %d particles x %d points, best of %d update() calls of the unmodified reference "
                             "(oracle/_ref; single-threaded by construction -- `--impl reference` also runs one process "
                             "per core); grid: %s; host has %d cores" % (n_ref, n_total, n_pts, reps, how, os.cpu_count() or 0)}
        except Exception as e:  # the checker library is test infrastructure; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "evals/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % e}
        line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------- cfg3: computeGrid
def run_grid(args):
    """--workload cfg3: Grid3d::computeGrid (PointCloudTools.cpp:84-149) on map L, voxels/s, end to end from host points."""
    import torch
    import amcl3d_b200
    from amcl3d_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ.setdefault("AMCL3D_SYNTH_CACHE", "/tmp/amcl3d_synth_%d" % os.getuid())
    if world > 1 and rank != 0:
        dist.barrier()
    pts, bounds = synth.make_map("warehouse")
    if world > 1 and rank == 0:
        dist.barrier()
    stream = torch.cuda.Stream()
    ctx = amcl3d_b200.Context(local_rank, stream=stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(ctx.unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(uid.cpu().numpy(), rank, world)
    pts4 = amcl3d_b200.capi.as_xyzw(pts)
    n_vox = 1
    for a in grid_dims(bounds):
        n_vox *= a
    ms = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0r = time.perf_counter()
    with torch.cuda.stream(stream):
        for k in range(args.warmup + args.steps):
            grid = amcl3d_b200.Grid(ctx, bounds)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            grid.compute(pts4, 0.05, keep_dist=False)
            ctx.synchronize()
            if k >= args.warmup:
                ms.append(1e3 * (time.perf_counter() - t0))
            grid.close()
    clocks = sampler.stop(t0r, time.perf_counter())
    total = float(np.sum(ms))
    if world > 1:
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t[0])
    if rank == 0:
        peak, peak_kind = measured_peaks()
        value = n_vox * args.steps / (total * 1e-3)
        line = {"metric": "computegrid_voxels_per_s", "value": value, "unit": "voxels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg3: computeGrid on map L, %d voxels, %d map points, z-slabs over %d GPU(s), "
                                       "probability plane only" % (n_vox, len(pts), world)},
                "clocks": clocks,
                "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": int(len(pts)) * 16, "d2h_bytes_per_step": 0},
                "gpu_launches": int(ctx.launch_count()),
                "roofline": {"bound": "hbm", "kernel": "df_tile_kernel", "achieved": n_vox * 4 / (total / args.steps * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": n_vox * 4 / (total / args.steps * 1e-3) / 1e9 / peak,
                             "traffic": None, "definition": "4 B written per voxel (probability plane) / end-to-end time "
                                                            "per build; the kernel is a compute-bound exact ring search"},
                "cpu_baseline": None}
        try:
            from oracle.bindings import Reference
            cpts, cb = synth.map_room()
            R = Reference()
            G = R.grid()
            t0 = time.perf_counter()
            ok = G.open_from_cloud(cpts, cb, 0.05)
            dt = time.perf_counter() - t0
            nv = 1
            for a in grid_dims(cb):
                nv *= a
            line["cpu_baseline"] = {"value": nv / dt if ok else None, "unit": "voxels/s", "cores": 1, "kind": "reference",
                                    "sample": "the reference's computeGrid (kd-tree 1-NN per voxel) on map S, %d voxels" % nv}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "voxels/s", "cores": 1, "kind": "reference",
                                    "sample": "unavailable: %s" % e}
        emit(line)
    if world > 1:
        ctx.comm_destroy()
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--weak", action="store_true", help="weak scaling: every rank gets a full-size particle set "
                    "(default: strong scaling, ONE set of the workload's size partitioned over the ranks)")
    ap.add_argument("--strong", action="store_true", help="(default) kept for compatibility")
    ap.add_argument("--particles", type=int, default=None, help="particle count of the set (default: the workload's)")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check after the timed regions")
    ap.add_argument("--no-latency", action="store_true", help="skip the configs[0] / configs[1] latency sub-record")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: processes for the all-cores figure "
                    "(0 = one per host core, 1 = skip)")
    ap.add_argument("--ref-particles", type=int, default=0, help="particle subsample for the CPU reference (0 = auto)")
    ap.add_argument("--exact", type=int, default=-1, help="sum_mode option (0 auto, 1 exact one CTA, 2 fast, 3 exact segmented)")
    ap.add_argument("--splits", type=int, default=-1, help="weight_point_splits option")
    ap.add_argument("--block", type=int, default=-1, help="weight_block_threads option")
    ap.add_argument("--variant", type=int, default=-1, help="weight_variant option (0 v5, 4 v4)")
    ap.add_argument("--chunk", type=int, default=-1, help="weight_chunk_points option")
    ap.add_argument("--particle-order", type=int, default=-1, help="particle_order option (0 auto, 1 off, 2 on)")
    ap.add_argument("--cloud-order", type=int, default=-1, help="cloud_order option (0 auto, 1 caller's order, 2 Morton)")
    ap.add_argument("--peer-reduce", type=int, default=-1, help="peer_reduce option (0 auto = peer memory, 1 = NCCL)")
    ap.add_argument("--global-schedule", type=int, default=-1, help="global_schedule option (0 auto = the weighting work of a "
                    "sharded set is dealt out by pose over all ranks, 1 = every rank weighs its own shard)")
    ap.add_argument("--deal-chunk", type=int, default=-1, help="global_schedule_chunk option (particles per chunk dealt "
                    "round-robin to the ranks; 0 = contiguous slices of the pose order)")
    args = ap.parse_args()
    quiet_stdout()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.workload == "cfg3":
            args.workload = "cfg4"
        return run_reference(args)
    if args.workload == "cfg3":
        return run_grid(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
