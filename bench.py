#!/usr/bin/env python
"""bench.py -- particle*point evaluations/s of the amcl3d measurement update (ParticleFilter::update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one ParticleFilter::update over one synthetic sensor cloud (weighting of every particle against the
cloud, beacon likelihood, both normalisations, mean pose).  Default workload = BASELINE.json configs[1]:
map S (20 x 20 x 5 m @ 0.1 m), 10 000 particles x 10 000 points per GPU (weak scaling over particles; with
N > 1 the ranks exchange the ten partial sums of the update through NCCL).

Printed JSON (one line, rank 0):
  value    : evals/s with the cloud already resident in HBM (device-timed with CUDA events on the library's stream)
  e2e      : the same metric through the host-buffer C-ABI call amcl3d_cuda_pf_update (pinned cloud H2D + mean D2H
             inside the wall-clock timed region), plus update latency percentiles
  roofline : the weighting kernel alone (events recorded around it inside the library)
  cpu_baseline : the UNMODIFIED reference (oracle/_ref/libamcl3d_ref.so) timed on this box's host, bounded sample

--impl reference times that CPU reference as its own arm (rank 0 only).
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE weighting-kernel launch, from the `ncu --set full` captures
# summarised in profiles/r1_weight_v4_cfg2_ncu_full.txt (cfg2: 7.79 MB read + 0 written; the 8 MB grid is L2-resident)
# and profiles/r1_weight_v4_cfg4_ncu_full.txt (cfg4, one of the 64 chunk launches: 265.4 MB + 28.5 MB)
NCU_DRAM_TRAFFIC = {"cfg2": 7791872, "cfg4": 293814272}


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


def workload_description(name, w, n_gpus):
    dims = [int(round((w["bounds"][3 + a] - w["bounds"][a]) / w["bounds"][6])) for a in range(3)]
    return ("%s: map %dx%dx%d voxels @%.2f m, %d particles/GPU x %d points, %d beacons, %d GPU(s)"
            % (name, dims[0], dims[1], dims[2], w["bounds"][6], len(w["particles"]), len(w["cloud"]), len(w["ranges"]),
               n_gpus))


def reference_sample(w, max_particles):
    """Times the unmodified reference's update() on a bounded particle subsample of the workload."""
    from oracle.bindings import Reference
    R = Reference()
    G = R.grid()
    if not G.open_from_cloud(w["map_points"], w["bounds"], w["sensor_dev"]):
        raise RuntimeError("reference computeGrid failed (map over the 250 M-cell cap?)")
    G.set_cloud(w["cloud"])
    F = R.filter()
    n = min(max_particles, len(w["particles"]))
    F.set_particles(w["particles"][:n])
    return R, G, F, n


def _reference_shard_worker(workload, n_particles, n_points, n_sample, shard, reps, barrier, out):
    """One of P independent reference processes (the reference itself is single-threaded): its own Grid3d and
    ParticleFilter, a disjoint particle shard, update() timed between two barriers."""
    try:
        from amcl3d_b200 import synth
        w = synth.make_workload(workload, n_particles=n_particles, n_points=n_points)
        first = (shard * n_sample) % max(1, len(w["particles"]) - n_sample + 1)
        w["particles"] = w["particles"][first:first + n_sample]
        R, G, F, n = reference_sample(w, n_sample)
        F.time_update(G, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=1)
        barrier.wait(timeout=300)
        t0 = time.perf_counter()
        for _ in range(reps):
            F.time_update(G, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=1)
        out.put((shard, n, time.perf_counter() - t0))
    except Exception as e:  # report, never hang the parent
        out.put((shard, 0, float("inf")))
        try:
            barrier.abort()
        except Exception:
            pass
        print("reference shard %d failed: %s" % (shard, e), file=sys.stderr)


def reference_all_cores(args, n_pts, reps=3):
    """Aggregate rate of P = all host cores independent reference processes, each on its own particle shard.
    NOT the reference as shipped (which is one thread): reported beside the single-core figure as the generous
    "every host thread" CPU number."""
    import multiprocessing as mp
    try:
        usable = len(os.sched_getaffinity(0))   # the cores this process may actually run on
    except Exception:
        usable = os.cpu_count() or 1
    procs_n = args.ref_procs if args.ref_procs > 0 else min(usable, 256)
    if procs_n <= 1:
        return None
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(procs_n), ctx.Queue()
    procs = [ctx.Process(target=_reference_shard_worker,
                         args=(args.workload, args.particles, args.points, args.ref_particles, k, reps, barrier, out))
             for k in range(procs_n)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(out.get(timeout=600))
    except Exception:
        pass
    for p in procs:
        p.join(timeout=5)
        if p.is_alive():
            p.terminate()
    if len(res) != procs_n or any(not np.isfinite(r[2]) for r in res):
        return None
    total = sum(r[1] for r in res) * n_pts * reps
    return {"value": total / max(r[2] for r in res), "unit": "evals/s", "cores": procs_n,
            "note": "%d independent single-threaded reference processes, one particle shard each, timed together "
                    "(the reference itself has no threads)" % procs_n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from amcl3d_b200 import synth
    w = synth.make_workload(args.workload, n_particles=args.particles, n_points=args.points)
    n_sample = args.ref_particles
    R, G, F, n = reference_sample(w, n_sample)
    n_pts = len(w["cloud"])
    times = []
    for i in range(args.warmup + args.steps):
        t = F.time_update(G, w["ranges"], w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=1)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3 * float(np.mean(times))
    value = n * n_pts / float(np.mean(times))
    sample = "%d of %d particles x %d points per step (linear in particles, ParticleFilter.cpp:129)" % (
        n, len(w["particles"]), n_pts)
    del F, G, R
    all_cores = None
    try:
        all_cores = reference_all_cores(args, n_pts)
    except Exception as e:
        print("all-cores reference figure unavailable: %s" % e, file=sys.stderr)
    single = {"value": value, "ms_per_step": ms, "cores": 1}
    cores = 1
    if all_cores:
        # headline of this arm: every host core busy with the unmodified reference (one process per core, one particle
        # shard each); the single-thread figure -- what the reference as shipped delivers -- stays beside it
        value, cores = all_cores["value"], all_cores["cores"]
        ms = 1e3 * cores * n * n_pts / value
        sample = "%d processes x (%s)" % (cores, sample)
    line = {
        "impl": "reference", "metric": "particle_point_evals_per_s", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_description(args.workload, w, args.gpus), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "reference", "sample": sample,
                         "threads_note": "the reference is single-threaded by construction (Node.cpp:71-75): all "
                                         "host threads = one unmodified reference process per core, each on its own "
                                         "particle shard; host has %d cores" % (os.cpu_count() or 0),
                         "single_thread": single},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


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

    stream = torch.cuda.Stream()
    ctx = amcl3d_b200.Context(local_rank, stream=stream.cuda_stream)
    info = ctx.device_info()

    # ---- workload (identical cloud and map on every rank; particles differ by rank through the seed offset)
    w = synth.make_workload(args.workload, n_particles=args.particles, n_points=args.points)
    if args.strong and world > 1:
        # strong scaling: ONE particle set of the workload's size, block-partitioned over the ranks
        from amcl3d_b200 import shard
        first, count = shard.partition(len(w["particles"]), rank, world)
        w["particles"] = np.ascontiguousarray(w["particles"][first:first + count])
    elif world > 1:
        rng = np.random.default_rng(1000 + rank)
        w["particles"][1:, :4] += rng.normal(0, 1e-3, (len(w["particles"]) - 1, 4)).astype(np.float32)
    if args.sort_particles:
        # experiment: particles pre-ordered by pose on the host (yaw, then a Morton curve over x, y, z)
        P = w["particles"]
        if args.sort_particles == "yaw":
            order = np.argsort(P[:, 3], kind="stable")
        else:
            cell = float(args.sort_particles.split(":")[1]) if ":" in args.sort_particles else 0.25
            yaw_cell = cell / 10.0
            q = np.concatenate([P[:, :3] / cell, P[:, 3:4] / yaw_cell], axis=1)
            q = np.floor(q - q.min(0)).astype(np.uint64)
            key = np.zeros(len(P), np.uint64)
            for bit in range(12):
                for a in range(4):
                    key |= ((q[:, a] >> np.uint64(bit)) & np.uint64(1)) << np.uint64(4 * bit + a)
            order = np.argsort(key, kind="stable")
        w["particles"] = np.ascontiguousarray(P[order])
    n_part, n_pts = len(w["particles"]), len(w["cloud"])
    grid = amcl3d_b200.Grid(ctx, w["bounds"])
    t_grid = time.perf_counter()
    grid.compute(w["map_points"], w["sensor_dev"], keep_dist=False)
    t_grid = time.perf_counter() - t_grid
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(ctx.unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(uid.cpu().numpy(), rank, world)
    pf = amcl3d_b200.Filter(ctx)
    pf.upload(w["particles"])
    if args.exact >= 0:
        ctx.set_option("sum_mode", args.exact)
    if args.splits >= 0:
        ctx.set_option("weight_point_splits", args.splits)
    if args.block > 0:
        ctx.set_option("weight_block_threads", args.block)
    if args.variant >= 0:
        ctx.set_option("weight_variant", args.variant)
    if args.l2fetch > 0:
        ctx.set_option("l2_fetch_granularity", args.l2fetch)
    if args.particle_order >= 0:
        ctx.set_option("particle_order", args.particle_order)
    if args.chunk > 0:
        ctx.set_option("weight_chunk_points", args.chunk)
    if args.peer_reduce >= 0:
        ctx.set_option("peer_reduce", args.peer_reduce)
    ctx.set_option("kernel_timing", 1)

    # a handful of distinct clouds (fresh measurement every step), in pinned host memory
    n_clouds = 4
    clouds = []
    for k in range(n_clouds):
        c = synth.sensor_cloud(w["map_points"], w["pose"], n_pts, synth.WORKLOADS[args.workload]["radius"], seed=100 + k)
        if args.morton:
            c = np.ascontiguousarray(c[synth.morton_order(c, args.morton)])
        t = torch.from_numpy(c).pin_memory()
        clouds.append(t)
    ranges = w["ranges"]
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(k):
        pf.update_staged(grid, ranges, w["alpha"], w["sigma_range"], w["roll"], w["pitch"], want_mean=False)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-timed, cloud resident in HBM
    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            pf.stage_cloud(clouds[k % n_clouds].numpy())
            step_resident(k)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        kernel_ms = []
        launches0 = ctx.launch_count()
        t_region0 = time.perf_counter()
        align = torch.zeros(1, device="cuda")
        for k in range(args.steps):
            pf.stage_cloud(clouds[k % n_clouds].numpy())   # untimed: `value` is quoted with inputs resident
            flush.zero_()                                   # evict the grid / cloud from L2 between timed steps
            if world > 1:
                dist.all_reduce(align)                      # untimed: the ranks' streams start the step together
            ev[k][0].record(stream)
            step_resident(k)
            ev[k][1].record(stream)
            ev[k][1].synchronize()
            kernel_ms.append(ctx.last_kernel_ms())
        barrier()
        t_region1 = time.perf_counter()
        launches = ctx.launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in ev]
        in_map = pf.last_in_map_evals()

        # ---- end to end through the host-buffer entry point: pinned H2D + update + mean D2H, wall clock
        e2e_ms = []
        for k in range(max(3, args.warmup // 2)):
            pf.update(grid, clouds[k % n_clouds].numpy(), ranges, w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
        barrier()
        for k in range(args.steps):
            flush.zero_()
            barrier()                                       # untimed; includes torch.cuda.synchronize()
            t0 = time.perf_counter()
            mean = pf.update(grid, clouds[k % n_clouds].numpy(), ranges, w["alpha"], w["sigma_range"], w["roll"], w["pitch"])
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
        barrier()
    clocks = sampler.stop(t_region0, t_region1)

    total_ms = float(np.sum(step_ms))
    total_e2e_ms = float(np.sum(e2e_ms))
    if world > 1:
        t = torch.tensor([total_ms, total_e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, total_e2e_ms = float(t[0]), float(t[1])
    n_part_total = float(n_part) * world
    if world > 1:
        t = torch.tensor([float(n_part)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        n_part_total = float(t[0])
    evals_per_step = n_part_total * n_pts
    value = evals_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = evals_per_step * args.steps / (total_e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_kind = measured_peaks()
        k_ms = float(np.mean(kernel_ms))
        rate_in_map = in_map / (k_ms * 1e-3)
        roofline = {
            "bound": "hbm", "kernel": "weight_v4_kernel",
            "achieved": rate_in_map * SECTOR_BYTES / 1e9, "peak": peak, "unit": "GB/s",
            "frac": rate_in_map * SECTOR_BYTES / 1e9 / peak, "traffic": NCU_DRAM_TRAFFIC.get(args.workload),
            "peak_kind": peak_kind + " HBM copy bandwidth (MEASURED_PEAKS.json)",
            "definition": "sector-granular gather traffic: 32 B x in-map evaluations per launch / kernel time "
                          "(SURVEY.md 8d); 4 B per evaluation are algorithmically needed",
            "algorithmic_gbs": rate_in_map * ALGO_BYTES_PER_EVAL / 1e9,
            "in_map_evals_per_launch": in_map, "evals_per_launch": float(n_part) * n_pts, "kernel_ms": k_ms,
            "kernel_share_of_step": k_ms / (total_ms / args.steps) if world == 1 else None,
        }
        # the device's own random-gather rooflines (amcl3d_cuda_probe_gather): 4-byte loads at random addresses, one
        # 32-byte sector each, over a footprint the size of the probability plane (L2-resident for map S) and over 4 GiB
        # (HBM-resident)
        plane_bytes = 4
        for a in range(3):
            plane_bytes *= int(round((w["bounds"][3 + a] - w["bounds"][a]) / w["bounds"][6]))
        try:
            l2_gbs, _ = ctx.probe_gather(min(max(plane_bytes, 1 << 20), 48 << 20), 1)
            hbm_gbs, _ = ctx.probe_gather(4 << 30, 1)
            in_l2 = plane_bytes <= info["l2_bytes"] // 2
            roofline.update({
                "gather_peak_l2_gbs": l2_gbs, "gather_peak_hbm_gbs": hbm_gbs,
                "gather_regime": "l2" if in_l2 else "hbm+l2",
                "frac_of_gather_peak": roofline["achieved"] / (l2_gbs if in_l2 else hbm_gbs),
                "gather_peak_kind": "measured in this run: random 4-byte read-only loads, sectors/s x 32 B "
                                    "(tools/gather_probe.py); above 1 in the hbm+l2 regime means L2 reuse",
            })
        except Exception as e:
            roofline["gather_peak_error"] = str(e)
        cpu = None
        try:
            if world > 1:
                raise RuntimeError("measured at N = 1 only")
            R, G, F, n_ref = reference_sample(w, args.ref_particles)
            reps = max(1, int(round(15.0 / max(1e-3, 60e-9 * n_ref * n_pts))))
            reps = min(reps, 20)
            F.time_update(G, ranges, w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=1)
            t_best = F.time_update(G, ranges, w["alpha"], w["sigma_range"], w["roll"], w["pitch"], reps=reps)
            cpu = {"value": n_ref * n_pts / t_best, "unit": "evals/s", "cores": 1, "kind": "reference",
                   "sample": "%d of %d particles x %d points, best of %d update() calls of the unmodified reference "
                             "(oracle/_ref; single-threaded by construction -- `--impl reference` also reports one "
                             "process per core), host has %d cores" % (n_ref, n_part, n_pts, reps, os.cpu_count() or 0)}
        except Exception as e:  # the checker library is test infrastructure; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "evals/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % e}
        line = {
            "metric": "particle_point_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_description(args.workload, w, world),
                       "particles_total": int(n_part_total),
                       "l2": "flushed between timed steps by a %d MiB device write" % (L2_FLUSH_BYTES >> 20),
                       "sum_mode": ctx.get_option("sum_mode"), "point_splits": ctx.get_option("weight_point_splits"),
                       "particle_order": ctx.get_option("particle_order"),
                       "partial_sums_exchange": ("peer memory inside the update kernels" if ctx.comm_peer_active()
                                                 else "ncclAllReduce") if world > 1 else "none (one GPU)",
                       "grid_build_s": t_grid, "sm_count": info["sm_count"], "l2_bytes": info["l2_bytes"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": n_pts * 16 + len(ranges) * 16,
                    "d2h_bytes_per_step": 48, "update_p50_ms": float(np.percentile(e2e_ms, 50)),
                    "update_p90_ms": float(np.percentile(e2e_ms, 90)), "update_p99_ms": float(np.percentile(e2e_ms, 99)),
                    "mean_pose": [float(v) for v in mean]},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    pf.close()
    grid.close()
    if world > 1:
        ctx.comm_destroy()
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg4", "cfg5"])
    ap.add_argument("--particles", type=int, default=None, help="particles per GPU (default: the workload's)")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: processes for the all-cores figure "
                    "(0 = one per host core, 1 = skip)")
    ap.add_argument("--ref-particles", type=int, default=1000, help="particle subsample for the CPU reference")
    ap.add_argument("--exact", type=int, default=-1, help="sum_mode option (0 auto, 1 exact, 2 fast)")
    ap.add_argument("--splits", type=int, default=-1, help="weight_point_splits option")
    ap.add_argument("--block", type=int, default=0, help="weight_block_threads option")
    ap.add_argument("--variant", type=int, default=-1, help="weight_variant option (0 v4, 4 v3, 3 v3 unroll 8, 1 v2, 2 v1)")
    ap.add_argument("--l2fetch", type=int, default=0, help="l2_fetch_granularity option (32, 64, 128 bytes)")
    ap.add_argument("--chunk", type=int, default=0, help="weight_chunk_points option")
    ap.add_argument("--particle-order", type=int, default=-1, help="particle_order option (0 auto, 1 off, 2 on)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: partition the workload's particle set "
                    "over the ranks instead of giving every rank a full-size set")
    ap.add_argument("--peer-reduce", type=int, default=-1, help="peer_reduce option (0 auto = peer memory, 1 = NCCL)")
    ap.add_argument("--sort-particles", default="", help="experiment: 'yaw' or 'morton[:cell_m]' host pre-ordering")
    ap.add_argument("--morton", type=float, default=0.0, help="experiment: Morton-order the cloud (cell size in m)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
