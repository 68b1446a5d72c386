#!/usr/bin/env python
"""Folds the counters bench.py quotes out of an `ncu --set full` report into profiles/ncu_summary.json (run HERE, no GPU):

    python tools/ncu_to_json.py <report.ncu-rep> <key> <evals_per_launch> <committed summary file> [launch index]

<key> is the bench workload the capture belongs to ("cfg4", "cfg2", ...).  bench.py loads the JSON and reports
roofline.traffic = dram_bytes_per_launch, roofline.ncu = the whole record, roofline.issue_frac from thread_inst_per_eval."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
            "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)


def main():
    rep, key, evals, summary = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    d = dict(zip(head, rows[2 + which]))
    u = dict(zip(head, units))

    def val(name):
        return float(d[name].replace(",", "")) * unit_scale(u[name])

    dur = val("gpu__time_duration.sum")
    lts = val("lts__t_sectors_srcunit_tex_op_read.sum")
    l1s = val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
    req = val("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
    inst = val("smsp__inst_executed.sum")
    lanes = val("smsp__thread_inst_executed_per_inst_executed.ratio")
    rec = {
        "file": summary, "kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size"),
        "evals_per_launch": evals, "duration_us_under_ncu": dur * 1e6,
        "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "l2_sectors_read_per_launch": lts, "l2_to_l1_gbs_under_ncu": lts * 32 / dur / 1e9,
        "l1_sectors_per_request": l1s / req, "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
        "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"),
        "warp_inst_per_launch": inst, "thread_inst_per_eval": inst * lanes / evals, "warp_inst_per_warp_eval": inst * 32 / evals,
        "active_lanes_per_inst": lanes, "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": int(val("launch__registers_per_thread")),
        "stall_long_scoreboard_per_issue": val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "pipe_fma_pct": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    }
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    cur = json.load(open(path)) if os.path.exists(path) else {}
    cur[key] = rec
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
