#!/usr/bin/env python
"""Measures the random-gather roofline of the GPU with amcl3d_cuda_probe_gather: sector GB/s for footprints from
L2-resident (map S = 8 MB) to HBM-resident (map L = 6.4 GB), with 1/2/4/8 lanes sharing a sector.
Usage (on the GPU box): python tools/gather_probe.py [--csv gpurun_out/gather_probe.csv]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()
    import amcl3d_b200
    ctx = amcl3d_b200.Context(0)
    rows = ["footprint_mib,lanes_per_sector,sector_gbs,warp_requests_per_s"]
    for mib in (8, 32, 64, 128, 512, 2048, 6144):
        for lanes in (1, 2, 4, 8):
            gbs, req = ctx.probe_gather(mib << 20, lanes)
            rows.append("%d,%d,%.1f,%.4g" % (mib, lanes, gbs, req))
            print(rows[-1], flush=True)
    ctx.close()
    if args.csv:
        os.makedirs(os.path.dirname(args.csv), exist_ok=True)
        open(args.csv, "w").write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
