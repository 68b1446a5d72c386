#!/usr/bin/env python
"""Extracts the steady-state loop of a weighting kernel from the built object (cuobjdump -sass, run HERE, no GPU) and writes
an excerpt + opcode histogram to profiles/:

    python tools/sass_excerpt.py 'weight_v5_kernelILi256ELb0ELb0ELb0ELi0E' profiles/r2_weight_v5_linear_sass.txt

The steady-state loop = the innermost backward branch whose body holds at least four gathers (LDG.E.CONSTANT) and packed
fp32 (FFMA2); points per iteration = its gather count.  Instructions inside the verification path (the BSSY/BSYNC region guarded by the near-face test) are listed
but counted separately: a warp enters it for one group in ~5 on map L, one in ~50 on map S."""
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "amcl3d_b200", "lib", "weight.cu.o")
INS = re.compile(r"^\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*")


def main():
    pat, out = sys.argv[1], sys.argv[2]
    points_per_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    names = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
    fn = [l.split("Function :")[1].strip() for l in names.splitlines() if "Function :" in l and pat in l]
    if not fn:
        raise SystemExit("no kernel matches " + pat)
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], OBJ], capture_output=True, text=True).stdout
    ins = []
    for l in sass.splitlines():
        m = INS.match(l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    best = None
    for k, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_index:
                body = ins[addr_index[tgt]:k + 1]
                n_ldg = sum("LDG.E.CONSTANT" in b for _, b in body)
                if n_ldg >= 4 and any("FFMA2" in b for _, b in body):
                    if best is None or len(body) < len(best):
                        best = body
                        points_per_iter = n_ldg
    if best is None:
        raise SystemExit("no loop found")
    # verification region: BSSY Bn, T opens a reconvergence scope early (the compiler hoists it above the straight-path
    # arithmetic); the code that only near-face lanes run starts behind the first forward branch to the scope's end T
    # (the near-face test) and ends at the BSYNC at T
    pending, end, hist, vhist = None, None, {}, {}
    lines = []
    for a, t in best:
        op = t.split()[0] if not t.startswith("@") else t.split()[1]
        op = op.split(".")[0]
        inside = end is not None and a < end
        if end is not None and a >= end:
            end = None
            pending = None
        (vhist if inside else hist)[op] = (vhist if inside else hist).get(op, 0) + 1
        lines.append("%s /*%04x*/ %s" % ("V" if inside else " ", a, t))
        m = re.match(r"(?:@!?U?P\d\s+)?BSSY\S*\s+B\d,\s*(0x[0-9a-f]+)", t)
        if m and pending is None and end is None:
            pending = int(m.group(1), 16)
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and pending is not None and end is None and pending - 0x20 <= int(m.group(1), 16) <= pending and int(m.group(1), 16) > a:
            end = int(m.group(1), 16)
    n_hot = sum(hist.values())
    with open(out, "w") as f:
        f.write("# %s\n# steady-state loop: %d instructions on the straight path for %d points = %.1f warp-instructions per "
                "point and lane-group (V = verification path, %d more instructions, entered only near voxel faces)\n"
                % (fn[0], n_hot, points_per_iter, n_hot / points_per_iter, sum(vhist.values())))
        f.write("# straight-path opcode histogram: %s\n" % ", ".join("%s %d" % kv for kv in sorted(hist.items(), key=lambda kv: -kv[1])))
        f.write("# verification-path opcode histogram: %s\n" % ", ".join("%s %d" % kv for kv in sorted(vhist.items(), key=lambda kv: -kv[1])))
        f.write("\n".join(lines) + "\n")
    print(out, "straight", n_hot, "verify", sum(vhist.values()), "per point %.1f" % (n_hot / points_per_iter))


if __name__ == "__main__":
    main()
