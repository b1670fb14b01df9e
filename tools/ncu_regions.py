#!/usr/bin/env python3
"""Group tools/ncu_by_line.py output into source regions (functions / marked phases of grl_kernels.cu).

usage: ncu_regions.py <ncu source csv> <nvdisasm dump> <mangled kernel> <grl_kernels.cu> [games]
Regions are the top-level function definitions of the file plus the `// ----` phase markers inside the turn kernel."""
import re
import subprocess
import sys
import os


def main():
    csvf, dis, kernel, src = sys.argv[1:5]
    games = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_by_line.py"), csvf, dis, kernel, src, "--all"],
                         capture_output=True, text=True).stdout
    lines = open(src).read().split("\n")
    marks = []  # (line, name)
    for i, l in enumerate(lines, 1):
        m = re.match(r"(?:__device__|__global__|template|static|cudaError_t)", l)
        if l.startswith("__device__") or l.startswith("__global__"):
            nm = re.search(r"(\w+)\s*\(", l if "(" in l else lines[i])
            marks.append((i, "fn " + (nm.group(1) if nm else l[:40])))
        elif re.match(r"\s+// ---- ", l):
            marks.append((i, "   " + l.strip()[8:60]))
        elif re.match(r"  // (engine legal-action mask|observation planes: Serializer)", l):
            marks.append((i, "   " + l.strip()[3:60]))
    marks.sort()
    tot = re.search(r"total warp instructions (\d+), samples (\d+)", out)
    total_inst = int(tot.group(1))
    agg = {}
    for l in out.split("\n"):
        m = re.match(r"\s*(\d+)\s+([\d.]+)\s+([\d.]+)\s", l)
        if not m:
            continue
        ln, ip, sp = int(m.group(1)), float(m.group(2)), float(m.group(3))
        name = "?"
        for (ml, nm) in marks:
            if ml <= ln:
                name = f"{ml:5d} {nm}"
            else:
                break
        a = agg.setdefault(name, [0.0, 0.0])
        a[0] += ip
        a[1] += sp
    print(f"total warp instructions {total_inst} = {total_inst / games:.0f} per game")
    print(f"{'region':70s} inst%  inst/game  stall%")
    for k in sorted(agg):
        print(f"{k:70s} {agg[k][0]:5.1f}  {agg[k][0] / 100 * total_inst / games:8.0f}  {agg[k][1]:5.1f}")


main()
