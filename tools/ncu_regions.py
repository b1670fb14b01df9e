#!/usr/bin/env python3
"""Per-region instruction / stall-sample shares of one kernel from an ncu capture.

usage: ncu_regions.py <ncu --page source --csv dump> <nvdisasm -g -c dump of the cubin> <mangled kernel substring> [games]

ncu's source-page CSV has per-SASS-instruction counters but no line numbers; `nvdisasm -g -c` prints
`//## File "...", line N [inlined at ...]` markers for the same instruction sequence.  The two are joined by
instruction order within the kernel; every instruction is charged to the innermost (file, line) it came from, and lines
are grouped into regions: the function definitions of csrc/*.cuh|*.cu plus the `// ----` phase markers of the turn kernel.
"""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "generalsreinforcementlearning_b200", "csrc")


def region_marks(path):
    marks = []
    lines = open(path).read().split("\n")
    for i, l in enumerate(lines, 1):
        if l.startswith("__device__") or l.startswith("__global__") or l.startswith("    grl_turn_kernel("):
            nm = re.search(r"(\w+)\s*\(", l if "(" in l else lines[i])
            marks.append((i, "fn " + (nm.group(1) if nm else l[:40])))
        elif re.match(r"\s+// ---- ", l):
            marks.append((i, "   " + l.strip()[8:64]))
        elif re.match(r"  // (engine legal-action mask|observation planes: Serializer)", l):
            marks.append((i, "   " + l.strip()[3:64]))
    return marks


def main():
    ncu_csv, dis, kernel = sys.argv[1:4]
    games = int(sys.argv[4]) if len(sys.argv) > 4 else 65536
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and kernel in l)
    cur = ("?", 0)
    seq = []
    for l in lines[start + 1:]:
        if l.startswith("\t.section") and seq:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            f = os.path.basename(m.group(1))
            if f.endswith((".cu", ".cuh")):
                cur = (f, int(m.group(2)))
            else:  # a CUDA header: charge the line of OUR file that inlined it
                chain = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
                ours = [(os.path.basename(a), int(b)) for a, b in chain if a.endswith((".cu", ".cuh"))]
                if ours:
                    cur = ours[0]
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            seq.append(cur)
    rows = list(csv.reader(open(ncu_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = hdr.index("Instructions Executed")
    si = hdr.index("# Samples") if "# Samples" in hdr else None
    body = [r for r in rows[hdr_i + 1:] if len(r) > ci]
    if len(body) != len(seq):
        print(f"warning: {len(body)} ncu rows vs {len(seq)} disassembled instructions", file=sys.stderr)
    marks = {}
    inst, samp = defaultdict(int), defaultdict(int)
    for k in range(min(len(body), len(seq))):
        f, ln = seq[k]
        if f not in marks:
            p = os.path.join(CSRC, f)
            marks[f] = region_marks(p) if os.path.exists(p) else []
        name = f"{f}:?"
        for (ml, nm) in marks[f]:
            if ml <= ln:
                name = f"{f}:{ml:<5d} {nm}"
            else:
                break
        inst[name] += int(float(body[k][ci] or 0))
        if si is not None:
            samp[name] += int(float(body[k][si] or 0))
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print(f"total warp instructions {ti} = {ti / games:.0f} per game; {len(seq)} SASS instructions ({len(seq) * 16 / 1024:.0f} KB)")
    print(f"{'region':88s} inst%  inst/game  stall%")
    for k in sorted(inst, key=lambda n: (n.split(":")[0], int(re.search(r":(\d+|\?)", n).group(1).replace("?", "0")))):
        print(f"{k:88s} {100 * inst[k] / ti:5.1f}  {inst[k] / games:8.0f}  {100 * samp[k] / ts:5.1f}")


main()
