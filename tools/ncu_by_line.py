#!/usr/bin/env python3
"""Aggregate an ncu source-page CSV (SASS level) by CUDA source line.

usage: ncu_by_line.py <ncu --page source --csv dump> <nvdisasm -g -c dump> <kernel mangled name> [source.cu]

ncu's CSV export has per-SASS-instruction counters but no line numbers; nvdisasm -g prints
`//## File "...", line N` markers for the same instruction sequence.  The two are joined by
instruction order within the kernel.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    ncu_csv, dis, kernel = sys.argv[1:4]
    src = sys.argv[4] if len(sys.argv) > 4 else None
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text." + kernel))
    cur = None
    inlined = None
    seq = []
    for l in lines[start + 1:]:
        if l.startswith("\t.section") or l.startswith("//-----"):
            if seq:
                break
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            # instructions inlined from CUDA headers are charged to the line of OUR file that inlined them
            if m.group(1).endswith(".cu") or not m.group(4):
                if m.group(1).endswith(".cu"):
                    cur = int(m.group(2))
            else:
                cur = int(m.group(4)) if m.group(3).endswith(".cu") else cur
            inlined = int(m.group(4)) if m.group(4) else None
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            seq.append((cur, inlined, l.strip()))
    rows = list(csv.reader(open(ncu_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = hdr.index("Instructions Executed")
    si = hdr.index("# Samples") if "# Samples" in hdr else None
    body = [r for r in rows[hdr_i + 1:] if len(r) > ci]
    if len(body) != len(seq):
        print(f"warning: {len(body)} ncu rows vs {len(seq)} disassembled instructions", file=sys.stderr)
    n = min(len(body), len(seq))
    inst = defaultdict(int)
    samp = defaultdict(int)
    for k in range(n):
        line = seq[k][0]
        inst[line] += int(float(body[k][ci] or 0))
        if si is not None:
            samp[line] += int(float(body[k][si] or 0))
    total_i = sum(inst.values()) or 1
    total_s = sum(samp.values()) or 1
    text = open(src).read().split("\n") if src else None
    print(f"total warp instructions {total_i}, samples {total_s}")
    print(f"{'line':>5} {'inst%':>6} {'stall%':>6}  source")
    for line in sorted(inst, key=lambda l: -(inst[l] / total_i + samp[l] / total_s)):
        pi, ps = 100 * inst[line] / total_i, 100 * samp[line] / total_s
        if pi < 0.3 and ps < 0.3 and "--all" not in sys.argv:
            continue
        s = text[line - 1].strip()[:110] if text and line and line <= len(text) else ""
        print(f"{line!s:>5} {pi:8.4f} {ps:8.4f}  {s}")


if __name__ == "__main__":
    main()
