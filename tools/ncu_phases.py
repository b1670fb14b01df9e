#!/usr/bin/env python3
"""Per-phase instruction / stall-sample shares of grl_turn_kernel from an ncu capture.

usage: ncu_phases.py <ncu --page source --csv dump> <nvdisasm -g -c dump> <mangled kernel> <grl_kernels.cu> [games]

ncu's CSV has per-SASS counters without line numbers; nvdisasm -g has the line markers for the
same instruction sequence (no inline chains).  Instructions that come from inlined helpers or
CUDA intrinsic headers are attributed to the nearest preceding instruction that maps into the
kernel body, which is approximate under instruction scheduling but good enough for shares.
"""
import csv
import re
import sys
from collections import defaultdict

PHASES = [
    ("prologue", "grl_turn_kernel(const __grid_constant__"),
    ("load slab (TMA)", "---- stage the slab in shared memory"),
    ("mask words -> regs", "---- mask words into registers"),
    ("policy sampling", "---- the synthetic policy reads the PRE-turn"),
    ("fog of war", "---- fog of war, from LAST turn"),
    ("actions (lane 0)", "---- actions: serial by definition"),
    ("eliminations", "---- eliminations + tile turnover"),
    ("production", "---- production over the cached lists"),
    ("stats / game over", "---- end of turn: stats, game over"),
    ("reward", "---- reward: CalculateRewardWithConfig"),
    ("slab write-back", "---- write the state back"),
    ("scalar read-outs", "---- scalar read-outs"),
    ("legal mask", "// engine legal-action mask, packed"),
    ("observation planes", "// observation planes: Serializer.StateToTensor"),
]


def main():
    ncu_csv, dis, kernel, srcfile = sys.argv[1:5]
    games = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
    src = open(srcfile).read().split("\n")
    marks = []
    for name, needle in PHASES:
        marks.append((name, next(i + 1 for i, l in enumerate(src) if needle in l)))
    kernel_end = next(i + 1 for i, l in enumerate(src) if l.startswith("// Reset: freshly uploaded slabs"))
    marks.append(("end", kernel_end))
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text." + kernel))
    cur_file, cur_line, seq = None, None, []
    for l in lines[start + 1:]:
        if (l.startswith("\t.section") or l.startswith("//-----")) and seq:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_file, cur_line = m.group(1), int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            seq.append((cur_file, cur_line))
    rows = list(csv.reader(open(ncu_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = rows[hdr_i + 1:hdr_i + 1 + len(seq)]
    inst, samp = defaultdict(int), defaultdict(int)
    phase = "prologue"
    # the plane-major observation writer is an inlined helper defined above the kernel
    obs_lo = next((i + 1 for i, l in enumerate(src) if "void obs_plane_major(" in l), None)
    obs_hi = marks[0][1]
    for (f, line), r in zip(seq, body):
        if f and f.endswith("grl_kernels.cu") and obs_lo and obs_lo <= line < obs_hi:
            phase = "observation planes"
        elif f and f.endswith("grl_kernels.cu") and marks[0][1] <= line < kernel_end:
            for (name, lo), (_, hi) in zip(marks, marks[1:]):
                if lo <= line < hi:
                    phase = name
        inst[phase] += int(float(r[ci] or 0))
        samp[phase] += int(float(r[si] or 0))
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"{'phase':22s} {'warp-inst/game':>15s} {'inst %':>7s} {'stall-sample %':>15s}")
    for name, _ in marks[:-1]:
        print(f"{name:22s} {inst[name] / games:15.1f} {100 * inst[name] / ti:7.1f} {100 * samp[name] / max(ts, 1):15.1f}")
    print(f"{'total':22s} {ti / games:15.1f}")


if __name__ == "__main__":
    main()
