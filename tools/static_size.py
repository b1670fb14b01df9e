#!/usr/bin/env python3
"""Static SASS size of grl_turn_kernel per phase (instruction-cache footprint).
usage: static_size.py <nvdisasm -g -c dump> <mangled kernel> <grl_kernels.cu>"""
import re
import sys
from collections import defaultdict

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_phases as npz


def main():
    dis, kernel, srcfile = sys.argv[1:4]
    src = open(srcfile).read().split("\n")
    marks = []
    for name, needle in npz.PHASES:
        hit = [i + 1 for i, l in enumerate(src) if needle in l]
        if hit:
            marks.append((name, hit[0]))
    kernel_end = next(i + 1 for i, l in enumerate(src) if "__noinline__ void policy_phase" in l and i + 1 > marks[0][1])
    marks.append(("end", kernel_end))
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text." + kernel))
    cnt = defaultdict(int)
    phase, n = "prologue", 0
    for l in lines[start + 1:]:
        if (l.startswith("\t.section") or l.startswith("//-----")) and n:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            f, line = m.group(1), int(m.group(2))
            if f.endswith("grl_kernels.cu") and marks[0][1] <= line < kernel_end:
                for (name, lo), (_, hi) in zip(marks, marks[1:]):
                    if lo <= line < hi:
                        phase = name
            elif f.endswith("grl_kernels.cu") and line >= kernel_end:
                phase = "out-of-line functions"
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            cnt[phase] += 1
            n += 1
    for k, v in cnt.items():
        print(f"{k:24s} {v:5d} instrs {v * 16 / 1024:6.1f} KB")
    print(f"{'total':24s} {n:5d} instrs {n * 16 / 1024:6.1f} KB")


if __name__ == "__main__":
    main()
