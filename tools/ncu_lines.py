#!/usr/bin/env python3
"""Per-source-line instruction / stall shares inside one file of a kernel (companion of ncu_regions.py).
usage: ncu_lines.py <ncu source csv> <nvdisasm -g -c dump> <mangled kernel substring> <file basename> [games] [min%]"""
import csv, os, re, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ncu_csv, dis, kernel, want = sys.argv[1:5]
games = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
minp = float(sys.argv[6]) if len(sys.argv) > 6 else 0.3
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and kernel in l)
cur, seq = ("?", 0), []
for l in lines[start + 1:]:
    if l.startswith("\t.section") and seq:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1))
        if f.endswith((".cu", ".cuh")):
            cur = (f, int(m.group(2)))
        else:
            ours = [(os.path.basename(a), int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', l) if a.endswith((".cu", ".cuh"))]
            if ours:
                cur = ours[0]
        continue
    mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)", l)
    if mm:
        seq.append((cur, mm.group(1)))
rows = list(csv.reader(open(ncu_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > ci]
inst, samp, ops = defaultdict(int), defaultdict(int), defaultdict(lambda: defaultdict(int))
ti = ts = 0
for k in range(min(len(body), len(seq))):
    (f, ln), op = seq[k]
    n = int(float(body[k][ci] or 0)); s = int(float(body[k][si] or 0))
    ti += n; ts += s
    if f == want:
        inst[ln] += n; samp[ln] += s; ops[ln][op.split(".")[0]] += n
src = open(os.path.join(ROOT, "generalsreinforcementlearning_b200", "csrc", want)).read().split("\n")
for ln in sorted(inst):
    pi, ps = 100 * inst[ln] / ti, 100 * samp[ln] / ts
    if pi >= minp or ps >= minp:
        top = ",".join(f"{o}:{c / games:.0f}" for o, c in sorted(ops[ln].items(), key=lambda x: -x[1])[:5])
        print(f"{ln:5d} {pi:5.1f}% {inst[ln] / games:7.0f}/game stall {ps:5.1f}%  {src[ln - 1].strip()[:80]:80s} {top}")
