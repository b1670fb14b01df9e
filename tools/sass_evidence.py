#!/usr/bin/env python3
"""profiles/<tag>_sass_tma.txt: what the built libgrlcuda.so contains, from `cuobjdump -sass` — architectures, and per
turn-kernel instantiation the counts of the SASS mnemonics that prove the Blackwell data path (B200_PROFILING.md):
UBLKCP (cp.async.bulk, the TMA slab loads), UBLKPF (cp.async.bulk.prefetch.L2), SYNCS (mbarrier), STG.E.EF.128
(128-bit evict-first streaming stores of the observation planes), ATOMS / REDUX / SHFL (bit-stream and lane-group
machinery), plus code size.  usage: python tools/sass_evidence.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generalsreinforcementlearning_b200 import build as grl_build  # noqa: E402

PATTERNS = ["UBLKCP", "UBLKPF", "SYNCS", "STG.E.EF.128", "STG.E.EF", "LDS.128", "ATOMS", "REDUX", "SHFL", "UTCHMMA|UTCQMMA|UTCMMA", "HMMA|IMMA"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    lib = grl_build.LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    per = collections.OrderedDict()
    name = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            per[name]["_total"] += 1
            for p in PATTERNS:
                if re.match(p.replace(".", r"\.") + r"(\.|$)", op) or re.match("(" + p + r")(\.|$)", op):
                    per[name][p] += 1
    out = [f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   (library source hash {grl_build.source_hash()})",
           f"# cubin architectures: {', '.join(archs)}",
           "# columns: SASS instructions (code KB) | " + " | ".join(PATTERNS), ""]
    demangle = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.split("\n")
    tot = collections.Counter()
    for (n, c), d in zip(per.items(), demangle):
        if "grl_turn_kernel" not in n and "grl_mapgen_kernel" not in n and "grl_reset_kernel" not in n:
            continue
        short = re.sub(r"\(GrlKParams.*", "", d).replace("void ", "")
        out.append(f"{short:70s} {c['_total']:6d} ({c['_total'] * 16 / 1024:5.0f} KB) | " + " | ".join(f"{c[p]:4d}" for p in PATTERNS))
        tot.update(c)
    out.append("")
    out.append(f"{'all kernels above':70s} {tot['_total']:6d} ({tot['_total'] * 16 / 1024:5.0f} KB) | " + " | ".join(f"{tot[p]:4d}" for p in PATTERNS))
    out.append("# turn kernel template arguments: <players template, W, H, lanes per game, DO_STEP, DO_OUT, GYM>; no tensor-core"
               " instruction anywhere (nothing here is a contraction)")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_tma.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:12]), "\n...\n", out[-3])


main()
