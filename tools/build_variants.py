#!/usr/bin/env python3
"""Build comparison variants of libgrlcuda.so (-D switches) next to the product library for A/B timing on the GPU box:
build/libgrlcuda_<name>.so, loaded with GRL_LIB_PATH=<path> by tools/phase_bench.py / tools/gym_bench.py.  A switch lives in
the sources only while it is being measured; profiles/ records the outcome and the loser is deleted."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generalsreinforcementlearning_b200 import build as b

VARIANTS = {
    "base": [],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    for n in names:
        defs = VARIANTS[n] if n in VARIANTS else n.split(",")[1:]
        out = os.path.join(ROOT, "build", f"libgrlcuda_{n.split(',')[0]}.so")
        b.build(force=True, defines=defs, out=out)
        print("built", out, defs)
