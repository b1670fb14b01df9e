#!/usr/bin/env python3
"""Build scheduling variants of libgrlcuda.so for A/B timing on the GPU box (profiles/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generalsreinforcementlearning_b200 import build as b

VARIANTS = {
    "np": [],
    "np_u4": ["GRL_OBS_UNROLL=4"],
    "pers_sync": ["GRL_PERSISTENT=1", "GRL_CTA_SYNC=1"],
    "pers": ["GRL_PERSISTENT=1"],
    "w4": ["GRL_WARPS_PER_CTA=4"],
    "w16": ["GRL_WARPS_PER_CTA=16"],
    "mb5": ["GRL_MIN_BLOCKS=5"],
    "mb6": ["GRL_MIN_BLOCKS=6"],
    "straddle_merge": ["GRL_STRADDLE_INLINE=0"],
    "staged_scalars": ["GRL_STAGE_SCALARS=1"],
    "obs_incr": ["GRL_OBS_INCR=1"],
    "linear_st": ["GRL_LINEAR_STCS=0"],
    "obs_nojoin": ["GRL_OBS_JOIN=0"],
    "gt1_ballot": ["GRL_GT1_BALLOT=1"],
    "straddle_elem": ["GRL_STRADDLE_INLINE=2"],
    "obs_planewise": ["GRL_OBS_PLANEWISE=1"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    for n in names:
        out = os.path.join(ROOT, "build", f"libgrlcuda_{n}.so")
        b.build(force=True, defines=VARIANTS[n], out=out)
        print("built", out)
