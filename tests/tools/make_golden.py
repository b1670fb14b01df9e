#!/usr/bin/env python3
"""Generate tests/golden/*.npz: small fixed-seed fixtures of the turn engine.

The reference is Go and cannot run in this image, so these vectors are produced by the CPU
oracle (oracle/grl_oracle.c), which is itself pinned by the reference's own known-answer
tests (tests/kats.py) and seeded mapgen golden counts.  They freeze the oracle's behaviour:
a later edit of the oracle OR of the CUDA kernels that changes any trajectory fails
tests/test_golden.py.  Usage: python tests/tools/make_golden.py  (rewrites tests/golden/).

Per config (W,H,P): 16 games, seeds 12345+i, 120 turns of the counter-based random-legal-move
policy (policy seed 2024):
  boards_*   owner/army/type of the generated maps           (mapgen + Go math/rand)
  hash       [T][B] uint64 full-state digest after each turn (grl_state_hash)
  reward     [T][B][P] fp32 bits, done [T][B], winner [T][B], step_error [T][B]
  mask       [T][B][P][words] packed engine masks
  obs_hash   [T][B*P] uint64 digest of each observation tensor (grl_buffer_hash)
  obs_last   the final observation tensors in full
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = [(5, 5, 2), (8, 8, 4), (10, 10, 2), (15, 15, 2), (20, 20, 2), (20, 20, 4)]
B, T, SEED, POLICY_SEED = 16, 120, 12345, 2024


def rollout(lib, W, H, P):
    from generalsreinforcementlearning_b200 import _abi
    from helpers import new_engine

    e = new_engine(lib, W, H, P, B)
    seeds = np.arange(B, dtype=np.int64) + SEED
    e.reset_seeded(seeds)
    st = e.get_state()
    res = dict(boards_owner=st["owner"].astype(np.int8), boards_army=st["army"].astype(np.int16),
               boards_type=st["type"].astype(np.int8))
    out = e.alloc_outputs_host()
    keys = ("hash", "reward", "done", "winner", "step_error", "mask", "obs_hash")
    acc = {k: [] for k in keys}
    for _ in range(T):
        e.step_fused(None, e.outputs(**out), _abi.STEP_FLAG_RANDOM_POLICY, POLICY_SEED)
        acc["hash"].append(e.state_hash().copy())
        acc["reward"].append(out["reward"].view(np.uint32).copy())
        acc["done"].append(out["done"].copy())
        acc["winner"].append(out["winner"].copy())
        acc["step_error"].append(out["step_error"].copy())
        acc["mask"].append(out["mask_bits"].copy())
        acc["obs_hash"].append(e.buffer_hash(out["obs"], 9 * W * H, B * P).copy())
    for k in keys:
        res[k] = np.stack(acc[k])
    res["obs_last"] = out["obs"].copy()
    e.close()
    return res


def main():
    from generalsreinforcementlearning_b200._abi import BoundLibrary

    lib = BoundLibrary(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for (W, H, P) in CONFIGS:
        res = rollout(lib, W, H, P)
        path = os.path.join(ROOT, "tests", "golden", f"rollout_{W}x{H}x{P}p.npz")
        np.savez_compressed(path, **res)
        print(path, os.path.getsize(path), "bytes; finished games:", int(res["done"][-1].sum()),
              "error turns:", int((res["step_error"] != 0).sum()))


if __name__ == "__main__":
    main()
