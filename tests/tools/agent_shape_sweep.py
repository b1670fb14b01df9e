#!/usr/bin/env python3
"""The in-step random agent of the CUDA gym step against the oracle's sampler over many board shapes (the generic
instantiation's word-boundary cases included: 32, 33, 64, 65 and 1,024 tiles), fog on and off, through the device-side
auto-reset.  One JSON line per shape.  usage: python tests/tools/agent_shape_sweep.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from generalsreinforcementlearning_b200 import load_library
from generalsreinforcementlearning_b200._abi import BoundLibrary
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

SHAPES = [(8, 4), (4, 8), (11, 3), (3, 11), (8, 8), (13, 5), (5, 13), (7, 7), (9, 9), (12, 12), (16, 16), (17, 19), (23, 9),
          (25, 25), (31, 32), (32, 31), (32, 32), (10, 10), (15, 15), (20, 20)]


def main():
    cuda = load_library()
    oracle = BoundLibrary(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    for W, H in SHAPES:
        for fog in (True, False):
            B, max_turns, steps = 48, 12, 40
            mk = lambda lib, **kw: GeneralsVecEnv(B, W, H, max_turns=max_turns, seed=W * 100 + H, fog_of_war=fog, lib=lib,
                                                  auto_reset="device", **kw)
            g, o = mk(cuda), mk(oracle, host_threads=1)
            og, _ = g.reset()
            oo, _ = o.reset()
            resets = 0
            for t in range(steps):
                a = o.sample_actions()
                rg, ro = g.step(None), o.step(a)
                ctx = f"{W}x{H} fog={fog} step {t}"
                assert torch.equal(rg[4]["action"].cpu(), a), f"{ctx}: action"
                assert torch.equal(rg[0].cpu(), ro[0]), f"{ctx}: observation"
                assert torch.equal(rg[1].cpu().view(torch.int64), ro[1].view(torch.int64)), f"{ctx}: reward"
                assert torch.equal(rg[2].cpu(), ro[2]) and torch.equal(rg[3].cpu(), ro[3]), f"{ctx}: flags"
                assert torch.equal(rg[4]["valid_actions_mask"].cpu(), ro[4]["valid_actions_mask"]), f"{ctx}: mask"
                assert np.array_equal(g.engine.state_hash(), o.engine.state_hash()), f"{ctx}: state"
                resets += int((ro[2] | ro[3]).sum())
            g.close()
            o.close()
            print(json.dumps(dict(board=[W, H], fog=fog, envs=B, steps=steps, episodes_finished=resets, mismatches=0)), flush=True)


if __name__ == "__main__":
    main()
