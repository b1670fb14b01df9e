#!/usr/bin/env python3
"""Full-size soak of the gym path: two GeneralsVecEnv instances — the CUDA library and the CPU oracle — are driven
with the same actions (the oracle's own `sample_actions`, compared with the device's every step) through whole
episodes including truncation and automatic re-seeding; reward (float64 bits), terminated, truncated, invalid flags,
turn counters and full-state digests are compared EVERY step, observation and mask planes by digest every 20th step.
One JSON line per configuration (profiles/r1h_gym_soak.jsonl).

GRL_SOAK_RESET=device: the device-side auto-reset; GRL_SOAK_AGENT=in_step: on eight steps of nine the CUDA env draws its
random agent inside the gym step's own launch (step(None)) and must play what the oracle's sampler drew from its mask bytes.

usage: python tests/tools/gym_soak.py [W B max_turns steps] ..."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from generalsreinforcementlearning_b200 import load_library
from generalsreinforcementlearning_b200._abi import BoundLibrary
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv


def soak(cuda, oracle, W, B, max_turns, steps, mode="host", in_step=False):
    g = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=777, lib=cuda, auto_reset=mode)
    o = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=777, lib=oracle, host_threads=0, auto_reset=mode)
    og, _ = g.reset()
    oo, _ = o.reset()
    t0 = time.time()
    N = W * W
    finished = rejected = 0
    for t in range(steps):
        ctx = f"{W}x{W} step {t}"
        a = o.sample_actions()
        if in_step and t % 9 != 4:
            # the CUDA env draws its random agent inside the step's launch: it must play what the oracle's sampler drew
            rg = g.step(None)
            assert torch.equal(rg[4]["action"].cpu(), a), f"{ctx}: the action drawn inside the step"
        else:
            assert torch.equal(g.sample_actions().cpu(), a), f"{ctx}: sampled actions"
            if t % 9 == 4:
                a = a.clone()
                a[::97] = (a[::97] + 1) % (N * 5)      # some arbitrary (often masked-out) indices
            rg = g.step(a.to(g.device))
        ro = o.step(a)
        for k, name in ((1, "reward"), (2, "terminated"), (3, "truncated")):
            x, y = rg[k].cpu(), ro[k]
            if x.dtype == torch.float64:
                x, y = x.view(torch.int64), y.view(torch.int64)
            assert torch.equal(x, y), f"{ctx}: {name}"
        for name in ("invalid_action", "turn", "winner", "step_error"):
            assert torch.equal(rg[4][name].cpu(), ro[4][name]), f"{ctx}: info[{name}]"
        assert torch.equal(g._stats.cpu(), o._stats), f"{ctx}: PlayerState"
        assert np.array_equal(g.engine.state_hash(), o.engine.state_hash()), f"{ctx}: state"
        if t % 20 == 19:
            assert np.array_equal(g.engine.buffer_hash(g._obs, 9 * N, B * 2), o.engine.buffer_hash(o._obs.numpy(), 9 * N, B * 2)), f"{ctx}: obs"
            mg = g._mask.view(torch.uint8)
            words = (N * 5) // 4   # whole words of every row (the digest kernel reads 32-bit words)
            if (N * 5) % 4 == 0:
                assert np.array_equal(g.engine.buffer_hash(mg, words, B * 2), o.engine.buffer_hash(o._mask.view(torch.uint8).numpy(), words, B * 2)), f"{ctx}: mask"
            else:
                assert torch.equal(mg[:4096].cpu(), o._mask.view(torch.uint8)[:4096]), f"{ctx}: mask"
        if mode == "device" and t % 20 == 19:
            assert np.array_equal(g.engine.buffer_hash(g._final_obs, 9 * N, B), o.engine.buffer_hash(o._final_obs.numpy(), 9 * N, B)), f"{ctx}: final obs"
            assert torch.equal(g._episode_dev.cpu(), o._episode_dev), f"{ctx}: episode counters"
        finished += int((ro[2] | ro[3]).sum())
        rejected += int(ro[4]["invalid_action"].sum())
    g.close()
    o.close()
    return dict(board=[W, W], envs=B, max_turns=max_turns, steps=steps, auto_reset=mode, agent="in_step" if in_step else "sampler", env_steps_compared=B * steps, episodes_finished=finished,
                rejected_actions=rejected, mismatches=0, seconds=round(time.time() - t0, 1))


def main():
    cuda = load_library()
    oracle = BoundLibrary(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    args = [int(v) for v in sys.argv[1:]]
    configs = [tuple(args[i:i + 4]) for i in range(0, len(args), 4)] or [(15, 65536, 120, 300), (20, 32768, 90, 200), (10, 65536, 60, 200)]
    mode = os.environ.get("GRL_SOAK_RESET", "host")
    for cfg in configs:
        print(json.dumps(soak(cuda, oracle, *cfg, mode=mode, in_step=os.environ.get("GRL_SOAK_AGENT") == "in_step")), flush=True)


if __name__ == "__main__":
    main()
