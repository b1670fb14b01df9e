#!/usr/bin/env python3
"""Marginal cost of the fused gym step's read-outs: the launch timed with and without the observation tensor
(CUDA events, device-resident).  usage: python tools/gym_parts.py W B"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv


def main():
    W, B = int(sys.argv[1]), int(sys.argv[2])
    env = GeneralsVecEnv(B, W, W, max_turns=500, seed=3, auto_reset="device")
    env.reset()
    for _ in range(30):
        env.step(env.sample_actions())
    torch.cuda.synchronize()
    res = {}
    for name in ("full", "no_obs"):
        tot, T = 0.0, 40
        for _ in range(T):
            a = env.sample_actions()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            env._opp_draws += 1
            env._flip ^= 1
            o = env._out[env._flip]
            e0.record()
            env.engine.gym_step(env.max_turns, env._base_seed * 1000003 + env._opp_draws, action=a, opponent_action=None,
                                obs=env._obs if name == "full" else None, mask=env._mask, stats=env._stats, actions=env._actions,
                                prev_stats=env._prev_stats, turns=env._turns, calls=env._calls, reward=o["reward"],
                                terminated=o["terminated"], truncated=o["truncated"], valid=o["valid"], done=env._done,
                                winner=o["winner"], step_error=o["step_error"], n_finished=env._nfin)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = round(tot / T, 4)
    print(json.dumps({"board": W, "envs": B, "lib": os.environ.get("GRL_LIB_PATH", "default"), "gym_step_ms": res}))


if __name__ == "__main__":
    main()
