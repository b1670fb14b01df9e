#!/usr/bin/env python3
"""Time the turn kernel's instantiations separately on one GPU (CUDA events, kernel's stream):
step only (no read-outs), read-outs only (observe), fused, fused without obs, fused obs only.
usage: python tools/phase_bench.py [W H P B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from generalsreinforcementlearning_b200 import load_library
from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config


def main():
    W, H, P, B = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (20, 20, 2, 65536)
    lib = load_library()
    dev = torch.device("cuda:0")
    e = BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, host_threads=0))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    e.set_stream(stream.cuda_stream)
    seeds = np.arange(B, dtype=np.int64) + 12345
    T = 60
    e.reset_seeded(seeds)
    rec = torch.empty((T + 5, B, e.A, 8), dtype=torch.uint8, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    for t in range(T + 5):
        e.sample_actions(2024, rec[t])
        e.step_fused(rec[t], e.outputs(done=done))
    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    mask = torch.empty((B, P, e.mask_words), dtype=torch.int32, device=dev)
    reward = torch.empty((B, P), dtype=torch.float32, device=dev)
    modes = {
        "fused(obs+mask+reward+done)": lambda t: e.step_fused(rec[t], e.outputs(obs=obs, mask_bits=mask, reward=reward, done=done)),
        "step only": lambda t: e.step(rec[t]),
        "fused(mask+reward+done)": lambda t: e.step_fused(rec[t], e.outputs(mask_bits=mask, reward=reward, done=done)),
        "fused(obs only)": lambda t: e.step_fused(rec[t], e.outputs(obs=obs)),
        "observe(obs+mask) no step": lambda t: e.observe(e.outputs(obs=obs, mask_bits=mask, reward=reward, done=done)),
        "observe(obs) no step": lambda t: e.observe(e.outputs(obs=obs)),
    }
    res = {}
    for name, fn in modes.items():
        e.reset_seeded(seeds)
        for t in range(5):
            fn(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for t in range(5, 5 + T):
            fn(t)
        e1.record(stream)
        torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) / T, 4)
    from bench import algorithmic_bytes_per_env_step
    gym = {}
    if P == 2 and "--no-gym" not in sys.argv:   # the fused gym step (grl_gym_step), device-timed launch by launch
        from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
        e.close()
        env = GeneralsVecEnv(B, W, H, max_turns=500, seed=3, auto_reset="device")
        env.reset()
        for _ in range(30):
            env.step(env.sample_actions())
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(T):
            a = env.sample_actions()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            env._opp_draws += 1
            env._flip ^= 1
            o = env._out[env._flip]
            e0.record()
            env.engine.gym_step(env.max_turns, env._base_seed * 1000003 + env._opp_draws, action=a, opponent_action=None,
                                obs=env._obs, mask=env._mask, stats=env._stats, actions=env._actions, prev_stats=env._prev_stats,
                                turns=env._turns, calls=env._calls, reward=o["reward"], terminated=o["terminated"],
                                truncated=o["truncated"], valid=o["valid"], done=env._done, winner=o["winner"],
                                step_error=o["step_error"], n_finished=env._nfin)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        a = algorithmic_bytes_per_env_step(W, H, P)
        gym_bytes = a["read"] + a["write"] - a["mask"] + 5 * W * H * P + 16 * P + 24
        gym = {"gym_step_ms": round(tot / T, 4), "gym_bytes_per_env_step": gym_bytes, "gym_GBs": round(gym_bytes * B / (tot / T) / 1e6, 1)}
        env.close()
    alg = algorithmic_bytes_per_env_step(W, H, P)["total"]
    fused = res["fused(obs+mask+reward+done)"]
    print(json.dumps({"config": [W, H, P, B], "bytes_per_env_step": alg,
                      "fused_GBs": round(alg * B / fused / 1e6, 1), "fused_Msteps": round(B / fused / 1e3, 1),
                      "prefetch": os.environ.get("GRL_PREFETCH_DIST", "default"), "ms_per_launch": res, **gym}))


if __name__ == "__main__":
    main()
