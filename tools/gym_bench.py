#!/usr/bin/env python3
"""Steps per second through the generals_gym-compatible vector env (GeneralsVecEnv) on one GPU:
random legal agent actions (torch.multinomial over the mask), the random opponent, observation,
mask, reward, auto-reset — everything a training loop calls per step.  The reference's own
gym path measures 12 steps/s for one env and 250.8 steps/s for 16 envs."""
import json
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch

from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv


def main():
    out = []
    for (B, W) in ((1024, 15), (16384, 15), (65536, 15), (65536, 20)):
        env = GeneralsVecEnv(B, W, W, max_turns=500, seed=3)
        obs, info = env.reset()
        act = env.sample_actions
        for _ in range(5):
            obs, r, term, trunc, info = env.step(act())
        torch.cuda.synchronize()
        steps = 60
        t_policy = t_env = 0.0
        t0 = time.perf_counter()
        for _ in range(steps):
            a0 = time.perf_counter()
            action = act()
            torch.cuda.synchronize()
            a1 = time.perf_counter()
            obs, r, term, trunc, info = env.step(action)
            torch.cuda.synchronize()
            t_policy += a1 - a0
            t_env += time.perf_counter() - a1
        dt = time.perf_counter() - t0
        rec = {"envs": B, "board": W, "env_steps_per_s": round(B * steps / dt), "ms_per_vector_step": round(1e3 * dt / steps, 3),
               "ms_env_step": round(1e3 * t_env / steps, 3), "ms_random_policy": round(1e3 * t_policy / steps, 3),
               "env_only_steps_per_s": round(B * steps / t_env)}
        # steady state of a long run: episode ends are spread over time, so EVERY step re-seeds B/max_turns envs
        # (device map generation, turn-0 set-up, read-outs of the new games) on top of the step itself
        env._calls.copy_(torch.randint(0, env.max_turns, (B,), device=env._calls.device, dtype=torch.int32))
        for _ in range(5):
            env.step(act())
        torch.cuda.synchronize()
        t_env = 0.0
        resets = 0
        for _ in range(steps):
            action = act()
            torch.cuda.synchronize()
            a1 = time.perf_counter()
            obs, r, term, trunc, info = env.step(action)
            torch.cuda.synchronize()
            t_env += time.perf_counter() - a1
            resets += int((term | trunc).sum())
        rec.update({"steady_ms_env_step": round(1e3 * t_env / steps, 3), "steady_env_only_steps_per_s": round(B * steps / t_env),
                    "steady_resets_per_step": round(resets / steps, 1)})
        out.append(rec)
        env.close()
        # the same steady state with the device-side auto-reset: no host read in step(), one sync at the end
        env = GeneralsVecEnv(B, W, W, max_turns=500, seed=3, auto_reset="device")
        env.reset()
        env._calls.copy_(torch.randint(0, env.max_turns, (B,), device=env._calls.device, dtype=torch.int32))
        for _ in range(5):
            env.step(env.sample_actions())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            env.step(env.sample_actions())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rec.update({"device_reset_ms_per_vector_step": round(1e3 * dt / steps, 3), "device_reset_env_steps_per_s": round(B * steps / dt)})
        env.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
