#!/usr/bin/env python3
"""ParallelEnvPool over GeneralsVecEnv: env-steps/s INTO the replay ring in HBM (state, action, reward, next_state, done
of every env, every step), the reference's collection benchmark (documentation/claude/parallel-experience-collection-plan.md:
10x10, 300-step cap; 18.4 / 69.8 / 138.3 / 250.8 steps/s with 1 / 4 / 8 / 16 worker threads over gRPC).
usage: python tools/pool_bench.py [W B steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
from generalsreinforcementlearning_b200.parallel_env import RANDOM_AGENT, ParallelEnvPool, ReplayBuffer


def bench(W, B, steps, agent="in_step"):
    """agent: "in_step" = the random agent drawn inside the vector step's launch (RANDOM_AGENT), "sampler" = a
    grl_gym_sample launch of its own per step."""
    vec = GeneralsVecEnv(B, W, W, max_turns=300, seed=5, auto_reset="device")
    buf = ReplayBuffer(capacity=8 * B)
    pool = ParallelEnvPool(B, vec_env=vec, batch_action_fn=RANDOM_AGENT if agent == "in_step" else (lambda s, m: vec.sample_actions()),
                           replay_buffer=buf, max_steps_per_episode=300, seed=5)
    pool.run(1)
    vec._calls.copy_(torch.randint(0, 300, (B,), device=vec.device, dtype=torch.int32))   # episode ends spread over time
    pool.run(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    pool.run(steps)
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    s, a, r, ns, d = buf.sample_tensors(min(4096, len(buf) - 2 * B))
    out = {"board": [W, W], "envs": B, "steps": steps, "agent": agent, "ms_per_vector_step": ms, "host_enqueue_ms": host_ms,
           "transitions_per_s": B / (ms * 1e-3), "episodes": pool.total_episodes, "buffer_rows": len(buf),
           "sampled_batch": list(s.shape), "bytes_per_transition": int(2 * s[0].numel() * 4 + 8 + 4 + 1),
           "reference": "250.8 steps/s with 16 worker threads over gRPC (10x10, 300-step cap)"}
    pool.close()
    return out


if __name__ == "__main__":
    args = [int(v) for v in sys.argv[1:]]
    cfgs = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [(10, 65536, 60), (15, 65536, 60), (10, 16, 200)]
    for cfg in cfgs:
        for agent in ("sampler", "in_step"):
            print(json.dumps(bench(*cfg, agent=agent)), flush=True)
