#!/usr/bin/env python3
"""Full-size parity soak: BASELINE-size batches played to the 500-turn cap on the GPU and on the CPU
oracle with the same recorded actions, compared EVERY turn (full-state digest of every game, reward
bits, done, winner, step_error, packed masks; observation digests every 25th turn), then re-seeded
and played on.  Prints one JSON line per configuration (profiles/r1_parity_soak.jsonl).

usage: python tests/tools/parity_soak.py [W H P B T] ...   (default: the four BASELINE shapes)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from generalsreinforcementlearning_b200 import load_library
from generalsreinforcementlearning_b200._abi import BoundLibrary
from helpers import new_engine


def soak(cuda, oracle, W, H, P, B, T):
    gc = new_engine(cuda, W, H, P, B, host_threads=0)
    oc = new_engine(oracle, W, H, P, B, host_threads=0)
    dev = torch.device("cuda:0")
    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    go = gc.alloc_outputs_host()
    del go["obs"]
    oo = oc.alloc_outputs_host()
    oo_small = {k: v for k, v in oo.items() if k != "obs"}
    t0 = time.time()
    errors = finished = compared = 0
    for episode, turns in ((0, T), (1, 60)):
        seeds = np.arange(B, dtype=np.int64) + 12345 + episode * B
        gc.reset_seeded(seeds)
        oc.reset_seeded(seeds)
        assert np.array_equal(gc.state_hash(), oc.state_hash()), "maps differ"
        for t in range(turns):
            acts = oc.sample_actions(2024 + episode)
            with_obs = t % 25 == 24
            gc.step_fused(acts, gc.outputs(obs=obs, **go))
            oc.step_fused(acts, oc.outputs(**(oo if with_obs else oo_small)))
            ctx = f"{W}x{H}x{P}p episode {episode} turn {t}"
            assert np.array_equal(gc.state_hash(), oc.state_hash()), ctx
            for k in go:
                a, b = go[k], oo[k]
                if a.dtype == np.float32:
                    a, b = a.view(np.uint32), b.view(np.uint32)
                assert np.array_equal(a, b), f"{ctx}: {k}"
            if with_obs:
                assert np.array_equal(gc.buffer_hash(obs, 9 * W * H, B * P), oc.buffer_hash(oo["obs"], 9 * W * H, B * P)), ctx
            errors += int((oo["step_error"] != 0).sum())
            compared += B
        finished += int(oo["done"].sum())
    assert np.array_equal(gc.stats(), oc.stats())
    return dict(board=[W, H], players=P, games=B, turns=T + 60, game_turns_compared=compared, error_turns=errors,
                games_finished=finished, mismatches=0, seconds=round(time.time() - t0, 1))


def main():
    cuda = load_library()
    oracle = BoundLibrary(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    args = [int(v) for v in sys.argv[1:]]
    configs = [tuple(args[i:i + 5]) for i in range(0, len(args), 5)] or [
        (10, 10, 2, 65536, 500), (15, 15, 2, 65536, 500), (20, 20, 2, 65536, 500), (20, 20, 4, 32768, 500)]
    for cfg in configs:
        print(json.dumps(soak(cuda, oracle, *cfg)), flush=True)


if __name__ == "__main__":
    main()
