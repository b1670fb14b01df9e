#!/usr/bin/env python3
"""Soak of the launch overlap: whole episodes stepped as ONE chain of unsynchronised launches (every launch overlaps its
predecessor on the device: programmatic dependent launch + per-warp epoch words, csrc/grl_turn.cuh), against the oracle
stepping the same turns.  Compared at the end of every chain: every env's full-state digest, the counters, and the last
launch's reward / done / mask planes and observation digests.  One JSON line per configuration.

usage: python tests/tools/chain_soak.py [W H P B chain_length chains] ..."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from generalsreinforcementlearning_b200 import _abi, load_library
from generalsreinforcementlearning_b200._abi import BoundLibrary
from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config


def soak(cuda, oracle, W, H, P, B, L, chains):
    dev = torch.device("cuda:0")
    gc = BatchedEngine(cuda, make_config(cuda, num_envs=B, width=W, height=H, num_players=P, host_threads=0))
    oc = BatchedEngine(oracle, make_config(oracle, num_envs=B, width=W, height=H, num_players=P, host_threads=0))
    seeds = np.arange(B, dtype=np.int64) + 31337
    gc.reset_seeded(seeds)
    oc.reset_seeded(seeds)
    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    mask = torch.empty((B, P, gc.mask_words), dtype=torch.int32, device=dev)
    reward = torch.empty((B, P), dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    outs = gc.outputs(obs=obs, mask_bits=mask, reward=reward, done=done)
    oo = oc.alloc_outputs_host()
    small = {k: oo[k] for k in ("mask_bits", "reward", "done")}
    t0 = time.time()
    launches = 0
    for c in range(chains):
        for t in range(L):  # no host synchronisation inside a chain
            if t % 5 == 2:
                gc.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 99 + c)
            else:
                gc.step_fused(None, outs, _abi.STEP_FLAG_RANDOM_POLICY, 99 + c)
        for t in range(L):
            if t % 5 == 2:
                oc.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 99 + c)
            else:
                oc.step_fused(None, oc.outputs(**(oo if t == L - 1 else small)), _abi.STEP_FLAG_RANDOM_POLICY, 99 + c)
        launches += L
        ctx = f"{W}x{H}x{P}p chain {c}"
        assert np.array_equal(gc.state_hash(), oc.state_hash()), ctx
        assert np.array_equal(reward.cpu().numpy().view(np.uint32), oo["reward"].view(np.uint32)), ctx
        assert np.array_equal(done.cpu().numpy(), oo["done"]) and np.array_equal(mask.cpu().numpy().view(np.uint32), oo["mask_bits"]), ctx
        assert np.array_equal(gc.buffer_hash(obs, 9 * W * H, B * P), oc.buffer_hash(oo["obs"], 9 * W * H, B * P)), ctx
        assert np.array_equal(gc.stats(), oc.stats()), ctx
    st = gc.stats()
    return dict(board=[W, H], players=P, games=B, chain_length=L, chains=chains, overlapped_launches=launches - chains,
                game_turns=int(st[0]), games_finished=int(oo["done"].sum()), mismatches=0, seconds=round(time.time() - t0, 1))


def main():
    cuda = load_library()
    oracle = BoundLibrary(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    args = [int(v) for v in sys.argv[1:]]
    configs = [tuple(args[i:i + 6]) for i in range(0, len(args), 6)] or [
        (10, 10, 2, 65536, 100, 5), (15, 15, 2, 65536, 100, 4), (20, 20, 2, 65536, 100, 3), (20, 20, 4, 32768, 100, 2)]
    for cfg in configs:
        print(json.dumps(soak(cuda, oracle, *cfg)), flush=True)


if __name__ == "__main__":
    main()
