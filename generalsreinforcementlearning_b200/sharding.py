"""Game sharding across the GPUs of one box, and the only collectives of the system.

Games are independent in the reference (one ``Engine`` per game, no shared state:
internal/grpc/gameserver/game_manager.go:20-54), so the step path needs NO collective:
rank r of G owns the contiguous env range ``[r*B/G, (r+1)*B/G)`` (SURVEY.md 8e:
``gpu = env / (B_total / G)``), with its own slabs and its own stream.  ``env_id_base``
keys both the map seeds and the synthetic policy, so a sharded run reproduces the
single-device run game for game.

Collectives exist only for the learner hand-off (BASELINE config 5):
  * ``all_reduce_stats``   episode statistics summed over ranks,
  * ``gather_experience``  fixed-size experience records gathered to the learner rank —
    the record layout follows experience.proto:25-60 / collector.go:57-75
    (state, action, reward, next_state, done, action_mask, player, turn).
``torch.distributed`` is the plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    total_envs: int
    first: int   # global id of this rank's env 0  (== grl_config.env_id_base)
    count: int   # envs resident on this rank

    def owner_of(self, env_id: int) -> int:
        """Rank that owns a global env id."""
        return owner_rank(self.total_envs, self.world, env_id)

    def local(self, env_id: int) -> int:
        if not (self.first <= env_id < self.first + self.count):
            raise IndexError(f"env {env_id} is not resident on rank {self.rank}")
        return env_id - self.first

    def seeds(self, base_seed: int, episode: int = 0) -> np.ndarray:
        """Map seeds of this shard: seed_i = base_seed + global env id (+ total_envs per
        completed episode, so a re-seeded env never repeats a map of the run)."""
        ids = np.arange(self.first, self.first + self.count, dtype=np.int64)
        return ids + int(base_seed) + int(episode) * self.total_envs


def shard_for(total_envs: int, world: int, rank: int) -> Shard:
    """Contiguous block partition; the first ``total_envs % world`` ranks hold one extra env."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    if total_envs < world:
        raise ValueError("fewer envs than ranks")
    q, r = divmod(total_envs, world)
    first = rank * q + min(rank, r)
    return Shard(rank, world, total_envs, first, q + (1 if rank < r else 0))


def owner_rank(total_envs: int, world: int, env_id: int) -> int:
    if not (0 <= env_id < total_envs):
        raise IndexError(env_id)
    q, r = divmod(total_envs, world)
    cut = r * (q + 1)
    return env_id // (q + 1) if env_id < cut else r + (env_id - cut) // q


def shard_from_env(total_envs: int) -> Shard:
    """Shard of this process under torchrun (RANK / WORLD_SIZE)."""
    return shard_for(total_envs, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")))


def create_sharded_engine(lib, shard: Shard, device: int = 0, **config):
    """One BatchedEngine holding this rank's envs (``env_id_base`` = the shard offset)."""
    from .engine import BatchedEngine, make_config

    cfg = make_config(lib, num_envs=shard.count, env_id_base=shard.first, device=device, **config)
    return BatchedEngine(lib, cfg)


# ---------------------------------------------------------------------------------------
# collectives (learner hand-off only)
# ---------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist

    return dist


def all_reduce_stats(stats, device=None):
    """Sum the lifetime counters of grl_stats ([env-steps, error turns, games finished,
    rejected steps]) — or any small numeric vector — over all ranks.  Returns numpy int64."""
    import torch

    dist = _dist()
    t = torch.as_tensor(np.asarray(stats).astype(np.int64))
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


EXPERIENCE_FIELDS = ("state", "next_state", "action", "reward", "done", "mask_bits", "player", "turn", "env_id")


def pack_experience(prev_obs, obs, prev_mask_bits, action_index, reward, done, turn, env_id_base: int = 0):
    """Build the experience records of one step from the fused step's planes, the way
    SimpleCollector.OnStateTransition does (internal/experience/collector.go:30-98): one
    record per player that submitted a move (action_index >= 0; aborted turns emit none),
    state/mask from BEFORE the step, next_state/reward/done from after it.

    All arguments are torch tensors on one device:
      prev_obs, obs [B,P,9,H,W] f32; prev_mask_bits [B,P,Wm] i32; action_index [B,P] i32;
      reward [B,P] f32; done [B] u8; turn: int or [B] tensor.
    Returns a dict of tensors with a leading record dimension (selection runs on the device)."""
    import torch

    B, P = action_index.shape
    sel = (action_index >= 0).reshape(-1).nonzero(as_tuple=True)[0]
    env = torch.div(sel, P, rounding_mode="floor")
    ply = sel - env * P
    flat = lambda t: t.reshape(B * P, *t.shape[2:])  # noqa: E731
    if not torch.is_tensor(turn):
        turn = torch.full((B,), int(turn), dtype=torch.int32, device=action_index.device)
    return {
        "state": flat(prev_obs).index_select(0, sel),
        "next_state": flat(obs).index_select(0, sel),
        "action": action_index.reshape(-1).index_select(0, sel),
        "reward": reward.reshape(-1).index_select(0, sel),
        "done": done.index_select(0, env),
        "mask_bits": flat(prev_mask_bits).index_select(0, sel),
        "player": ply.to(torch.int32),
        "turn": turn.to(torch.int32).index_select(0, env),
        "env_id": (env + env_id_base).to(torch.int32),
    }


def gather_experience(records: Dict[str, "object"], capacity: int, dst: int = 0) -> Optional[Dict[str, "object"]]:
    """Gather up to ``capacity`` records per rank to the learner rank ``dst``.

    Fixed-size exchange: every rank contributes a [capacity, ...] block per field
    plus its valid count, so the collective's shape never depends on the data; the learner
    compacts.  One ``gather`` per field (grouped ncclSend/ncclRecv over NVLink on GPUs).  Returns the
    concatenated records on ``dst`` and None elsewhere."""
    import torch

    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = int(records["action"].shape[0])
    keep = min(n, capacity)
    dev = records["action"].device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([keep], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(counts, mine)
    else:
        counts[0] = keep
    out = {}
    for name in EXPERIENCE_FIELDS:
        t = records[name]
        block = torch.zeros((capacity,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        block[:keep] = t[:keep]
        if world > 1:
            parts = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
            dist.gather(block, parts, dst=dst)
        else:
            parts = [block]
        if rank == dst:
            out[name] = torch.cat([parts[r][: int(counts[r])] for r in range(world)], 0)
    return out if rank == dst else None
