"""Game sharding across the GPUs of one box, and the only collectives of the system.

Games are independent in the reference (one ``Engine`` per game, no shared state:
internal/grpc/gameserver/game_manager.go:20-54), so the step path needs NO collective:
rank r of G owns the contiguous env range ``[r*B/G, (r+1)*B/G)`` (SURVEY.md 8e:
``gpu = env / (B_total / G)``), with its own slabs and its own stream.  ``env_id_base``
keys both the map seeds and the synthetic policy, so a sharded run reproduces the
single-device run game for game.

Collectives exist only for the learner hand-off (BASELINE config 5):
  * ``all_reduce_stats``   episode statistics summed over ranks,
  * ``gather_experience``  experience records gathered to the learner rank as one byte buffer per rank with its
    actual row count (no padding); the record fields follow experience.proto:25-60 / collector.go:57-75
    (state, action, reward, next_state, done, action_mask, player, turn) — either as float32 tensors
    (``pack_experience``) or compact (``pack_experience_packed``: the packed observation records before and after the
    step, 13x fewer bytes over NVLink; ``expand_experience`` rebuilds tensors and mask on the learner).
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


def pack_experience_packed(prev_packed, next_packed, action_index, reward, done, turn, env_id_base: int = 0, limit: int = 0):
    """The same records in their COMPACT form, for the learner hand-off: instead of two float32 tensors per record
    (2 x 14,400 B at 20x20) a record carries the packed observation records of its game before and after the step
    (grl_step_outputs.obs_packed: 2 x 1,120 B) — everything Serializer.StateToTensor and GenerateActionMask read, so
    the learner rebuilds state, next_state and the action mask with ``expand_experience``.

      prev_packed, next_packed [B, RW] i32 (device); action_index [B,P] i32; reward [B,P] f32; done [B] u8.
    ``limit`` > 0 keeps the first ``limit`` records (a per-turn sample)."""
    import torch

    B, P = action_index.shape
    sel = (action_index >= 0).reshape(-1).nonzero(as_tuple=True)[0]
    if limit > 0:
        sel = sel[:limit]
    env = torch.div(sel, P, rounding_mode="floor")
    ply = sel - env * P
    if not torch.is_tensor(turn):
        turn = torch.full((B,), int(turn), dtype=torch.int32, device=action_index.device)
    return {
        "state_packed": prev_packed.index_select(0, env),
        "next_packed": next_packed.index_select(0, env),
        "action": action_index.reshape(-1).index_select(0, sel),
        "reward": reward.reshape(-1).index_select(0, sel),
        "done": done.index_select(0, env),
        "player": ply.to(torch.int32),
        "turn": turn.to(torch.int32).index_select(0, env),
        "env_id": (env + env_id_base).to(torch.int32),
    }


def serializer_mask_from_packed(packed: np.ndarray, player: np.ndarray, W: int, H: int, P: int) -> np.ndarray:
    """Serializer.GenerateActionMask (internal/experience/serializer.go:112-176) from packed observation records: tiles
    the player OWNS (true ownership) with army >= 2, directions up, down, left, right, target in bounds and not a
    mountain.  packed [n, RW] uint32, player [n]; returns bool [n, W*H*4]."""
    n, N = packed.shape[0], W * H
    NW, NA = (N + 31) // 32, (N + 7) & ~7
    bits = lambda words: np.unpackbits(np.ascontiguousarray(words).view(np.uint8), axis=1, bitorder="little")[:, :N].astype(bool)  # noqa: E731
    own_all = packed[:, : P * NW].reshape(n, P, NW)
    own = bits(own_all[np.arange(n), player])
    mnt = bits(packed[:, 2 * P * NW: 2 * P * NW + NW])
    army = np.ascontiguousarray(packed[:, (2 * P + 2) * NW: (2 * P + 2) * NW + NA // 2]).view(np.uint16)[:, :N]
    src = (own & (army >= 2)).reshape(n, H, W)
    free = ~mnt.reshape(n, H, W)
    mask = np.zeros((n, H, W, 4), bool)
    mask[:, 1:, :, 0] = src[:, 1:, :] & free[:, :-1, :]     # up
    mask[:, :-1, :, 1] = src[:, :-1, :] & free[:, 1:, :]    # down
    mask[:, :, 1:, 2] = src[:, :, 1:] & free[:, :, :-1]     # left
    mask[:, :, :-1, 3] = src[:, :, :-1] & free[:, :, 1:]    # right
    return mask.reshape(n, N * 4)


def expand_experience(records: Dict[str, "object"], lib, W: int, H: int, P: int, threads: int = 0) -> Dict[str, np.ndarray]:
    """Learner side of ``pack_experience_packed``: rebuild the float32 ``state`` / ``next_state`` tensors of each
    record's player (grl_expand_obs, bit-identical to the ``obs`` plane of the step) and its serializer action mask."""
    host = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in records.items()}
    n = int(host["action"].shape[0])
    RW = int(lib.obs_packed_words(W, H, P))
    out = {k: v for k, v in host.items() if k not in ("state_packed", "next_packed")}
    ply = host["player"].astype(np.int64)
    for src, dst in (("state_packed", "state"), ("next_packed", "next_state")):
        pk = np.ascontiguousarray(host[src]).view(np.uint32).reshape(n, RW)
        full = np.empty((n, P, 9, H, W), np.float32)
        lib.check(lib.expand_obs(W, H, P, pk.ctypes.data, n, full.ctypes.data, threads), "expand_obs")
        out[dst] = full[np.arange(n), ply]
        if src == "state_packed":
            out["mask_bits"] = serializer_mask_from_packed(pk, ply, W, H, P)
    return out


def gather_rows(buf, dst: int = 0):
    """Variable-size gather of row blocks to rank ``dst``: ONE all_gather of the row counts, then each rank that has rows
    sends exactly those rows (grouped ncclSend/ncclRecv over NVLink on GPUs; gloo send/recv in the CPU tests) — no
    padding travels.  ``buf`` [n, R] (any dtype, the same R on every rank).  Returns (rows [sum n, R] in rank order on
    ``dst`` / None elsewhere, counts as a python list)."""
    import torch

    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return buf, [int(buf.shape[0])]
    world, rank = dist.get_world_size(), dist.get_rank()
    n = int(buf.shape[0])
    counts_t = torch.zeros(world, dtype=torch.int64, device=buf.device)
    dist.all_gather_into_tensor(counts_t, torch.tensor([n], dtype=torch.int64, device=buf.device))
    counts = [int(c) for c in counts_t.tolist()]
    buf = buf.contiguous()
    if rank == dst:
        out = torch.empty((sum(counts),) + tuple(buf.shape[1:]), dtype=buf.dtype, device=buf.device)
        ops, off = [], 0
        for r, c in enumerate(counts):
            if r == dst:
                out[off:off + c] = buf
            elif c > 0:
                ops.append(dist.P2POp(dist.irecv, out[off:off + c], r))
            off += c
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return out, counts
    if n > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, buf, dst)]):
            w.wait()
    return None, counts


def gather_experience(records: Dict[str, "object"], capacity: Optional[int] = None, dst: int = 0) -> Optional[Dict[str, "object"]]:
    """Gather every rank's experience records to the learner rank ``dst`` (BASELINE config 5; the reference's
    hand-off is experience_service.go:287-378 over gRPC).

    The records of a rank travel as ONE contiguous byte buffer [n, record_bytes] — every field's row bytes side by side
    — through ``gather_rows``: a count exchange plus one variable-size transfer per rank, nothing zero-padded.
    ``capacity`` (optional) bounds what a rank may contribute per call; records beyond it are NOT sent and are reported:
    the result carries ``"dropped"`` (int64 [world], records each rank held back) so a lossy hand-off is never silent.
    Returns the concatenated records (fields as given, in rank order) on ``dst`` and None elsewhere."""
    import torch

    dist = _dist()
    names = [k for k in records if k not in ("dropped", "counts")]
    n = int(records[names[0]].shape[0])
    keep = n if capacity is None else min(n, int(capacity))
    dev = records[names[0]].device
    cols, layout = [], []
    for name in names:
        t = records[name][:keep].contiguous()
        b = t.reshape(keep, -1).view(torch.uint8) if t.dtype != torch.uint8 else t.reshape(keep, -1)
        layout.append((name, t.dtype, tuple(t.shape[1:]), b.shape[1]))
        cols.append(b)
    buf = torch.cat(cols, dim=1) if keep or cols else torch.zeros((0, 0), dtype=torch.uint8, device=dev)
    rows, counts = gather_rows(buf, dst)
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    dropped = torch.tensor([n - keep], dtype=torch.int64, device=dev)
    if world > 1:
        all_dropped = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_dropped, dropped)
        dropped = all_dropped
    if rank != dst:
        return None
    out, off = {}, 0
    for name, dtype, shape, width in layout:
        col = rows[:, off:off + width].contiguous()
        out[name] = (col if dtype == torch.uint8 else col.view(dtype)).reshape((rows.shape[0],) + shape)
        off += width
    out["dropped"] = dropped
    out["counts"] = torch.tensor(counts, dtype=torch.int64)
    return out
