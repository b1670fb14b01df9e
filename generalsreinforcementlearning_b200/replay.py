"""Replay / state codec over ``grl_get_state`` / ``grl_set_state``.

The reference only has the in-memory ``GameState.Clone`` (internal/game/state.go:37-70).  Because
the turn engine is deterministic given (seeds, action sequence), a replay is just that pair:

  *.grlreplay   zip container: ``meta.json`` (board, players, seeds, ABI version, digests every
                ``digest_every`` turns) + ``actions.npy`` uint8 [T][B][A][8] (grl_action records,
                including the SKIP_ENV flag) — written with ``ReplayWriter``, verified with ``replay()``.
  *.grlstate    checkpoint: every plane of ``grl_state_planes`` for all envs (npz) — ``save_state`` /
                ``load_state`` round-trip bit for bit, including the cached OwnedTiles sets and the
                changed / visibility-changed tile sets that the next turn depends on.
"""
from __future__ import annotations

import io
import json
import zipfile
from typing import Dict, Optional

import numpy as np

from . import _abi
from .engine import BatchedEngine

FORMAT_VERSION = 1


def _to_host_actions(actions, B: int, A: int) -> np.ndarray:
    if actions is None:
        return np.zeros((B, A, 8), np.uint8)
    if hasattr(actions, "detach"):
        actions = actions.detach().cpu().numpy()
    a = np.ascontiguousarray(actions)
    return a.view(np.uint8).reshape(B, A, 8).copy()


class ReplayWriter:
    """Record a run: ``w = ReplayWriter(engine, seeds); ...; w.record(actions) after each step; w.save(path)``."""

    def __init__(self, engine: BatchedEngine, seeds, digest_every: int = 50):
        self.engine = engine
        self.seeds = np.asarray(seeds, np.int64).copy()
        self.digest_every = int(digest_every)
        self.steps = []
        self.digests: Dict[int, list] = {}

    def record(self, actions) -> None:
        """Call right AFTER the engine stepped with ``actions`` (host array, torch tensor or None)."""
        self.steps.append(_to_host_actions(actions, self.engine.B, self.engine.A))
        t = len(self.steps)
        if self.digest_every and t % self.digest_every == 0:
            self.digests[t] = [int(v) for v in self.engine.state_hash()]

    def save(self, path: str) -> None:
        e = self.engine
        self.digests[len(self.steps)] = [int(v) for v in e.state_hash()]
        meta = dict(format=FORMAT_VERSION, abi=_abi.GRL_OK, width=e.W, height=e.H, players=e.P, num_envs=e.B,
                    max_actions=e.A, fog_of_war=int(e.cfg.fog_of_war), turns=len(self.steps),
                    seeds=[int(s) for s in self.seeds], digests={str(k): v for k, v in self.digests.items()})
        buf = io.BytesIO()
        np.save(buf, np.stack(self.steps) if self.steps else np.zeros((0, e.B, e.A, 8), np.uint8))
        with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
            z.writestr("meta.json", json.dumps(meta))
            z.writestr("actions.npy", buf.getvalue())


def load_replay(path: str):
    with zipfile.ZipFile(path) as z:
        meta = json.loads(z.read("meta.json"))
        actions = np.load(io.BytesIO(z.read("actions.npy")))
    return meta, actions


def replay(engine: BatchedEngine, path: str, until: Optional[int] = None) -> int:
    """Re-run a recorded game batch on ``engine`` (any implementation of the ABI) and check every
    recorded digest.  Returns the number of turns replayed; raises on the first divergence."""
    meta, actions = load_replay(path)
    for k in ("width", "height", "players", "num_envs", "max_actions"):
        have = {"width": engine.W, "height": engine.H, "players": engine.P, "num_envs": engine.B, "max_actions": engine.A}[k]
        if have != meta[k]:
            raise ValueError(f"replay was recorded with {k}={meta[k]}, engine has {have}")
    engine.reset_seeded(np.asarray(meta["seeds"], np.int64))
    T = meta["turns"] if until is None else min(until, meta["turns"])
    for t in range(T):
        engine.step(actions[t].view(_abi.ACTION_DTYPE).reshape(engine.B, engine.A))
        want = meta["digests"].get(str(t + 1))
        if want is not None:
            got = engine.state_hash()
            if not np.array_equal(got, np.asarray(want, np.uint64)):
                bad = np.nonzero(got != np.asarray(want, np.uint64))[0]
                raise AssertionError(f"replay diverged at turn {t + 1} in {len(bad)} env(s), first env {int(bad[0])}")
    return T


def save_state(engine: BatchedEngine, path: str) -> None:
    st = engine.get_state()
    np.savez_compressed(path, __format=np.int32(FORMAT_VERSION), __dims=np.array([engine.W, engine.H, engine.P, engine.B]), **st)


def load_state(engine: BatchedEngine, path: str) -> None:
    z = np.load(path)
    W, H, P, B = (int(v) for v in z["__dims"])
    if (W, H, P, B) != (engine.W, engine.H, engine.P, engine.B):
        raise ValueError(f"checkpoint is {W}x{H}x{P}p x{B}, engine is {engine.W}x{engine.H}x{engine.P}p x{engine.B}")
    engine.set_state({k: z[k] for k in z.files if not k.startswith("__")})
