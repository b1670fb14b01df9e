"""Replay / state codec over ``grl_get_state`` / ``grl_set_state``.

The reference only has the in-memory ``GameState.Clone`` (internal/game/state.go:37-70).  Because
the turn engine is deterministic given (seeds, action sequence), a replay is just that pair:

  *.grlreplay   zip container: ``meta.json`` (ABI version, the full rule configuration — board, players, fog,
                production constants, env_id_base —, seeds, per-step flags and policy seeds, digests every
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

FORMAT_VERSION = 2   # 2: ABI version and the full rule configuration in meta.json; per-step flags / policy seeds

# every grl_config field that shapes a trajectory: a replay only means something on an engine with the same rules
_CONFIG_KEYS = ("width", "height", "num_players", "num_envs", "max_actions", "fog_of_war", "env_id_base", "city_ratio",
                "city_start_army", "min_general_spacing", "production_general", "production_city", "production_normal",
                "normal_growth_interval")


def _config_of(engine: BatchedEngine) -> Dict[str, int]:
    return {k: int(getattr(engine.cfg, k)) for k in _CONFIG_KEYS}


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
        self.flags = []          # per step: (grl_step flags, policy_seed)
        self.digests: Dict[int, list] = {}

    def record(self, actions, flags: int = 0, policy_seed: int = 0) -> None:
        """Call right AFTER the engine stepped with ``actions`` (host array, torch tensor or None) — or, for a step
        taken with GRL_STEP_FLAG_RANDOM_POLICY, with the step's ``flags`` and ``policy_seed`` (the counter-based policy is
        a pure function of them, so the step replays without its moves being stored)."""
        self.steps.append(_to_host_actions(None if flags & _abi.STEP_FLAG_RANDOM_POLICY else actions, self.engine.B, self.engine.A))
        self.flags.append((int(flags), int(policy_seed)))
        t = len(self.steps)
        if self.digest_every and t % self.digest_every == 0:
            self.digests[t] = [int(v) for v in self.engine.state_hash()]

    def save(self, path: str) -> None:
        e = self.engine
        self.digests[len(self.steps)] = [int(v) for v in e.state_hash()]
        meta = dict(format=FORMAT_VERSION, abi=int(e.lib.abi_version()), config=_config_of(e), turns=len(self.steps),
                    step_flags=[list(f) for f in self.flags],
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
    if meta.get("format") != FORMAT_VERSION:
        raise ValueError(f"replay format {meta.get('format')} (this codec reads format {FORMAT_VERSION})")
    # ABI revisions only ADD entry points and trailing struct fields (include/grlcuda.h): a recording made with an older
    # revision replays on a newer library — its digests catch any drift — but not the other way round
    if meta["abi"] > int(engine.lib.abi_version()):
        raise ValueError(f"replay was recorded with ABI version {meta['abi']}, engine library has {int(engine.lib.abi_version())}")
    have = _config_of(engine)
    diff = {k: (meta["config"].get(k), have[k]) for k in _CONFIG_KEYS if meta["config"].get(k) != have[k]}
    if diff:   # a configuration mismatch, not a divergence
        raise ValueError("replay was recorded under a different configuration: "
                         + ", ".join(f"{k}={a} (engine has {b})" for k, (a, b) in diff.items()))
    engine.reset_seeded(np.asarray(meta["seeds"], np.int64))
    T = meta["turns"] if until is None else min(until, meta["turns"])
    for t in range(T):
        flags, policy_seed = meta["step_flags"][t]
        if flags & _abi.STEP_FLAG_RANDOM_POLICY:
            engine.step(None, flags, policy_seed)
        else:
            engine.step(actions[t].view(_abi.ACTION_DTYPE).reshape(engine.B, engine.A), flags)
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
