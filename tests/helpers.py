"""Shared test helpers: hand-built game states in the style of the reference's Go tests."""
from __future__ import annotations

import numpy as np

from generalsreinforcementlearning_b200 import _abi
from generalsreinforcementlearning_b200.engine import BatchedEngine, make_actions, make_config, set_action

NORMAL, GENERAL, CITY, MOUNTAIN = 0, 1, 2, 3


def new_engine(lib, W, H, P, B=1, **cfg) -> BatchedEngine:
    cfg.setdefault("max_actions", max(2, P))
    cfg.setdefault("host_threads", 1)
    return BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, **cfg))


def blank_state(W, H, P, count=1):
    """core.NewBoard + fresh players (engine_initializer.go:113-143 / action_mask_test.go:16-55)."""
    N = W * H
    return dict(
        owner=np.full((count, N), -1, np.int32),
        army=np.zeros((count, N), np.int32),
        type=np.zeros((count, N), np.int32),
        visible=np.zeros((count, N), np.uint32),
        owned=np.zeros((count, P, N), np.uint8),
        changed=np.zeros((count, N), np.uint8),
        vis_changed=np.zeros((count, N), np.uint8),
        turn=np.zeros(count, np.int32),
        game_over=np.zeros(count, np.int32),
        alive=np.ones((count, P), np.int32),
        army_count=np.zeros((count, P), np.int32),
        general_idx=np.full((count, P), -1, np.int32),
        step_error=np.zeros(count, np.int32),
    )


def put(state, W, x, y, owner=-1, army=0, type_=NORMAL, env=0):
    i = y * W + x
    state["owner"][env, i] = owner
    state["army"][env, i] = army
    state["type"][env, i] = type_
    return i


def full_stats(state):
    """What Engine.updatePlayerStats does at turn 0 (stats.go:33-63): the Go tests call it
    after editing the board by hand."""
    count, P, N = state["owned"].shape
    for c in range(count):
        for p in range(P):
            own = state["owner"][c] == p
            state["owned"][c, p] = own.astype(np.uint8)
            state["army_count"][c, p] = int(state["army"][c][own].sum())
            gens = np.nonzero(own & (state["type"][c] == GENERAL))[0]
            state["general_idx"][c, p] = int(gens[-1]) if len(gens) else -1
            state["alive"][c, p] = 1 if len(gens) else 0
    return state


def full_fog(state, W, H):
    """performFullVisibilityUpdateOptimized (visibility_optimized.go:33-53) for test setup."""
    count, P, N = state["owned"].shape
    state["visible"][:] = 0
    for c in range(count):
        for p in range(P):
            if not state["alive"][c, p]:
                continue
            for i in np.nonzero(state["owned"][c, p])[0]:
                x, y = i % W, i // W
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        nx, ny = x + dx, y + dy
                        if 0 <= nx < W and 0 <= ny < H:
                            state["visible"][c, ny * W + nx] |= np.uint32(1 << p)
    return state


def one_action(engine, player, fx, fy, tx, ty, move_all=True, env=0, slot=0, actions=None):
    if actions is None:
        actions = make_actions(engine.B, engine.A)
    set_action(actions, env, slot, player, fx, fy, tx, ty, move_all)
    return actions


def compare_states(a, b, ctx=""):
    """Bit-exact comparison of two get_state() dicts (general_idx uses the canonical
    highest-index tie-break in both implementations)."""
    for k in a:
        if not np.array_equal(a[k], b[k]):
            bad = np.argwhere(np.asarray(a[k]) != np.asarray(b[k]))
            raise AssertionError(f"{ctx}: state plane {k!r} differs at {bad[:8].tolist()} "
                                 f"({len(bad)} cells): {np.asarray(a[k])[tuple(bad[0])]} vs "
                                 f"{np.asarray(b[k])[tuple(bad[0])]}")


def ctypes_fn(lib, name, restype, argtypes):
    fn = getattr(lib.cdll, lib.prefix + name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn
