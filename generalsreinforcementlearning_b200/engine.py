"""Host-side mirror of the reference's ``game.Engine`` surface over the C ABI.

``BatchedEngine`` drives B games through one ``grl_env``; method names follow the
reference (internal/game/engine.go): ``step`` = ``Engine.Step``,
``get_legal_action_mask`` = ``Engine.GetLegalActionMask``, ``compute_player_visibility``
= ``Engine.ComputePlayerVisibility``, ``is_game_over`` / ``get_winner``,
``game_state`` = ``Engine.GameState()``.  The class is a thin ctypes veneer: all
game logic lives behind the ABI (CUDA kernels in the product library).

PyTorch is only plumbing here (device buffers); numpy arrays are host buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import ACTION_DTYPE, BoundLibrary, Config, GymAutoresetIO, GymOutputs, GymStepIO, StatePlanes, StepOutputs


def _ptr(buf) -> Optional[int]:
    """Address of a numpy array, torch tensor, raw int, or None."""
    if buf is None:
        return None
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        if not buf.flags["C_CONTIGUOUS"]:
            raise ValueError("buffer must be C-contiguous")
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):
        if not buf.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return buf.data_ptr()
    raise TypeError(f"unsupported buffer type {type(buf)!r}")


def make_config(lib: BoundLibrary, **overrides) -> Config:
    cfg = Config()
    lib.check(lib.default_config(C.byref(cfg)), "default_config")
    for k, v in overrides.items():
        if k.startswith("reward_"):
            setattr(cfg.reward, k[len("reward_"):], v)
        else:
            if not hasattr(cfg, k):
                raise AttributeError(f"grl_config has no field {k!r}")
            setattr(cfg, k, v)
    return cfg


def make_actions(num_envs: int, max_actions: int) -> np.ndarray:
    return np.zeros((num_envs, max_actions), dtype=ACTION_DTYPE)


def set_action(actions: np.ndarray, env: int, slot: int, player: int, fx: int, fy: int, tx: int, ty: int,
               move_all: bool = True) -> None:
    """core.MoveAction{PlayerID, FromX, FromY, ToX, ToY, MoveAll} (core/action.go:23-35)."""
    a = actions[env, slot]
    a["player_id"], a["from_x"], a["from_y"], a["to_x"], a["to_y"] = player, fx, fy, tx, ty
    a["move_all"] = 1 if move_all else 0
    a["present"] = 1


class BatchedEngine:
    def __init__(self, lib: BoundLibrary, cfg: Config):
        self.lib = lib
        self.cfg = cfg
        self._h = C.c_void_p()
        lib.check(lib.create(C.byref(cfg), C.byref(self._h)), "create")
        self.B = cfg.num_envs
        self.W, self.H, self.P = cfg.width, cfg.height, cfg.num_players
        self.N = self.W * self.H
        self.A = cfg.max_actions
        self.mask_words = (4 * self.N + 31) // 32

    # -- lifecycle ---------------------------------------------------------
    def close(self) -> None:
        if self._h:
            self.lib.destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self) -> None:
        self.lib.check(self.lib.sync(self._h), "sync")

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Run on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        self.lib.check(self.lib.set_stream(self._h, cuda_stream), "set_stream")

    def use_torch_stream(self) -> None:
        """Issue this engine's work on torch's CURRENT CUDA stream, so that kernels reading or writing
        torch tensors are ordered with the torch ops around them (the engine's own stream is
        non-blocking: without this, hand-offs through device tensors need explicit synchronisation)."""
        import torch

        if self.lib.prefix == "grl_" and torch.cuda.is_available():
            with torch.cuda.device(self.cfg.device):
                self.set_stream(torch.cuda.current_stream().cuda_stream)

    # -- reset ---------------------------------------------------------------
    def reset_seeded(self, seeds: Sequence[int], env_ids: Optional[Sequence[int]] = None) -> None:
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        ids = None if env_ids is None else np.ascontiguousarray(env_ids, dtype=np.int32)
        n = len(seeds)
        self.lib.check(self.lib.reset_seeded(self._h, _ptr(ids), n, _ptr(seeds)), "reset_seeded")

    def reset_boards(self, owner, army, type_, env_ids: Optional[Sequence[int]] = None) -> None:
        owner = np.ascontiguousarray(owner, dtype=np.int32).reshape(-1, self.N)
        army = np.ascontiguousarray(army, dtype=np.int32).reshape(-1, self.N)
        type_ = np.ascontiguousarray(type_, dtype=np.int32).reshape(-1, self.N)
        ids = None if env_ids is None else np.ascontiguousarray(env_ids, dtype=np.int32)
        n = owner.shape[0]
        self.lib.check(
            self.lib.reset_boards(self._h, _ptr(ids), n, _ptr(owner), _ptr(army), _ptr(type_)), "reset_boards"
        )

    def mapgen(self, seed: int):
        owner = np.zeros(self.N, np.int32)
        army = np.zeros(self.N, np.int32)
        type_ = np.zeros(self.N, np.int32)
        self.lib.check(
            self.lib.mapgen(C.byref(self.cfg), int(seed), _ptr(owner), _ptr(army), _ptr(type_)), "mapgen"
        )
        return owner, army, type_

    # -- stepping ------------------------------------------------------------
    def step(self, actions=None, flags: int = 0, policy_seed: int = 0) -> None:
        """Engine.Step for every env (internal/game/engine.go:75)."""
        self.lib.check(self.lib.step(self._h, _ptr(actions), flags, policy_seed), "step")

    def outputs(self, obs=None, mask_bits=None, reward=None, done=None, winner=None, step_error=None,
                action_index=None, obs_packed=None) -> StepOutputs:
        o = StepOutputs()
        o.obs, o.mask_bits, o.reward = _ptr(obs), _ptr(mask_bits), _ptr(reward)
        o.done, o.winner, o.step_error = _ptr(done), _ptr(winner), _ptr(step_error)
        o.action_index, o.obs_packed = _ptr(action_index), _ptr(obs_packed)
        return o

    @property
    def packed_words(self) -> int:
        """32-bit words of one packed observation record (grl_obs_packed_words)."""
        return int(self.lib.obs_packed_words(self.W, self.H, self.P))

    def expand_obs(self, packed: np.ndarray, out: Optional[np.ndarray] = None, threads: int = 0) -> np.ndarray:
        """Host-side expansion of packed observation records [n, packed_words] into Serializer.StateToTensor's
        float32 tensors [n, P, 9, H, W] (grl_expand_obs: no device work, bit-identical to the ``obs`` plane)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1, self.packed_words)
        n = packed.shape[0]
        if out is None:
            out = np.empty((n, self.P, _abi.GRL_OBS_CHANNELS, self.H, self.W), np.float32)
        self.lib.check(self.lib.expand_obs(self.W, self.H, self.P, _ptr(packed), n, _ptr(out), threads), "expand_obs")
        return out

    def alloc_outputs_host(self) -> Dict[str, np.ndarray]:
        B, P, N = self.B, self.P, self.N
        return dict(
            obs=np.zeros((B, P, _abi.GRL_OBS_CHANNELS, self.H, self.W), np.float32),
            mask_bits=np.zeros((B, P, self.mask_words), np.uint32),
            reward=np.zeros((B, P), np.float32),
            done=np.zeros(B, np.uint8),
            winner=np.zeros(B, np.int8),
            step_error=np.zeros(B, np.uint8),
            action_index=np.zeros((B, P), np.int32),
        )

    def step_fused(self, actions, out: StepOutputs, flags: int = 0, policy_seed: int = 0) -> None:
        self.lib.check(
            self.lib.step_fused(self._h, _ptr(actions), flags, policy_seed, C.byref(out)), "step_fused"
        )

    def observe(self, out: StepOutputs) -> None:
        self.lib.check(self.lib.observe(self._h, C.byref(out)), "observe")

    def gym_observe(self, max_turns: int, obs=None, mask=None, stats=None) -> None:
        """generals_gym read-outs for every (env, player): obs [B,P,9,H,W] f32, mask [B,P,N*5] u8,
        stats [B,P,4] i32 (python/generals_gym/generals_env.py:291-387)."""
        o = GymOutputs()
        o.obs, o.mask, o.stats = _ptr(obs), _ptr(mask), _ptr(stats)
        self.lib.check(self.lib.gym_observe(self._h, int(max_turns), C.byref(o)), "gym_observe")

    def gym_observe_envs(self, max_turns: int, env_ids, obs=None, mask=None, stats=None) -> None:
        """The gym read-outs of the listed envs only (host int32 ids); other envs' rows are left untouched."""
        ids = np.ascontiguousarray(env_ids, dtype=np.int32)
        o = GymOutputs()
        o.obs, o.mask, o.stats = _ptr(obs), _ptr(mask), _ptr(stats)
        self.lib.check(self.lib.gym_observe_envs(self._h, int(max_turns), ids.ctypes.data, len(ids), C.byref(o)),
                       "gym_observe_envs")

    def gym_autoreset(self, max_turns: int, base_seed: int, **planes) -> None:
        """Re-seed every env whose episode ended in the last gym_step, on the device (grl_gym_autoreset).  Keyword
        planes: terminated, truncated, episode, turns, calls, obs, mask, stats, final_obs, n_reset."""
        io = GymAutoresetIO()
        for k, v in planes.items():
            if k in ("obs", "mask", "stats"):
                setattr(io.out, k, _ptr(v))
            else:
                setattr(io, k, _ptr(v))
        self.lib.check(self.lib.gym_autoreset(self._h, int(max_turns), int(base_seed), C.byref(io)), "gym_autoreset")

    def gym_sample(self, seed: int, mask, player: int, action) -> None:
        """A uniformly random valid Discrete(N*5) action per env from the gym mask plane (int64 [B])."""
        self.lib.check(self.lib.gym_sample(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(mask), int(player), _ptr(action)),
                       "gym_sample")

    def replay_push_rows(self, obs, views: int, view: int, obs_floats: int, capacity: int, next_states=None, next_row0: int = 0,
                         states=None, state_row0: int = 0, done=None, final_obs=None) -> None:
        """One vector step's observation rows into a replay ring (grl_replay_push_rows): ``next_states`` rows from
        ``obs`` (``final_obs`` where ``done``), ``states`` rows of the next transition, in one pass."""
        io = _abi.ReplayRowsIO()
        io.obs, io.final_obs, io.done = _ptr(obs), _ptr(final_obs), _ptr(done)
        io.next_states, io.states = _ptr(next_states), _ptr(states)
        io.capacity, io.next_row0, io.state_row0 = int(capacity), int(next_row0), int(state_row0)
        io.views, io.view, io.obs_floats = int(views), int(view), int(obs_floats)
        self.lib.check(self.lib.replay_push_rows(self._h, C.byref(io)), "replay_push_rows")

    def gym_encode(self, action_idx, player: int, slot: int, mask, skip_invalid: bool, actions, valid) -> None:
        """Discrete(N*5) indices -> grl_action slots, rejecting indices the gym mask forbids."""
        self.lib.check(self.lib.gym_encode(self._h, _ptr(action_idx), int(player), int(slot), _ptr(mask),
                                           1 if skip_invalid else 0, _ptr(actions), _ptr(valid)), "gym_encode")

    def gym_step(self, max_turns: int, opponent_seed: int, **planes) -> None:
        """One GeneralsEnv.step() for every env (grl_gym_step).  Keyword planes: action, opponent_action, obs, mask,
        stats, actions, prev_stats, turns, calls, reward, terminated, truncated, valid, done, winner, step_error,
        n_finished; with ``action`` absent (or None) player 0 is the random agent: ``agent_seed`` (int) keys its draw and
        ``sampled_action`` (int64 [B]) receives the index it played."""
        io = GymStepIO()
        io.agent_seed = int(planes.pop("agent_seed", 0)) & 0xFFFFFFFFFFFFFFFF
        for k, v in planes.items():
            if k in ("obs", "mask", "stats"):
                setattr(io.out, k, _ptr(v))
            else:
                setattr(io, k, _ptr(v))
        self.lib.check(self.lib.gym_step(self._h, int(max_turns), int(opponent_seed), C.byref(io)), "gym_step")

    def sample_actions(self, policy_seed: int, actions=None):
        if actions is None:
            actions = make_actions(self.B, self.A)
        self.lib.check(self.lib.sample_actions(self._h, policy_seed, _ptr(actions)), "sample_actions")
        return actions

    # -- read-outs -------------------------------------------------------------
    def mask(self, variant: int = _abi.MASK_ENGINE_URDL, out=None):
        if out is None:
            if variant in (_abi.MASK_ENGINE_URDL, _abi.MASK_SERIALIZER_UDLR):
                out = np.zeros((self.B, self.P, self.N * 4), np.uint8)
            elif variant == _abi.MASK_ENGINE_URDL_BITS:
                out = np.zeros((self.B, self.P, self.mask_words), np.uint32)
            else:
                out = np.zeros((self.B, self.P, 2, self.mask_words), np.uint32)
        self.lib.check(self.lib.mask(self._h, variant, _ptr(out)), "mask")
        return out

    def get_legal_action_mask(self, env: int, player_id: int) -> np.ndarray:
        """Engine.GetLegalActionMask (engine.go:271-280): []bool of W*H*4, dirs U,R,D,L."""
        if player_id < 0 or player_id >= self.P:
            return np.zeros(self.N * 4, bool)
        return self.mask(_abi.MASK_ENGINE_URDL)[env, player_id].astype(bool)

    def visibility(self):
        vis = np.zeros((self.B, self.P, self.N), np.uint8)
        fog = np.zeros((self.B, self.P, self.N), np.uint8)
        self.lib.check(self.lib.visibility(self._h, _ptr(vis), _ptr(fog)), "visibility")
        return vis, fog

    def compute_player_visibility(self, env: int, player_id: int):
        """Engine.ComputePlayerVisibility (visibility.go:153): (VisibleTiles, FogTiles)."""
        vis, fog = self.visibility()
        return vis[env, player_id].astype(bool), fog[env, player_id].astype(bool)

    def get_state(self, first: int = 0, count: Optional[int] = None) -> Dict[str, np.ndarray]:
        """Engine.GameState() for envs [first, first+count) as planar arrays."""
        count = self.B - first if count is None else count
        planes = StatePlanes()
        out: Dict[str, np.ndarray] = {}
        for name, dt, shape in _abi.STATE_FIELDS:
            dims = {"N": (count, self.N), "PN": (count, self.P, self.N), "P": (count, self.P), "": (count,)}[shape]
            arr = np.zeros(dims, dt)
            out[name] = arr
            setattr(planes, name, _ptr(arr))
        self.lib.check(self.lib.get_state(self._h, first, count, C.byref(planes)), "get_state")
        return out

    def set_state(self, state: Dict[str, np.ndarray], first: int = 0) -> None:
        planes = StatePlanes()
        keep = []
        count = None
        for name, dt, _ in _abi.STATE_FIELDS:
            if name in state and state[name] is not None:
                arr = np.ascontiguousarray(state[name], dtype=dt)
                keep.append(arr)
                count = arr.shape[0] if count is None else count
                setattr(planes, name, _ptr(arr))
        self.lib.check(self.lib.set_state(self._h, first, count, C.byref(planes)), "set_state")

    def state_hash(self) -> np.ndarray:
        out = np.zeros(self.B, np.uint64)
        self.lib.check(self.lib.state_hash(self._h, _ptr(out)), "state_hash")
        return out

    def buffer_hash(self, buf, row_words: int, rows: int) -> np.ndarray:
        out = np.zeros(rows, np.uint64)
        self.lib.check(self.lib.buffer_hash(self._h, _ptr(buf), row_words, rows, _ptr(out)), "buffer_hash")
        return out

    def stats(self) -> np.ndarray:
        out = np.zeros(4, np.uint64)
        self.lib.check(self.lib.stats(self._h, _ptr(out)), "stats")
        return out

    def launch_count(self) -> int:
        out = np.zeros(1, np.uint64)
        self.lib.check(self.lib.launch_count(self._h, _ptr(out)), "launch_count")
        return int(out[0])

    def is_game_over(self) -> np.ndarray:
        return self.get_state()["game_over"].astype(bool)

    def get_winner(self) -> np.ndarray:
        return self.get_state()["winner"]
