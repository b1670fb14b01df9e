"""``generals_gym``-compatible VECTOR environment backed directly by the C ABI.

Contract mirrored from the reference's ``python/generals_gym/generals_env.py`` (one game
per gRPC round trip there; B games per kernel launch here):

  observation  float32 ``(9, H, W)`` in [0, 1] per env        generals_env.py:111-116, 291-342
  action       ``Discrete(W*H*5)``: tile*5 + {up,right,down,left,half}      :118-120, 389-441
  mask         ``info["valid_actions_mask"]`` bool ``(W*H*5,)``             :344-387
  reward       the client-side shaping of ``_calculate_reward``             :499-561
  terminated   game status left IN_PROGRESS; truncated at ``max_turns``     :277-279
  invalid action (mask false): no turn is taken, reward -0.1                :226-229

The agent is player 0; the opponent is the reference's default random opponent (a uniformly
random full move, ``_submit_random_opponent_action`` :443-497) drawn on the device, or a second
action tensor for self-play.  Observations, masks, rewards and flags are torch tensors on the
engine's device — nothing crosses PCIe per step except one 4-byte "did any episode end" flag.  The tensors step()
returns are the env's own (double-buffered) output planes: the observation, mask and turn are valid until the next
step(), reward / terminated / truncated / the other info entries until the one after; clone what must live longer.
Finished or truncated envs are re-seeded
automatically on the device (``info["final_observation"]`` keeps the pre-reset view), as Gymnasium vector
envs do.

This module needs neither ``gymnasium`` (absent from the image) nor gRPC.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import _abi
from .engine import BatchedEngine, make_config

# direction order of the client: up, right, down, left (generals_env.py:369, 413)
_DX = (0, 1, 0, -1)
_DY = (-1, 0, 1, 0)


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (shape/dtype/bounds only)."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Discrete:
    def __init__(self, n):
        self.n = int(n)


class GeneralsVecEnv:
    metadata = {"render_modes": ["ansi"], "render_fps": 4}

    def __init__(self, num_envs: int, board_width: int = 15, board_height: int = 15, max_players: int = 2,
                 fog_of_war: bool = True, max_turns: int = 500, device: int = 0, seed: int = 12345,
                 self_play: bool = False, lib=None, host_threads: int = 0, auto_reset: str = "host"):
        import torch

        if max_players != 2:
            raise ValueError("generals_gym drives 2-player games (generals_env.py:68)")
        if lib is None:
            from . import load_library

            lib = load_library()
        self.torch = torch
        self.num_envs, self.W, self.H, self.P = num_envs, board_width, board_height, max_players
        self.board_size = self.N = board_width * board_height
        self.max_turns, self.self_play = max_turns, self_play
        # Finished envs are always re-seeded on the device (grl_gym_autoreset).  "device": NO host read — step() never
        # synchronises, so a training loop can enqueue steps ahead; info["final_observation"] is a dense [B, 9, H, W]
        # plane whose rows are valid where info["final_env_mask"] is set, and info["turn"] already shows 0 for the
        # re-seeded envs.  "host": the step reads one 4-byte flag and, when episodes ended, hands out the compact
        # info["final_observation"] / ["final_env_ids"] form with the finished envs' last turn counters in info["turn"].
        # "host_reset": the same contract with the re-seeding driven from the host (grl_reset_seeded +
        # grl_gym_observe_envs) — the independent path the tests check the device-side one against.
        if auto_reset not in ("host", "device", "host_reset"):
            raise ValueError("auto_reset is 'host', 'device' or 'host_reset'")
        self.auto_reset = auto_reset
        self.engine = BatchedEngine(lib, make_config(lib, num_envs=num_envs, width=board_width, height=board_height,
                                                     num_players=max_players, device=device, max_actions=max_players,
                                                     fog_of_war=1 if fog_of_war else 0, host_threads=host_threads))
        self.on_device = lib.prefix == "grl_"
        self.engine.use_torch_stream()   # the env hands tensors between torch ops and engine kernels every step
        self.device = torch.device("cuda", device) if self.on_device else torch.device("cpu")
        self.single_observation_space = Box(0.0, 1.0, (9, board_height, board_width), np.float32)
        self.single_action_space = Discrete(self.N * 5)
        self.observation_space = Box(0.0, 1.0, (num_envs, 9, board_height, board_width), np.float32)
        self.action_space = Discrete(self.N * 5)
        B, P, N, dev = num_envs, self.P, self.N, self.device
        self._obs = torch.zeros((B, P, 9, self.H, self.W), dtype=torch.float32, device=dev)
        self._mask = torch.zeros((B, P, N * 5), dtype=torch.bool, device=dev)   # the kernel writes 0/1 bytes
        self._stats = torch.zeros((B, P, 4), dtype=torch.int32, device=dev)
        self._prev_stats = torch.zeros_like(self._stats)
        self._actions = torch.zeros((B, P, 8), dtype=torch.uint8, device=dev)  # grl_action[B][P]
        self._done = torch.zeros(B, dtype=torch.uint8, device=dev)
        self._turns = torch.zeros(B, dtype=torch.int32, device=dev)
        # per-step result planes, double-buffered: what step() returns stays valid until the step after next, so the
        # hot loop hands out the kernel's own output tensors instead of cloning them (flags are written as 0/1 bytes
        # straight into bool tensors)
        self._out = [dict(reward=torch.zeros(B, dtype=torch.float64, device=dev),
                          terminated=torch.zeros(B, dtype=torch.bool, device=dev),
                          truncated=torch.zeros(B, dtype=torch.bool, device=dev),
                          valid=torch.zeros(B, dtype=torch.bool, device=dev),
                          winner=torch.zeros(B, dtype=torch.int8, device=dev),
                          step_error=torch.zeros(B, dtype=torch.uint8, device=dev)) for _ in range(2)]
        self._flip = 0
        self._nfin = torch.zeros(1, dtype=torch.int32, device=dev)
        self._opp_draws = 0
        if auto_reset != "host_reset":
            self._episode_dev = torch.zeros(B, dtype=torch.int64, device=dev)
            self._final_obs = torch.zeros((B, 9, self.H, self.W), dtype=torch.float32, device=dev)
        self._sample_draws = 0
        self._sampled = [torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(2)]
        self._calls = torch.zeros(B, dtype=torch.int32, device=dev)   # step() calls this episode (incl. rejected actions)
        self._episode = np.zeros(B, dtype=np.int64)
        self._base_seed = int(seed)
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(int(seed))

    # ------------------------------------------------------------------ helpers
    def _seeds(self, env_ids: np.ndarray) -> np.ndarray:
        return self._base_seed + env_ids.astype(np.int64) + self._episode[env_ids] * self.num_envs

    def _refresh(self, env_ids=None):
        if env_ids is None:
            self.engine.gym_observe(self.max_turns, self._obs, self._mask, self._stats)
        else:   # only the re-seeded envs: the other rows already hold the step's read-outs
            self.engine.gym_observe_envs(self.max_turns, env_ids, self._obs, self._mask, self._stats)

    # ------------------------------------------------------------------ gym API
    def reset(self, seed: Optional[int] = None, options=None) -> Tuple[Any, Dict[str, Any]]:
        if seed is not None:   # a seeded reset restarts every random stream: maps, the opponent's and the agent's draws
            self._base_seed = int(seed)
            self._gen.manual_seed(int(seed))
            self._opp_draws = self._sample_draws = 0
        self._episode[:] = 0
        if self.auto_reset != "host_reset":
            self._episode_dev.zero_()
        ids = np.arange(self.num_envs)
        self.engine.reset_seeded(self._seeds(ids))
        self._turns.zero_()
        self._calls.zero_()
        self._refresh()
        return self._obs[:, 0], {"valid_actions_mask": self._mask[:, 0], "turn": self._turns.clone()}

    def step(self, action, opponent_action=None):
        """action: int64 [B] indices into Discrete(N*5) for player 0, or None for the reference's random agent
        (python/generals_agent/random_agent.py) drawn inside the step's launch — ``step(None)`` plays exactly
        ``step(sample_actions())`` and reports the indices in ``info["action"]`` (and ``opponent_action`` for
        player 1 under self-play; otherwise the reference's default random opponent, a uniformly
        random legal full move drawn on the device).  Returns (obs, reward, terminated, truncated, info).

        The whole step is ONE library call (``grl_gym_step``): action decoding with the client-side
        rejection of masked-out indices (that env takes no turn, reward -0.1, generals_env.py:226-229),
        the turn, the gym read-outs, the client's reward (:499-561, float64) and the episode flags."""
        t = self.torch
        agent = {}
        if action is None:
            # the random agent inside the step's own launch: the very index sample_actions() would have returned (same
            # draw counter, same seed), without the sampler's launch and its pass over the N*5 mask bytes;
            # info["action"] is what each env played
            self._sample_draws += 1
            agent = dict(agent_seed=self._base_seed * 7919 + self._sample_draws,
                         sampled_action=self._sampled[self._sample_draws & 1])
        else:
            action = t.as_tensor(action, device=self.device)
            if action.dtype != t.int64 or not action.is_contiguous():
                action = action.to(t.int64).contiguous()
        oa = None
        if opponent_action is not None:
            oa = t.as_tensor(opponent_action, device=self.device).to(t.int64).contiguous()
        self._opp_draws += 1
        self._flip ^= 1
        o = self._out[self._flip]
        self.engine.gym_step(self.max_turns, self._base_seed * 1000003 + self._opp_draws, action=action, opponent_action=oa,
                             obs=self._obs, mask=self._mask, stats=self._stats, actions=self._actions,
                             prev_stats=self._prev_stats, turns=self._turns, calls=self._calls, reward=o["reward"],
                             terminated=o["terminated"], truncated=o["truncated"], valid=o["valid"], done=self._done,
                             winner=o["winner"], step_error=o["step_error"], n_finished=self._nfin, **agent)
        terminated, truncated, reward = o["terminated"], o["truncated"], o["reward"]
        # info tensors are the env's own planes: valid until the next step() (turn, mask) or the one after (the rest)
        info: Dict[str, Any] = {"invalid_action": ~o["valid"], "winner": o["winner"], "step_error": o["step_error"]}
        if agent:
            info["action"] = agent["sampled_action"]
        turn = self._turns
        if self.auto_reset == "device":
            info["final_env_mask"] = terminated | truncated
            self._autoreset(terminated, truncated)
            info["final_observation"] = self._final_obs
        elif self.auto_reset == "host":
            # the re-seeding is enqueued BEFORE the one host read of the step (how many episodes ended?), so the device
            # works through it while the host waits; with the count known the id list needs no second synchronisation
            turn = self._turns.clone()       # the finished envs' turn counters restart in the auto-reset
            self._autoreset(terminated, truncated)
            n = int(self._nfin.item())
            if n > 0:
                ids = t.nonzero_static(terminated | truncated, size=n).squeeze(1)
                info["final_env_ids"] = ids
                info["final_observation"] = self._final_obs[ids]
        elif int(self._nfin.item()) > 0:     # "host_reset": the host drives the re-seeding
            finished = terminated | truncated
            ids = finished.nonzero(as_tuple=True)[0]
            info["final_env_ids"] = ids
            turn = self._turns.clone()
            info["final_observation"] = self._obs[ids, 0].clone()
            ids_np = ids.cpu().numpy()
            self._episode[ids_np] += 1
            self.engine.reset_seeded(self._seeds(ids_np), ids_np.astype(np.int32))
            self._turns[ids] = 0
            self._calls[ids] = 0
            self._refresh(ids_np)
        info["turn"] = turn
        info["valid_actions_mask"] = self._mask[:, 0]
        return self._obs[:, 0], reward, terminated, truncated, info

    def _autoreset(self, terminated, truncated):
        self.engine.gym_autoreset(self.max_turns, self._base_seed, terminated=terminated, truncated=truncated,
                                  episode=self._episode_dev, turns=self._turns, calls=self._calls, obs=self._obs,
                                  mask=self._mask, stats=self._stats, final_obs=self._final_obs)

    def reset_envs(self, env_mask) -> None:
        """Start a new episode in the envs flagged in ``env_mask`` (bool [B]) — what a caller that caps episode length
        itself does between steps (``ParallelEnvPool.max_steps_per_episode``).  Same re-seeding as the auto-reset: the
        env's episode counter advances and its observation / mask rows are replaced; no host read unless the host
        drives the re-seeding (``auto_reset="host_reset"``)."""
        t = self.torch
        env_mask = t.as_tensor(env_mask, device=self.device).to(t.bool)
        if self.auto_reset != "host_reset":
            self._autoreset(env_mask, t.zeros_like(env_mask))
            return
        ids = env_mask.nonzero(as_tuple=True)[0]
        if ids.numel() == 0:
            return
        ids_np = ids.cpu().numpy()
        self._episode[ids_np] += 1
        self.engine.reset_seeded(self._seeds(ids_np), ids_np.astype(np.int32))
        self._turns[ids] = 0
        self._calls[ids] = 0
        self._refresh(ids_np)

    def sample_actions(self, generator=None, player: int = 0):
        """A uniformly random VALID action per env (envs without one get action 0, which is rejected): the random
        agent of the reference (python/generals_agent/random_agent.py) for every env in one kernel launch.  With a
        torch ``generator`` the draw goes through ``torch.multinomial`` instead (slower: a [B, N*5] float pass)."""
        t = self.torch
        if generator is not None:
            m = self._mask[:, player].to(t.float32)
            none = m.sum(1) == 0
            m[:, 0] += none.to(t.float32)
            return t.multinomial(m, 1, generator=generator).squeeze(1)
        self._sample_draws += 1
        out = self._sampled[self._sample_draws & 1]
        self.engine.gym_sample(self._base_seed * 7919 + self._sample_draws, self._mask, player, out)
        return out

    def opponent_view(self):
        """Player 1's observation and mask (self-play)."""
        return self._obs[:, 1], self._mask[:, 1]

    def render(self, env: int = 0, player: int = 0) -> str:
        from .render import render_board

        return render_board(self.engine, env, player)

    def close(self):
        self.engine.close()
