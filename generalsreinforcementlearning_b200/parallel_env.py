"""``ParallelEnvPool`` / ``ReplayBuffer`` of the reference's ``generals_gym`` package, on the device.

The reference collects experience with one ``GeneralsEnv`` per worker THREAD, each stepping its game over gRPC
(~80 ms per step) and pushing ``(state, action, reward, next_state, done)`` into a thread-safe list
(``python/generals_gym/vector_env.py:27-192``, ``replay_buffer.py:13-55``; 18 / 70 / 138 / 251 env-steps/s with
1 / 4 / 8 / 16 workers).  Here the N environments are the N slots of ONE ``GeneralsVecEnv``: a step of
all of them is one kernel launch (``grl_gym_step``) plus the device-side auto-reset, the transitions of a step are
written into the replay ring in HBM by a handful of tensor copies, and nothing crosses PCIe unless the caller asks for
episode results or samples to the host.

Same constructor arguments, methods and properties as the reference classes, so ``train_dqn_parallel.py`` switches by
changing its imports and its env factory:

    ParallelEnvPool(num_envs, env_factory, action_fn, replay_buffer, max_steps_per_episode, max_env_retries, seed)
        .start() .stop(join_timeout) .total_env_steps .total_episodes .alive_workers .pop_episode_results()
    ReplayBuffer(capacity) .push(state, action, reward, next_state, done) .sample(batch_size) .total_pushed len()

What differs, because the envs are one batch: ``env_factory(worker_id)`` is called once (worker 0) and returns a
``GeneralsVecEnv`` with ``num_envs`` slots (or ``vec_env=`` is passed); ``action_fn`` keeps the reference's per-env
signature ``(state, valid_mask, worker_id, rng) -> int`` (evaluated on the host, env by env — for small pools and for
tests), and ``batch_action_fn(states, valid_masks) -> actions`` — device tensors in, an int64 tensor out — is the path
that scales.  ``ReplayBuffer.sample`` returns the reference's list of tuples; ``sample_tensors`` the same batch as five
device tensors.
"""
from __future__ import annotations

import logging
import random
import threading
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)

# (state, valid_mask, worker_id, rng) -> action index          vector_env.py:23-24
ActionFn = Callable[[np.ndarray, np.ndarray, int, random.Random], int]


class ReplayBuffer:
    """Ring-buffer replay memory (replay_buffer.py:13-55) whose rows live in tensors on ``device`` (HBM when the pool's
    env is on the GPU).  Storage is allocated at the first push, when the observation shape is known."""

    def __init__(self, capacity: int, device=None):
        if capacity <= 0:
            raise ValueError(f"capacity must be positive, got {capacity}")
        import torch

        self.torch = torch
        self.capacity = int(capacity)
        self.device = torch.device(device) if device is not None else None
        self._states = self._next_states = self._actions = self._rewards = self._dones = None
        self._size = 0
        self._write_idx = 0
        self._total_pushed = 0
        self._lock = threading.Lock()
        self._gen = None
        self._sampler_seed = 0x5EED
        self._pending = None   # (first row, count) of a vector step whose second half has not arrived

    # ------------------------------------------------------------------ storage
    def _allocate(self, obs_shape, device):
        t = self.torch
        self.device = self.device or device
        self._states = t.zeros((self.capacity, *obs_shape), dtype=t.float32, device=self.device)
        self._next_states = t.zeros_like(self._states)
        self._actions = t.zeros(self.capacity, dtype=t.int64, device=self.device)
        self._rewards = t.zeros(self.capacity, dtype=t.float32, device=self.device)
        self._dones = t.zeros(self.capacity, dtype=t.bool, device=self.device)
        self._gen = t.Generator(device=self.device)
        self._gen.manual_seed(self._sampler_seed)

    def seed(self, seed: int) -> None:
        """Seed of the sampler's own generator (the reference draws from the global ``random``)."""
        self._sampler_seed = int(seed)
        if self._gen is not None:
            self._gen.manual_seed(int(seed))

    # ------------------------------------------------------------------ reference API
    def push(self, state, action, reward, next_state, done) -> None:
        """Add one experience, evicting the oldest when full."""
        t = self.torch
        s = t.as_tensor(state, dtype=t.float32)
        self.push_batch(s.unsqueeze(0), t.as_tensor([int(action)]), t.as_tensor([float(reward)]),
                        t.as_tensor(next_state, dtype=t.float32).unsqueeze(0), t.as_tensor([bool(done)]))

    def push_batch(self, states, actions, rewards, next_states, dones, next_states_if_done=None) -> None:
        """n experiences at once (one per env of a vector step): five ring writes, no host round trip.
        ``next_states_if_done`` (optional, same shape as ``next_states``): rows taken instead of ``next_states`` where
        ``dones`` is set — the final observations of envs that were re-seeded inside the step."""
        t = self.torch
        n = int(states.shape[0])
        if n == 0:
            return
        with self._lock:
            if self._states is None:
                self._allocate(tuple(states.shape[1:]), states.device)
            dev = self.device
            if n > self.capacity:   # only the newest `capacity` rows can stay
                keep = slice(n - self.capacity, n)
                states, actions, rewards, next_states, dones = (x[keep] for x in (states, actions, rewards, next_states, dones))
                if next_states_if_done is not None:
                    next_states_if_done = next_states_if_done[keep]
                self._total_pushed += n - self.capacity
                n = self.capacity
            dones = dones.to(device=dev, dtype=t.bool)
            if next_states_if_done is not None:
                next_states = t.where(dones.view(-1, *([1] * (states.dim() - 1))), next_states_if_done.to(dev), next_states.to(dev))
            first = min(n, self.capacity - self._write_idx)   # rows up to the end of the ring, then the wrapped rest
            for lo, hi, at in ((0, first, self._write_idx), (first, n, 0)):
                if hi <= lo:
                    continue
                dst = slice(at, at + hi - lo)
                self._states[dst].copy_(states[lo:hi])
                self._next_states[dst].copy_(next_states[lo:hi])
                self._actions[dst].copy_(actions[lo:hi])
                self._rewards[dst].copy_(rewards[lo:hi])
                self._dones[dst].copy_(dones[lo:hi])
            self._write_idx = (self._write_idx + n) % self.capacity
            self._size = min(self.capacity, self._size + n)
            self._total_pushed += n

    def begin_step(self, states):
        """First half of a vector step's push: the n states go straight into the ring rows the step's transitions will
        occupy (no staging copy of the observation plane, which the env overwrites while stepping).  Returns the ticket
        ``finish_step`` takes.  Until then the rows are not sampled."""
        n = int(states.shape[0])
        if n > self.capacity:
            raise ValueError(f"a vector step of {n} envs needs a replay capacity of at least {n}")
        with self._lock:
            if self._states is None:
                self._allocate(tuple(states.shape[1:]), states.device)
            first = min(n, self.capacity - self._write_idx)
            pieces = [(lo, hi, at) for lo, hi, at in ((0, first, self._write_idx), (first, n, 0)) if hi > lo]
            for lo, hi, at in pieces:
                self._states[at:at + hi - lo].copy_(states[lo:hi])
            # rows being written must not be sampled: shrink the readable prefix when the write runs into it
            if self._size == self.capacity or self._write_idx + n > self._size:
                self._pending = (self._write_idx, n)
            return pieces, n

    def finish_step(self, ticket, actions, rewards, next_states, dones, next_states_if_done=None) -> None:
        """Second half: actions, rewards, done flags and next states (``next_states_if_done`` rows where the episode
        ended) of the rows ``begin_step`` reserved; the rows become visible to ``sample``."""
        t = self.torch
        pieces, n = ticket
        with self._lock:
            dones = dones.to(device=self.device, dtype=t.bool)
            shape1 = [1] * (self._states.dim() - 1)
            for lo, hi, at in pieces:
                dst = slice(at, at + hi - lo)
                if next_states_if_done is not None:   # one pass: the final observation where the episode ended
                    t.where(dones[lo:hi].view(-1, *shape1), next_states_if_done[lo:hi], next_states[lo:hi], out=self._next_states[dst])
                else:
                    self._next_states[dst].copy_(next_states[lo:hi])
                self._actions[dst].copy_(actions[lo:hi])
                self._rewards[dst].copy_(rewards[lo:hi])
                self._dones[dst].copy_(dones[lo:hi])
            self._write_idx = (self._write_idx + n) % self.capacity
            self._size = min(self.capacity, self._size + n)
            self._total_pushed += n
            self._pending = None

    # ---- the pool's path over grl_replay_push_rows: the observation rows are written by ONE pass of a kernel (next_states of
    # the step that ended and states of the step to come), the small planes by commit_vector_step
    def ensure_storage(self, obs_shape, device) -> None:
        with self._lock:
            if self._states is None:
                self._allocate(tuple(obs_shape), device)

    @property
    def write_row(self) -> int:
        return self._write_idx

    def commit_vector_step(self, n: int, actions, rewards, dones) -> None:
        """Actions, rewards and done flags of the n rows at ``write_row`` (their states / next_states are in place); the rows
        become visible to ``sample``, the n rows after them — whose states are already written — stay hidden."""
        t = self.torch
        with self._lock:
            first = min(n, self.capacity - self._write_idx)
            for lo, hi, at in ((0, first, self._write_idx), (first, n, 0)):
                if hi <= lo:
                    continue
                dst = slice(at, at + hi - lo)
                self._actions[dst].copy_(actions[lo:hi])
                self._rewards[dst].copy_(rewards[lo:hi])
                self._dones[dst].copy_(dones[lo:hi])
            self._write_idx = (self._write_idx + n) % self.capacity
            self._size = min(self.capacity, self._size + n)
            self._total_pushed += n
            self._pending = (self._write_idx, n)

    def sample_tensors(self, batch_size: int):
        """A uniformly random batch (without replacement) as (states, actions, rewards, next_states, dones) tensors on
        the buffer's device."""
        t = self.torch
        with self._lock:
            if batch_size > self._size:
                raise ValueError("Sample larger than population or is negative")   # what random.sample raises
            size, shift = self._size, 0
            if self._pending is not None:
                # the rows of a step in flight hold a new state but the old action/reward: sample the others
                end = self._pending[0] + self._pending[1]
                if self._size == self.capacity:
                    shift, size = end % self.capacity, self.capacity - self._pending[1]
                elif end > self.capacity:
                    # a ring that is not full yet whose step in flight already wraps (capacity not a multiple of the
                    # vector step): its first rows, complete transitions of the oldest step, have new states too
                    shift = end - self.capacity
                    size = self._size - shift
                if batch_size > size:
                    raise ValueError("Sample larger than population or is negative")
            if size <= (1 << 16) or 4 * batch_size > size:
                idx = t.randperm(size, generator=self._gen, device=self.device)[:batch_size]
            else:   # a large ring: independent draws (a permutation of millions of rows per batch would cost more than
                    # the batch; duplicates are a 1-in-size/batch event)
                idx = t.randint(size, (batch_size,), generator=self._gen, device=self.device)
            if shift:
                idx = (idx + shift) % self.capacity
            return (self._states[idx], self._actions[idx], self._rewards[idx], self._next_states[idx], self._dones[idx])

    def sample(self, batch_size: int) -> List[Tuple]:
        """The reference's form: a list of (state, action, reward, next_state, done) tuples with numpy states, an int
        action, a float reward and a bool done."""
        s, a, r, ns, d = (x.cpu().numpy() for x in self.sample_tensors(batch_size))
        return [(s[i], int(a[i]), float(r[i]), ns[i], bool(d[i])) for i in range(len(a))]

    @property
    def total_pushed(self) -> int:
        """Monotonic count of pushes; doubles as a global env-step counter."""
        with self._lock:
            return self._total_pushed

    def __len__(self) -> int:
        with self._lock:
            return self._size


# batch_action_fn=RANDOM_AGENT: the reference's random agent (python/generals_agent/random_agent.py) drawn inside the vector
# step's own launch — the indices vec.sample_actions() would return, without its launch
RANDOM_AGENT = "random_agent"


class ParallelEnvPool:
    """N environments stepped together (vector_env.py:27-192): one collector thread drives a ``GeneralsVecEnv``,
    pushes every env's transition of every step into the shared replay buffer and keeps the per-episode results."""

    _RESULTS = 1 << 20   # finished episodes kept on the device between two pop_episode_results() calls

    def __init__(self, num_envs: int, env_factory: Optional[Callable[[int], Any]] = None, action_fn: Optional[ActionFn] = None,
                 replay_buffer: Optional[ReplayBuffer] = None, max_steps_per_episode: int = 200, max_env_retries: int = 3,
                 seed: int = 42, batch_action_fn: Optional[Callable[[Any, Any], Any]] = None, vec_env: Any = None):
        if (action_fn is None) == (batch_action_fn is None):
            raise ValueError("give action_fn (per env, the reference's signature) or batch_action_fn (all envs at once)")
        if replay_buffer is None:
            raise ValueError("a replay buffer is required")
        if vec_env is None and env_factory is None:
            raise ValueError("give env_factory or vec_env")
        self.num_envs = int(num_envs)
        self.env_factory = env_factory
        self.action_fn = action_fn
        self.batch_action_fn = batch_action_fn
        self.replay_buffer = replay_buffer
        self.max_steps_per_episode = int(max_steps_per_episode)
        self.max_env_retries = int(max_env_retries)
        self.seed = int(seed)
        self._vec = vec_env
        self._stop_event = threading.Event()
        self._threads: List[threading.Thread] = []
        self._stats_lock = threading.Lock()
        self._alive_workers = 0
        self._obs = self._mask = None
        # vector_env.py:138-141: a private generator per worker, seeded seed * 1000 + worker_id
        self._rngs = [random.Random(self.seed * 1000 + i) for i in range(self.num_envs)] if action_fn is not None else None

    # ------------------------------------------------------------------ set-up
    def _create_env(self):
        """vector_env.py:114-135: create the environment with retries."""
        last = None
        for attempt in range(self.max_env_retries):
            try:
                env = self.env_factory(0)
                if getattr(env, "num_envs", None) != self.num_envs:
                    raise ValueError(f"env_factory must return a vector env of {self.num_envs} envs (got {getattr(env, 'num_envs', None)})")
                return env
            except ValueError:
                raise
            except Exception as exc:  # noqa: BLE001 - the reference retries on anything
                last = exc
                logger.warning("env creation attempt %d/%d failed: %s", attempt + 1, self.max_env_retries, exc)
        raise RuntimeError(f"failed to create environment: {last}")

    def _ensure_started_state(self):
        if self._obs is not None:
            return
        if self._vec is None:
            self._vec = self._create_env()
        vec = self._vec
        t = vec.torch
        self._t = t
        dev = vec.device
        B = self.num_envs
        self._obs, info = vec.reset(seed=self.seed)
        self._mask = info["valid_actions_mask"]
        self._ep_reward = t.zeros(B, dtype=t.float64, device=dev)
        self._ep_len = t.zeros(B, dtype=t.int32, device=dev)
        self._episodes = t.zeros((), dtype=t.int64, device=dev)
        # finished episodes wait on the device: (reward, length, worker) rows and a running count; row _RESULTS is the
        # slot the rows of unfinished envs are scattered to, so that every step writes a fixed-shape index set
        R = self._RESULTS
        self._res_reward = t.zeros(R + 1, dtype=t.float64, device=dev)
        self._res_len = t.zeros(R + 1, dtype=t.int32, device=dev)
        self._res_env = t.zeros(R + 1, dtype=t.int32, device=dev)
        self._res_count = t.zeros((), dtype=t.int64, device=dev)
        self._res_popped = 0
        self._env_ids = t.arange(B, dtype=t.int32, device=dev)
        self._cap_resets = self.max_steps_per_episode < vec.max_turns
        self._dense_final = vec.auto_reset == "device"
        # the observation rows go into the ring through grl_replay_push_rows (one pass of one kernel per step) when the
        # final observations are a dense plane and the ring can hold two vector steps; otherwise through tensor copies
        self._native_rows = self._dense_final and self.replay_buffer.capacity >= 2 * B and \
            (self.replay_buffer.device is None or self.replay_buffer.device == dev)
        if self._native_rows:
            buf = self.replay_buffer
            buf.ensure_storage(tuple(self._obs.shape[1:]), dev)
            self._obs_floats = int(np.prod(self._obs.shape[1:]))
            vec.engine.replay_push_rows(vec._obs, vec.P, 0, self._obs_floats, buf.capacity, states=buf._states,
                                        state_row0=buf.write_row)

    # ------------------------------------------------------------------ one step of every env
    def _actions(self):
        t = self._t
        if self.batch_action_fn == RANDOM_AGENT:   # drawn inside the vector step's launch (GeneralsVecEnv.step(None))
            return None
        if self.batch_action_fn is not None:
            a = self.batch_action_fn(self._obs, self._mask)
            return t.as_tensor(a, device=self._vec.device).to(t.int64)
        states, masks = self._obs.cpu().numpy(), self._mask.cpu().numpy().astype(bool)
        acts = [int(self.action_fn(states[i], masks[i], i, self._rngs[i])) for i in range(self.num_envs)]
        return t.as_tensor(acts, dtype=t.int64, device=self._vec.device)

    def step_once(self) -> None:
        """One step of all envs: actions, ``GeneralsVecEnv.step``, the transitions into the replay buffer, episode
        accounting on the device.  (``start()`` runs this in a loop on the collector thread.)"""
        self._ensure_started_state()
        t, vec = self._t, self._vec
        actions = self._actions()
        if self._native_rows:
            return self._step_native(actions)
        ticket = self.replay_buffer.begin_step(self._obs)   # the env's observation plane is overwritten by the step
        next_obs, reward, terminated, truncated, info = vec.step(actions)
        if actions is None:
            actions = info["action"]
        done = terminated | truncated                       # vector_env.py:172: done = terminated or truncated
        final = None
        if self._dense_final:
            final = info["final_observation"]               # dense plane, rows valid where the episode ended
        elif "final_env_ids" in info:                       # compact form: scatter into a dense copy of the new plane
            final = next_obs.clone()
            final[info["final_env_ids"]] = info["final_observation"]
        self.replay_buffer.finish_step(ticket, actions, reward, next_obs, done, next_states_if_done=final)
        self._ep_reward += reward
        self._ep_len += 1
        ended = done
        if self._cap_resets:   # vector_env.py:164: an episode also ends at max_steps_per_episode (not a `done` transition)
            capped = (self._ep_len >= self.max_steps_per_episode) & ~done
            vec.reset_envs(capped)
            ended = done | capped
        self._account(reward, ended)
        self._obs, self._mask = next_obs, info["valid_actions_mask"]

    def _step_native(self, actions) -> None:
        t, vec, buf = self._t, self._vec, self.replay_buffer
        next_obs, reward, terminated, truncated, info = vec.step(actions)
        if actions is None:
            actions = info["action"]
        done = terminated | truncated
        self._ep_reward += reward
        self._ep_len += 1
        ended = done
        if self._cap_resets:
            capped = (self._ep_len >= self.max_steps_per_episode) & ~done
            vec.reset_envs(capped)   # their last views land in the final-observation plane, like the finished envs'
            ended = done | capped
        w = buf.write_row
        vec.engine.replay_push_rows(vec._obs, vec.P, 0, self._obs_floats, buf.capacity, next_states=buf._next_states, next_row0=w,
                                    states=buf._states, state_row0=(w + self.num_envs) % buf.capacity, done=ended,
                                    final_obs=info["final_observation"])
        buf.commit_vector_step(self.num_envs, actions, reward, done)
        self._account(reward, ended)
        self._obs, self._mask = next_obs, info["valid_actions_mask"]

    def _account(self, reward, ended) -> None:
        """Finished episodes -> result rows (fixed-shape scatter: unfinished envs write the spare row)."""
        t, vec = self._t, self._vec
        R = self._RESULTS
        pos = self._res_count + t.cumsum(ended.to(t.int64), 0) - 1
        slot = t.where(ended, pos % R, t.full_like(pos, R))
        self._res_reward[slot] = self._ep_reward
        self._res_len[slot] = self._ep_len
        self._res_env[slot] = self._env_ids
        n_ended = ended.sum()
        self._res_count += n_ended
        self._episodes += n_ended
        zero = t.zeros((), dtype=t.float64, device=vec.device)
        self._ep_reward = t.where(ended, zero, self._ep_reward)
        self._ep_len = t.where(ended, t.zeros((), dtype=t.int32, device=vec.device), self._ep_len)

    def run(self, steps: int) -> None:
        """``steps`` vector steps on the calling thread (benchmarks, tests, learners that interleave by hand)."""
        for _ in range(steps):
            self.step_once()

    # ------------------------------------------------------------------ reference API
    def start(self) -> None:
        """Start collecting (the reference spawns one daemon thread per environment; one drives all of them here)."""
        if self._threads:
            raise RuntimeError("Pool already started")
        self._stop_event.clear()
        self._ensure_started_state()
        with self._stats_lock:
            self._alive_workers = self.num_envs
        th = threading.Thread(target=self._worker_loop, name="env-collector", daemon=True)
        self._threads.append(th)
        th.start()
        logger.info("Started the collector of %d envs", self.num_envs)

    def _worker_loop(self) -> None:
        try:
            while not self._stop_event.is_set():
                self.step_once()
        except Exception as exc:  # noqa: BLE001
            logger.error("collector dying: %s", exc)
        finally:
            with self._stats_lock:
                self._alive_workers = 0

    def stop(self, join_timeout: float = 10.0) -> None:
        """Signal the collector to stop and join it."""
        self._stop_event.set()
        for th in self._threads:
            th.join(timeout=join_timeout)
            if th.is_alive():
                logger.warning("Worker %s did not stop within %.1fs", th.name, join_timeout)
        self._threads = []
        with self._stats_lock:
            self._alive_workers = 0

    def close(self) -> None:
        self.stop()
        if self._vec is not None:
            self._vec.close()
            self._vec = None

    @property
    def total_env_steps(self) -> int:
        return self.replay_buffer.total_pushed

    @property
    def total_episodes(self) -> int:
        return int(self._episodes.item()) if self._obs is not None else 0

    @property
    def alive_workers(self) -> int:
        with self._stats_lock:
            return self._alive_workers

    def pop_episode_results(self) -> List[Tuple[float, int, int]]:
        """Drain and return the (episode_reward, episode_length, worker_id) of episodes finished since the last call."""
        if self._obs is None:
            return []
        with self._stats_lock:
            count = int(self._res_count.item())
            R = self._RESULTS
            lo = max(self._res_popped, count - R)   # rows older than the ring were overwritten
            if count == lo:
                return []
            idx = self._t.arange(lo, count, device=self._vec.device) % R
            rew, ln, env = (x[idx].cpu().numpy() for x in (self._res_reward, self._res_len, self._res_env))
            self._res_popped = count
        return [(float(rew[i]), int(ln[i]), int(env[i])) for i in range(len(rew))]
