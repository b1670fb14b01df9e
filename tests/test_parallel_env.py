"""ParallelEnvPool / ReplayBuffer (python/generals_gym/vector_env.py, replay_buffer.py) over GeneralsVecEnv.

The CPU tests drive the pool over the oracle's copy of the ABI; the GPU test is the same through libgrlcuda.so, with the
collected transitions compared against the oracle-driven pool."""
import random
import threading
import time

import numpy as np
import pytest

from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
from generalsreinforcementlearning_b200.parallel_env import ParallelEnvPool, ReplayBuffer


def test_replay_buffer_is_the_reference_ring():
    """replay_buffer.py:13-55: capacity check, ring eviction of the oldest, total_pushed, sample without replacement."""
    with pytest.raises(ValueError):
        ReplayBuffer(0)
    buf = ReplayBuffer(4)
    assert len(buf) == 0 and buf.total_pushed == 0
    for k in range(6):
        buf.push(np.full((9, 3, 3), k, np.float32), k, 0.5 * k, np.full((9, 3, 3), k + 1, np.float32), k == 5)
    assert len(buf) == 4 and buf.total_pushed == 6
    batch = buf.sample(4)
    assert sorted(a for _, a, _, _, _ in batch) == [2, 3, 4, 5], "the two oldest were evicted"
    for s, a, r, ns, d in batch:
        assert s.shape == (9, 3, 3) and (s == a).all() and (ns == a + 1).all() and r == 0.5 * a and d == (a == 5)
        assert isinstance(a, int) and isinstance(d, bool) and isinstance(r, float)
    with pytest.raises(ValueError):
        buf.sample(5)


def test_replay_buffer_thread_safety():
    """python/test_parallel_env.py:19-51 (test_replay_buffer_thread_safety): four pushers, one sampler."""
    capacity, n_threads, pushes = 500, 4, 300
    buf = ReplayBuffer(capacity)
    errors = []

    def pusher(tid):
        try:
            for i in range(pushes):
                buf.push(np.zeros((9, 2, 2), np.float32), tid, float(i), np.zeros((9, 2, 2), np.float32), False)
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    def sampler():
        try:
            for _ in range(50):
                if len(buf) >= 16:
                    assert len(buf.sample(16)) == 16
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    ths = [threading.Thread(target=pusher, args=(k,)) for k in range(n_threads)] + [threading.Thread(target=sampler)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors
    assert buf.total_pushed == n_threads * pushes and len(buf) == capacity


def _random_action_fn(state, valid_mask, worker_id, rng):   # python/test_parallel_env.py:75-79
    valid = np.where(valid_mask)[0]
    return int(rng.choice(list(valid))) if len(valid) else 0


def test_parallel_env_pool_collects_like_the_reference(oracle_lib):
    """python/test_parallel_env.py:54-130 (test_parallel_env_pool) with the reference's per-env action function: episodes
    finish, workers stop cleanly, sampled transitions have the reference's shapes and types, results drain once."""
    board, num_envs = 5, 2
    buf = ReplayBuffer(capacity=10000)
    pool = ParallelEnvPool(num_envs=num_envs, action_fn=_random_action_fn, replay_buffer=buf, max_steps_per_episode=50, seed=42,
                           env_factory=lambda wid: GeneralsVecEnv(num_envs, board, board, fog_of_war=False, max_turns=100,
                                                                  lib=oracle_lib, host_threads=1, auto_reset="device"))
    start = time.time()
    pool.start()
    with pytest.raises(RuntimeError):
        pool.start()
    try:
        while pool.total_episodes < 4:
            assert time.time() - start < 90 and pool.alive_workers == num_envs
            time.sleep(0.01)
    finally:
        pool.stop(join_timeout=10.0)
    assert pool.alive_workers == 0
    assert pool.total_env_steps == buf.total_pushed >= 4 and len(buf) > 0
    for state, action, reward, next_state, done in buf.sample(min(32, len(buf))):
        assert state.shape == (9, board, board) and next_state.shape == (9, board, board) and state.dtype == np.float32
        assert isinstance(action, int) and isinstance(done, (bool, np.bool_)) and np.isfinite(reward)
    results = pool.pop_episode_results()
    assert len(results) >= 4 and {w for _, _, w in results} <= {0, 1}
    assert all(1 <= length <= 50 for _, length, _ in results), "episodes end at max_steps_per_episode at the latest"
    assert pool.pop_episode_results() == []
    pool.close()


def _drive(lib, B, steps, auto_reset="device", cap=200, max_turns=12, agent="first_valid"):
    """A pool stepped synchronously with a deterministic batch policy; returns everything it produced."""
    import torch

    vec = GeneralsVecEnv(B, 6, 6, max_turns=max_turns, lib=lib, host_threads=1, auto_reset=auto_reset, seed=77)
    buf = ReplayBuffer(capacity=B * (steps + 1))   # one spare vector step: the native path writes the next states ahead

    def policy(states, masks):   # the first valid action of every env (0 when none): needs no generator
        return masks.to(torch.int8).argmax(dim=1)

    if agent == "sampler":       # the random agent as a launch of its own
        policy = lambda states, masks: vec.sample_actions()  # noqa: E731
    elif agent == "in_step":     # the random agent drawn inside the vector step
        from generalsreinforcementlearning_b200.parallel_env import RANDOM_AGENT as policy

    pool = ParallelEnvPool(B, vec_env=vec, batch_action_fn=policy, replay_buffer=buf, max_steps_per_episode=cap, seed=77)
    pool.run(steps)
    rows = [x[: B * steps].cpu().numpy().copy() for x in (buf._states, buf._actions, buf._rewards, buf._next_states, buf._dones)]
    out = dict(rows=rows, n=len(buf), total=buf.total_pushed, episodes=pool.total_episodes, results=pool.pop_episode_results())
    pool.close()
    return out


def test_pool_transitions_are_consistent(oracle_lib):
    """Every row of a vector step is one env's transition: next_state of step t is state of step t+1 unless the episode
    ended there (then it is the final observation, and the next row starts a new episode at turn 0); episode results
    add up; the compact-info auto-reset mode pushes the same rows as the dense one."""
    B, steps = 7, 40
    a = _drive(oracle_lib, B, steps)
    assert a["n"] == a["total"] == B * steps
    s, act, r, ns, d = a["rows"]
    s, ns, d = s.reshape(steps, B, *s.shape[1:]), ns.reshape(steps, B, *ns.shape[1:]), d.reshape(steps, B)
    assert d.any(), "episodes end within 12 turns"
    for t in range(steps - 1):
        cont = ~d[t]
        assert np.array_equal(ns[t][cont], s[t + 1][cont]), f"step {t}: next_state is the next step's state"
        # plane 7 is the turn fraction: a new episode starts at 0, the final observation of the old one does not
        assert (s[t + 1][d[t], 7] == 0).all() and (ns[t][d[t], 7] > 0).all()
    lengths = sorted(l for _, l, _ in a["results"])
    assert a["episodes"] == len(a["results"]) == int(d.sum()) and lengths[-1] <= 12
    total_reward = sum(x for x, _, _ in a["results"])
    r = r.reshape(steps, B)
    # the rewards of finished episodes are sums over exactly their rows
    done_env_steps = 0.0
    for b in range(B):
        ends = np.flatnonzero(d[:, b])
        if len(ends):
            done_env_steps += r[: ends[-1] + 1, b].astype(np.float64).sum()
    assert abs(total_reward - done_env_steps) < 1e-4
    # auto_reset="device" takes the rows through grl_replay_push_rows, "host" (compact final observations) through tensor
    # copies: the same ring either way
    b2 = _drive(oracle_lib, B, steps, auto_reset="host")
    for x, y in zip(a["rows"], b2["rows"]):
        assert np.array_equal(x, y)
    assert a["results"] == b2["results"]


def test_sampling_skips_the_rows_of_a_step_in_flight_that_wraps_before_the_ring_is_full():
    """A ring whose capacity is not a multiple of the vector step: the third step of four envs into ten rows writes its
    states into rows 8, 9, 0, 1 while the ring holds only eight transitions — rows 0 and 1 then pair a new state with
    the first step's action until the step is finished, and must not be sampled meanwhile."""
    import torch

    buf = ReplayBuffer(capacity=10)
    mk = lambda step: (torch.full((4, 3), float(step)), torch.full((4,), step, dtype=torch.int64),
                       torch.full((4,), float(step)), torch.full((4, 3), float(step)), torch.zeros(4, dtype=torch.bool))
    for step in (1, 2):
        s, a, r, ns, d = mk(step)
        buf.finish_step(buf.begin_step(s), a, r, ns, d)
    assert len(buf) == 8
    s, a, r, ns, d = mk(3)
    ticket = buf.begin_step(s)
    for _ in range(20):
        bs, ba, br, bns, bd = buf.sample_tensors(6)
        assert torch.equal(bs[:, 0].to(torch.int64), ba), "a sampled row pairs its own state with its own action"
    with pytest.raises(ValueError):
        buf.sample_tensors(7)                       # six complete rows are readable while the step is in flight
    buf.finish_step(ticket, a, r, ns, d)
    assert len(buf) == 10
    bs, ba, *_ = buf.sample_tensors(10)
    assert torch.equal(bs[:, 0].to(torch.int64), ba) and sorted(ba.tolist()) == [1, 1, 2, 2, 2, 2, 3, 3, 3, 3]


def test_sampled_rows_are_always_whole_transitions():
    """Whatever the capacity / vector-step ratio and whenever sample() is called — between steps or while one is in
    flight — a sampled row pairs the state, action, reward and next state of ONE transition, and the readable population
    is exactly the finished transitions still in the ring."""
    import torch

    rng = np.random.default_rng(11)
    for _ in range(40):
        n = int(rng.integers(1, 7))
        cap = int(rng.integers(n, 4 * n + 3))
        buf = ReplayBuffer(capacity=cap, device="cpu")
        buf.seed(int(rng.integers(1 << 30)))
        for step in range(1, 14):
            s = torch.full((n, 2), float(step))
            ticket = buf.begin_step(s)
            readable = min(len(buf), cap - n) if len(buf) + n > cap else len(buf)   # rows the step in flight does not touch
            if readable:
                bs, ba, br, bns, _ = buf.sample_tensors(readable)
                assert torch.equal(bs[:, 0].to(torch.int64), ba) and torch.equal(bns[:, 0], br), (cap, n, step)
                assert int(ba.max()) < step
            with pytest.raises(ValueError):
                buf.sample_tensors(readable + 1)
            buf.finish_step(ticket, torch.full((n,), step, dtype=torch.int64), torch.full((n,), float(step)),
                            torch.full((n, 2), float(step)), torch.zeros(n, dtype=torch.bool))
            bs, ba, br, bns, _ = buf.sample_tensors(len(buf))
            assert torch.equal(bs[:, 0].to(torch.int64), ba) and torch.equal(bns[:, 0], br), (cap, n, step)
            assert len(buf) == min(cap, step * n)


def test_sampled_rows_are_whole_transitions_on_the_kernel_path():
    """The same invariant for the pool's path over grl_replay_push_rows (its two row streams emulated with index
    writes): a step's pass writes the next_states of the n rows at write_row and the states of the n rows after them;
    commit_vector_step publishes the former and keeps the latter — new states under old actions — away from sample()."""
    import torch

    rng = np.random.default_rng(12)
    for _ in range(40):
        n = int(rng.integers(1, 7))
        cap = int(rng.integers(2 * n, 5 * n + 3))
        buf = ReplayBuffer(capacity=cap, device="cpu")
        buf.seed(int(rng.integers(1 << 30)))
        buf.ensure_storage((2,), torch.device("cpu"))
        rows = lambda r0: (r0 + torch.arange(n)) % cap   # noqa: E731
        buf._states[rows(buf.write_row)] = 1.0                                   # the states of the first transitions
        for step in range(1, 14):
            w = buf.write_row
            buf._next_states[rows(w)] = float(step)                              # one pass: next_states of this step ...
            buf._states[rows(w + n)] = float(step + 1)                           # ... and states of the next one
            buf.commit_vector_step(n, torch.full((n,), step, dtype=torch.int64), torch.full((n,), float(step)),
                                   torch.zeros(n, dtype=torch.bool))
            readable = min(step * n, cap - n)
            assert len(buf) == min(cap, step * n)
            bs, ba, br, bns, _ = buf.sample_tensors(readable)
            assert torch.equal(bs[:, 0].to(torch.int64), ba) and torch.equal(bns[:, 0], br), (cap, n, step)
            with pytest.raises(ValueError):
                buf.sample_tensors(readable + 1)


def test_pool_random_agent_in_step_equals_sampler(oracle_lib):
    """batch_action_fn=RANDOM_AGENT (the agent drawn inside grl_gym_step) collects the very rows — states, actions,
    rewards, next states, done flags, episode results — that batch_action_fn=vec.sample_actions() does."""
    a = _drive(oracle_lib, 9, 45, agent="sampler")
    b = _drive(oracle_lib, 9, 45, agent="in_step")
    for x, y in zip(a["rows"], b["rows"]):
        assert np.array_equal(x, y)
    assert a["rows"][4].any() and (a["rows"][1] > 0).any()
    assert a["results"] == b["results"] and a["episodes"] == b["episodes"] > 0


def test_replay_push_rows_matches_numpy(oracle_lib):
    """grl_replay_push_rows on the oracle binding against plain indexing: ring wrap, final rows where done, either
    destination absent."""
    from helpers import new_engine

    B, P, F, cap = 6, 2, 9 * 4 * 4, 16
    e = new_engine(oracle_lib, 4, 4, P, B)
    rng = np.random.default_rng(3)
    obs = rng.random((B, P, F), dtype=np.float32)
    final = rng.random((B, F), dtype=np.float32)
    done = np.array([0, 1, 0, 0, 1, 0], np.uint8)
    ns, st = np.zeros((cap, F), np.float32), np.zeros((cap, F), np.float32)
    e.replay_push_rows(obs, P, 1, F, cap, next_states=ns, next_row0=13, states=st, state_row0=3, done=done, final_obs=final)
    for b in range(B):
        assert np.array_equal(ns[(13 + b) % cap], final[b] if done[b] else obs[b, 1])
        assert np.array_equal(st[(3 + b) % cap], obs[b, 1])
    assert not ns[3:13].any() and not st[9:].any() and not st[:3].any()
    ns2 = np.zeros_like(ns)
    e.replay_push_rows(obs, P, 0, F, cap, next_states=ns2, next_row0=0)          # no done flags: every row from obs
    assert np.array_equal(ns2[:B], obs[:, 0])
    with pytest.raises(RuntimeError):
        e.replay_push_rows(obs, P, 2, F, cap, next_states=ns2)                   # view out of range
    with pytest.raises(RuntimeError):
        e.replay_push_rows(obs, P, 0, F, B - 1, next_states=ns2)                 # a ring smaller than one vector step


def test_pool_caps_episodes_at_max_steps(oracle_lib):
    """vector_env.py:164: an episode is cut at max_steps_per_episode; the cut is not a `done` transition
    (:172 done = terminated or truncated) and the env starts a new episode."""
    a = _drive(oracle_lib, 5, 30, cap=4, max_turns=100)
    s, act, r, ns, d = a["rows"]
    assert not d.any()
    assert a["episodes"] == len(a["results"]) == 5 * (30 // 4)
    assert all(length == 4 for _, length, _ in a["results"])
    s = s.reshape(30, 5, *s.shape[1:])
    for t in range(30):
        assert (s[t][:, 7] == np.float32((t % 4) / 100)).all(), "turn fraction restarts every four steps"


@pytest.mark.gpu
def test_cuda_pool_collects_what_the_oracle_pool_collects(cuda_lib, oracle_lib):
    """The same pool over libgrlcuda.so (ring in HBM) and over the oracle: identical rows, results and counters."""
    for cap, max_turns in ((200, 12), (5, 100)):
        g = _drive(cuda_lib, 300, 30, cap=cap, max_turns=max_turns)
        o = _drive(oracle_lib, 300, 30, cap=cap, max_turns=max_turns)
        for x, y in zip(g["rows"], o["rows"]):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
        assert g["results"] == o["results"] and g["episodes"] == o["episodes"] and g["total"] == o["total"] == 9000
    # the random agent drawn inside the CUDA gym step collects what the oracle pool collects with its sampler
    g = _drive(cuda_lib, 300, 30, agent="in_step")
    o = _drive(oracle_lib, 300, 30, agent="sampler")
    for x, y in zip(g["rows"], o["rows"]):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert g["results"] == o["results"] and g["episodes"] == o["episodes"] > 0
    # and the tensor-copy path (compact final observations) on the GPU equals the kernel path
    h = _drive(cuda_lib, 300, 30, auto_reset="host")
    k = _drive(cuda_lib, 300, 30, auto_reset="device")
    for x, y in zip(h["rows"], k["rows"]):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
