"""grl_gym_step — one ``GeneralsEnv.step()`` (python/generals_gym/generals_env.py:210-289) for every env.

The CUDA library runs it as ONE launch (the turn kernel's gym instantiation: action decoding with the
client-side rejection, the random opponent's draw, the turn, the client's observation / N*5 mask /
PlayerState read-outs, the float64 reward and the episode flags); the oracle runs the same contract on the
CPU.  Every output plane is compared bit for bit at every step, including rejected actions, truncation,
self-play opponents and boards on each lane-group instantiation."""
import os
import subprocess
import sys

import numpy as np
import pytest

from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _planes(torch, dev, B, P, N, H, W):
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
    return dict(obs=z((B, P, 9, H, W), torch.float32), mask=z((B, P, N * 5), torch.uint8), stats=z((B, P, 4), torch.int32),
                actions=z((B, P, 8), torch.uint8), prev_stats=z((B, P, 4), torch.int32), turns=z(B, torch.int32),
                calls=z(B, torch.int32), reward=z(B, torch.float64), terminated=z(B, torch.uint8),
                truncated=z(B, torch.uint8), valid=z(B, torch.uint8), done=z(B, torch.uint8), winner=z(B, torch.int8),
                step_error=z(B, torch.uint8), n_finished=z(1, torch.int32))


def _pick(mask_row, rng, want_invalid):
    """k-th set (or, for a rejected action, unset) entry of one env's mask."""
    idx = np.flatnonzero(mask_row == (0 if want_invalid else 1))
    if len(idx) == 0:
        return int(rng.integers(0, len(mask_row)))
    return int(idx[rng.integers(0, len(idx))])


def gym_rollout(lib, W, H, P, B, steps, max_turns, self_play, seed=77, fog=1, init=None):
    """Drive grl_gym_step with actions drawn (host RNG) from the library's own mask plane; returns the trace of
    every output plane after every step."""
    import torch

    on_device = lib.prefix == "grl_"
    dev = torch.device("cuda", 0) if on_device else torch.device("cpu")
    e = BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, max_actions=P, host_threads=1,
                                       fog_of_war=fog))
    if on_device:
        e.use_torch_stream()
    e.reset_seeded(np.arange(B, dtype=np.int64) + 12345)
    if init is not None:
        e.set_state(init)
    N = W * H
    pl = _planes(torch, dev, B, P, N, H, W)
    e.gym_observe(max_turns, pl["obs"], pl["mask"], pl["stats"])
    rng = np.random.default_rng(seed)
    trace = []
    for t in range(steps):
        m = pl["mask"].cpu().numpy()
        act = np.zeros(B, np.int64)
        opp = np.zeros(B, np.int64)
        for b in range(B):
            bad = rng.random() < 0.08
            act[b] = _pick(m[b, 0], rng, bad)
            opp[b] = _pick(m[b, 1], rng, rng.random() < 0.05)
        if t % 7 == 3:
            act[0], opp[1 % B] = -1, N * 5 + 3          # out-of-range indices are rejected too
        a = torch.as_tensor(act, device=dev)
        oa = torch.as_tensor(opp, device=dev) if self_play else None
        e.gym_step(max_turns, 1000 + t, action=a, opponent_action=oa, **pl)
        if on_device:
            torch.cuda.synchronize()
        snap = {k: v.cpu().numpy().copy() for k, v in pl.items() if k not in ("actions", "prev_stats")}
        snap["hash"] = e.state_hash().copy()
        trace.append(snap)
    launches = e.launch_count()
    e.close()
    return trace, launches


def _compare(ta, tb, what):
    assert len(ta) == len(tb)
    for t, (a, b) in enumerate(zip(ta, tb)):
        for k in a:
            x, y = a[k], b[k]
            if x.dtype == np.float32:
                x, y = x.view(np.uint32), y.view(np.uint32)
            elif x.dtype == np.float64:
                x, y = x.view(np.uint64), y.view(np.uint64)
            if not np.array_equal(x, y):
                bad = np.argwhere(x != y)
                raise AssertionError(f"{what}: plane {k!r} differs at step {t}, first at {bad[0].tolist()} "
                                     f"({len(bad)} cells): {a[k][tuple(bad[0])]} vs {b[k][tuple(bad[0])]}")


CASES = [  # W, H, P, B, steps, max_turns, self_play
    (10, 10, 2, 203, 45, 30, False),    # 4 lanes per game, partial last warp, truncation by turns and by calls
    (15, 15, 2, 130, 40, 500, True),    # 8 lanes per game, self-play opponent indices
    (20, 20, 2, 96, 60, 500, False),    # one game per warp
    (20, 20, 4, 40, 50, 35, False),     # four players: players 2,3 keep the synthetic policy's half-move bit
    (15, 15, 4, 38, 30, 40, False),     # four views per game on the run writer (P == PT = 4), partial last warp
    (15, 15, 3, 37, 30, 40, False),     # P < PT: run-time-scheduled writer; odd byte offsets of the mask blocks
    (7, 13, 3, 33, 40, 500, True),      # generic geometry
    (5, 5, 2, 64, 80, 500, False),      # tiny boards finish: terminated, +-100, eliminated-opponent bonus
]


def test_oracle_gym_step_is_deterministic(oracle_lib):
    a, _ = gym_rollout(oracle_lib, 8, 8, 2, 24, 30, 20, False)
    b, _ = gym_rollout(oracle_lib, 8, 8, 2, 24, 30, 20, False)
    _compare(a, b, "oracle twice")
    assert any(s["truncated"].any() for s in a) and any((s["valid"] == 0).any() for s in a)
    rej = [(s["reward"][s["valid"] == 0]) for s in a if (s["valid"] == 0).any()]
    assert all((r == -0.1).all() for r in rej), "a rejected action costs -0.1 and takes no turn"


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,P,B,steps,max_turns,self_play", CASES)
def test_cuda_gym_step_matches_oracle(cuda_lib, oracle_lib, W, H, P, B, steps, max_turns, self_play):
    g, launches = gym_rollout(cuda_lib, W, H, P, B, steps, max_turns, self_play)
    o, _ = gym_rollout(oracle_lib, W, H, P, B, steps, max_turns, self_play)
    _compare(g, o, f"cuda vs oracle {W}x{H}x{P}p")
    # reset (mapgen + turn-0) and the initial read-out aside, a step is ONE launch; state_hash adds one per step
    assert launches <= 3 + 2 * steps + 2, f"{launches} launches for {steps} fused steps"
    if (W, H) == (5, 5):
        assert any(s["terminated"].any() for s in g), "5x5 games end within 80 steps"


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,P,B", [(20, 20, 2, 40), (15, 15, 2, 70), (10, 10, 2, 90), (6, 9, 2, 33)])
def test_cuda_gym_step_matches_oracle_without_fog(cuda_lib, oracle_lib, W, H, P, B):
    """fog_of_war = 0: every tile is in sight for every player (each writer takes its own no-fog branch)."""
    g, _ = gym_rollout(cuda_lib, W, H, P, B, 30, 500, False, fog=0)
    o, _ = gym_rollout(oracle_lib, W, H, P, B, 30, 500, False, fog=0)
    _compare(g, o, f"cuda vs oracle {W}x{H}x{P}p, no fog")


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,P,B", [(10, 10, 2, 96), (15, 15, 2, 64), (20, 20, 2, 40), (8, 8, 3, 64), (10, 10, 4, 64)])
def test_cuda_gym_step_matches_oracle_on_dense_battle_states(cuda_lib, oracle_lib, W, H, P, B):
    """Hand-built crowded boards (several general-type tiles per player, stale cached lists, arbitrary visibility bits —
    test_cuda_parity.dense_battle_state): eliminations, the +50 bonus, terminated episodes and orphaned tiles all occur
    within a few steps, and every writer has to agree with the oracle on states no map generator produces."""
    from test_cuda_parity import dense_battle_state

    init = dense_battle_state(np.random.default_rng(W * 7 + P), W, H, P, B)
    g, _ = gym_rollout(cuda_lib, W, H, P, B, 40, 500, False, init=init)
    o, _ = gym_rollout(oracle_lib, W, H, P, B, 40, 500, False, init=init)
    _compare(g, o, f"cuda vs oracle {W}x{H}x{P}p, dense battles")
    assert any(s["terminated"].any() for s in o), "games end on these boards"


def _partial_observe(lib, W, H, P, B, on_device):
    """Play, observe everything, play on, re-observe a subset: those rows equal a full read-out, the others keep
    the earlier one (grl_gym_observe_envs — what the vector env calls after re-seeding finished envs)."""
    import torch

    from generalsreinforcementlearning_b200 import _abi

    dev = torch.device("cuda", 0) if on_device else torch.device("cpu")
    e = BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, max_actions=P, host_threads=1))
    if on_device:
        e.use_torch_stream()
    e.reset_seeded(np.arange(B, dtype=np.int64) + 99)
    N = W * H
    mk = lambda: (torch.zeros((B, P, 9, H, W), dtype=torch.float32, device=dev), torch.zeros((B, P, N * 5), dtype=torch.uint8, device=dev),
                  torch.zeros((B, P, 4), dtype=torch.int32, device=dev))
    for _ in range(12):
        e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 3)
    part = mk()
    e.gym_observe(500, *part)
    before = [t.clone() for t in part]
    for _ in range(9):
        e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 4)
    ids = np.array(sorted({0, B - 1, B // 2, 7 % B, 33 % B}), dtype=np.int32)
    e.gym_observe_envs(500, ids, *part)
    full = mk()
    e.gym_observe(500, *full)
    if on_device:
        torch.cuda.synchronize()
    sel = torch.zeros(B, dtype=torch.bool)
    sel[torch.as_tensor(ids.astype(np.int64))] = True
    for got, new, old, name in zip(part, full, before, ("obs", "mask", "stats")):
        got, new, old = got.cpu(), new.cpu(), old.cpu()
        assert torch.equal(got[sel], new[sel]), f"{name}: listed envs must hold the new read-out"
        assert torch.equal(got[~sel], old[~sel]), f"{name}: other envs must be left untouched"
        assert not torch.equal(new[~sel], old[~sel]), "the state did move on"
    e.close()


def test_oracle_gym_observe_envs(oracle_lib):
    _partial_observe(oracle_lib, 8, 8, 2, 40, False)


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,P,B", [(15, 15, 2, 300), (20, 20, 2, 65), (9, 7, 3, 40)])
def test_cuda_gym_observe_envs(cuda_lib, W, H, P, B):
    _partial_observe(cuda_lib, W, H, P, B, True)


def _sampled_actions(lib, W, H, P, B, on_device, turns=25):
    import torch

    from generalsreinforcementlearning_b200 import _abi

    dev = torch.device("cuda", 0) if on_device else torch.device("cpu")
    e = BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, max_actions=P, host_threads=1,
                                       env_id_base=1000))
    if on_device:
        e.use_torch_stream()   # the planes are filled by torch kernels and by the engine in turn
    e.reset_seeded(np.arange(B, dtype=np.int64) + 5)
    for _ in range(turns):
        e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 3)
    N = W * H
    mask = torch.zeros((B, P, N * 5), dtype=torch.uint8, device=dev)
    e.gym_observe(500, None, mask, None)
    out = []
    for p in range(P):
        for seed in (1, 2, 2 ** 63 + 11):
            a = torch.full((B,), -7, dtype=torch.int64, device=dev)
            e.gym_sample(seed, mask, p, a)
            out.append(a.cpu().numpy().copy())
    m = mask.cpu().numpy()
    e.close()
    return m, out


def test_oracle_gym_sample_draws_valid_actions(oracle_lib):
    m, draws = _sampled_actions(oracle_lib, 8, 8, 2, 64, False)
    P = m.shape[1]
    k = 0
    for p in range(P):
        per_seed = []
        for _ in range(3):
            a = draws[k]; k += 1
            has = m[:, p].any(1)
            assert (m[np.arange(len(a)), p, a][has] == 1).all(), "a drawn action is valid wherever one exists"
            assert (a[~has] == 0).all()
            per_seed.append(a)
        assert not np.array_equal(per_seed[0], per_seed[1]), "different seeds, different draws"
    # uniform over the valid entries: over many envs every position class of the k-th-set-bit draw is used
    first = np.array([np.flatnonzero(m[b, 0])[0] if m[b, 0].any() else 0 for b in range(m.shape[0])])
    assert (draws[0] != first).any(), "not simply the first valid action"


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,P,B", [(15, 15, 2, 300), (20, 20, 2, 70), (9, 7, 3, 33)])
def test_cuda_gym_sample_matches_oracle(cuda_lib, oracle_lib, W, H, P, B):
    mg, g = _sampled_actions(cuda_lib, W, H, P, B, True)
    mo, o = _sampled_actions(oracle_lib, W, H, P, B, False)
    assert np.array_equal(mg, mo)
    for x, y in zip(g, o):
        assert np.array_equal(x, y)


def _drive_pair(a, b, steps, compare_final=True, compact_info=False, b_random_agent=False):
    """Two vector envs stepped with the same actions; every returned tensor compared each step.  With b_random_agent env b
    draws its agent's action inside the step (step(None), grl_gym_step_io.action == NULL) while env a plays what its
    sampler (grl_gym_sample) returned: the two must be the same indices, step after step."""
    import torch

    oa, _ = a.reset()
    ob, _ = b.reset()
    resets = 0
    for t in range(steps):
        assert torch.equal(oa.cpu(), ob.cpu()), f"observation, step {t}"
        act = a.sample_actions().cpu()
        if b_random_agent:
            ra, rb = a.step(act.to(a.device)), b.step(None)
            assert torch.equal(act, rb[4]["action"].cpu()), f"the action drawn inside the step, step {t}"
            assert torch.equal(ra[4]["invalid_action"].cpu(), rb[4]["invalid_action"].cpu()), f"valid flags, step {t}"
        else:
            assert torch.equal(act, b.sample_actions().cpu()), f"sampled actions, step {t}"
            if t % 5 == 2:
                act = act.clone()
                act[::7] = 0
            ra, rb = a.step(act.to(a.device)), b.step(act.to(b.device))
        oa, ob = ra[0], rb[0]
        for k, name in ((1, "reward"), (2, "terminated"), (3, "truncated")):
            assert torch.equal(ra[k].cpu(), rb[k].cpu()), f"{name}, step {t}"
        assert torch.equal(ra[4]["valid_actions_mask"].cpu(), rb[4]["valid_actions_mask"].cpu()), f"mask, step {t}"
        assert torch.equal(a._turns.cpu(), b._turns.cpu()) and torch.equal(a._calls.cpu(), b._calls.cpu()), f"counters, step {t}"
        fin = (ra[2] | ra[3]).cpu()
        resets += int(fin.sum())
        if compare_final and fin.any():
            fa = ra[4]["final_observation"].cpu()
            fb = rb[4]["final_observation"].cpu()
            fa = fa[fin] if fa.shape[0] == fin.shape[0] else fa
            fb = fb[fin] if fb.shape[0] == fin.shape[0] else fb
            assert torch.equal(fa, fb), f"final observation, step {t}"
        if compact_info:   # both envs hand out the compact form: ids, final views and the pre-reset turn counters
            assert ("final_env_ids" in ra[4]) == bool(fin.any()) and ("final_env_ids" in rb[4]) == bool(fin.any())
            assert torch.equal(ra[4]["turn"].cpu(), rb[4]["turn"].cpu()), f"info turn, step {t}"
            if fin.any():
                assert torch.equal(ra[4]["final_env_ids"].cpu(), rb[4]["final_env_ids"].cpu()), f"final env ids, step {t}"
                assert torch.equal(ra[4]["final_env_ids"].cpu(), fin.nonzero(as_tuple=True)[0])
                assert torch.equal(ra[4]["final_observation"].cpu(), rb[4]["final_observation"].cpu())
    return resets


def test_vector_env_device_autoreset_equals_host_autoreset(oracle_lib):
    """auto_reset='device' (grl_gym_autoreset: no host read, dense final_observation) and 'host' (the same re-seeding,
    compact final_observation / final_env_ids after one flag read) re-seed the same envs with the same seeds as the
    host-driven path ('host_reset': grl_reset_seeded + grl_gym_observe_envs) and hand out the same tensors."""
    from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

    a = GeneralsVecEnv(40, 8, 8, max_turns=9, seed=5, lib=oracle_lib, host_threads=1, auto_reset="host_reset")
    b = GeneralsVecEnv(40, 8, 8, max_turns=9, seed=5, lib=oracle_lib, host_threads=1, auto_reset="device")
    assert _drive_pair(a, b, 40) >= 120
    a.close()
    b.close()
    a = GeneralsVecEnv(40, 8, 8, max_turns=9, seed=5, lib=oracle_lib, host_threads=1, auto_reset="host_reset")
    b = GeneralsVecEnv(40, 8, 8, max_turns=9, seed=5, lib=oracle_lib, host_threads=1)   # the default: "host"
    assert b.auto_reset == "host"
    assert _drive_pair(a, b, 40, compact_info=True) >= 120
    a.close()
    b.close()


def test_vector_env_random_agent_step_equals_sampled_step(oracle_lib):
    """step(None) — the random agent drawn by grl_gym_step itself (io.action == NULL, agent_seed, sampled_action) — plays
    exactly step(sample_actions()): same indices, same transitions, through several generations of episodes."""
    from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

    for fog in (True, False):
        a = GeneralsVecEnv(40, 8, 8, max_turns=9, seed=5, fog_of_war=fog, lib=oracle_lib, host_threads=1, auto_reset="device")
        b = GeneralsVecEnv(40, 8, 8, max_turns=9, seed=5, fog_of_war=fog, lib=oracle_lib, host_threads=1, auto_reset="device")
        assert _drive_pair(a, b, 40, b_random_agent=True) >= 120
        a.close()
        b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,B,max_turns,fog", [(15, 15, 300, 30, True), (20, 20, 96, 40, True), (10, 10, 1000, 25, True),
                                                 (15, 15, 130, 30, False), (6, 9, 70, 20, True), (32, 32, 9, 30, True)])
def test_cuda_random_agent_step_matches_sampler_and_oracle(cuda_lib, oracle_lib, W, H, B, max_turns, fog):
    """The in-launch random agent of the CUDA gym step (gym_random_agent: the k-th valid entry from the direction masks
    in registers) against the oracle stepping what ITS sampler drew from the mask bytes, and against the CUDA sampler
    kernel: the same indices and the same transitions on every baked board, a generic one and the largest one."""
    from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

    mk = lambda lib, **kw: GeneralsVecEnv(B, W, H, max_turns=max_turns, seed=43, fog_of_war=fog, lib=lib, auto_reset="device", **kw)
    o, g = mk(oracle_lib, host_threads=1), mk(cuda_lib)
    assert _drive_pair(o, g, 2 * max_turns + 5, b_random_agent=True) >= B
    import numpy as np
    assert np.array_equal(g.engine.state_hash(), o.engine.state_hash())
    g1, g2 = mk(cuda_lib), mk(cuda_lib)
    _drive_pair(g1, g2, max_turns + 3, b_random_agent=True)     # CUDA sampler kernel + step == CUDA step(None)
    for e in (o, g, g1, g2):
        e.close()


def test_seeded_reset_replays_the_same_episode(oracle_lib):
    """reset(seed=s) restarts the maps and both random streams (the opponent's and the in-step agent's draws): the same
    env replays the same trajectory; another seed does not."""
    import torch

    from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

    env = GeneralsVecEnv(24, 8, 8, max_turns=9, seed=1, lib=oracle_lib, host_threads=1, auto_reset="device")

    def run(seed):
        obs, _ = env.reset(seed=seed)
        trace = [obs.clone()]
        for _ in range(25):
            obs, r, te, tr, info = env.step(None)
            trace += [obs.clone(), r.clone(), te.clone(), tr.clone(), info["action"].clone()]
        return trace

    a, b, c = run(5), run(5), run(6)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert not all(torch.equal(x, y) for x, y in zip(a, c))
    env.close()


def _random_agent_call(lib, P, sampled):
    import torch

    dev = torch.device("cuda", 0) if lib.prefix == "grl_" else torch.device("cpu")
    e = BatchedEngine(lib, make_config(lib, num_envs=8, width=8, height=8, num_players=P, max_actions=P, host_threads=1))
    if lib.prefix == "grl_":
        e.use_torch_stream()
    e.reset_seeded(np.arange(8, dtype=np.int64) + 3)
    pl = _planes(torch, dev, 8, P, 64, 8, 8)
    e.gym_observe(20, pl["obs"], pl["mask"], pl["stats"])
    try:
        e.gym_step(20, 1, agent_seed=5, sampled_action=torch.zeros(8, dtype=torch.int64, device=dev) if sampled else None, **pl)
    finally:
        e.close()


def test_random_agent_step_needs_its_output_plane(oracle_lib):
    """action == NULL without sampled_action is an argument error, not a draw that nobody can read."""
    with pytest.raises(RuntimeError):
        _random_agent_call(oracle_lib, 2, sampled=False)
    _random_agent_call(oracle_lib, 2, sampled=True)
    _random_agent_call(oracle_lib, 3, sampled=True)     # the oracle draws for any number of players


@pytest.mark.gpu
def test_cuda_random_agent_step_argument_errors(cuda_lib):
    """libgrlcuda.so: the same argument error, and a clear GRL_ERR_UNSUPPORTED for envs of more than two players (the
    in-launch agent is instantiated for the two-player template; grl_gym_sample + action serves the others)."""
    with pytest.raises(RuntimeError):
        _random_agent_call(cuda_lib, 2, sampled=False)
    _random_agent_call(cuda_lib, 2, sampled=True)
    with pytest.raises(RuntimeError, match="two-player"):
        _random_agent_call(cuda_lib, 3, sampled=True)


@pytest.mark.gpu
@pytest.mark.parametrize("W,B,max_turns", [(15, 300, 7), (20, 96, 9), (10, 1000, 6)])
def test_cuda_device_autoreset_matches_oracle(cuda_lib, oracle_lib, W, B, max_turns):
    """The device-side auto-reset (compaction, map generation, turn-0 set-up and read-outs sized by a device-side count)
    against the oracle's, through several generations of episodes."""
    from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

    g = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=31, lib=cuda_lib, auto_reset="device")
    o = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=31, lib=oracle_lib, host_threads=1, auto_reset="device")
    assert _drive_pair(g, o, 4 * max_turns + 3) >= 3 * B
    import numpy as np
    assert np.array_equal(g.engine.state_hash(), o.engine.state_hash())
    # and the device path equals the CUDA host path
    h = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=31, lib=cuda_lib, auto_reset="host_reset")
    g2 = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=31, lib=cuda_lib, auto_reset="device")
    _drive_pair(g2, h, 2 * max_turns + 2)
    # and the compact-info mode over the device path equals the host-driven one, info entries included
    h2 = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=31, lib=cuda_lib, auto_reset="host_reset")
    g3 = GeneralsVecEnv(B, W, W, max_turns=max_turns, seed=31, lib=cuda_lib, auto_reset="host")
    _drive_pair(g3, h2, 2 * max_turns + 2, compact_info=True)
    for e in (g, o, h, g2, h2, g3):
        e.close()
