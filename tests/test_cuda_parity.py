"""Bit-exact parity of the CUDA turn engine against the CPU oracle on identical inputs:
full state (board, cached lists, visibility, tile sets, player stats), observation planes,
packed legal masks, rewards, done/winner, step errors and action indices — every turn."""
import os

import numpy as np
import pytest

from generalsreinforcementlearning_b200 import _abi
from generalsreinforcementlearning_b200.engine import make_actions
from helpers import CITY, GENERAL, MOUNTAIN, NORMAL, blank_state, compare_states, full_fog, full_stats, new_engine

pytestmark = pytest.mark.gpu

OUT_KEYS = ("obs", "mask_bits", "reward", "done", "winner", "step_error", "action_index")


def compare_outputs(a, b, ctx):
    for k in OUT_KEYS:
        x, y = a[k], b[k]
        if k in ("obs", "reward"):
            x, y = x.view(np.uint32), y.view(np.uint32)  # bit-exact float32
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            raise AssertionError(f"{ctx}: output {k!r} differs at {bad[:6].tolist()} ({len(bad)} cells): "
                                 f"{a[k][tuple(bad[0])]} vs {b[k][tuple(bad[0])]}")


def corrupt_actions(rng, actions, W, H, rate):
    """Inject invalid moves (every validation error class) so error turns occur."""
    B, A = actions.shape
    hit = rng.random((B, A)) < rate
    for b, s in zip(*np.nonzero(hit)):
        a = actions[b, s]
        kind = rng.integers(0, 6)
        if kind == 0:
            a["from_x"] = rng.integers(-2, W + 2)
            a["from_y"] = rng.integers(-2, H + 2)
        elif kind == 1:
            a["to_x"] = rng.integers(-2, W + 2)
        elif kind == 2:
            a["to_x"], a["to_y"] = a["from_x"], a["from_y"]
        elif kind == 3:
            a["to_x"] = a["from_x"] + rng.integers(-2, 3)
            a["to_y"] = a["from_y"] + rng.integers(-2, 3)
        elif kind == 4:
            a["player_id"] = rng.integers(-1, 9)
        else:
            a["from_x"], a["from_y"] = rng.integers(0, W), rng.integers(0, H)
            a["to_x"], a["to_y"] = a["from_x"], min(H - 1, a["from_y"] + 1)
        a["present"] = 1


def rollout_compare(cuda_lib, oracle_lib, W, H, P, B, T, seed, err_rate=0.0, policy_in_kernel=False,
                    init=None, check_state_every=1, **cfg):
    gc = new_engine(cuda_lib, W, H, P, B, **cfg)
    oc = new_engine(oracle_lib, W, H, P, B, **cfg)
    if init is None:
        seeds = np.arange(B, dtype=np.int64) + 12345 + seed * 1000
        gc.reset_seeded(seeds)
        oc.reset_seeded(seeds)
    else:
        gc.set_state(init)
        oc.set_state(init)
    compare_states(gc.get_state(), oc.get_state(), "after reset")
    go, oo = gc.alloc_outputs_host(), oc.alloc_outputs_host()
    gc.observe(gc.outputs(**go))
    oc.observe(oc.outputs(**oo))
    compare_outputs(go, oo, "observe after reset")
    rng = np.random.default_rng(seed)
    for t in range(T):
        if policy_in_kernel:
            gc.step_fused(None, gc.outputs(**go), flags=_abi.STEP_FLAG_RANDOM_POLICY, policy_seed=seed + 7)
            oc.step_fused(None, oc.outputs(**oo), flags=_abi.STEP_FLAG_RANDOM_POLICY, policy_seed=seed + 7)
        else:
            actions = oc.sample_actions(seed + 7)
            if t % 3 == 0:  # the device-side sampler must draw the same moves
                assert np.array_equal(gc.sample_actions(seed + 7), actions), f"turn {t}: sampled actions differ"
            if err_rate:
                corrupt_actions(rng, actions, W, H, err_rate)
            gc.step_fused(actions, gc.outputs(**go))
            oc.step_fused(actions, oc.outputs(**oo))
        compare_outputs(go, oo, f"turn {t}")
        if t % check_state_every == 0 or t == T - 1:
            compare_states(gc.get_state(), oc.get_state(), f"turn {t}")
            assert np.array_equal(gc.state_hash(), oc.state_hash()), f"turn {t}: state digests differ"
    assert np.array_equal(gc.stats(), oc.stats())
    return gc, oc


@pytest.mark.parametrize("W,H,P", [(10, 10, 2), (15, 15, 2), (20, 20, 2), (20, 20, 4), (5, 5, 2), (7, 13, 3),
                                    (12, 9, 1), (32, 32, 8), (14, 11, 5)])
def test_seeded_rollout_parity(cuda_lib, oracle_lib, W, H, P):
    gc, _ = rollout_compare(cuda_lib, oracle_lib, W, H, P, B=48, T=120, seed=W * 100 + P)
    assert gc.stats()[0] > 0


@pytest.mark.parametrize("W,H,P", [(10, 10, 2), (20, 20, 2), (20, 20, 4), (15, 15, 3)])
def test_error_turn_parity(cuda_lib, oracle_lib, W, H, P):
    """Invalid actions abort the turn after moves and before production."""
    gc, _ = rollout_compare(cuda_lib, oracle_lib, W, H, P, B=64, T=150, seed=3, err_rate=0.08)
    assert gc.stats()[1] > 20  # error turns really happened


@pytest.mark.parametrize("W,H,P", [(10, 10, 2), (20, 20, 2), (20, 20, 4), (15, 15, 2)])
def test_in_kernel_policy_parity(cuda_lib, oracle_lib, W, H, P):
    rollout_compare(cuda_lib, oracle_lib, W, H, P, B=64, T=200, seed=11, policy_in_kernel=True, check_state_every=10)


def dense_battle_state(rng, W, H, P, B):
    """Crowded boards with big armies, extra general-type tiles (Q4), stale cached lists and
    arbitrary visibility bits: eliminations, resurrections (Q9), mutual captures (Q10) and
    orphaned tiles (Q7) all occur within a few turns of random play."""
    N = W * H
    s = blank_state(W, H, P, B)
    for b in range(B):
        owner = rng.integers(-1, P, N)
        army = rng.integers(0, 60, N)
        type_ = np.zeros(N, np.int64)
        type_[rng.random(N) < 0.08] = CITY
        type_[rng.random(N) < 0.08] = MOUNTAIN
        owner[type_ == MOUNTAIN] = -1
        army[type_ == MOUNTAIN] = 0
        for p in range(P):
            for _ in range(rng.integers(1, 3)):
                i = rng.integers(0, N)
                owner[i], type_[i], army[i] = p, GENERAL, rng.integers(1, 8)
        s["owner"][b], s["army"][b], s["type"][b] = owner, army, type_
    full_stats(s)
    full_fog(s, W, H)
    # stale lists: drop some owned tiles, keep the lists disjoint
    drop = rng.random(s["owned"].shape) < 0.1
    s["owned"][drop] = 0
    s["visible"] ^= (rng.random(s["visible"].shape) < 0.05).astype(np.uint32) * np.uint32(rng.integers(1, 1 << P))
    s["vis_changed"][:] = (rng.random(s["vis_changed"].shape) < 0.03)
    s["turn"][:] = rng.integers(0, 60, B)
    return s


@pytest.mark.parametrize("W,H,P", [(5, 5, 2), (6, 6, 4), (8, 8, 3), (10, 10, 2), (9, 7, 8)])
def test_dense_battle_parity(cuda_lib, oracle_lib, W, H, P):
    rng = np.random.default_rng(W * 31 + P)
    init = dense_battle_state(rng, W, H, P, B=96)
    gc, _ = rollout_compare(cuda_lib, oracle_lib, W, H, P, B=96, T=80, seed=5, err_rate=0.03, init=init,
                            max_actions=max(2, P))
    st = gc.stats()
    assert st[2] > 0, "no game finished: the elimination path was not exercised"


@pytest.mark.parametrize("W,H,P,A", [(8, 8, 2, 6), (10, 10, 4, 12), (10, 10, 2, 12), (15, 15, 3, 9), (20, 20, 4, 12)])
def test_multiple_actions_per_player(cuda_lib, oracle_lib, W, H, P, A):
    """UI-style submission: several moves per player per turn (up to the 12-slot cap), applied in stable
    player-id order — on the generic kernel and on packed lane groups, where a group's lanes decode
    more than one slot each."""
    B = 45
    rng = np.random.default_rng(9 + A)
    init = dense_battle_state(rng, W, H, P, B)
    gc = new_engine(cuda_lib, W, H, P, B, max_actions=A)
    oc = new_engine(oracle_lib, W, H, P, B, max_actions=A)
    gc.set_state(init)
    oc.set_state(init)
    go, oo = gc.alloc_outputs_host(), oc.alloc_outputs_host()
    for t in range(40):
        actions = make_actions(B, A)
        st = oc.get_state()
        for b in range(B):
            for sl in range(A):
                if rng.random() < 0.2:
                    continue
                p = int(rng.integers(0, P))
                mine = np.nonzero((st["owner"][b] == p) & (st["army"][b] > 1))[0]
                if len(mine) == 0:
                    continue
                i = int(rng.choice(mine))
                dx, dy = [(0, -1), (1, 0), (0, 1), (-1, 0)][int(rng.integers(0, 4))]
                a = actions[b, sl]
                a["player_id"], a["from_x"], a["from_y"] = p, i % W, i // W
                a["to_x"], a["to_y"], a["move_all"], a["present"] = i % W + dx, i // W + dy, rng.integers(0, 2), 1
        gc.step_fused(actions, gc.outputs(**go))
        oc.step_fused(actions, oc.outputs(**oo))
        compare_outputs(go, oo, f"turn {t}")
        compare_states(gc.get_state(), oc.get_state(), f"turn {t}")


def test_readout_variants_match_oracle(cuda_lib, oracle_lib):
    W, H, P, B = 15, 15, 3, 32
    gc, oc = rollout_compare(cuda_lib, oracle_lib, W, H, P, B, T=60, seed=21, err_rate=0.05)
    for variant in (_abi.MASK_ENGINE_URDL, _abi.MASK_SERIALIZER_UDLR, _abi.MASK_ENGINE_URDL_BITS,
                    _abi.MASK_ENGINE_HALF_BITS):
        assert np.array_equal(gc.mask(variant), oc.mask(variant)), variant
    gv, gf = gc.visibility()
    ov, of = oc.visibility()
    assert np.array_equal(gv, ov) and np.array_equal(gf, of)
    # observation digest helper agrees with the oracle's on the same planes
    go = gc.alloc_outputs_host()
    gc.observe(gc.outputs(**go))
    rows, row_words = B * P, 9 * W * H
    assert np.array_equal(gc.buffer_hash(go["obs"], row_words, rows), oc.buffer_hash(go["obs"], row_words, rows))


def test_step_and_fused_step_agree(cuda_lib):
    """grl_step (turn kernel alone) and grl_step_fused leave identical state."""
    W, H, P, B = 20, 20, 2, 128
    a = new_engine(cuda_lib, W, H, P, B)
    b = new_engine(cuda_lib, W, H, P, B)
    seeds = np.arange(B) + 777
    a.reset_seeded(seeds)
    b.reset_seeded(seeds)
    out = b.alloc_outputs_host()
    for t in range(60):
        a.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 99)
        b.step_fused(None, b.outputs(**out), _abi.STEP_FLAG_RANDOM_POLICY, 99)
        assert np.array_equal(a.state_hash(), b.state_hash()), t
    compare_states(a.get_state(), b.get_state(), "step vs fused")


def test_device_buffers_with_torch(cuda_lib, oracle_lib):
    """Caller-allocated DEVICE buffers (torch tensors): no staging, asynchronous on the env stream."""
    import torch

    W, H, P, B = 20, 20, 2, 256
    gc = new_engine(cuda_lib, W, H, P, B)
    oc = new_engine(oracle_lib, W, H, P, B)
    seeds = np.arange(B) + 4242
    gc.reset_seeded(seeds)
    oc.reset_seeded(seeds)
    dev = torch.device("cuda:0")
    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    mask = torch.empty((B, P, gc.mask_words), dtype=torch.int32, device=dev)
    reward = torch.empty((B, P), dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    winner = torch.empty(B, dtype=torch.int8, device=dev)
    serr = torch.empty(B, dtype=torch.uint8, device=dev)
    aidx = torch.empty((B, P), dtype=torch.int32, device=dev)
    acts = torch.empty((B, gc.A, 8), dtype=torch.uint8, device=dev)
    oo = oc.alloc_outputs_host()
    for t in range(50):
        gc.sample_actions(5, acts)
        gc.step_fused(acts, gc.outputs(obs, mask, reward, done, winner, serr, aidx))
        gc.sync()
        host_actions = acts.cpu().numpy().view(_abi.ACTION_DTYPE).reshape(B, gc.A)
        oc.step_fused(host_actions, oc.outputs(**oo))
        go = dict(obs=obs.cpu().numpy(), mask_bits=mask.cpu().numpy().view(np.uint32), reward=reward.cpu().numpy(),
                  done=done.cpu().numpy(), winner=winner.cpu().numpy(), step_error=serr.cpu().numpy(),
                  action_index=aidx.cpu().numpy())
        compare_outputs(go, oo, f"turn {t}")
    assert gc.launch_count() >= 100


def test_partial_reset_and_unreset_envs(cuda_lib, oracle_lib):
    W, H, P, B = 10, 10, 2, 16
    gc = new_engine(cuda_lib, W, H, P, B)
    oc = new_engine(oracle_lib, W, H, P, B)
    ids = [3, 7, 8, 15]
    seeds = [5, 6, 7, 8]
    gc.reset_seeded(seeds, ids)
    oc.reset_seeded(seeds, ids)
    go, oo = gc.alloc_outputs_host(), oc.alloc_outputs_host()
    for t in range(10):
        gc.step_fused(None, gc.outputs(**go), _abi.STEP_FLAG_RANDOM_POLICY, 3)
        oc.step_fused(None, oc.outputs(**oo), _abi.STEP_FLAG_RANDOM_POLICY, 3)
        for k in ("reward", "done", "step_error", "action_index"):
            assert np.array_equal(go[k], oo[k]), (t, k)
        assert np.array_equal(go["obs"][ids], oo["obs"][ids])
    # never-reset envs reject steps like finished games (turn_processor.go:95-113)
    assert (go["step_error"][[0, 1, 2]] == _abi.STEP_GAME_OVER).all()
    assert gc.stats()[3] == oc.stats()[3] == 12 * 10
    # re-seed a subset mid-flight
    gc.reset_seeded([99], [7])
    oc.reset_seeded([99], [7])
    compare_states(gc.get_state(3, 13), oc.get_state(3, 13), "after partial reset")


def test_set_state_rejects_unreachable_states(cuda_lib):
    e = new_engine(cuda_lib, 5, 5, 2)
    s = blank_state(5, 5, 2)
    s["army"][0, 0] = 70000
    with pytest.raises(RuntimeError, match="uint16"):
        e.set_state(s)
    s = blank_state(5, 5, 2)
    s["owned"][0, 0, 3] = s["owned"][0, 1, 3] = 1
    with pytest.raises(RuntimeError, match="at most one"):
        e.set_state(s)


def test_army_overflow_is_flagged_not_wrapped(cuda_lib):
    W = H = 5
    e = new_engine(cuda_lib, W, H, 2)
    s = blank_state(W, H, 2)
    s["owner"][0, 0], s["army"][0, 0], s["type"][0, 0] = 0, 65535, GENERAL
    s["owner"][0, 24], s["army"][0, 24], s["type"][0, 24] = 1, 5, GENERAL
    full_stats(s)
    e.set_state(s)
    e.step(None)
    t = e.get_state()
    assert t["army"][0, 0] == 65535
    assert t["step_error"][0] == _abi.STEP_ARMY_OVERFLOW


@pytest.mark.parametrize("W,H,P", [(10, 10, 2), (10, 10, 3), (10, 10, 4), (15, 15, 2), (15, 15, 3), (15, 15, 4),
                                    (20, 20, 2), (20, 20, 3), (20, 20, 4), (20, 20, 5)])
def test_baked_board_instantiations(cuda_lib, oracle_lib, W, H, P):
    """Every baked instantiation (board x player template; five players fall through to the generic kernel), with a
    batch that leaves the last warp partly filled, invalid moves (error turns) mixed in, and the in-kernel policy."""
    gc, _ = rollout_compare(cuda_lib, oracle_lib, W, H, P, B=37, T=60, seed=W * 10 + P, err_rate=0.05)
    assert gc.stats()[1] > 0, "the run must contain error turns"
    rollout_compare(cuda_lib, oracle_lib, W, H, P, B=5, T=40, seed=W + P, policy_in_kernel=True)


def test_lane_groups_dense_endgames(cuda_lib, oracle_lib):
    """Packed groups through eliminations, tile turnover and game endings: dense 10x10 battles where each
    group of a warp is at a different stage (some games over, some aborting, some eliminating)."""
    for LG in (8, 9):   # two seeds
        W = H = 10
        P, B = 4, 29
        rng = np.random.default_rng(LG)
        s = blank_state(W, H, P, B)
        for c in range(B):
            owners = rng.integers(-1, P, W * H)
            s["owner"][c] = owners
            s["army"][c] = np.where(owners >= 0, rng.integers(1, 60, W * H), rng.integers(0, 3, W * H))
            s["type"][c] = rng.choice([NORMAL, NORMAL, NORMAL, CITY, MOUNTAIN], W * H)
            for p in range(P):  # one general each, weakly defended so captures happen
                i = int(rng.integers(0, W * H))
                s["owner"][c, i], s["army"][c, i], s["type"][c, i] = p, int(rng.integers(1, 4)), GENERAL
            mnt = s["type"][c] == MOUNTAIN
            s["owner"][c][mnt], s["army"][c][mnt] = -1, 0
        full_stats(s)
        full_fog(s, W, H)
        gc, _ = rollout_compare(cuda_lib, oracle_lib, W, H, P, B=B, T=150, seed=77 + LG, init=s, err_rate=0.02)
        assert gc.stats()[2] > 0, "some games must finish"


@pytest.mark.parametrize("W,H,P", [(10, 10, 2), (15, 15, 2), (20, 20, 2), (20, 20, 4), (5, 5, 2), (3, 3, 2), (32, 32, 8),
                                    (25, 25, 4), (7, 13, 3), (2, 2, 2)])
def test_device_mapgen_matches_host_and_oracle(cuda_lib, oracle_lib, W, H, P, monkeypatch):
    """Seeded resets generate their maps on the device (grl_mapgen_gpu.cu).  Thousands of seeds —
    negative, zero, above 2^31 and consecutive — must give the boards of the host generator
    (GRL_HOST_MAPGEN=1) and of the oracle's separately written generator, tile for tile."""
    B = 3000
    rng = np.random.default_rng(W * 31 + P)
    seeds = np.concatenate([np.arange(1000, dtype=np.int64) + 12345, rng.integers(-2**62, 2**62, 1990, dtype=np.int64),
                            np.array([0, -1, 2**31 - 1, 2**31, -(2**31), 89482311, 2**31 - 2, 1, -(2**31 - 1), 2**32 + 5], np.int64)])
    dev = new_engine(cuda_lib, W, H, P, B)
    dev.reset_seeded(seeds)
    monkeypatch.setenv("GRL_HOST_MAPGEN", "1")
    host = new_engine(cuda_lib, W, H, P, B, host_threads=0)
    host.reset_seeded(seeds)
    orc = new_engine(oracle_lib, W, H, P, B, host_threads=0)
    orc.reset_seeded(seeds)
    a, b, c = dev.get_state(), host.get_state(), orc.get_state()
    for k in a:
        assert np.array_equal(a[k], b[k]), f"device vs host mapgen: {k}"
        assert np.array_equal(a[k], c[k]), f"device mapgen vs oracle: {k}"
    assert np.array_equal(dev.state_hash(), orc.state_hash())
    # partial reset by env id goes through the same device path
    ids = np.arange(0, B, 7, dtype=np.int32)
    dev.reset_seeded(seeds[: len(ids)] + 99, ids)
    orc.reset_seeded(seeds[: len(ids)] + 99, ids)
    assert np.array_equal(dev.state_hash(), orc.state_hash())


def test_device_mapgen_reports_unplaceable_generals(cuda_lib, oracle_lib, monkeypatch):
    """mapgen/generator.go:166-253: when a general finds no tile far enough from the others the generator fails.  Eight
    generals at Manhattan distance >= 2 do not fit a 3x3 board (five do), so every seed fails: the device generator, the
    host generator and the oracle all report map generation failed; the device path leaves such envs as empty boards
    without generals, i.e. finished games, never half-written ones."""
    seeds = np.arange(16, dtype=np.int64) + 5
    dev = new_engine(cuda_lib, 3, 3, 8, 16)
    with pytest.raises(RuntimeError, match="map generation failed"):
        dev.reset_seeded(seeds)
    s = dev.get_state()
    assert (s["owner"] == -1).all() and (s["army"] == 0).all() and (s["type"] == 0).all() and s["game_over"].all()
    orc = new_engine(oracle_lib, 3, 3, 8, 16, host_threads=0)
    with pytest.raises(RuntimeError, match="map generation failed"):
        orc.reset_seeded(seeds)
    monkeypatch.setenv("GRL_HOST_MAPGEN", "1")
    host = new_engine(cuda_lib, 3, 3, 8, 16, host_threads=0)
    with pytest.raises(RuntimeError, match="map generation failed"):
        host.reset_seeded(seeds)


@pytest.mark.parametrize("W,H,P", [(10, 10, 2), (15, 15, 2), (20, 20, 2), (20, 20, 4), (9, 9, 2)])
def test_fog_disabled_parity(cuda_lib, oracle_lib, W, H, P):
    """GameState.FogOfWarEnabled == false (the reference's tests build such engines by hand): every tile is
    visible to every player in the observation planes; the baked plane-major / linear / packed writers
    and the generic one must all agree with the oracle."""
    rollout_compare(cuda_lib, oracle_lib, W, H, P, B=41, T=50, seed=5 + P, err_rate=0.03, fog_of_war=0)
    g, o = new_engine(cuda_lib, W, H, P, 9, fog_of_war=0), new_engine(oracle_lib, W, H, P, 9, fog_of_war=0)
    for e in (g, o):
        e.reset_seeded(np.arange(9, dtype=np.int64) + 3)
        for _ in range(20):
            e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 1)
    for e in (g, o):
        e.obs_gym = np.zeros((9, P, 9, H, W), np.float32)
        e.mask_gym = np.zeros((9, P, W * H * 5), np.uint8)
        e.gym_observe(100, e.obs_gym, e.mask_gym, None)
    assert np.array_equal(g.obs_gym.view(np.uint32), o.obs_gym.view(np.uint32)) and np.array_equal(g.mask_gym, o.mask_gym)
    assert (g.obs_gym[:, :, 0] == 1.0).all(), "without fog every tile is visible"


def test_pinned_host_results_are_written_in_place(cuda_lib, oracle_lib):
    """Result planes in PINNED host memory (what a trainer's host loop allocates) are written by the kernel
    directly — no staging copy — and must hold the oracle's values when the call returns; pageable buffers
    of the same call take the staged path."""
    import torch

    W, H, P, B = 20, 20, 2, 9000   # >= 8192: the staged planes of the call go through the pipelined sub-ranges
    gc, oc = new_engine(cuda_lib, W, H, P, B, host_threads=0), new_engine(oracle_lib, W, H, P, B, host_threads=0)
    seeds = np.arange(B, dtype=np.int64) + 4242
    gc.reset_seeded(seeds)
    oc.reset_seeded(seeds)
    pin = lambda *shape, dt: torch.zeros(shape, dtype=dt).pin_memory()  # noqa: E731
    reward, done = pin(B, P, dt=torch.float32), pin(B, dt=torch.uint8)
    winner, err, aidx = pin(B, dt=torch.int8), pin(B, dt=torch.uint8), pin(B, P, dt=torch.int32)
    mask = np.zeros((B, P, gc.mask_words), np.uint32)          # pageable: staged
    acts_pinned = torch.zeros((B, gc.A, 8), dtype=torch.uint8).pin_memory()
    oo = oc.alloc_outputs_host()
    rng = np.random.default_rng(1)
    for t in range(25):
        acts = oc.sample_actions(31)
        corrupt_actions(rng, acts, W, H, 0.05)
        acts_pinned.copy_(torch.from_numpy(acts.view(np.uint8).reshape(B, gc.A, 8)))
        reward.fill_(-7.0)
        done.fill_(9)
        gc.step_fused(acts_pinned, gc.outputs(mask_bits=mask, reward=reward, done=done, winner=winner, step_error=err,
                                              action_index=aidx))
        oc.step_fused(acts, oc.outputs(**oo))
        assert np.array_equal(reward.numpy().view(np.uint32), oo["reward"].view(np.uint32)), f"turn {t}"
        assert np.array_equal(done.numpy(), oo["done"]) and np.array_equal(winner.numpy(), oo["winner"])
        assert np.array_equal(err.numpy(), oo["step_error"]) and np.array_equal(aidx.numpy(), oo["action_index"])
        assert np.array_equal(mask, oo["mask_bits"])
    assert np.array_equal(gc.state_hash(), oc.state_hash())


@pytest.mark.parametrize("W,H,P,B", [(15, 15, 2, 8202), (15, 15, 4, 8197), (15, 15, 3, 8200)])
def test_observation_run_writer_through_host_subranges(cuda_lib, oracle_lib, W, H, P, B):
    """15x15 observation tensors in HOST memory for a ragged batch >= 8192: the call is pipelined as sub-ranges, whole
    warps take the compile-time-scheduled run writer (four games per run, P == PT), the last, partial warp and P < PT the
    run-time-scheduled one; every float must equal the oracle's, block edges between sub-ranges and writers included."""
    gc, oc = new_engine(cuda_lib, W, H, P, B, host_threads=0), new_engine(oracle_lib, W, H, P, B, host_threads=0)
    seeds = np.arange(B, dtype=np.int64) + 99
    gc.reset_seeded(seeds)
    oc.reset_seeded(seeds)
    go, oo = gc.alloc_outputs_host(), oc.alloc_outputs_host()
    for t in range(8):
        go["obs"].fill(-3.0)
        acts = oc.sample_actions(5)
        gc.step_fused(acts, gc.outputs(**go))
        oc.step_fused(acts, oc.outputs(**oo))
        compare_outputs(go, oo, f"turn {t}")
    assert np.array_equal(gc.state_hash(), oc.state_hash())


@pytest.mark.parametrize("W,H,P,fog", [(20, 20, 2, 1), (15, 15, 2, 1), (10, 10, 2, 1), (20, 20, 4, 1), (7, 13, 3, 1), (15, 15, 2, 0)])
def test_packed_observation_records(cuda_lib, oracle_lib, W, H, P, fog):
    """grl_step_outputs.obs_packed (the host-delivery read-out): the kernel's records equal the oracle's word for word,
    and grl_expand_obs turns them into the very tensors the same launch wrote into `obs`."""
    B = 301   # leaves the last warp partly filled on every lane-group size
    gc = new_engine(cuda_lib, W, H, P, B, fog_of_war=fog)
    oc = new_engine(oracle_lib, W, H, P, B, fog_of_war=fog)
    seeds = np.arange(B, dtype=np.int64) + 4321
    gc.reset_seeded(seeds)
    oc.reset_seeded(seeds)
    go, oo = gc.alloc_outputs_host(), oc.alloc_outputs_host()
    gp, op = np.zeros((B, gc.packed_words), np.uint32), np.zeros((B, oc.packed_words), np.uint32)
    for t in range(45):
        gc.step_fused(None, gc.outputs(obs_packed=gp, **go), _abi.STEP_FLAG_RANDOM_POLICY, 5)
        oc.step_fused(None, oc.outputs(obs_packed=op, **oo), _abi.STEP_FLAG_RANDOM_POLICY, 5)
        assert np.array_equal(gp, op), f"packed records differ at turn {t}"
        if t % 9 == 0:
            assert np.array_equal(gc.expand_obs(gp).view(np.uint32), go["obs"].view(np.uint32)), t
    # packed records alone (no fp32 planes in the launch), from device memory too
    import torch
    dp = torch.zeros((B, gc.packed_words), dtype=torch.int32, device="cuda")
    gc.observe(gc.outputs(obs_packed=dp))
    gc.sync()
    assert np.array_equal(dp.cpu().numpy().view(np.uint32), op)
