"""Callers either side of the turn path: the generals_gym read-outs and vector
env, the SKIP_ENV turn barrier, the replay/state codec and the ASCII renderer.

CPU legs run the host logic against the oracle library (same ABI); GPU legs run the CUDA
read-outs against the oracle and the vector env on the device."""
import math
import os

import numpy as np
import pytest

from generalsreinforcementlearning_b200 import _abi, render, replay
from generalsreinforcementlearning_b200.engine import make_actions, set_action
from helpers import CITY, GENERAL, MOUNTAIN, NORMAL, blank_state, full_fog, full_stats, new_engine, put


# ---- a literal numpy restatement of the reference CLIENT code, from the proto view ------------
def proto_view(st, vis, fog, c, p):
    """server.go:556-582 for one env c and player p."""
    N = st["owner"].shape[1]
    view = []
    for i in range(N):
        t = dict(type=int(st["type"][c, i]), owner=int(st["owner"][c, i]), army=int(st["army"][c, i]),
                 visible=bool(vis[c, p, i]), fog=bool(fog[c, p, i]))
        if not t["visible"] and not t["fog"]:
            t.update(type=NORMAL, owner=-1, army=0)
        elif t["fog"] and not t["visible"]:
            t.update(owner=-1, army=0)
        view.append(t)
    return view


def client_observation(view, W, H, player_id, turn_count, max_turns):
    """generals_env.py:291-342 _get_observation."""
    obs = np.zeros((9, H, W), np.float32)
    for y in range(H):
        for x in range(W):
            tile = view[y * W + x]
            if tile["visible"]:
                obs[0, y, x] = 1.0
            if tile["owner"] == player_id:
                obs[1, y, x] = 0.5
            elif tile["owner"] >= 0:
                obs[1, y, x] = 1.0
            if tile["army"] > 0:
                obs[2, y, x] = np.log(tile["army"] + 1) / 10.0
            obs[{NORMAL: 3, MOUNTAIN: 4, CITY: 5, GENERAL: 6}[tile["type"]], y, x] = 1.0
    obs[7, :, :] = min(turn_count / max_turns, 1.0)
    return obs


def client_mask(view, W, H, player_id):
    """generals_env.py:344-387 _get_valid_actions_mask."""
    mask = np.zeros(W * H * 5, bool)
    for y in range(H):
        for x in range(W):
            idx = y * W + x
            if view[idx]["owner"] != player_id or view[idx]["army"] <= 1:
                continue
            for d, (dx, dy) in enumerate([(0, -1), (1, 0), (0, 1), (-1, 0)]):
                nx, ny = x + dx, y + dy
                if 0 <= nx < W and 0 <= ny < H and view[ny * W + nx]["type"] != MOUNTAIN:
                    mask[idx * 5 + d] = True
                    mask[idx * 5 + 4] = True
    return mask


def gym_readouts(e, max_turns):
    obs = np.zeros((e.B, e.P, 9, e.H, e.W), np.float32)
    mask = np.zeros((e.B, e.P, e.N * 5), np.uint8)
    stats = np.zeros((e.B, e.P, 4), np.int32)
    e.gym_observe(max_turns, obs, mask, stats)
    return obs, mask, stats


def _played(lib, W, H, P, B, turns, seed=12345):
    e = new_engine(lib, W, H, P, B)
    e.reset_seeded(np.arange(B, dtype=np.int64) + seed)
    for _ in range(turns):
        e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 5)
    return e


@pytest.mark.parametrize("W,H,P,turns", [(8, 8, 2, 0), (8, 8, 2, 37), (10, 10, 2, 120), (9, 7, 3, 60)])
def test_oracle_gym_readouts_match_the_reference_client(oracle_lib, W, H, P, turns):
    B, max_turns = 6, 100
    e = _played(oracle_lib, W, H, P, B, turns)
    obs, mask, stats = gym_readouts(e, max_turns)
    st = e.get_state()
    vis, fog = e.visibility()
    for c in range(B):
        for p in range(P):
            view = proto_view(st, vis, fog, c, p)
            want = client_observation(view, W, H, p, int(st["turn"][c]), max_turns)
            assert np.array_equal(obs[c, p].view(np.uint32), want.view(np.uint32)), (c, p)
            assert np.array_equal(mask[c, p].astype(bool), client_mask(view, W, H, p)), (c, p)
            assert stats[c, p, 0] == st["army_count"][c, p] and stats[c, p, 1] == st["owned"][c, p].sum()
            assert stats[c, p, 2] == st["alive"][c, p] and stats[c, p, 3] == st["general_idx"][c, p]
    assert obs.min() >= 0.0 and obs.max() <= 1.0


def _skip_roundtrip(lib):
    """SKIP_ENV leaves an env untouched while its neighbours take their turn."""
    B = 8
    e = new_engine(lib, 10, 10, 2, B)
    ref = new_engine(lib, 10, 10, 2, B)
    seeds = np.arange(B, dtype=np.int64) + 7
    e.reset_seeded(seeds)
    ref.reset_seeded(seeds)
    out, rout = e.alloc_outputs_host(), ref.alloc_outputs_host()
    rng = np.random.default_rng(3)
    lag = np.zeros(B, int)  # turns each env of `e` is behind `ref`
    history = []
    for t in range(30):
        acts = ref.sample_actions(11)
        history.append(acts.copy())
        ref.step_fused(acts, ref.outputs(**rout))
        # e replays ref's action stream, but a random subset of envs sits out each call
        mine = make_actions(B, e.A)
        skip = rng.random(B) < 0.3
        for b in range(B):
            turn_b = t - lag[b]
            mine[b] = history[turn_b][b]
            if skip[b]:
                mine[b, 0]["flags"] = _abi.ACTION_FLAG_SKIP_ENV
                lag[b] += 1
        before = e.state_hash().copy()
        e.step_fused(mine, e.outputs(**out))
        after = e.state_hash()
        assert np.array_equal(after[skip], before[skip]), "a skipped env must not change"
        assert (after[~skip] != before[~skip]).all()
    # every env followed ref's trajectory, just later: compare against a replay of its own length
    chk = new_engine(lib, 10, 10, 2, B)
    chk.reset_seeded(seeds)
    for t in range(30):
        acts = history[t].copy()
        for b in range(B):
            if t >= 30 - lag[b]:
                acts[b, 0]["flags"] = _abi.ACTION_FLAG_SKIP_ENV
        chk.step(acts)
    assert np.array_equal(chk.state_hash(), e.state_hash())
    return e


def test_skip_env_turn_barrier_oracle(oracle_lib):
    _skip_roundtrip(oracle_lib)


def test_replay_and_state_codec_roundtrip(oracle_lib, tmp_path):
    B = 12
    e = new_engine(oracle_lib, 10, 10, 2, B)
    seeds = np.arange(B, dtype=np.int64) + 99
    e.reset_seeded(seeds)
    w = replay.ReplayWriter(e, seeds, digest_every=10)
    for t in range(45):
        if t % 9 == 4:   # a step taken with the in-kernel policy is recorded by its flags and seed alone
            e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 1000 + t)
            w.record(None, _abi.STEP_FLAG_RANDOM_POLICY, 1000 + t)
            continue
        a = e.sample_actions(4)
        if t % 7 == 3:
            a[t % B, 0]["flags"] = _abi.ACTION_FLAG_SKIP_ENV
        e.step(a)
        w.record(a)
    path = str(tmp_path / "run.grlreplay")
    w.save(path)
    meta, actions = replay.load_replay(path)
    assert meta["turns"] == 45 and actions.shape == (45, B, e.A, 8)
    assert meta["abi"] == oracle_lib.abi_version() == 3 and meta["config"]["normal_growth_interval"] == 25
    e2 = new_engine(oracle_lib, 10, 10, 2, B)
    assert replay.replay(e2, path) == 45
    assert np.array_equal(e2.state_hash(), e.state_hash())
    # an engine with other rules is a configuration mismatch, not a divergence
    for other in (dict(fog_of_war=0), dict(normal_growth_interval=10), dict(env_id_base=5), dict(production_city=2)):
        with pytest.raises(ValueError, match="different configuration"):
            replay.replay(new_engine(oracle_lib, 10, 10, 2, B, **other), path)
    # a recording of an OLDER ABI revision replays (revisions only add), one of a newer revision is refused
    import io, json, zipfile

    def rewrite(name, **changes):
        out = str(tmp_path / name)
        buf = io.BytesIO()
        np.save(buf, actions)
        with zipfile.ZipFile(out, "w") as z:
            z.writestr("meta.json", json.dumps(dict(meta, **changes)))
            z.writestr("actions.npy", buf.getvalue())
        return out

    assert replay.replay(new_engine(oracle_lib, 10, 10, 2, B), rewrite("older.grlreplay", abi=meta["abi"] - 1)) == 45
    with pytest.raises(ValueError, match="ABI version"):
        replay.replay(new_engine(oracle_lib, 10, 10, 2, B), rewrite("newer.grlreplay", abi=meta["abi"] + 1))
    # a tampered action stream is caught by the recorded digests
    actions[5, 0, 0, 3] ^= 1
    bad = str(tmp_path / "bad.grlreplay")
    buf = io.BytesIO()
    np.save(buf, actions)
    with zipfile.ZipFile(bad, "w") as z:
        z.writestr("meta.json", json.dumps(meta))
        z.writestr("actions.npy", buf.getvalue())
    with pytest.raises(AssertionError):
        replay.replay(new_engine(oracle_lib, 10, 10, 2, B), bad)
    # checkpoint: every plane round-trips, and the restored engine continues identically
    ck = str(tmp_path / "ck.npz")
    replay.save_state(e, ck)
    e3 = new_engine(oracle_lib, 10, 10, 2, B)
    e3.reset_seeded(seeds + 1000)
    replay.load_state(e3, ck)
    assert np.array_equal(e3.state_hash(), e.state_hash())
    for _ in range(10):
        a = e.sample_actions(4)
        e.step(a)
        e3.step(a)
    assert np.array_equal(e3.state_hash(), e.state_hash())


def test_ascii_board_matches_rendering_go(oracle_lib):
    """Engine.Board (rendering.go:34-143): a hand-built 4x3 board, rendered for player 0 and unfogged."""
    W, H, P = 4, 3, 2
    e = new_engine(oracle_lib, W, H, P)
    s = blank_state(W, H, P)
    put(s, W, 0, 0, 0, 7, GENERAL)
    put(s, W, 1, 0, 0, 12, NORMAL)
    put(s, W, 2, 0, -1, 0, MOUNTAIN)
    put(s, W, 3, 0, -1, 40, CITY)
    put(s, W, 0, 1, 0, 150, NORMAL)
    put(s, W, 1, 1, -1, 0, NORMAL)
    put(s, W, 3, 2, 1, 3, GENERAL)
    put(s, W, 2, 2, 1, 55, CITY)
    full_stats(s)
    full_fog(s, W, H)
    e.set_state(s)
    R, G, B_, Wh, X = "\033[31m", "\033[90m", "\033[34m", "\033[37m", "\033[0m"
    txt = render.render_board(e, 0, player_id=-1)
    lines = txt.split("\n")
    assert lines[0] == "     0 1 2 3"
    # an owned normal tile with army 10..99 prints BOTH digits: fmt "%*d" with width 1 (rendering.go:132)
    assert lines[1] == f" 0 {R}A♔{X} {R}A12{X} {G} ▲{X} {Wh} ⬢{X} "
    assert lines[2] == f" 1 {R}A+{X} {G} ·{X} {G} ·{X} {G} ·{X} "
    assert lines[3] == f" 2 {G} ·{X} {G} ·{X} {B_}B⬢{X} {B_}B♔{X} "
    assert lines[5] == "·=empty ⬢=city ♔=general ▲=mountain A-H=players"
    fogged = render.render_board(e, 0, player_id=0).split("\n")
    assert fogged[1] == f" 0 {R}A♔{X} {R}A12{X} {G} ▲{X} {G} {X} "        # (3,0) is outside player 0's 3x3 reach
    assert fogged[3] == f" 2 {G} ·{X} {G} ·{X} {G} {X} {G} {X} "


def test_vector_env_contract_on_the_oracle(oracle_lib):
    """GeneralsVecEnv over the oracle library (host tensors): spaces, masks, invalid actions, auto-reset."""
    import torch

    from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

    B = 16
    env = GeneralsVecEnv(B, 8, 8, max_turns=25, lib=oracle_lib, host_threads=1, seed=5)
    obs, info = env.reset()
    assert tuple(obs.shape) == (B, 9, 8, 8) and obs.dtype == torch.float32
    assert env.single_action_space.n == 8 * 8 * 5
    mask = info["valid_actions_mask"]
    assert mask.shape == (B, 320) and mask.any(1).all()
    g = torch.Generator().manual_seed(0)
    total_reward = torch.zeros(B, dtype=torch.float64)
    saw_reset = False
    for t in range(60):
        probs = mask.to(torch.float32)
        action = torch.multinomial(probs, 1, generator=g).squeeze(1)
        if t == 3:
            action[0] = int((~mask[0]).nonzero()[0])       # an invalid action for env 0
        turns_before = info["turn"].clone()
        obs, reward, term, trunc, info = env.step(action)
        if t == 3:
            assert bool(info["invalid_action"][0]) and float(reward[0]) == -0.1
            assert int(info["turn"][0]) == int(turns_before[0]), "an invalid action takes no turn"
            assert not bool(info["invalid_action"][1:].any())
        total_reward += reward
        mask = info["valid_actions_mask"]
        assert float(obs.min()) >= 0.0 and float(obs.max()) <= 1.0
        if bool((term | trunc).any()):
            saw_reset = True
            ids = info["final_env_ids"]
            assert (env._turns[ids] == 0).all(), "finished envs are re-seeded"
            # env 0 lost one call to the rejected action: it is cut after 25 step() calls with 24 turns played
            assert (info["turn"][ids] >= 24).all() or bool(term[ids].any())
            assert info["final_observation"].shape[0] == len(ids)
    assert saw_reset, "max_turns=25 must truncate within 60 steps"
    assert torch.isfinite(total_reward).all()
    env.close()


# ---------------------------------------------------------------------------------- GPU legs
@pytest.mark.gpu
@pytest.mark.parametrize("W,H,P,B,turns", [(10, 10, 2, 256, 60), (15, 15, 2, 128, 90), (20, 20, 4, 64, 120), (7, 13, 3, 32, 40)])
def test_cuda_gym_readouts_match_oracle(cuda_lib, oracle_lib, W, H, P, B, turns):
    g, o = _played(cuda_lib, W, H, P, B, turns), _played(oracle_lib, W, H, P, B, turns)
    assert np.array_equal(g.state_hash(), o.state_hash())
    for a, b, name in zip(gym_readouts(g, 500), gym_readouts(o, 500), ("obs", "mask", "stats")):
        if a.dtype == np.float32:
            a, b = a.view(np.uint32), b.view(np.uint32)
        assert np.array_equal(a, b), name


@pytest.mark.gpu
def test_skip_env_turn_barrier_cuda(cuda_lib, oracle_lib):
    g = _skip_roundtrip(cuda_lib)
    o = _skip_roundtrip(oracle_lib)
    assert np.array_equal(g.state_hash(), o.state_hash())
    for k, v in g.get_state().items():
        assert np.array_equal(v, o.get_state()[k]), k


@pytest.mark.gpu
def test_vector_env_on_device(cuda_lib):
    """4,096 games stepped through the gym contract on the GPU; tensors never leave the device."""
    import torch

    from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

    B = 4096
    env = GeneralsVecEnv(B, 15, 15, max_turns=40, seed=21)
    obs, info = env.reset()
    assert obs.is_cuda and obs.shape == (B, 9, 15, 15)
    g = torch.Generator(device=obs.device).manual_seed(1)
    episodes = 0
    for t in range(50):
        action = torch.multinomial(info["valid_actions_mask"].to(torch.float32), 1, generator=g).squeeze(1)
        obs, reward, term, trunc, info = env.step(action)
        assert reward.is_cuda and not bool(info["invalid_action"].any())
        episodes += int((term | trunc).sum())
    assert episodes >= B, "every env hits max_turns=40 once within 50 steps"
    # the same seeds replayed give the same observations (determinism of the device path)
    env2 = GeneralsVecEnv(B, 15, 15, max_turns=40, seed=21)
    o2, i2 = env2.reset()
    assert torch.equal(o2, GeneralsVecEnv.reset(env, seed=21)[0])
    env.close()
    env2.close()


def test_go_rand_and_demo_policy(oracle_lib):
    """GoRand (python) == the oracle's C generator == the canonical Go outputs; the demo policy
    of demo_helpers.go:12-62 only ever proposes moves the engine accepts."""
    import ctypes as C

    from generalsreinforcementlearning_b200.demo_policy import GoRand, demo_actions_for
    from helpers import ctypes_fn

    r = GoRand(1)
    assert [r.int63() for _ in range(3)] == [5577006791947779410, 8674665223082153551, 6129484611666145821]
    r = GoRand(1)
    assert [r.intn(100) for _ in range(10)] == [81, 87, 47, 59, 81, 18, 25, 40, 56, 0]
    fn = ctypes_fn(oracle_lib, "test_gorand", C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p])
    for seed in (12345, 42, -7, 0, (1 << 31) + 5):
        out = np.zeros(50, np.int64)
        assert fn(seed, 0, 0, 50, out.ctypes.data) == 0
        r = GoRand(seed)
        assert [r.int63() for _ in range(50)] == out.tolist(), seed
        assert fn(seed, 1, 37, 50, out.ctypes.data) == 0
        r = GoRand(seed)
        assert [r.intn(37) for _ in range(50)] == out.tolist(), seed
    r = GoRand(1)  # rand.New(rand.NewSource(1)).Float64() / Float32(): the well-known first values
    assert abs(r.float64() - 0.6046602879796196) < 1e-16
    r = GoRand(1)
    assert abs(r.float32() - 0.6046603) < 1e-7
    e = new_engine(oracle_lib, 10, 10, 2, 1)
    e.reset_seeded([12345])
    rng = GoRand(99)
    played = 0
    for _ in range(200):
        acts = demo_actions_for(e, 0, rng)
        played += int(acts["present"].sum())
        before = e.get_state()["turn"][0]
        e.step(acts)
        st = e.get_state()
        assert st["turn"][0] == before + 1
        # a move legal at submission can only be invalidated by the OTHER player's move of the same turn
        assert st["step_error"][0] in (0, _abi.STEP_NOT_OWNED, _abi.STEP_INSUFFICIENT_ARMY)
    assert 60 < played < 200, "each player moves with probability 0.3 per turn"


def _python_decode(a, W, H):
    """generals_env.py:389-441 for a valid index."""
    frm, info = a // 5, a % 5
    fx, fy = frm % W, frm // W
    dirs = [(0, -1), (1, 0), (0, 1), (-1, 0)]
    if info < 4:
        tx, ty = fx + dirs[info][0], fy + dirs[info][1]
    else:
        for dx, dy in dirs:
            tx, ty = fx + dx, fy + dy
            if 0 <= tx < W and 0 <= ty < H:
                break
    return fx, fy, tx, ty, info != 4


def _gym_encode_case(lib, to_dev):
    W, H, P, B = 7, 6, 2, 64
    e = _played(lib, W, H, P, B, 25)
    obs, mask, stats = gym_readouts(e, 100)
    rng = np.random.default_rng(9)
    idx = rng.integers(-3, W * H * 5 + 3, B).astype(np.int64)
    for b in range(0, B, 2):  # half of the envs get a valid index (if they have one)
        ok = np.nonzero(mask[b, 0])[0]
        if len(ok):
            idx[b] = ok[rng.integers(len(ok))]
    acts = make_actions(B, e.A)
    acts["player_id"][:, 1], acts["present"][:, 1] = 1, 1   # a pre-filled opponent slot must survive
    valid = np.zeros(B, np.uint8)
    d_idx, d_mask, d_acts, d_valid = to_dev(idx), to_dev(mask), to_dev(acts.view(np.uint8).reshape(B, e.A, 8)), to_dev(valid)
    e.gym_encode(d_idx, 0, 0, d_mask, True, d_acts, d_valid)
    acts = np.asarray(d_acts.cpu() if hasattr(d_acts, "cpu") else d_acts).reshape(B, e.A, 8).view(_abi.ACTION_DTYPE).reshape(B, e.A)
    valid = np.asarray(d_valid.cpu() if hasattr(d_valid, "cpu") else d_valid)
    assert valid.sum() > 0 and (valid == 0).sum() > 0
    for b in range(B):
        a = int(idx[b])
        ok = 0 <= a < W * H * 5 and bool(mask[b, 0, a])
        assert bool(valid[b]) == ok
        r = acts[b, 0]
        if ok:
            fx, fy, tx, ty, mv = _python_decode(a, W, H)
            assert (r["player_id"], r["from_x"], r["from_y"], r["to_x"], r["to_y"], r["move_all"], r["present"], r["flags"]) == (
                0, fx, fy, tx, ty, int(mv), 1, 0)
        else:
            assert r["present"] == 0 and r["flags"] == _abi.ACTION_FLAG_SKIP_ENV
        assert acts[b, 1]["present"] == 1 and acts[b, 1]["player_id"] == 1


def test_gym_encode_oracle(oracle_lib):
    _gym_encode_case(oracle_lib, lambda a: a)


@pytest.mark.gpu
def test_gym_encode_cuda(cuda_lib):
    import torch

    _gym_encode_case(cuda_lib, lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())
