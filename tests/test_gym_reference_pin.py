"""The generals_gym contract pinned against the reference's own client code.

tests/golden/gym_ref/*.npz hold outputs of the UNMODIFIED ``python/generals_gym/generals_env.py``
(``_get_observation``, ``_get_valid_actions_mask``, ``_action_index_to_game_action``,
``_calculate_reward`` and whole ``reset()``/``step()`` episodes), produced in the build container by
tests/tools/make_gym_fixtures.py, which imports the reference client behind a ``gymnasium`` stub and serves
it ``generals_pb`` ``GameState`` messages.  Nothing here reads /root/reference.

  readouts_*   given engine states (all planes, loaded with grl_set_state) -> for every player the
               client's observation tensor, N*5 mask, PlayerState numbers, and the decoding of EVERY
               action index
  episodes_*   whole episodes: the indices the agent and the opponent chose, and what ``step()``
               returned (observation, reward, terminated, truncated, invalid_action, next mask)

Tolerance: none.  float32 observations and float64 rewards are compared bit for bit.
CPU legs check the oracle; ``-m gpu`` legs check ``grl_gym_observe`` / ``grl_gym_encode`` /
``grl_gym_step`` of libgrlcuda.so against the same files.
"""
import glob
import os

import numpy as np
import pytest

from generalsreinforcementlearning_b200 import _abi
from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config
from test_gym_step import _planes

HERE = os.path.dirname(os.path.abspath(__file__))
GYM_REF = os.path.join(HERE, "golden", "gym_ref")
READOUTS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GYM_REF, "readouts_*.npz")))
EPISODES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GYM_REF, "episodes_*.npz")))


def test_fixture_inventory():
    """The committed set the generator writes (READOUT_CASES / EPISODE_CASES): nothing silently missing."""
    assert len(READOUTS) == 9 and len(EPISODES) == 8, (READOUTS, EPISODES)


def _device(lib):
    import torch

    return torch.device("cuda", 0) if lib.prefix == "grl_" else torch.device("cpu")


def _engine(lib, B, W, H, P, fog):
    e = BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, max_actions=max(2, P),
                                       host_threads=1, fog_of_war=int(fog)))
    if lib.prefix == "grl_":
        e.use_torch_stream()
    return e


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize]) if a.dtype.kind == "f" else a


def check_readouts(lib, name):
    import torch

    d = np.load(os.path.join(GYM_REF, name + ".npz"))
    W, H, P, fog, max_turns = (int(v) for v in d["meta"])
    N = W * H
    B = d["obs"].shape[0]
    dev = _device(lib)
    e = _engine(lib, B, W, H, P, fog)
    e.reset_seeded(np.arange(B, dtype=np.int64) + 1)
    e.set_state({k[len("state_"):]: d[k] for k in d.files if k.startswith("state_")})
    obs = torch.zeros((B, P, 9, H, W), dtype=torch.float32, device=dev)
    mask = torch.zeros((B, P, N * 5), dtype=torch.uint8, device=dev)
    stats = torch.zeros((B, P, 4), dtype=torch.int32, device=dev)
    e.gym_observe(max_turns, obs, mask, stats)
    e.sync()
    got_obs, got_mask, got_stats = obs.cpu().numpy(), mask.cpu().numpy(), stats.cpu().numpy()
    assert np.array_equal(_bits(got_obs), _bits(d["obs"])), f"{name}: _get_observation differs"
    assert np.array_equal(got_mask.astype(bool), d["mask"]), f"{name}: _get_valid_actions_mask differs"
    # PlayerState.army_count / tile_count / status as the client's reward reads them (generals_env.py:526-554)
    assert np.array_equal(got_stats[:, :, :3], d["stats"]), f"{name}: PlayerState numbers differ"
    assert got_obs.min() >= 0.0 and got_obs.max() <= 1.0            # the Box(0, 1) the client declares (:111-116)

    # _action_index_to_game_action for every index of every player's action space
    actions = torch.zeros((B, e.A, 8), dtype=torch.uint8, device=dev)
    valid = torch.zeros(B, dtype=torch.uint8, device=dev)
    want = d["decode"]                                               # [B][P][N*5][valid, fx, fy, tx, ty, half]
    for p in range(P):
        for a in range(N * 5):
            idx = torch.full((B,), a, dtype=torch.int64, device=dev)
            actions.zero_()
            e.gym_encode(idx, p, 0, mask, False, actions, valid)
            e.sync()
            rec = actions.cpu().numpy().view(_abi.ACTION_DTYPE)[:, 0, 0]
            v = valid.cpu().numpy()
            w = want[:, p, a]
            assert np.array_equal(v, w[:, 0].astype(np.uint8)), (name, p, a)
            assert np.array_equal(rec["present"], w[:, 0].astype(np.uint8)), (name, p, a)
            ok = w[:, 0] == 1
            for f, col in (("from_x", 1), ("from_y", 2), ("to_x", 3), ("to_y", 4)):
                assert np.array_equal(rec[f][ok], w[ok, col]), (name, p, a, f)
            assert np.array_equal(rec["move_all"][ok], 1 - w[ok, 5]), (name, p, a)     # MoveAll = !half
            assert (rec["player_id"][ok] == p).all()
    e.close()


def check_episodes(lib, name):
    import torch

    d = np.load(os.path.join(GYM_REF, name + ".npz"))
    W, H, P, fog, max_turns, n_ep = (int(v) for v in d["meta"])
    N, B = W * H, n_ep
    dev = _device(lib)
    e = _engine(lib, B, W, H, P, fog)
    ep = [{k[len(f"ep{b}_"):]: d[k] for k in d.files if k.startswith(f"ep{b}_")} for b in range(B)]
    if "seed" in ep[0]:
        e.reset_seeded(np.array([int(x["seed"][0]) for x in ep], np.int64))
        st = e.get_state()
        for b in range(B):       # the episode started from this very map
            for k in ("owner", "army", "type"):
                assert np.array_equal(st[k][b], ep[b][f"init_{k}"][0]), (name, b, k)
    else:
        e.reset_boards(*(np.concatenate([x[f"init_{k}"] for x in ep]) for k in ("owner", "army", "type")))
    pl = _planes(torch, dev, B, P, N, H, W)
    e.gym_observe(max_turns, pl["obs"], pl["mask"], pl["stats"])
    e.sync()
    T = [len(x["action"]) for x in ep]
    obs0, mask0 = pl["obs"].cpu().numpy(), pl["mask"].cpu().numpy()
    for b in range(B):           # reset(): the first observation and info["valid_actions_mask"]
        assert np.array_equal(_bits(obs0[b, 0]), _bits(ep[b]["obs"][0])), (name, b, "reset obs")
        assert np.array_equal(mask0[b, 0].astype(bool), ep[b]["mask"][0]), (name, b, "reset mask")
    seen = dict(invalid=0, terminated=0, truncated=0, steps=0, server_rejected=0)
    for t in range(max(T)):
        act = np.array([x["action"][t] if t < T[b] else -1 for b, x in enumerate(ep)], np.int64)
        opp = np.array([x["opp_action"][t] if t < T[b] else -1 for b, x in enumerate(ep)], np.int64)
        e.gym_step(max_turns, 0, action=torch.as_tensor(act, device=dev), opponent_action=torch.as_tensor(opp, device=dev),
                   **pl)
        e.sync()
        out = {k: pl[k].cpu().numpy() for k in ("obs", "mask", "reward", "terminated", "truncated", "valid")}
        for b in range(B):
            if t >= T[b]:
                continue
            x, ctx = ep[b], (name, b, t)
            assert out["valid"][b] == 1 - x["invalid"][t], ctx
            assert np.array_equal(_bits(out["reward"][b:b + 1]), _bits(x["reward"][t:t + 1])), \
                (ctx, out["reward"][b], x["reward"][t])
            assert out["terminated"][b] == x["terminated"][t], ctx
            # the class's own rule (:279) OR Gymnasium's TimeLimit on step() calls, which the registered env adds
            # (max_episode_steps, :607-611) and grl_gym_step folds in
            assert out["truncated"][b] == (x["truncated"][t] | (t + 1 >= max_turns)), ctx
            assert np.array_equal(_bits(out["obs"][b, 0]), _bits(x["obs"][t + 1])), (ctx, "obs")
            assert np.array_equal(out["mask"][b, 0].astype(bool), x["mask"][t + 1]), (ctx, "mask")
            seen["invalid"] += int(x["invalid"][t]); seen["terminated"] += int(x["terminated"][t])
            seen["truncated"] += int(x["truncated"][t]); seen["steps"] += 1
            seen["server_rejected"] += int(x["server_rejected"][t] != 0)
    e.close()
    return seen


@pytest.mark.parametrize("name", READOUTS)
def test_oracle_readouts_match_reference_client(oracle_lib, name):
    check_readouts(oracle_lib, name)


@pytest.mark.parametrize("name", EPISODES)
def test_oracle_episodes_match_reference_client(oracle_lib, name):
    seen = check_episodes(oracle_lib, name)
    assert seen["steps"] > 0
    if name == "episodes_5x5x2p":
        assert seen["terminated"] >= 3 and seen["invalid"] > 100
    if name == "episodes_10x10x2p_trunc":
        assert seen["truncated"] == 2


def test_fixtures_cover_the_awkward_cases():
    """What the fixtures were built to contain (so a regenerated set that lost a case is noticed)."""
    tot = dict(server_rejected=0, bonus=0, win=0, loss=0)
    for name in EPISODES:
        d = np.load(os.path.join(GYM_REF, name + ".npz"))
        for b in range(int(d["meta"][5])):
            r = d[f"ep{b}_reward"]
            tot["server_rejected"] += int((d[f"ep{b}_server_rejected"] != 0).sum())
            tot["bonus"] += int(((r > 40) & (r < 60)).sum())     # +50: an opponent left while the game went on
            tot["win"] += int((r == 100.0).sum())
            tot["loss"] += int((r == -100.0).sum())
    assert tot["server_rejected"] >= 5 and tot["bonus"] >= 2 and tot["win"] >= 3 and tot["loss"] >= 1, tot


@pytest.mark.gpu
@pytest.mark.parametrize("name", READOUTS)
def test_cuda_readouts_match_reference_client(cuda_lib, name):
    check_readouts(cuda_lib, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", EPISODES)
def test_cuda_episodes_match_reference_client(cuda_lib, name):
    check_episodes(cuda_lib, name)
