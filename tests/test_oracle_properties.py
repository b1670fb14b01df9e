"""Property tests of the oracle (hypothesis): invariants of the turn engine that hold for ANY
board and ANY action stream — the reference's tests check them on hand-picked cases only
(core/movement_test.go, engine_test.go).  Boards are arbitrary (not mapgen output)."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from generalsreinforcementlearning_b200 import _abi
from generalsreinforcementlearning_b200.engine import make_actions, set_action
from helpers import CITY, GENERAL, MOUNTAIN, NORMAL, blank_state, full_fog, full_stats, new_engine


@st.composite
def boards(draw):
    W, H, P = draw(st.integers(2, 9)), draw(st.integers(2, 9)), draw(st.integers(2, 4))
    N = W * H
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    s = blank_state(W, H, P)
    s["owner"][0] = rng.integers(-1, P, N)
    s["type"][0] = rng.choice([NORMAL, NORMAL, NORMAL, CITY, MOUNTAIN], N)
    s["army"][0] = rng.integers(0, 80, N)
    for p in range(P):
        i = int(rng.integers(0, N))
        s["owner"][0, i], s["type"][0, i], s["army"][0, i] = p, GENERAL, int(rng.integers(1, 9))
    m = s["type"][0] == MOUNTAIN
    s["owner"][0][m], s["army"][0][m] = -1, 0
    full_stats(s)
    full_fog(s, W, H)
    return W, H, P, s, seed


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(boards(), st.integers(1, 25))
def test_engine_invariants_under_arbitrary_play(oracle_lib, board, turns):
    W, H, P, s, seed = board
    e = new_engine(oracle_lib, W, H, P, max_actions=P)
    e.set_state(s)
    rng = np.random.default_rng(seed + 1)
    N = W * H
    mountains = s["type"][0] == MOUNTAIN
    for t in range(turns):
        before = e.get_state()
        acts = make_actions(1, P)
        for p in range(P):
            if rng.random() < 0.8:  # mostly plausible moves, sometimes nonsense
                fx, fy = int(rng.integers(-1, W + 1)), int(rng.integers(-1, H + 1))
                dx, dy = [(0, 1), (0, -1), (1, 0), (-1, 0), (0, 0), (1, 1)][int(rng.integers(0, 6))]
                set_action(acts, 0, p, p, fx, fy, fx + dx, fy + dy, bool(rng.integers(2)))
        e.step(acts)
        st_ = e.get_state()
        if before["game_over"][0]:
            assert st_["step_error"][0] == _abi.STEP_GAME_OVER
            for k in ("owner", "army", "turn", "visible"):
                assert np.array_equal(st_[k], before[k]), "a finished game never changes"
            continue
        assert st_["turn"][0] == before["turn"][0] + 1
        assert (st_["army"][0] >= 0).all()
        assert (st_["owner"][0][mountains] == -1).all() and (st_["army"][0][mountains] == 0).all()
        assert np.array_equal(st_["type"], before["type"]), "terrain never changes (movement.go:38)"
        assert ((st_["owner"][0] >= -1) & (st_["owner"][0] < P)).all()
        # a tile sits in at most one cached list, and cached lists never contain mountains
        assert (st_["owned"][0].sum(axis=0) <= 1).all()
        assert not st_["owned"][0][:, mountains].any()
        # armies only appear through production: total army grows by at most one per producing tile
        grown = int(st_["army"][0].sum()) - int(before["army"][0].sum())
        assert grown <= N, "no move creates armies"
        if st_["step_error"][0] != 0:
            assert grown <= 0, "an aborted turn skips production (engine.go:111-113), moves only destroy armies"
        # ArmyCount is the sum over the cached list; Alive <=> the list holds a general-type tile
        for p in range(P):
            lst = st_["owned"][0, p].astype(bool)
            if st_["step_error"][0] == 0 and st_["changed"][0].any():
                assert st_["army_count"][0, p] == int(st_["army"][0][lst].sum())
                assert (st_["owner"][0][lst] == p).all(), "a rebuilt list holds only tiles the player owns"
            gens = lst & (st_["type"][0] == GENERAL)
            if st_["step_error"][0] == 0 and st_["changed"][0].any():
                assert bool(st_["alive"][0, p]) == bool(gens.any())
        alive = int(st_["alive"][0].sum())
        if st_["step_error"][0] == 0:
            assert bool(st_["game_over"][0]) == (alive <= 1)
            assert st_["winner"][0] == (int(np.argmax(st_["alive"][0])) if (alive == 1 and st_["game_over"][0]) else -1)


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 60), st.integers(0, 60), st.booleans(), st.sampled_from([-1, 0, 1]))
def test_single_move_arithmetic(oracle_lib, a_from, a_to, move_all, to_owner):
    """core.ApplyMoveAction (movement.go:23-89) for every army pair: conservation on friendly moves,
    larger-minus-smaller on attacks, ties defend."""
    e = new_engine(oracle_lib, 3, 1, 2)
    s = blank_state(3, 1, 2)
    s["owner"][0] = [0, to_owner, 1]
    s["army"][0] = [a_from, a_to, 5]
    s["type"][0] = [NORMAL, NORMAL, GENERAL]
    s["alive"][0] = [1, 1]
    e.set_state(s)
    acts = make_actions(1, 2)
    set_action(acts, 0, 0, 0, 0, 0, 1, 0, move_all)
    e.step(acts)
    out = e.get_state()
    moved = a_from - 1 if move_all else max(1, a_from // 2)
    assert out["army"][0, 0] == a_from - moved
    if to_owner == 0:
        assert (out["owner"][0, 1], out["army"][0, 1]) == (0, a_to + moved)
    elif moved > a_to:
        assert (out["owner"][0, 1], out["army"][0, 1]) == (0, moved - a_to)
    else:
        assert (out["owner"][0, 1], out["army"][0, 1]) == (to_owner, a_to - moved)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(boards(), st.integers(1, 12), st.integers(0, 2**31 - 1))
def test_gym_step_invariants_on_arbitrary_boards(oracle_lib, board, steps, aseed):
    """The gym contract on ANY board (python/generals_gym/generals_env.py): observations stay in [0, 1], exactly one type
    plane is set per tile, the mask agrees with what the env accepts (an action the mask allows is never rejected, one it
    forbids always is and leaves the game untouched at a cost of -0.1), `any` == OR of the four directions, and a
    sampled action is always a masked-in one."""
    W, H, P, s, seed = board
    N = W * H
    e = new_engine(oracle_lib, W, H, P, max_actions=P)
    e.set_state(s)
    z = lambda shape, dt: np.zeros(shape, dt)
    pl = dict(obs=z((1, P, 9, H, W), np.float32), mask=z((1, P, N * 5), np.uint8), stats=z((1, P, 4), np.int32),
              actions=z((1, P, 8), np.uint8), prev_stats=z((1, P, 4), np.int32), turns=z(1, np.int32), calls=z(1, np.int32),
              reward=z(1, np.float64), terminated=z(1, np.uint8), truncated=z(1, np.uint8), valid=z(1, np.uint8),
              done=z(1, np.uint8), winner=z(1, np.int8), step_error=z(1, np.uint8), n_finished=z(1, np.int32))
    e.gym_observe(500, pl["obs"], pl["mask"], pl["stats"])
    rng = np.random.default_rng(aseed)
    for t in range(steps):
        obs, mask = pl["obs"][0], pl["mask"][0]
        assert obs.min() >= 0.0 and obs.max() <= 1.0
        assert np.array_equal(obs[:, 3:7].sum(1), np.ones((P, H, W), np.float32)), "one type plane per tile"
        m5 = mask.reshape(P, N, 5)
        assert np.array_equal(m5[:, :, 4], m5[:, :, :4].max(2)), "half-move flag == any direction"
        sampled = z(1, np.int64)
        e.gym_sample(int(rng.integers(0, 2**62)), pl["mask"], 0, sampled)
        if mask[0].any():
            assert mask[0, sampled[0]] == 1, "sampled actions are valid"
        forbid = rng.random() < 0.3 and not mask[0].all()
        a = int(rng.choice(np.flatnonzero(mask[0] == 0))) if forbid else int(sampled[0])
        before = e.state_hash().copy()
        accepted = bool(mask[0, a]) if 0 <= a < N * 5 else False   # the mask plane is rewritten by the step
        e.gym_step(500, int(rng.integers(0, 2**62)), action=np.array([a], np.int64), opponent_action=None, **pl)
        assert bool(pl["valid"][0]) == accepted
        if not accepted:
            assert pl["reward"][0] == -0.1 and np.array_equal(before, e.state_hash()), "a rejected action takes no turn"
            assert not pl["terminated"][0]
        if pl["terminated"][0]:
            assert pl["done"][0] == 1 and abs(pl["reward"][0]) == 100.0
            break
    e.close()
