"""Pins the CPU oracle against the reference's own known-answer tests and seeded goldens
(SURVEY Appendix B/C).  CPU only."""
import ctypes as C

import numpy as np
import pytest

import kats
from helpers import CITY, GENERAL, MOUNTAIN, NORMAL, blank_state, ctypes_fn, new_engine, put

KATS = [getattr(kats, n) for n in sorted(dir(kats)) if n.startswith("kat_")]


@pytest.mark.parametrize("kat", KATS, ids=lambda f: f.__name__)
def test_kat(oracle_lib, kat):
    kat(oracle_lib)


# ---- Go math/rand ----------------------------------------------------
def _draw(lib, seed, kind, n, count):
    fn = ctypes_fn(lib, "test_gorand", C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p])
    out = np.zeros(count, np.int64)
    assert fn(seed, kind, n, count, out.ctypes.data) == 0
    return out.tolist()


def test_gorand_canonical_values(oracle_lib):
    # rand.New(rand.NewSource(1)): the canonical Go outputs
    assert _draw(oracle_lib, 1, 0, 0, 3) == [5577006791947779410, 8674665223082153551, 6129484611666145821]
    assert _draw(oracle_lib, 1, 1, 100, 10) == [81, 87, 47, 59, 81, 18, 25, 40, 56, 0]
    assert _draw(oracle_lib, 12345, 0, 0, 3) == [7828158075477027098, 5950071357434416446, 6808766918387264829]
    assert _draw(oracle_lib, 12345, 1, 20, 8) == [3, 3, 4, 16, 1, 15, 2, 6]
    assert _draw(oracle_lib, 42, 1, 10, 8) == [5, 7, 8, 0, 3, 5, 7, 6]


def test_gorand_seed_normalisation(oracle_lib):
    # rng.go Seed: seed %= 2^31-1; negative += 2^31-1; 0 -> 89482311
    m = (1 << 31) - 1
    assert _draw(oracle_lib, 5, 0, 0, 4) == _draw(oracle_lib, 5 + m, 0, 0, 4)
    assert _draw(oracle_lib, -3, 0, 0, 4) == _draw(oracle_lib, m - 3, 0, 0, 4)
    assert _draw(oracle_lib, 0, 0, 0, 4) == _draw(oracle_lib, 89482311, 0, 0, 4)


# ---- mapgen seeded goldens (internal/game/mapgen/generator_test.go) ------------------------
def _mountains(lib, W, H, veins, lo, hi, seed=12345):
    fn = ctypes_fn(lib, "test_place_mountains", C.c_int,
                   [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p])
    t = np.zeros(W * H, np.int32)
    assert fn(W, H, veins, lo, hi, seed, t.ctypes.data) == 0
    return t


def test_mapgen_mountain_goldens(oracle_lib):
    assert (_mountains(oracle_lib, 20, 20, 5, 3, 5) == MOUNTAIN).sum() == 22   # generator_test.go:84
    assert (_mountains(oracle_lib, 10, 10, 0, 3, 2) == MOUNTAIN).sum() == 0    # :103
    assert (_mountains(oracle_lib, 30, 30, 1, 5, 5) == MOUNTAIN).sum() == 5    # :124
    assert (_mountains(oracle_lib, 3, 3, 10, 1, 1) == MOUNTAIN).sum() == 9     # :148


def test_mapgen_city_golden(oracle_lib):
    fn = ctypes_fn(oracle_lib, "test_place_cities", C.c_int,
                   [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p])
    t = np.zeros(400, np.int32)
    a = np.zeros(400, np.int32)
    assert fn(20, 20, 20, 50, 12345, t.ctypes.data, a.ctypes.data) == 0
    assert (t == CITY).sum() == 20                                             # :175
    assert (a[t == CITY] == 50).all()
    # generator_test.go:178-194 NoCitiesIfRatioIsTooHigh: 100 / 101 = 0 cities
    t, a = np.zeros(100, np.int32), np.zeros(100, np.int32)
    assert fn(10, 10, 101, 40, 12345, t.ctypes.data, a.ctypes.data) == 0 and (t == CITY).sum() == 0
    # :196-218 CitiesNotOnPreExistingMountains: ratio 5 asks for 20 cities; (1,1) and (5,5) stay mountains
    t, a = np.zeros(100, np.int32), np.zeros(100, np.int32)
    t[1 * 10 + 1] = t[5 * 10 + 5] = MOUNTAIN
    assert fn(10, 10, 5, 40, 12345, t.ctypes.data, a.ctypes.data) == 0
    assert t[11] == MOUNTAIN and t[55] == MOUNTAIN and (t == CITY).sum() == 20 and (a[t == MOUNTAIN] == 0).all()
    # :220-244 MaxAttemptsForCitiesWhenNoSpace: a board full of mountains takes no city and the loop ends
    t, a = np.full(25, MOUNTAIN, np.int32), np.zeros(25, np.int32)
    assert fn(5, 5, 1, 40, 12345, t.ctypes.data, a.ctypes.data) == 0 and (t == MOUNTAIN).all()


def test_default_map_config(oracle_lib):
    """generator_test.go:18-34 TestDefaultMapConfig: ratio 20, city army 40, spacing 5 (grl_default_config carries
    them); veins (w*h)/50, vein length 3 .. w/4 are what both generators derive — checked through a 20x15 map made with
    the explicit values against the default path, which must agree tile for tile."""
    from generalsreinforcementlearning_b200.engine import make_config
    cfg = make_config(oracle_lib, num_envs=1, width=20, height=15, num_players=2)
    assert (cfg.city_ratio, cfg.city_start_army, cfg.min_general_spacing) == (20, 40, 5)
    gen = ctypes_fn(oracle_lib, "test_generate_map", C.c_int, [C.c_int] * 9 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p])
    W, H = 20, 15
    o, a, t = (np.zeros(W * H, np.int32) for _ in range(3))
    assert gen(W, H, 2, 20, 40, 5, (W * H) // 50, 3, W // 4, 4242, o.ctypes.data, a.ctypes.data, t.ctypes.data) == 0
    e = new_engine(oracle_lib, W, H, 2, 1)
    e.reset_seeded([4242])
    s = e.get_state()
    assert np.array_equal(s["owner"][0], o) and np.array_equal(s["army"][0], a) and np.array_equal(s["type"][0], t)


def test_mapgen_full_25x25_golden(oracle_lib):
    """generator_test.go:394-455 TestGenerateMap_FullIntegration: 25x25, 4 players, seed 12345,
    ratio 30, 10 veins of 4..8, spacing 6, city army 35 -> 58 mountains, 20 cities, 4 generals."""
    fn = ctypes_fn(oracle_lib, "test_generate_map", C.c_int,
                   [C.c_int] * 9 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p])
    owner, army, type_ = (np.zeros(625, np.int32) for _ in range(3))
    assert fn(25, 25, 4, 30, 35, 6, 10, 4, 8, 12345, owner.ctypes.data, army.ctypes.data, type_.ctypes.data) == 0
    assert (type_ == MOUNTAIN).sum() == 58                                     # :451
    assert (type_ == CITY).sum() == 20                                         # :447
    assert (army[type_ == CITY] == 35).all() and (owner[type_ == CITY] == -1).all()
    gens = [int(np.nonzero((type_ == GENERAL) & (owner == p))[0][0]) for p in range(4)]
    assert (type_ == GENERAL).sum() == 4 and (army[type_ == GENERAL] == 2).all()
    for i in range(4):
        for j in range(i + 1, 4):
            (xi, yi), (xj, yj) = (gens[i] % 25, gens[i] // 25), (gens[j] % 25, gens[j] // 25)
            assert abs(xi - xj) + abs(yi - yj) >= 6                            # :456-466
    assert gens == [139, 462, 344, 403]  # recorded by the survey's independent restatement (App. B)


def test_mapgen_general_fallbacks(oracle_lib):
    fn = ctypes_fn(oracle_lib, "test_place_generals", C.c_int,
                   [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p])
    # PanicOnImpossibleSpacing (generator_test.go:336-348): 3 generals, spacing 5 on 3x3 -> error
    t = np.zeros(9, np.int32)
    o = np.zeros(9, np.int32)
    assert fn(3, 3, 3, 5, 12345, t.ctypes.data, o.ctypes.data) != 0
    # FallbackGeneralPlacement (:350-392): only (0,0) and (2,2) usable, spacing 4
    t = np.full(9, MOUNTAIN, np.int32)
    t[0] = t[8] = NORMAL
    assert fn(3, 3, 2, 4, 12345, t.ctypes.data, o.ctypes.data) == 0
    assert t[0] == GENERAL and t[8] == GENERAL and sorted([o[0], o[8]]) == [0, 1]
    # BasicGeneralPlacementAndSpacing (:247-275): 20x20, 4 players, spacing 3
    t = np.zeros(400, np.int32)
    o = np.zeros(400, np.int32)
    assert fn(20, 20, 4, 3, 12345, t.ctypes.data, o.ctypes.data) == 0
    g = np.nonzero(t == GENERAL)[0]
    assert len(g) == 4
    for i in range(4):
        for j in range(i + 1, 4):
            assert abs(g[i] % 20 - g[j] % 20) + abs(g[i] // 20 - g[j] // 20) >= 3


def test_mapgen_small_board_never_fails(oracle_lib):
    """generator_test.go:35-48: spacing clamps to 4 on 5x5 and seeds 0..49 all succeed."""
    e = new_engine(oracle_lib, 5, 5, 2)
    for seed in range(50):
        owner, army, type_ = e.mapgen(seed)
        g = np.nonzero(type_ == GENERAL)[0]
        assert len(g) == 2
        (x0, y0), (x1, y1) = [(i % 5, i // 5) for i in g]
        assert abs(x0 - x1) + abs(y0 - y1) >= 4


def test_mapgen_default_config_invariants(oracle_lib):
    for (W, H, P) in ((10, 10, 2), (15, 15, 2), (20, 20, 2), (20, 20, 4)):
        e = new_engine(oracle_lib, W, H, P)
        for seed in (12345, 12346, 7):
            owner, army, type_ = e.mapgen(seed)
            assert (type_ == GENERAL).sum() == P
            assert (type_ == CITY).sum() == (W * H) // 20
            assert (army[type_ == CITY] == 40).all()
            assert (army[type_ == MOUNTAIN] == 0).all() and (owner[type_ == MOUNTAIN] == -1).all()
            assert (owner[type_ != GENERAL] == -1).all()


# ---- rewards (internal/experience/rewards_test.go:17-142) -----------------------------------
def _reward(lib, e, p, prev_owner, prev_army):
    fn = ctypes_fn(lib, "test_reward", C.c_float, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p])
    po = np.ascontiguousarray(prev_owner, np.int32)
    pa = np.ascontiguousarray(prev_army, np.int32)
    return fn(e._h, 0, p, po.ctypes.data, pa.ctypes.data)


def _test_state(W, H):
    s = blank_state(W, H, 2)
    put(s, W, 0, 0, 0, 10, GENERAL)
    put(s, W, W - 1, H - 1, 1, 10, GENERAL)
    s["owned"][0, 0, 0] = 1
    s["owned"][0, 1, W * H - 1] = 1
    s["turn"][0] = 1
    return s


def test_reward_territory(oracle_lib):
    f32 = np.float32
    e = new_engine(oracle_lib, 3, 3, 2, fog_of_war=0)
    prev = _test_state(3, 3)
    curr = _test_state(3, 3)
    curr["owner"][0, 2] = 0
    e.set_state(curr)
    r = _reward(oracle_lib, e, 0, prev["owner"][0], prev["army"][0])
    # rewards_test.go:17-31 expects exactly TerritoryGained; with equal armies the advantage term is 0
    assert f32(r) == f32(0.01)
    prev["owner"][0, 1] = 0
    curr["owner"][0, 1] = 1
    e.set_state(curr)
    r = _reward(oracle_lib, e, 0, prev["owner"][0], prev["army"][0])
    assert f32(r) == f32(f32(0.01) + f32(-0.01))   # :33-39 gained one, lost one


def test_reward_city_general_army(oracle_lib):
    f32 = np.float32
    e = new_engine(oracle_lib, 5, 5, 2, fog_of_war=0)
    prev, curr = _test_state(5, 5), _test_state(5, 5)
    put(prev, 5, 2, 2, -1, 40, CITY)
    put(curr, 5, 2, 2, 0, 40, CITY)
    e.set_state(curr)
    r = _reward(oracle_lib, e, 0, prev["owner"][0], prev["army"][0])
    adv = f32(f32(50 - 10) / f32(60))
    exp = f32(0.1) + f32(0.01) + f32(40) * f32(0.001) + adv * f32(0.05)
    assert abs(r - exp) < 0.01                                        # rewards_test.go:42-68
    # exact float32 evaluation order (rewards.go:57-84)
    acc = f32(0)
    acc = f32(acc + f32(f32(1) * f32(0.01)))
    acc = f32(acc + f32(f32(40) * f32(0.001)))
    acc = f32(acc + f32(f32(1) * f32(0.1)))
    acc = f32(acc + f32(f32(0) * f32(-0.1)))
    acc = f32(acc + f32(f32(0) * f32(0.5)))
    acc = f32(acc + f32(f32(0) * f32(-0.5)))
    acc = f32(acc + f32(adv * f32(0.05)))
    assert f32(r) == acc

    e3 = new_engine(oracle_lib, 3, 3, 2, fog_of_war=0)
    prev, curr = _test_state(3, 3), _test_state(3, 3)
    put(prev, 3, 2, 2, 1, 8, GENERAL)
    put(curr, 3, 2, 2, 0, 8, GENERAL)
    e3.set_state(curr)
    # curr: both generals belong to player 0 -> player 1 has no general but Alive flags are still set
    r = _reward(oracle_lib, e3, 0, prev["owner"][0], prev["army"][0])
    exp = f32(0.5) + f32(0.01) + f32(8) * f32(0.001) + f32(1.0) * f32(0.05)
    assert abs(r - exp) < 0.01                                        # :70-95

    prev, curr = _test_state(3, 3), _test_state(3, 3)
    curr["army"][0, 0] = 15
    e3.set_state(curr)
    r = _reward(oracle_lib, e3, 0, prev["owner"][0], prev["army"][0])
    assert abs(r - (5 * 0.001 + (5 / 25) * 0.05)) < 0.01              # :97-118


def test_reward_city_changes(oracle_lib):
    """rewards_test.go:144-171 (TestCountCityChanges): city (1,1) neutral -> player 0 and city (3,3) player 0 -> player 1
    count as one gained and one lost; through CalculateRewardWithConfig that is 0.1 - 0.1 with no territory or army
    change, plus the advantage term of the new position ((50 - 50) / 100 = 0), in the reference's order of additions."""
    f32 = np.float32
    e = new_engine(oracle_lib, 5, 5, 2, fog_of_war=0)
    prev, curr = _test_state(5, 5), _test_state(5, 5)
    put(prev, 5, 1, 1, -1, 40, CITY)
    put(curr, 5, 1, 1, 0, 40, CITY)
    put(prev, 5, 3, 3, 0, 40, CITY)
    put(curr, 5, 3, 3, 1, 40, CITY)
    e.set_state(curr)
    r = _reward(oracle_lib, e, 0, prev["owner"][0], prev["army"][0])
    acc = f32(0)
    for term in (f32(0) * f32(0.01), f32(0) * f32(0.001), f32(1) * f32(0.1), f32(1) * f32(-0.1), f32(0) * f32(0.5),
                 f32(0) * f32(-0.5), f32(0) * f32(0.05)):
        acc = f32(acc + f32(term))
    assert f32(r) == acc
    # the other side of the same transition: player 1 gained a city, a tile and 40 armies, and lost nothing
    r1 = _reward(oracle_lib, e, 1, prev["owner"][0], prev["army"][0])
    acc = f32(0)
    for term in (f32(1) * f32(0.01), f32(40) * f32(0.001), f32(1) * f32(0.1), f32(0) * f32(-0.1), f32(0) * f32(0.5),
                 f32(0) * f32(-0.5), f32(f32(0) / f32(100)) * f32(0.05)):
        acc = f32(acc + f32(term))
    assert f32(r1) == acc


def test_reward_army_advantage(oracle_lib):
    f32 = np.float32
    e = new_engine(oracle_lib, 3, 3, 2, fog_of_war=0)
    s = _test_state(3, 3)
    s["owner"][0, 1], s["army"][0, 1] = 0, 5
    s["owner"][0, 2], s["army"][0, 2] = 0, 3
    s["army"][0, 8] = 8
    e.set_state(s)
    # prev == curr isolates the advantage term: (18-8)/26 * 0.05 (rewards_test.go:120-142)
    r0 = _reward(oracle_lib, e, 0, s["owner"][0], s["army"][0])
    r1 = _reward(oracle_lib, e, 1, s["owner"][0], s["army"][0])
    assert f32(r0) == f32(f32(f32(10) / f32(26)) * f32(0.05))
    assert f32(r1) == f32(f32(f32(-10) / f32(26)) * f32(0.05))


def test_reward_terminal(oracle_lib):
    """rewards.go:48-56: +1 / -1 only when a single winner exists."""
    e = new_engine(oracle_lib, 3, 3, 2)
    s = _test_state(3, 3)
    s["alive"][0] = [1, 0]
    e.set_state(s)
    assert _reward(oracle_lib, e, 0, s["owner"][0], s["army"][0]) == 1.0
    assert _reward(oracle_lib, e, 1, s["owner"][0], s["army"][0]) == -1.0
    s["alive"][0] = [0, 0]   # draw: falls through to the shaped reward
    e.set_state(s)
    assert _reward(oracle_lib, e, 0, s["owner"][0], s["army"][0]) == 0.0


# ---- serializer action index (serializer_test.go:126-183, collector_test.go:197-242) -----------
def test_action_index(oracle_lib):
    from helpers import full_stats, one_action
    W = H = 5
    e = new_engine(oracle_lib, W, H, 2)
    cases = [((1, 1, 1, 0), 24), ((2, 3, 3, 3), 71), ((2, 2, 1, 2), (2 * 5 + 2) * 4 + 2),
             ((2, 2, 2, 3), (2 * 5 + 2) * 4 + 1)]
    for (fx, fy, tx, ty), want in cases:
        s = blank_state(W, H, 2)
        put(s, W, 0, 0, 0, 10, GENERAL)
        put(s, W, 4, 4, 1, 10, GENERAL)
        put(s, W, fx, fy, 0, 5, NORMAL)
        full_stats(s)
        e.set_state(s)
        out = e.alloc_outputs_host()
        e.step_fused(one_action(e, 0, fx, fy, tx, ty, True), e.outputs(**out))
        assert out["step_error"][0] == 0
        assert out["action_index"][0].tolist() == [want, -1]
