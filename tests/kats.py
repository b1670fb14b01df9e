"""Known-answer tests transliterated from the reference's own Go tests.

Every function takes a bound library (the CPU oracle or the CUDA product — same C ABI)
and asserts the reference's expected values.  test_oracle_kat.py pins the oracle with
them; test_cuda_kat.py runs the identical cases through libgrlcuda.so on the GPU.
"""
from __future__ import annotations

import numpy as np

from generalsreinforcementlearning_b200 import _abi
from helpers import (CITY, GENERAL, MOUNTAIN, NORMAL, blank_state, full_fog, full_stats, new_engine, one_action,
                     put)


def _step_one_move(lib, tiles, action, W=3, H=3, P=2, lists=False, alive=None):
    """Build a board by hand, apply ONE MoveAction through Step, return (state, engine)."""
    e = new_engine(lib, W, H, P)
    s = blank_state(W, H, P)
    for (x, y, owner, army, type_) in tiles:
        put(s, W, x, y, owner, army, type_)
    if lists:
        full_stats(s)
    if alive is not None:
        s["alive"][0, :] = alive
    e.set_state(s)
    e.step(one_action(e, *action))
    return e.get_state(), e


# ---- core/movement_test.go:71-235 TestApplyMoveAction_BasicMovement -----------------------
MOVE_TABLE = [
    # name, from(owner,army), to(owner,army), move_all, exp_from, exp_to, exp_to_owner
    ("move all to own tile", (0, 10), (0, 5), True, 1, 14, 0),
    ("move half to own tile", (0, 10), (0, 5), False, 5, 10, 0),
    ("capture neutral tile", (0, 10), (-1, 3), True, 1, 6, 0),
    ("failed attack on enemy tile", (0, 5), (1, 10), True, 1, 6, 1),
    ("exact army match (no capture)", (0, 6), (1, 5), True, 1, 0, 1),
    ("move half with odd number", (0, 3), (0, 0), False, 2, 1, 0),
    ("move half of 2 (movement_test.go:485-503)", (0, 2), (0, 0), False, 1, 1, 0),
]


def kat_apply_move_table(lib):
    for name, (fo, fa), (to, ta), move_all, exp_from, exp_to, exp_owner in MOVE_TABLE:
        s, _ = _step_one_move(lib, [(0, 0, fo, fa, NORMAL), (1, 0, to, ta, NORMAL)], (0, 0, 0, 1, 0, move_all))
        assert s["step_error"][0] == 0, name
        assert s["army"][0, 0] == exp_from, name
        assert s["army"][0, 1] == exp_to, name
        assert s["owner"][0, 1] == exp_owner, name
        assert s["owner"][0, 0] == 0, name
        # movement.go:57-60: both tiles enter ChangedTiles
        assert s["changed"][0, 0] == 1 and s["changed"][0, 1] == 1, name
        # action_processor.go:78-87: only a capture marks the tile visibility-changed
        captured = exp_owner == 0 and to != 0
        assert s["vis_changed"][0, 1] == (1 if captured else 0), name
        assert s["turn"][0] == 1, name


# ---- core/movement_test.go:275-332 TestApplyMoveAction_InvalidMoves ---------------------------
def kat_apply_move_invalid(lib):
    """5x5, player 0 holds (0,0) with 10 armies, a mountain at (1,0): moving into the mountain, moving from a tile
    the player does not own, and moving with a single army are errors, and an error leaves the board as it was."""
    W = H = 5
    cases = [("move to mountain", [(0, 0, 0, 10, NORMAL), (1, 0, -1, 0, MOUNTAIN)], (0, 0, 0, 1, 0, True), _abi.STEP_TARGET_IS_MOUNTAIN),
             ("move from unowned tile", [(0, 0, 0, 10, NORMAL), (1, 0, -1, 0, MOUNTAIN)], (0, 3, 3, 3, 4, True), _abi.STEP_NOT_OWNED),
             ("move with insufficient army", [(0, 0, 0, 1, NORMAL)], (0, 0, 0, 1, 0, True), _abi.STEP_INSUFFICIENT_ARMY)]
    for name, tiles, action, code in cases:
        s, _ = _step_one_move(lib, tiles, action, W=W, H=H)
        assert s["step_error"][0] == code, name
        for (x, y, owner, army, type_) in tiles:
            i = y * W + x
            assert (s["owner"][0, i], s["army"][0, i], s["type"][0, i]) == (owner, army, type_), name
        assert s["changed"][0].sum() == 0, name


# ---- core/movement_test.go:237-273 capture details: city 50 vs 40 --------------------------
def kat_capture_city(lib):
    s, _ = _step_one_move(lib, [(0, 0, 0, 50, NORMAL), (1, 0, 1, 40, CITY)], (0, 0, 0, 1, 0, True))
    assert s["owner"][0, 1] == 0
    assert s["army"][0, 1] == 9  # (50-1) - 40
    assert s["type"][0, 1] == CITY  # movement.go:38: type never changes on capture
    assert s["army"][0, 0] == 1


# ---- core/movement_test.go:10-69 + internal/game/engine_test.go:184-248 ----------------------
def kat_general_capture_elimination_3x3(lib):
    tiles = [(0, 0, 0, 10, NORMAL), (1, 0, 1, 1, GENERAL), (1, 1, 1, 5, CITY), (2, 1, 1, 3, NORMAL)]
    s, _ = _step_one_move(lib, tiles, (0, 0, 0, 1, 0, True), lists=True, alive=1)
    # captured general: (10-1)-1 = 8, +1 production the same turn (stats ran after the elimination)
    assert s["owner"][0, 1] == 0 and s["army"][0, 1] == 9 and s["type"][0, 1] == GENERAL
    assert s["army"][0, 0] == 1
    assert s["owner"][0, 4] == 0 and s["army"][0, 4] == 6  # city turned over, +1
    assert s["owner"][0, 5] == 0 and s["army"][0, 5] == 3  # land turned over, no growth on turn 1
    assert s["alive"][0].tolist() == [1, 0]
    assert s["general_idx"][0, 1] == -1
    assert s["game_over"][0] == 1 and s["winner"][0] == 0


def kat_engine_elimination_5x5_seed12345(lib):
    """internal/game/engine_test.go:184-248 on the real seed-12345 5x5 map: 19 / 6 / 3."""
    W = H = 5
    e = new_engine(lib, W, H, 2)
    e.reset_seeded([12345])
    s = e.get_state()
    p1_gen = int(s["general_idx"][0, 1])
    assert p1_gen != -1 and s["general_idx"][0, 0] != -1
    put(s, W, 0, 0, 0, 20, NORMAL)
    new_gen = 1 * W + 0  # board.Idx(0, 1)
    if p1_gen != new_gen:
        s["owner"][0, p1_gen], s["type"][0, p1_gen], s["army"][0, p1_gen] = -1, NORMAL, 0
    put(s, W, 0, 1, 1, 1, GENERAL)
    put(s, W, 1, 1, 1, 5, CITY)
    put(s, W, 2, 2, 1, 3, NORMAL)
    assert s["general_idx"][0, 0] != new_gen
    full_stats(s)  # engine.updatePlayerStats() at turn 0
    e.set_state(s)
    e.step(one_action(e, 0, 0, 0, 0, 1, True))
    t = e.get_state()
    assert t["step_error"][0] == 0
    assert t["alive"][0].tolist() == [1, 0]
    assert t["general_idx"][0, 1] == -1
    assert t["owner"][0, new_gen] == 0 and t["army"][0, new_gen] == 19
    assert t["owner"][0, 6] == 0 and t["army"][0, 6] == 6
    assert t["owner"][0, 12] == 0 and t["army"][0, 12] == 3
    assert t["game_over"][0] == 1 and t["winner"][0] == 0


# ---- core/movement_test.go:333-465 TestProcessCaptures, at engine level ------------------------
def kat_process_captures(lib):
    W, H, P = 5, 5, 3

    def run(tiles, actions):
        e = new_engine(lib, W, H, P, max_actions=3)
        s = blank_state(W, H, P)
        for t in tiles:
            put(s, W, *t)
        full_stats(s)
        s["alive"][0, :] = 1
        e.set_state(s)
        acts = None
        for slot, a in enumerate(actions):
            acts = one_action(e, *a, slot=slot, actions=acts)
        e.step(acts)
        return e.get_state()

    # "multiple captures including general": P0 takes P1's city and P2's general, P1 takes P2 land.
    # Only the general capture yields an order (2 -> 0); P2's remaining list tiles go to P0.
    tiles = [
        (0, 0, 0, 10, GENERAL), (4, 0, 1, 10, GENERAL), (4, 4, 2, 1, GENERAL),
        (0, 1, 0, 60, NORMAL), (0, 2, 1, 40, CITY),   # P0 -> city of P1
        (3, 4, 0, 10, NORMAL),                         # P0 -> general of P2 at (4,4)
        (2, 2, 1, 10, NORMAL), (2, 3, 2, 5, NORMAL),   # P1 -> land of P2
        (1, 4, 2, 7, NORMAL),                          # bystander P2 land
    ]
    s = run(tiles, [(0, 3, 4, 4, 4, True), (1, 2, 2, 2, 3, True)])
    assert s["alive"][0].tolist() == [1, 1, 0]
    assert s["owner"][0, 4 * W + 4] == 0
    assert s["owner"][0, 4 * W + 1] == 0          # bystander land turned over to the capturer
    assert s["owner"][0, 3 * W + 2] == 1          # captured by P1 before the turnover: stays P1's
    assert s["game_over"][0] == 0

    # "no general captures": city + land captures, nobody eliminated
    tiles = [
        (0, 0, 0, 10, GENERAL), (4, 0, 1, 10, GENERAL), (4, 4, 2, 10, GENERAL),
        (0, 1, 0, 60, NORMAL), (0, 2, 1, 40, CITY), (2, 2, 1, 10, NORMAL), (2, 3, 2, 5, NORMAL),
    ]
    s = run(tiles, [(0, 0, 1, 0, 2, True), (1, 2, 2, 2, 3, True)])
    assert s["alive"][0].tolist() == [1, 1, 1]
    # 59 - 40 = 19, then +1: the city is still in P1's cached list at production time and P1 is
    # alive, so it produces for its new owner
    assert s["owner"][0, 2 * W + 0] == 0 and s["army"][0, 2 * W + 0] == 20

    # "neutral general capture (no elimination)"
    tiles = [(0, 0, 0, 10, GENERAL), (4, 0, 1, 10, GENERAL), (4, 4, 2, 10, GENERAL),
             (2, 2, 0, 10, NORMAL), (2, 3, -1, 1, GENERAL)]
    s = run(tiles, [(0, 2, 2, 2, 3, True)])
    assert s["alive"][0].tolist() == [1, 1, 1]
    assert s["owner"][0, 3 * W + 2] == 0

    # "duplicate general captures (same player)": P1 holds two general-type tiles (Q4); P0 takes one,
    # then P2 takes the other in the same turn.  First capture wins: one order (1 -> 0).
    tiles = [(0, 0, 0, 10, GENERAL), (4, 4, 2, 10, GENERAL),
             (1, 1, 1, 1, GENERAL), (3, 3, 1, 1, GENERAL), (3, 0, 1, 4, NORMAL),
             (0, 1, 0, 10, NORMAL), (3, 4, 2, 10, NORMAL)]
    s = run(tiles, [(0, 0, 1, 1, 1, True), (2, 3, 4, 3, 3, True)])
    assert s["owner"][0, 1 * W + 1] == 0
    assert s["owner"][0, 3 * W + 3] == 2          # P2's own capture stands
    assert s["owner"][0, 0 * W + 3] == 0          # P1's land goes to the FIRST capturer
    assert s["alive"][0].tolist() == [1, 0, 1]


# ---- core/action_test.go:23-186 TestMoveAction_Validate ----------------------------------------
def kat_validation(lib):
    W = H = 5

    def err_of(action, tiles):
        e = new_engine(lib, W, H, 2)
        s = blank_state(W, H, 2)
        for t in tiles:
            put(s, W, *t)
        e.set_state(s)
        e.step(one_action(e, *action))
        return int(e.get_state()["step_error"][0]), e

    base = [(2, 2, 0, 5, NORMAL)]
    assert err_of((0, 2, 2, 2, 1, True), base)[0] == _abi.STEP_OK
    oob = [(-1, 2, 0, 2), (5, 2, 4, 2), (2, -1, 2, 0), (2, 5, 2, 4),   # from out of bounds
           (0, 2, -1, 2), (4, 2, 5, 2), (2, 0, 2, -1), (2, 4, 2, 5)]   # to out of bounds
    for fx, fy, tx, ty in oob:
        assert err_of((0, fx, fy, tx, ty, True), base + [(0, 2, 0, 5, NORMAL), (4, 2, 0, 5, NORMAL),
                                                           (2, 0, 0, 5, NORMAL), (2, 4, 0, 5, NORMAL)])[0] \
            == _abi.STEP_INVALID_COORDINATES
    assert err_of((0, 2, 2, 2, 2, True), base)[0] == _abi.STEP_MOVE_TO_SELF
    assert err_of((0, 2, 2, 3, 3, True), base)[0] == _abi.STEP_NOT_ADJACENT   # diagonal
    assert err_of((0, 2, 2, 4, 2, True), base)[0] == _abi.STEP_NOT_ADJACENT   # same row, far
    assert err_of((0, 2, 2, 2, 1, True), [(2, 2, 1, 5, NORMAL)])[0] == _abi.STEP_NOT_OWNED
    assert err_of((0, 2, 2, 2, 1, True), [(2, 2, 0, 1, NORMAL)])[0] == _abi.STEP_INSUFFICIENT_ARMY
    assert err_of((0, 2, 2, 2, 1, True), [(2, 2, 0, 0, NORMAL)])[0] == _abi.STEP_INSUFFICIENT_ARMY
    code, e = err_of((0, 2, 2, 2, 1, True), base + [(2, 1, -1, 0, MOUNTAIN)])
    assert code == _abi.STEP_TARGET_IS_MOUNTAIN
    # precedence (action.go:56-105): bounds, self, adjacency, ownership, army, mountain
    assert err_of((0, 2, 2, 2, 1, True), [(2, 2, 1, 1, NORMAL), (2, 1, -1, 0, MOUNTAIN)])[0] == _abi.STEP_NOT_OWNED
    assert err_of((0, 2, 2, 2, 1, True), [(2, 2, 0, 1, NORMAL), (2, 1, -1, 0, MOUNTAIN)])[0] \
        == _abi.STEP_INSUFFICIENT_ARMY
    # an error turn leaves the board untouched but the turn counter advanced
    s = e.get_state()
    assert s["turn"][0] == 1 and s["army"][0, 2 * W + 2] == 5


# ---- internal/game/engine_test.go:65-102 -----------------------------------------------------------
def kat_basic_turn_and_game_over(lib):
    e = new_engine(lib, 5, 5, 1)
    e.reset_seeded([12345])
    s0 = e.get_state()
    assert s0["turn"][0] == 0 and s0["game_over"][0] == 0
    e.step(None)
    s1 = e.get_state()
    assert s1["turn"][0] == 1
    assert s1["army_count"][0, 0] == s0["army_count"][0, 0] + 1
    assert s1["game_over"][0] == 0
    # Step on a finished game -> ErrGameOver, nothing mutated (turn_processor.go:95-113)
    s1["game_over"][0] = 1
    e.set_state(s1)
    h = e.state_hash()
    e.step(None)
    s2 = e.get_state()
    assert s2["step_error"][0] == _abi.STEP_GAME_OVER
    assert s2["turn"][0] == 1
    assert np.array_equal(e.state_hash(), h)


# ---- internal/game/engine_test.go:104-182 production on/off the growth tick ---------------------------
def kat_production(lib):
    W = H = 5
    for turn_before, land_grows in ((24, True), (23, False)):
        e = new_engine(lib, W, H, 1)
        e.reset_seeded([12345])
        s = e.get_state()
        gen = int(s["general_idx"][0, 0])
        free = [i for i in range(W * H) if s["type"][0, i] == NORMAL and s["owner"][0, i] == -1]
        city, land = free[0], free[1]
        s["owner"][0, city], s["type"][0, city], s["army"][0, city] = 0, CITY, 5
        s["owner"][0, land], s["army"][0, land] = 0, 2
        full_stats(s)
        s["turn"][0] = turn_before  # Step increments to 25 / 24
        e.set_state(s)
        g0 = int(s["army"][0, gen])
        e.step(None)
        t = e.get_state()
        assert t["army"][0, gen] == g0 + 1
        assert t["army"][0, city] == 6
        assert t["army"][0, land] == (3 if land_grows else 2)


# ---- internal/game/engine_test.go:250-303 dead player's action ignored --------------------------------
def kat_dead_player_action(lib):
    W = H = 5
    e = new_engine(lib, W, H, 2)
    e.reset_seeded([12345])
    s = e.get_state()
    g1 = int(s["general_idx"][0, 1])
    s["owner"][0, g1], s["army"][0, g1], s["type"][0, g1] = -1, 0, NORMAL
    s["alive"][0, 1] = 0
    s["general_idx"][0, 1] = -1
    put(s, W, 0, 0, 0, 10, NORMAL)
    put(s, W, 1, 1, 1, 5, NORMAL)
    e.set_state(s)
    acts = one_action(e, 1, 1, 1, 1, 2, True, slot=0)
    acts = one_action(e, 0, 0, 0, 0, 1, True, slot=1, actions=acts)
    e.step(acts)
    t = e.get_state()
    assert t["step_error"][0] == 0
    assert t["owner"][0, 6] == 1 and t["army"][0, 6] == 5
    assert t["army"][0, 0] == 1
    assert t["owner"][0, 5] == 0 and t["army"][0, 5] == 9
    assert t["alive"][0].tolist() == [1, 0]


# ---- internal/game/action_mask_test.go:57-292 ------------------------------------------------------------
def _mask_case(lib, W, H, P, tiles, owned, alive=True, mountains=()):
    e = new_engine(lib, W, H, P)
    s = blank_state(W, H, P)
    for (x, y, army) in tiles:
        put(s, W, x, y, 0, army, NORMAL)
    for (x, y) in mountains:
        s["type"][0, y * W + x] = MOUNTAIN
    for (x, y) in owned:
        s["owned"][0, 0, y * W + x] = 1
    s["alive"][0, 0] = 1 if alive else 0
    e.set_state(s)
    return e


def kat_engine_mask(lib):
    # BasicScenario
    e = _mask_case(lib, 3, 3, 2, [(1, 1, 5)], [(1, 1)])
    m = e.get_legal_action_mask(0, 0)
    assert len(m) == 36
    base = (1 * 3 + 1) * 4
    assert m[base:base + 4].all() and m.sum() == 4
    # the packed variant carries the same bits
    bits = e.mask(_abi.MASK_ENGINE_URDL_BITS)[0, 0]
    unpacked = np.array([(bits[i >> 5] >> (i & 31)) & 1 for i in range(36)], bool)
    assert np.array_equal(unpacked, m)
    half = e.mask(_abi.MASK_ENGINE_HALF_BITS)[0, 0]
    assert np.array_equal(half[0], bits) and np.array_equal(half[1], bits)
    # EdgeTiles: corner (0,0): right and down only
    e = _mask_case(lib, 3, 3, 1, [(0, 0, 3)], [(0, 0)])
    m = e.get_legal_action_mask(0, 0)
    assert m[0:4].tolist() == [False, True, True, False]
    # InsufficientArmy
    e = _mask_case(lib, 3, 3, 1, [(0, 0, 1), (1, 1, 2)], [(0, 0), (1, 1)])
    m = e.get_legal_action_mask(0, 0)
    assert not m[0:4].any() and m[16:20].any()
    # Mountains all around
    ring = [(x, y) for x in range(3) for y in range(3) if (x, y) != (1, 1)]
    e = _mask_case(lib, 3, 3, 1, [(1, 1, 5)], [(1, 1)], mountains=ring)
    assert not e.get_legal_action_mask(0, 0).any()
    # DeadPlayer
    e = _mask_case(lib, 3, 3, 2, [(1, 1, 10)], [(1, 1)], alive=False)
    assert not e.get_legal_action_mask(0, 0).any()
    # InvalidPlayer
    assert len(e.get_legal_action_mask(0, -1)) == 36 and not e.get_legal_action_mask(0, -1).any()
    assert not e.get_legal_action_mask(0, 5).any()
    # ComplexScenario
    tiles = [(1, 1, 5), (2, 1, 1), (3, 3, 3), (0, 0, 2)]
    e = _mask_case(lib, 5, 5, 2, tiles, [(x, y) for x, y, _ in tiles], mountains=[(1, 2)])
    m = e.get_legal_action_mask(0, 0)
    assert 0 < m.sum() < 20
    assert not m[(1 * 5 + 1) * 4 + 2]
    assert m.sum() == 3 + 4 + 2  # (1,1): U,R,L; (3,3): all; (0,0): R,D
    # list-based, not ownership-based (rules/legal_moves.go:38): an owned tile missing from the
    # cached list offers no move
    e = _mask_case(lib, 3, 3, 2, [(1, 1, 5), (0, 0, 5)], [(1, 1)])
    m = e.get_legal_action_mask(0, 0)
    assert m.sum() == 4 and not m[0:4].any()


# ---- internal/experience/serializer_test.go:64-124,185-220 ---------------------------------------------------
def _detailed_state(W, H):
    """createTestGameState + createTestGameStateWithDetails (collector_test.go:298-357,
    serializer_test.go:11-59): fog disabled, every tile flagged visible."""
    s = blank_state(W, H, 2)
    put(s, W, 0, 0, 0, 10, GENERAL)
    put(s, W, W - 1, H - 1, 1, 10, GENERAL)
    s["owner"][0, 1], s["army"][0, 1] = 0, 5
    s["owner"][0, 2], s["army"][0, 2] = 0, 3
    if W > 2 and H > 2:
        put(s, W, 2, 2, -1, 40, CITY)
    put(s, W, 1, 1, 0, 0, MOUNTAIN)  # core.Tile{Type: TileMountain}: Owner zero-value = 0
    s["owned"][0, 0, 0] = 1
    s["owned"][0, 1, W * H - 1] = 1
    s["army_count"][0] = [10, 10]
    s["general_idx"][0] = [0, W * H - 1]
    s["turn"][0] = 1
    s["visible"][0, :] = 3
    return s


def kat_state_to_tensor(lib):
    W = H = 5
    e = new_engine(lib, W, H, 2, fog_of_war=0)
    e.set_state(_detailed_state(W, H))
    out = e.alloc_outputs_host()
    e.observe(e.outputs(**out))
    t = out["obs"][0, 0]
    assert t.size == 9 * 25
    assert t[0, 0, 0] == np.float32(10.0) / np.float32(1000.0)
    assert t[2, 0, 0] == 1.0
    assert t[6, 1, 1] == 1.0
    assert t[5, 2, 2] == 1.0
    assert t[4, 2, 2] == 1.0
    # mountains carry nothing but channel 6 (+7) (serializer.go:66-70)
    assert t[2, 1, 1] == 0.0 and t[7, 1, 1] == 1.0
    # player 1's view: own general in ch0/ch2, player 0's in ch1/ch3
    u = out["obs"][0, 1]
    assert u[0, 4, 4] == np.float32(10.0) / np.float32(1000.0) and u[2, 4, 4] == 1.0
    assert u[1, 0, 0] == np.float32(10.0) / np.float32(1000.0) and u[3, 0, 0] == 1.0
    assert u[5, 0, 0] == 1.0 and u[5, 4, 4] == 1.0  # generals light the "cities" channel


def kat_serializer_mask(lib):
    e = new_engine(lib, 3, 3, 2, fog_of_war=0)
    e.set_state(_detailed_state(3, 3))
    m = e.mask(_abi.MASK_SERIALIZER_UDLR)[0, 0].astype(bool)
    assert len(m) == 36
    assert m[3] and m[1] and not m[0] and not m[2]   # (0,0): right, down; not up, left
    assert not m[(0 * 3 + 1) * 4 + 1]                # (1,0) down into the mountain


def kat_tensor_under_fog(lib):
    e = new_engine(lib, 3, 3, 2, fog_of_war=1)
    s = _detailed_state(3, 3)
    s["visible"][0, :] = 0
    for i in (0, 1, 3, 4):
        s["visible"][0, i] = 1
    e.set_state(s)
    out = e.alloc_outputs_host()
    e.observe(e.outputs(**out))
    t = out["obs"][0, 0]
    assert t[7, 0, 0] == 1.0
    assert t[8, 2, 2] == 1.0
    assert t[1, 2, 2] == 0.0
    assert t[3, 2, 2] == 0.0 and t[5, 2, 2] == 0.0
    # ComputePlayerVisibility (visibility_optimized.go:166-195): non-normal tiles show as fog
    vis, fog = e.compute_player_visibility(0, 0)
    assert vis.tolist() == [True, True, False, True, True, False, False, False, False]
    assert fog[8] and not fog[4] and not fog[2]  # hidden enemy general leaks as a fog tile (Q14)


def kat_army_clip(lib):
    """serializer.go:84-88: min(army/1000, 1)."""
    e = new_engine(lib, 3, 3, 2, fog_of_war=0)
    s = _detailed_state(3, 3)
    s["army"][0, 0] = 1500
    put(s, 3, 2, 1, 1, 999, NORMAL)   # (note: at 3x3 the fixture's city overwrites P1's general at (2,2))
    e.set_state(s)
    out = e.alloc_outputs_host()
    e.observe(e.outputs(**out))
    assert out["obs"][0, 0, 0, 0, 0] == 1.0
    assert out["obs"][0, 0, 1, 1, 2] == np.float32(999.0) / np.float32(1000.0)
    assert out["obs"][0, 1, 0, 1, 2] == np.float32(999.0) / np.float32(1000.0)
    assert out["obs"][0, 1, 1, 0, 0] == 1.0
    # serializer_test.go:222-243 TestNormalizeArmyValue: 0 -> 0, 100 -> 0.1, 500 -> 0.5, 1000 -> 1, 2000 -> 1 (capped)
    e = new_engine(lib, 5, 5, 2, fog_of_war=0)
    s = _detailed_state(5, 5)
    table = [(0, np.float32(0.0)), (100, np.float32(0.1)), (500, np.float32(0.5)), (1000, np.float32(1.0)), (2000, np.float32(1.0))]
    for k, (army, _) in enumerate(table):
        put(s, 5, k, 3, 0, army, NORMAL)
    e.set_state(s)
    out = e.alloc_outputs_host()
    e.observe(e.outputs(**out))
    for k, (army, want) in enumerate(table):
        assert out["obs"][0, 0, 0, 3, k] == want, f"NormalizeArmyValue({army})"
        assert out["obs"][0, 1, 1, 3, k] == want, f"the enemy's view of {army}"


# ---- fog of war semantics (visibility_optimized.go) ----------------------------------------------
def kat_fog_lags_one_turn(lib):
    W = H = 7
    e = new_engine(lib, W, H, 2)
    s = blank_state(W, H, 2)
    put(s, W, 1, 1, 0, 10, GENERAL)
    put(s, W, 5, 5, 1, 10, GENERAL)
    full_stats(s)
    full_fog(s, W, H)
    e.set_state(s)
    e.step(one_action(e, 0, 1, 1, 2, 1, True))   # capture (2,1)
    t = e.get_state()
    assert t["owner"][0, 1 * W + 2] == 0
    assert t["vis_changed"][0, 1 * W + 2] == 1
    # visibility not yet extended to x=3 (turn_processor.go:124-127 ran before the capture)
    assert (t["visible"][0, 1 * W + 3] & 1) == 0
    e.step(None)
    t = e.get_state()
    assert (t["visible"][0, 1 * W + 3] & 1) == 1
    assert (t["visible"][0, 0 * W + 3] & 1) == 1 and (t["visible"][0, 2 * W + 3] & 1) == 1
    assert t["vis_changed"][0].sum() == 0


# ---- Tile.VisibleBitfield (core/board_test.go:282-370 TestBoard_VisibilityMap / TestTileBitfieldVisibility) -----------
def kat_visibility_bitfield(lib):
    """Tile.SetVisible / IsVisibleTo over the uint32 bitfield, through the state codec (grl_state_planes.visible IS
    Tile.VisibleBitfield) and the per-player read-out (grl_visibility = IsVisibleTo(p) per tile).  The reference's field
    holds 32 players and ignores ids -1 and 32; this ABI seats GRL_MAX_PLAYERS = 8, so ids 8..31 (and the two invalid
    ids) are not addressable at all: their bits are dropped by the codec, which is the only place they could enter."""
    W = H = 3
    P = 8
    e = new_engine(lib, W, H, P)
    s = blank_state(W, H, P)
    centre = 1 * W + 1
    # BasicVisibilityOperations: players 0, 3, 7 set, then 0 cleared
    s["visible"][0, centre] = (1 << 0) | (1 << 3) | (1 << 7)
    s["visible"][0, 0] = 0b10101010                      # BitfieldDirectManipulation
    s["visible"][0, 1] = 0xFFFFFFFF                      # AllPlayersVisible (32 bits set in the reference's field)
    s["visible"][0, 2] = 0xFFFFFFFF & ~(1 << 5)          # ... with one in the middle cleared
    s["visible"][0, 3] = (1 << 31) | (1 << 8)            # ids beyond the 8 seats: nothing addressable remains
    e.set_state(s)
    vis, fog = e.visibility()
    see = lambda t: [bool(vis[0, p, t]) for p in range(P)]  # noqa: E731
    assert see(centre) == [True, False, False, True, False, False, False, True]
    assert see(0) == [False, True, False, True, False, True, False, True]
    assert see(1) == [True] * 8
    assert see(2) == [True, True, True, True, True, False, True, True]
    assert see(3) == [False] * 8 and see(5) == [False] * 8      # unset defaults to false (TestBoard_VisibilityMap)
    assert not fog.any()                                          # normal tiles never show through fog
    t = e.get_state()
    assert t["visible"][0, centre] == 0b10001001 and t["visible"][0, 0] == 0b10101010
    assert t["visible"][0, 1] == 0xFF and t["visible"][0, 2] == 0xFF & ~(1 << 5) and t["visible"][0, 3] == 0
    s["visible"][0, centre] &= ~np.uint32(1)             # SetVisible(0, false)
    e.set_state(s)
    vis, _ = e.visibility()
    assert [bool(vis[0, p, centre]) for p in range(P)] == [False, False, False, True, False, False, False, True]
    e.close()
