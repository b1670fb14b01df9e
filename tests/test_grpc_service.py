"""The gRPC GameService / ExperienceService front end served over the CPU
oracle library: wire contract, validation order, turn barrier, fog-filtered views, experience
records.  When /root/reference is present (this container only) the reference's OWN generated
client stubs are pointed at the server as well."""
import os
import sys

import grpc
import numpy as np
import pytest

from generalsreinforcementlearning_b200 import _abi
from generalsreinforcementlearning_b200.grpc_schema import common, experience, game
from generalsreinforcementlearning_b200.grpc_service import Stub, serve, validate_move
from helpers import new_engine

GAME = "generals.game.v1.GameService"
EXP = "generals.experience.v1.ExperienceService"
DIRS = [(0, -1), (1, 0), (0, 1), (-1, 0)]  # engine mask order: up, right, down, left


@pytest.fixture()
def server(oracle_lib):
    srv, gs, port = serve("127.0.0.1:0", lib=oracle_lib, slots_per_pool=8, seed=1000)
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    yield gs, Stub(ch, GAME), Stub(ch, EXP), ch, port
    ch.close()
    srv.stop(0)
    gs.close()


def _start(stub, W=8, H=8, collect=False):
    cfg = game.GameConfig(width=W, height=H, max_players=2, fog_of_war=True, collect_experiences=collect)
    gid = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
    j = [stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name=n)) for n in ("alice", "bob")]
    return gid, j


def _legal_action(state, rng, turn):
    mask = np.array(state.action_mask, bool)
    W = state.board.width
    idx = np.nonzero(mask)[0]
    if len(idx) == 0:
        return None
    k = int(idx[rng.integers(len(idx))])
    t, d = k // 4, k % 4
    a = game.Action(type=common.ACTION_TYPE_MOVE, turn_number=turn, half=bool(rng.integers(2)))
    getattr(a, "from").x, getattr(a, "from").y = t % W, t // W
    a.to.x, a.to.y = t % W + DIRS[d][0], t // W + DIRS[d][1]
    return a


def test_lifecycle_and_validation_order(server):
    gs, stub, _, _, _ = server
    r = stub.CreateGame(game.CreateGameRequest())
    assert r.game_id == "game-1" and (r.config.width, r.config.height, r.config.max_players, r.config.fog_of_war) == (20, 20, 2, True)
    gid, (j0, j1) = _start(stub)
    assert (j0.player_id, j0.player_token) == (0, f"token-{gid}-0") and j1.player_id == 1
    # the first joiner saw the lobby placeholder, the second the running game
    assert j0.initial_state.current_phase == common.GAME_PHASE_LOBBY and j0.initial_state.status == common.GAME_STATUS_WAITING
    assert j1.initial_state.current_phase == common.GAME_PHASE_RUNNING and j1.initial_state.status == common.GAME_STATUS_IN_PROGRESS
    # re-join by name returns the same identity; a third player is refused
    assert stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="alice")).player_token == j0.player_token
    with pytest.raises(grpc.RpcError) as e:
        stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="carol"))
    assert e.value.code() == grpc.StatusCode.FAILED_PRECONDITION  # no longer in the lobby phase
    with pytest.raises(grpc.RpcError) as e:
        stub.JoinGame(game.JoinGameRequest(game_id="game-99", player_name="x"))
    assert e.value.code() == grpc.StatusCode.NOT_FOUND
    with pytest.raises(grpc.RpcError) as e:
        stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=0, player_token="nope"))
    assert e.value.code() == grpc.StatusCode.PERMISSION_DENIED
    # SubmitAction failures come back in-band, in the reference's precedence
    sub = lambda **kw: stub.SubmitAction(game.SubmitActionRequest(**kw))
    assert sub(game_id="game-99").error_code == common.ERROR_CODE_GAME_NOT_FOUND
    lobby = stub.CreateGame(game.CreateGameRequest()).game_id
    assert sub(game_id=lobby, player_id=0).error_code == common.ERROR_CODE_INVALID_PHASE
    assert sub(game_id=gid, player_id=0, player_token="bad").error_code == common.ERROR_CODE_INVALID_PLAYER
    st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=0, player_token=j0.player_token)).state
    a = _legal_action(st, np.random.default_rng(0), turn=5)
    r = sub(game_id=gid, player_id=0, player_token=j0.player_token, action=a)
    assert r.error_code == common.ERROR_CODE_INVALID_TURN and "expected 0, got 5" in r.error_message
    bad = game.Action(type=common.ACTION_TYPE_MOVE, turn_number=0)
    getattr(bad, "from").x = 0
    bad.to.x = 5
    r = sub(game_id=gid, player_id=0, player_token=j0.player_token, action=bad)
    assert r.error_code == common.ERROR_CODE_INVALID_TURN and "action validation failed" in r.error_message
    # idempotency: the cached response is returned, the action is not applied twice
    a.turn_number = 0
    r1 = sub(game_id=gid, player_id=0, player_token=j0.player_token, action=a, idempotency_key="k1")
    r2 = sub(game_id=gid, player_id=0, player_token=j0.player_token, action=a, idempotency_key="k1")
    assert r1.success and r2.success and r1.next_turn_number == r2.next_turn_number == 1
    assert stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=0, player_token=j0.player_token)).state.turn == 0
    # the turn runs when the LAST player has submitted (an empty request = no action this turn)
    assert sub(game_id=gid, player_id=1, player_token=j1.player_token).success
    assert stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=0, player_token=j0.player_token)).state.turn == 1


def test_served_games_follow_the_engine_exactly(server, oracle_lib):
    """Two games interleaved on one pool; each must equal a private engine fed the same moves."""
    _served_games_follow(server, oracle_lib)


@pytest.fixture()
def cuda_server(cuda_lib):
    srv, gs, port = serve("127.0.0.1:0", lib=cuda_lib, slots_per_pool=64, seed=1000)
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    yield gs, Stub(ch, GAME), Stub(ch, EXP), ch, port
    ch.close()
    srv.stop(0)
    gs.close()


@pytest.mark.gpu
def test_served_games_on_the_gpu_follow_the_oracle(cuda_server, oracle_lib):
    """The same service over libgrlcuda.so (device-resident pools); mirrors are oracle engines."""
    _served_games_follow(cuda_server, oracle_lib)


@pytest.mark.gpu
def test_experience_stream_on_the_gpu(cuda_server, oracle_lib):
    _experience_stream_records(cuda_server, oracle_lib)


def _served_games_follow(server, oracle_lib):
    gs, stub, _, _, _ = server
    rng = np.random.default_rng(7)
    games = [_start(stub, 8, 8) for _ in range(2)]
    mirrors = []
    for i, (gid, _) in enumerate(games):
        m = new_engine(oracle_lib, 8, 8, 2, 1)
        m.reset_seeded([1000 + int(gid.split("-")[1])])
        mirrors.append(m)
    for turn in range(40):
        for (gid, js), m in zip(games, mirrors):
            if rng.random() < 0.25:
                continue  # this game sits this round out: the other one must not be disturbed
            acts = np.zeros((1, m.A), _abi.ACTION_DTYPE)
            for j in js:
                st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)).state
                if st.current_phase != common.GAME_PHASE_RUNNING:
                    break
                a = _legal_action(st, rng, st.turn)
                req = game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)
                if a is not None:
                    req.action.CopyFrom(a)
                    f = getattr(a, "from")
                    s = acts[0, j.player_id]
                    s["player_id"], s["from_x"], s["from_y"], s["to_x"], s["to_y"] = j.player_id, f.x, f.y, a.to.x, a.to.y
                    s["move_all"], s["present"] = (0 if a.half else 1), 1
                resp = stub.SubmitAction(req)
                assert resp.success or resp.error_code == common.ERROR_CODE_UNSPECIFIED, resp.error_message
            else:
                m.step(acts)
    for (gid, js), m in zip(games, mirrors):
        g = gs.games[gid]
        served = g.pool.engine.get_state(g.slot, 1)
        want = m.get_state()
        for k in want:
            assert np.array_equal(served[k], want[k]), (gid, k)
        # and the fog-filtered proto view is what the client contract says
        vis, fog = m.visibility()
        for j in js:
            st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)).state
            assert st.turn == int(want["turn"][0]) and len(st.board.tiles) == 64
            for i, t in enumerate(st.board.tiles):
                v, f = bool(vis[0, j.player_id, i]), bool(fog[0, j.player_id, i])
                assert (t.visible, t.fog_of_war) == (v, f)
                if v:
                    assert (t.owner_id, t.army_count) == (int(want["owner"][0, i]), int(want["army"][0, i]))
                else:
                    assert (t.owner_id, t.army_count) == (-1, 0)
                    assert t.type == (common.TILE_TYPE_NORMAL if not f else {1: 2, 2: 3, 3: 4}[int(want["type"][0, i])])
            assert list(st.action_mask) == list(m.get_legal_action_mask(0, j.player_id))
            assert [p.tile_count for p in st.players] == [int(want["owned"][0, p].sum()) for p in range(2)]


def test_experience_stream_records(server, oracle_lib):
    _experience_stream_records(server, oracle_lib)


def _experience_stream_records(server, oracle_lib):
    gs, stub, xstub, _, _ = server
    gid, js = _start(stub, 6, 6, collect=True)
    m = new_engine(oracle_lib, 6, 6, 2, 1)
    m.reset_seeded([1000 + int(gid.split("-")[1])])
    rng = np.random.default_rng(3)
    expected = []
    for turn in range(12):
        prev = m.alloc_outputs_host()
        m.observe(m.outputs(obs=prev["obs"]))
        prev_mask = m.mask(_abi.MASK_SERIALIZER_UDLR)
        acts = np.zeros((1, m.A), _abi.ACTION_DTYPE)
        for j in js:
            st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)).state
            a = _legal_action(st, rng, st.turn)
            req = game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)
            if a is not None and not (turn == 4 and j.player_id == 1):  # player 1 passes on turn 4: no record for it
                req.action.CopyFrom(a)
                f = getattr(a, "from")
                s = acts[0, j.player_id]
                s["player_id"], s["from_x"], s["from_y"], s["to_x"], s["to_y"] = j.player_id, f.x, f.y, a.to.x, a.to.y
                s["move_all"], s["present"] = (0 if a.half else 1), 1
            assert stub.SubmitAction(req).success
        out = m.alloc_outputs_host()
        m.step_fused(acts, m.outputs(**out))
        for p in range(2):
            if out["action_index"][0, p] >= 0:
                expected.append((p, int(out["action_index"][0, p]), out["reward"][0, p], prev["obs"][0, p].copy(),
                                 out["obs"][0, p].copy(), prev_mask[0, p].astype(bool), int(m.get_state()["turn"][0])))
    got = []
    for b in xstub.StreamExperienceBatches(experience.StreamExperiencesRequest(game_ids=[gid], batch_size=5, follow=False)):
        assert len(b.experiences) <= 5
        got.extend(b.experiences)
    assert len(got) == len(expected) == 23
    for x, (p, a, r, s0, s1, mk, turn) in zip(got, expected):
        assert (x.game_id, x.player_id, x.action, x.turn, x.done) == (gid, p, a, turn, False)
        assert np.float32(x.reward).view(np.uint32) == np.float32(r).view(np.uint32)
        assert list(x.state.shape) == [9, 6, 6]
        assert np.array_equal(np.array(x.state.data, np.float32).view(np.uint32), s0.reshape(-1).view(np.uint32))
        assert np.array_equal(np.array(x.next_state.data, np.float32).view(np.uint32), s1.reshape(-1).view(np.uint32))
        assert list(x.action_mask) == list(mk)
    singles = list(xstub.StreamExperiences(experience.StreamExperiencesRequest(player_ids=[1], min_turn=3, follow=False)))
    assert singles and all(x.player_id == 1 and x.turn >= 3 for x in singles)
    stats = xstub.GetExperienceStats(experience.GetExperienceStatsRequest())
    assert stats.total_experiences == 23 and stats.total_games == 1 and stats.experiences_per_player[0] == 12
    dup = xstub.SubmitExperiences(experience.SubmitExperiencesRequest(experiences=got[:3] + got[:1]))
    assert (dup.accepted, dup.rejected) == (3, 1)
    # experience_service.go:381-448, 522-540: request-level checks abort, invalid experiences are counted as rejected
    for bad, msg in ((experience.SubmitExperiencesRequest(), "no experiences provided"),
                     (experience.SubmitExperiencesRequest(experiences=[experience.Experience(player_id=1)]), "game ID required"),
                     (experience.SubmitExperiencesRequest(experiences=[experience.Experience(game_id="g")] * 1001), "too many experiences (max 1000)")):
        with pytest.raises(grpc.RpcError) as e:
            xstub.SubmitExperiences(bad)
        assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT and e.value.details() == msg
    stateless = experience.Experience(game_id=gid, player_id=0)
    empty = experience.Experience(game_id=gid, player_id=0)
    empty.state.shape.extend([9, 6, 6]), empty.next_state.shape.extend([9, 6, 6])
    r = xstub.SubmitExperiences(experience.SubmitExperiencesRequest(experiences=[stateless, empty]))
    assert (r.accepted, r.rejected) == (0, 2)
    with pytest.raises(grpc.RpcError) as e:   # validateStreamRequest :510-520
        list(xstub.StreamExperienceBatches(experience.StreamExperiencesRequest(batch_size=1001)))
    assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT and e.value.details() == "invalid request: batch size too large (max 1000)"


def test_stream_game_updates(server):
    gs, stub, _, ch, _ = server
    gid, js = _start(stub, 8, 8)
    it = stub.StreamGame(game.StreamGameRequest(game_id=gid, player_id=0, player_token=js[0].player_token))
    first = next(it)
    assert first.WhichOneof("update") == "full_state" and first.full_state.turn == 0
    rng = np.random.default_rng(1)
    for j in js:
        st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)).state
        req = game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)
        req.action.CopyFrom(_legal_action(st, rng, 0))
        assert stub.SubmitAction(req).success
    upd = next(it)
    # a 2-move turn touches fewer than N/5 tiles: the server sends a delta (server.go:638-777)
    assert upd.WhichOneof("update") == "delta" and upd.delta.turn == 1
    assert 0 < len(upd.delta.tile_updates) < 64 // 5 and len(upd.delta.player_updates) == 2
    it.cancel()


def test_reference_stream_game_kats(server):
    """stream_test.go:44-176 (TestStreamGame): a stream opened from the lobby gets the full state at once, a game-started
    event when the second player joins, and a full state or delta after the turn both players passed on;
    :178-195 invalid credentials and :197-209 an unknown game are errors."""
    import threading

    gs, stub, _, ch, _ = server
    cfg = game.GameConfig(width=10, height=10, max_players=2, fog_of_war=True)
    gid = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
    j1 = stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="Player1"))
    it = stub.StreamGame(game.StreamGameRequest(game_id=gid, player_id=j1.player_id, player_token=j1.player_token))
    updates, done = [], threading.Event()

    def pump():
        try:
            for u in it:
                updates.append(u)
        except grpc.RpcError:
            pass
        done.set()

    th = threading.Thread(target=pump, daemon=True)
    th.start()

    def wait_for(pred, what):
        import time as _t
        t0 = _t.time()
        while _t.time() - t0 < 5.0:
            if pred():
                return
            _t.sleep(0.01)
        raise AssertionError(what)

    wait_for(lambda: len(updates) >= 1, "initial state")
    assert updates[0].WhichOneof("update") == "full_state"
    j2 = stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="Player2"))
    wait_for(lambda: any(u.WhichOneof("update") == "event" and u.event.WhichOneof("event") == "game_started" for u in updates[1:]),
             "Expected to find game started event")
    n_before = len(updates)
    for j in (j1, j2):   # no action this turn, from both players
        assert stub.SubmitAction(game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)).success
    wait_for(lambda: any(u.WhichOneof("update") in ("full_state", "delta") for u in updates[n_before:]),
             "Expected to receive either full state or delta update after turn processing")
    assert len(updates) > 2
    it.cancel()
    assert done.wait(1.0), "Stream did not finish within timeout"
    # invalid credentials / unknown game
    lobby = stub.CreateGame(game.CreateGameRequest()).game_id
    with pytest.raises(grpc.RpcError) as e:
        next(iter(stub.StreamGame(game.StreamGameRequest(game_id=lobby, player_id=999, player_token="invalid-token"))))
    assert "invalid player credentials" in e.value.details()
    with pytest.raises(grpc.RpcError) as e:
        next(iter(stub.StreamGame(game.StreamGameRequest(game_id="non-existent-game", player_id=0, player_token="token"))))
    assert "not found" in e.value.details()


def test_validate_move_precedence():
    W = H = 3
    owner = [0, -1, -1, -1, -1, -1, -1, -1, 1]
    army = [5, 0, 0, 0, 0, 0, 0, 0, 1]
    typ = [1, 3, 0, 0, 0, 0, 0, 0, 1]
    v = lambda *a: validate_move(owner, army, typ, W, H, *a)
    assert v(0, -1, 0, 0, 0) == _abi.STEP_INVALID_COORDINATES and v(0, 0, 0, 3, 0) == _abi.STEP_INVALID_COORDINATES
    assert v(0, 0, 0, 0, 0) == _abi.STEP_MOVE_TO_SELF and v(0, 0, 0, 1, 1) == _abi.STEP_NOT_ADJACENT
    assert v(1, 0, 0, 0, 1) == _abi.STEP_NOT_OWNED and v(1, 2, 2, 2, 1) == _abi.STEP_INSUFFICIENT_ARMY
    assert v(0, 0, 0, 1, 0) == _abi.STEP_TARGET_IS_MOUNTAIN and v(0, 0, 0, 0, 1) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/python/generals_pb"), reason="reference stubs not on this box")
def test_reference_generated_stubs_talk_to_this_server(server):
    """The reference's own protoc output (python/generals_pb) as the client: wire compatibility."""
    gs, _, _, _, port = server
    sys.path.insert(0, "/root/reference/python")
    try:
        from generals_pb.common.v1 import common_pb2
        from generals_pb.game.v1 import game_pb2, game_pb2_grpc
    finally:
        sys.path.pop(0)
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = game_pb2_grpc.GameServiceStub(ch)
    gid = stub.CreateGame(game_pb2.CreateGameRequest(config=game_pb2.GameConfig(width=7, height=5, max_players=2, fog_of_war=True))).game_id
    j = [stub.JoinGame(game_pb2.JoinGameRequest(game_id=gid, player_name=n)) for n in ("p0", "p1")]
    st = stub.GetGameState(game_pb2.GetGameStateRequest(game_id=gid, player_id=0, player_token=j[0].player_token)).state
    assert st.status == common_pb2.GAME_STATUS_IN_PROGRESS and st.current_phase == common_pb2.GAME_PHASE_RUNNING
    assert (st.board.width, st.board.height, len(st.board.tiles), len(st.action_mask)) == (7, 5, 35, 140)
    gens = [t for t in st.board.tiles if t.type == common_pb2.TILE_TYPE_GENERAL and t.visible]
    assert len(gens) == 1 and gens[0].owner_id == 0 and gens[0].army_count == 2  # engine_initializer: general army 2... 
    k = list(st.action_mask).index(True)
    t, d = k // 4, k % 4
    a = game_pb2.Action(type=common_pb2.ACTION_TYPE_MOVE, turn_number=0, half=False)
    getattr(a, "from").x, getattr(a, "from").y = t % 7, t // 7
    a.to.x, a.to.y = t % 7 + DIRS[d][0], t // 7 + DIRS[d][1]
    target = a.to.y * 7 + a.to.x
    before = (st.board.tiles[target].owner_id, st.board.tiles[target].army_count)
    assert stub.SubmitAction(game_pb2.SubmitActionRequest(game_id=gid, player_id=0, player_token=j[0].player_token, action=a)).success
    assert stub.SubmitAction(game_pb2.SubmitActionRequest(game_id=gid, player_id=1, player_token=j[1].player_token)).success
    st = stub.GetGameState(game_pb2.GetGameStateRequest(game_id=gid, player_id=0, player_token=j[0].player_token)).state
    after = (st.board.tiles[target].owner_id, st.board.tiles[target].army_count)
    # one army moved: an empty tile is captured with 1, a defended one (a 40-army city) loses 1
    assert st.turn == 1 and after == ((0, 1) if before[1] == 0 else (before[0], before[1] - 1))
    ch.close()


# ---- the reference's own server tests, transliterated (internal/grpc/gameserver/*_test.go) ------
def test_reference_server_kats(oracle_lib):
    srv, gs, port = serve("127.0.0.1:0", lib=oracle_lib, slots_per_pool=4, seed=5, max_games=7)
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = Stub(ch, GAME)
    try:
        # server_test.go:50-82 TestCreateGame
        r = stub.CreateGame(game.CreateGameRequest(config=game.GameConfig(width=20, height=20, max_players=4, fog_of_war=True)))
        assert r.game_id and (r.config.width, r.config.height, r.config.max_players) == (20, 20, 4)
        r2 = stub.CreateGame(game.CreateGameRequest())
        assert r2.game_id and r2.game_id != r.game_id
        assert (r2.config.width, r2.config.height, r2.config.max_players) == (20, 20, 2)
        # server_test.go:84-151 TestJoinGame
        gid = stub.CreateGame(game.CreateGameRequest(config=game.GameConfig(width=20, height=20, max_players=2))).game_id
        j = stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="player-1"))
        assert j.player_id == 0 and j.player_token and j.initial_state.game_id == gid
        assert j.initial_state.status == common.GAME_STATUS_WAITING
        j2 = stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="player-1"))
        assert (j2.player_id, j2.player_token) == (0, j.player_token)
        j3 = stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="player-2"))
        assert j3.player_id == 1 and j3.player_token and j3.player_token != j.player_token
        assert j3.initial_state.status == common.GAME_STATUS_IN_PROGRESS
        with pytest.raises(grpc.RpcError) as e:
            stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name="player-3"))
        assert "cannot join game" in e.value.details()
        with pytest.raises(grpc.RpcError) as e:
            stub.JoinGame(game.JoinGameRequest(game_id="non-existent", player_name="player-4"))
        assert "not found" in e.value.details()
        # server_test.go:153-211 TestGetGameState: a lobby game shows the placeholder board
        g10 = stub.CreateGame(game.CreateGameRequest(config=game.GameConfig(width=10, height=10, max_players=2))).game_id
        jj = stub.JoinGame(game.JoinGameRequest(game_id=g10, player_name="test-player"))
        st = stub.GetGameState(game.GetGameStateRequest(game_id=g10, player_id=jj.player_id, player_token=jj.player_token)).state
        assert st.game_id == g10 and st.status == common.GAME_STATUS_WAITING
        assert (st.board.width, st.board.height, len(st.board.tiles)) == (10, 10, 100)
        with pytest.raises(grpc.RpcError) as e:
            stub.GetGameState(game.GetGameStateRequest(game_id=g10, player_id=jj.player_id, player_token="invalid-token"))
        assert "invalid player credentials" in e.value.details()
        with pytest.raises(grpc.RpcError) as e:
            stub.GetGameState(game.GetGameStateRequest(game_id="non-existent", player_id=0, player_token="t"))
        assert "not found" in e.value.details()
        # idempotency_test.go:98-147 TestIdempotencyForErrors: a lobby game refuses actions, and caches the refusal
        a0 = game.Action(turn_number=0)
        e1 = stub.SubmitAction(game.SubmitActionRequest(game_id=g10, player_id=0, player_token=jj.player_token, action=a0,
                                                        idempotency_key="error-test-key"))
        e2 = stub.SubmitAction(game.SubmitActionRequest(game_id=g10, player_id=0, player_token=jj.player_token, action=a0,
                                                        idempotency_key="error-test-key"))
        assert not e1.success and e1.error_code == common.ERROR_CODE_INVALID_PHASE
        assert (e2.success, e2.error_code, e2.error_message) == (e1.success, e1.error_code, e1.error_message)
        # idempotency_test.go:14-96 TestIdempotencyWithSameKey + :149-200 TestIdempotencyAcrossPlayers
        jk = stub.JoinGame(game.JoinGameRequest(game_id=g10, player_name="Player2"))
        wait = game.Action(type=common.ACTION_TYPE_UNSPECIFIED, turn_number=0)
        s1 = stub.SubmitAction(game.SubmitActionRequest(game_id=g10, player_id=0, player_token=jj.player_token, action=wait,
                                                        idempotency_key="shared-key-123"))
        s2 = stub.SubmitAction(game.SubmitActionRequest(game_id=g10, player_id=0, player_token=jj.player_token, action=wait,
                                                        idempotency_key="shared-key-123"))
        assert s1.success and (s2.success, s2.error_code, s2.error_message, s2.next_turn_number) == (
            s1.success, s1.error_code, s1.error_message, s1.next_turn_number)
        s3 = stub.SubmitAction(game.SubmitActionRequest(game_id=g10, player_id=1, player_token=jk.player_token, action=wait,
                                                        idempotency_key="shared-key-123"))
        assert s3.success, "different players may use the same idempotency key"
        st = stub.GetGameState(game.GetGameStateRequest(game_id=g10, player_id=0, player_token=jj.player_token)).state
        assert st.turn == 1, "the game advanced only one turn"
        # stream_test.go:177-212
        with pytest.raises(grpc.RpcError) as e:
            next(stub.StreamGame(game.StreamGameRequest(game_id=g10, player_id=999, player_token="invalid-token")))
        assert "invalid player credentials" in e.value.details()
        with pytest.raises(grpc.RpcError) as e:
            next(stub.StreamGame(game.StreamGameRequest(game_id="non-existent-game", player_id=0, player_token="token")))
        assert "not found" in e.value.details()
        # max_games_test.go:13-46 (this server was started with max_games=7; 4 exist)
        for _ in range(3):
            assert stub.CreateGame(game.CreateGameRequest()).game_id
        with pytest.raises(grpc.RpcError) as e:
            stub.CreateGame(game.CreateGameRequest())
        assert e.value.code() == grpc.StatusCode.RESOURCE_EXHAUSTED and "server at capacity" in e.value.details()
    finally:
        ch.close()
        srv.stop(0)
        gs.close()


@pytest.mark.skipif(not os.path.isfile("/root/reference/python/experience_stream_client.py"), reason="reference client not on this box")
def test_reference_experience_stream_client_consumes_this_server(server):
    """The reference's OWN trainer-side client (python/experience_stream_client.py, unmodified) streams
    batches from this server and decodes them into the numpy records its trainers use."""
    gs, stub, _, _, port = server
    gid, js = _start(stub, 6, 6, collect=True)
    rng = np.random.default_rng(11)
    for turn in range(10):
        for j in js:
            st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)).state
            req = game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)
            a = _legal_action(st, rng, st.turn)
            if a is not None:
                req.action.CopyFrom(a)
            assert stub.SubmitAction(req).success
    sys.path.insert(0, "/root/reference/python")
    try:
        from experience_stream_client import ExperienceConfig, ExperienceStreamClient
    finally:
        sys.path.pop(0)
    client = ExperienceStreamClient(ExperienceConfig(server_address=f"127.0.0.1:{port}", game_ids=[gid], batch_size=8, follow=True))
    client.connect()
    client.start_streaming()
    try:
        batch = client.get_batch(16, timeout=10.0)
    finally:
        client.stop_event.set()
        client.disconnect()
    assert len(batch) == 16
    for x in batch:
        assert x["game_id"] == gid and x["state"].shape == (9, 6, 6) and x["next_state"].shape == (9, 6, 6)
        assert x["state"].dtype == np.float32 and x["action_mask"].shape == (6 * 6 * 4,)
        assert 0 <= x["action"] < 144 and x["action_mask"][x["action"]], "the recorded action was legal in the recorded state"
        assert x["state"][7].sum() > 0 and np.array_equal(x["state"][7] + x["state"][8], np.ones((6, 6), np.float32))
    assert client.get_stats()["total_batches"] >= 2


# ---- env slots and finished games are released (GameManager.cleanupGames, game_manager.go:236-343) -------------------
def _duel_board():
    """5x5: player 0's stack at (3,4) stands next to player 1's general at (4,4): one move ends the game."""
    owner, army, typ = np.full(25, -1, np.int32), np.zeros(25, np.int32), np.zeros(25, np.int32)
    for (x, y, o, a, t) in ((0, 0, 0, 3, 1), (3, 4, 0, 25, 0), (4, 4, 1, 4, 1), (1, 0, 1, 5, 0)):
        owner[y * 5 + x], army[y * 5 + x], typ[y * 5 + x] = o, a, t
    return owner[None], army[None], typ[None]


def _play_duel(gs, stub):
    gid, js = _start(stub, 5, 5)
    g = gs.games[gid]
    with g.pool.lock:
        g.pool.engine.reset_boards(*_duel_board(), env_ids=[g.slot])
        g.pool.refresh()
    strike = game.Action(type=common.ACTION_TYPE_MOVE, turn_number=0)
    getattr(strike, "from").x, getattr(strike, "from").y = 3, 4
    strike.to.x, strike.to.y = 4, 4
    assert stub.SubmitAction(game.SubmitActionRequest(game_id=gid, player_id=0, player_token=js[0].player_token, action=strike)).success
    assert stub.SubmitAction(game.SubmitActionRequest(game_id=gid, player_id=1, player_token=js[1].player_token)).success
    return gid, js


@pytest.mark.skipif(not os.path.isdir("/root/reference/python/generals_gym"), reason="the reference's client is not on this box")
def test_reference_gym_env_plays_against_this_server(server):
    """The reference's UNMODIFIED GeneralsEnv (python/generals_gym/generals_env.py — its real gRPC client and its own
    generated stubs) plays whole episodes against this server, as python/test_gym_minimal.py and test_gym_env.py do
    against the Go server.  After reset() and after every step() the observation and the action mask the client derived
    from the GameState messages must equal, bit for bit, the engine's own gym read-outs (grl_gym_observe_envs) of the
    game's env slot — the path GeneralsVecEnv hands out.  (Plane 7 is the client's own step counter over max_turns,
    generals_env.py:334-336, so it is compared with that.)  The opponent is an `opponent_agent` of the test that always
    submits — a random legal move or an empty request — so every turn runs as soon as both players have submitted, not
    on the turn timer the reference's 50 ms wait races with."""
    import random
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    import make_gym_fixtures as mk

    generals_env, game_pb2, common_pb2 = mk.load_reference()
    gs, stub, _, _, port = server
    rnd = random.Random(11)
    W = H = 6

    class Opponent:
        env = None

        def select_action(self, _agent_view):
            env = self.env
            st = env.stub.GetGameState(game_pb2.GetGameStateRequest(game_id=env.game_id, player_id=env.opponent_id,
                                                                    player_token=env.opponent_token)).state
            moves = []
            for t, tile in enumerate(st.board.tiles):
                if tile.owner_id != env.opponent_id or tile.army_count <= 1:
                    continue
                x, y = t % W, t // W
                for dx, dy in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                    nx, ny = x + dx, y + dy
                    if 0 <= nx < W and 0 <= ny < H and st.board.tiles[ny * W + nx].type != common_pb2.TILE_TYPE_MOUNTAIN:
                        a = game_pb2.Action(type=common_pb2.ACTION_TYPE_MOVE, half=False, turn_number=st.turn)
                        getattr(a, "from").CopyFrom(common_pb2.Coordinate(x=x, y=y))
                        a.to.CopyFrom(common_pb2.Coordinate(x=nx, y=ny))
                        moves.append(a)
            # submitted here, with the opponent's player id: the client's own submission for an opponent_agent leaves
            # player_id at 0 (generals_env.py:247-251), which the server — the Go one too, action_validator.go:80-92 —
            # refuses as invalid credentials
            req = game_pb2.SubmitActionRequest(game_id=env.game_id, player_id=env.opponent_id, player_token=env.opponent_token)
            if moves:
                req.action.CopyFrom(rnd.choice(moves))   # else: no action this turn
            env.stub.SubmitAction(req)   # (a turn in which a move fails in the engine answers "failed to process turn")
            return None

    class TurnStampingStub:
        """The client never fills Action.turn_number (generals_env.py:430-441), so from turn 1 on every move it submits is
        refused in-band as INVALID_TURN — by the Go server too (action_validator.go:102-108) — and it does not look at the
        response.  This shim stamps the turn of the last state the client fetched, which is all a working client would
        add; everything else on the wire is the reference client's own."""

        def __init__(self, stub):
            self._stub, self.turn, self.refused, self.failed_turn = stub, 0, 0, False

        def __getattr__(self, name):
            return getattr(self._stub, name)

        def GetGameState(self, req, **kw):
            r = self._stub.GetGameState(req, **kw)
            self.turn = r.state.turn
            return r

        def SubmitAction(self, req, **kw):
            if req.HasField("action"):
                req.action.turn_number = self.turn
            r = self._stub.SubmitAction(req, **kw)
            self.refused += r.error_code in (common.ERROR_CODE_INVALID_TURN, common.ERROR_CODE_INVALID_PLAYER)
            self.failed_turn |= "failed to process turn" in r.error_message
            return r

    opp = Opponent()
    env = generals_env.GeneralsEnv(server_address=f"127.0.0.1:{port}", board_width=W, board_height=H, max_players=2,
                                   fog_of_war=True, max_turns=40, turn_time_ms=600000, opponent_agent=opp)
    opp.env = env
    env.stub = TurnStampingStub(env.stub)
    rng = np.random.default_rng(5)
    steps = finished = 0
    try:
        for episode in range(3):
            obs, info = env.reset()
            g = gs.games[info["game_id"]]

            def check(obs, mask, what):
                pool = g.pool
                S, P, N = pool.S, pool.P, pool.N
                o = np.zeros((S, P, 9, H, W), np.float32)
                m = np.zeros((S, P, N * 5), np.uint8)
                st = np.zeros((S, P, 4), np.int32)
                with pool.lock:
                    pool.engine.gym_observe_envs(env.max_turns, [g.slot], o, m, st)
                    turn = int(pool.engine.get_state(g.slot, 1)["turn"][0])
                o, m = o[g.slot, env.player_id], m[g.slot, env.player_id].astype(bool)
                assert obs.dtype == np.float32 and obs.shape == (9, H, W)
                for plane in (0, 1, 2, 3, 4, 5, 6, 8):
                    assert np.array_equal(obs[plane].view(np.uint32), o[plane].view(np.uint32)), f"plane {plane}, {what}"
                assert (obs[7] == np.float32(min(env.turn_count / env.max_turns, 1.0))).all(), f"turn plane, {what}"
                assert np.array_equal(np.asarray(mask, bool), m), f"mask, {what}"
                return turn

            assert check(obs, info["valid_actions_mask"], "after reset") == 0
            for t in range(40):
                valid = np.flatnonzero(info["valid_actions_mask"])
                if len(valid) == 0:
                    break
                obs, reward, terminated, truncated, info = env.step(int(valid[rng.integers(len(valid))]))
                steps += 1
                assert np.isfinite(reward)
                if g.slot < 0:        # the game ended and the server released the env slot
                    assert terminated
                    finished += 1
                    break
                turn = check(obs, info["valid_actions_mask"], f"episode {episode} step {t}")
                assert info["turn"] == env.turn_count == t + 1 and turn == t + 1, "every step ran one turn"
                if env.stub.failed_turn:
                    # a move that was legal when submitted failed inside the turn (its source tile fell to the other
                    # player's move first): Engine.Step errors, the server answers "failed to process turn" and — as in
                    # the reference, game_manager.go:602-608 — keeps its own turn counter where it was, one behind the
                    # engine's, so every later action stamped with the state's turn is refused.  The episode ends here.
                    env.stub.failed_turn = False
                    break
                if terminated or truncated:
                    break
    finally:
        env.close()
    assert steps >= 60 and env.stub.refused == 0


@pytest.mark.skipif(not os.path.isfile("/root/reference/python/test_grpc_client.py"), reason="the reference's scripts are not on this box")
def test_reference_client_scripts_run_unmodified(oracle_lib, tmp_path):  # noqa: C901
    """python/test_grpc_client.py, python/test_gym_minimal.py and python/test_parallel_env.py — the reference's own smoke
    scripts, run as they are (they dial localhost:50051) against this server: connection, lifecycle, a move from the
    general, a GeneralsEnv reset and steps, and the reference's ParallelEnvPool collecting four episodes.  The only thing added is a `gymnasium` package with the four names generals_env.py imports (absent from the
    image)."""
    import subprocess
    import sys

    try:
        srv, gs, port = serve("localhost:50051", lib=oracle_lib, slots_per_pool=8, seed=7)
    except Exception as exc:   # the port is taken on this machine
        pytest.skip(f"cannot listen on localhost:50051: {exc}")
    if port != 50051:
        srv.stop(0)
        gs.close()
        pytest.skip("cannot listen on localhost:50051")
    stub_pkg = tmp_path / "gymnasium"
    stub_pkg.mkdir()
    (stub_pkg / "__init__.py").write_text(
        "from . import spaces\n"
        "class Env:\n    def reset(self, seed=None, options=None):\n        return None\n"
        "def register(**kw):\n    pass\n")
    (stub_pkg / "spaces.py").write_text(
        "class Box:\n    def __init__(self, low, high, shape, dtype):\n"
        "        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype\n"
        "class Discrete:\n    def __init__(self, n):\n        self.n = int(n)\n")
    envv = dict(os.environ, PYTHONPATH=str(tmp_path), PYTHONDONTWRITEBYTECODE="1")
    try:
        for script, needles in (("test_grpc_client.py", ["Server is responsive. Created test game: game-1", "Player 1 joined: ID=0",
                                                         "Player 2 joined: ID=1", "Status: GAME_STATUS_IN_PROGRESS", "Number of tiles: 100",
                                                         "All tests completed successfully!"]),
                                ("test_gym_minimal.py", ["Reset successful, game_id: game-", "Observation shape: (9, 5, 5)", "Step 1: Taking action",
                                                         "Test complete!"]),
                                # the reference's ParallelEnvPool (two worker threads, two gRPC envs) collecting into its
                                # ReplayBuffer from this server: about 13 s of the client's own waits
                                ("test_parallel_env.py", ["Game server is running", "4 episodes", "All workers stopped cleanly",
                                                          "Sampled 32 transitions with valid shapes/types", "All tests passed"])):
            proc = subprocess.run([sys.executable, script], cwd="/root/reference/python", env=envv, stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True, timeout=120)
            assert proc.returncode == 0, proc.stdout[-2000:]
            for needle in needles:
                assert needle in proc.stdout, f"{script}: {needle!r} missing from\n{proc.stdout[-2000:]}"
            assert "✗" not in proc.stdout or "Move failed" in proc.stdout, proc.stdout[-2000:]
            assert "Error during step" not in proc.stdout, proc.stdout[-2000:]
    finally:
        srv.stop(0)
        gs.close()


@pytest.mark.skipif(not os.path.isfile("/root/reference/python/test_gym_env.py"), reason="the reference's scripts are not on this box")
def test_more_reference_scripts_run_unmodified(oracle_lib, tmp_path):
    """python/test_gym_env.py (GeneralsEnv: reset, steps, action masks, observation channels), python/simple_game_client.py
    (a demo match with board rendering) and python/examples/test_experience_collection.py (a game with
    collect_experiences, ten turns of moves from both players, GetExperienceStats and StreamExperiences) — run as they
    are against this server on localhost:50051.  The last one ends in the script's own `sys.exit` without `import sys`
    after it has printed its verdict; python/simple_experience_test.py cannot run against any server (it reads
    game_pb2.ActionType, which the reference's stubs do not have)."""
    import subprocess
    import sys

    try:
        srv, gs, port = serve("localhost:50051", lib=oracle_lib, slots_per_pool=4, seed=7)
    except Exception as exc:   # the port is taken on this machine
        pytest.skip(f"cannot listen on localhost:50051: {exc}")
    if port != 50051:
        srv.stop(0)
        gs.close()
        pytest.skip("cannot listen on localhost:50051")
    stub_pkg = tmp_path / "gymnasium"
    stub_pkg.mkdir()
    (stub_pkg / "__init__.py").write_text(
        "from . import spaces\n"
        "class Env:\n    def reset(self, seed=None, options=None):\n        return None\n"
        "def register(**kw):\n    pass\n")
    (stub_pkg / "spaces.py").write_text(
        "class Box:\n    def __init__(self, low, high, shape, dtype):\n"
        "        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype\n"
        "class Discrete:\n    def __init__(self, n):\n        self.n = int(n)\n")
    envv = dict(os.environ, PYTHONPATH=f"{tmp_path}:/root/reference/python", PYTHONDONTWRITEBYTECODE="1")
    try:
        for script, rcs, needles in (
                ("test_gym_env.py", (0,), ["Game server is running", "Environment created successfully", "Reset successful",
                                           "All tests passed!", "Environment closed successfully", "Observation channels test complete"]),
                ("simple_game_client.py", (0,), ["Demo complete!"]),
                ("examples/test_experience_collection.py", (0, 1), ["Experiences collected: 20",
                                                                    "SUCCESS: Experience collection is working!", "Experience 1: Player 0"])):
            proc = subprocess.run([sys.executable, script], cwd="/root/reference/python", env=envv, stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True, timeout=120)
            assert proc.returncode in rcs, proc.stdout[-2000:]
            if proc.returncode == 1:
                assert "NameError: name 'sys' is not defined" in proc.stdout, proc.stdout[-2000:]
            for needle in needles:
                assert needle in proc.stdout, f"{script}: {needle!r} missing from\n{proc.stdout[-2000:]}"
            assert "✗" not in proc.stdout and "Error during step" not in proc.stdout, proc.stdout[-2000:]
        # the scripts abandoned more running games than one pool has slots: the pools grew instead of refusing them
        assert gs.more_pools
    finally:
        srv.stop(0)
        gs.close()


@pytest.mark.skipif(not os.path.isdir("/root/reference/python/generals_agent"), reason="the reference's agent SDK is not on this box")
def test_reference_agent_sdk_plays_a_match_against_this_server(server):
    """BASELINE config 0's client side: two of the reference's RandomAgents (python/generals_agent: AgentRunner,
    GameClient, GameSession, ExponentialBackoffPolling — unmodified, in their own threads) play some thirty turns of a 5x5 match (or all of it, when a general falls first) against
    this server, as scripts/run_random_match.py does against the Go server (that script itself builds GameConfig with a
    keyword the SDK does not have and cannot start).  Every move the agents submit carries the turn number of the state
    they polled, so — unlike the gym client's — their moves are accepted turn after turn, until a general falls."""
    import sys
    import threading

    sys.dont_write_bytecode = True
    if "/root/reference/python" not in sys.path:
        sys.path.insert(0, "/root/reference/python")
    from generals_agent import AgentRunner, ExponentialBackoffPolling, GameClient, GameConfig, GameConnection, RandomAgent

    gs, _, _, _, port = server
    addr = f"127.0.0.1:{port}"
    conn = GameConnection(addr)
    gid = GameClient(conn).create_game(GameConfig(width=5, height=5))
    res, runners = {}, {}

    def play(name):
        agent = RandomAgent(name=name)
        runner = runners[name] = AgentRunner(agent, server_address=addr, polling_strategy=ExponentialBackoffPolling(base_interval=0.01),
                                             enable_logging=False)
        try:
            runner.join_game(gid)
            runner.run(wait_for_players=2, timeout=30)
            res[name] = agent.move_count
        except Exception as exc:  # noqa: BLE001
            res[name] = repr(exc)
        finally:
            runner.disconnect()

    threads = [threading.Thread(target=play, args=(n,), daemon=True) for n in ("RandomAgent1", "RandomAgent2")]
    for th in threads:
        th.start()
    import time as _t
    t0 = _t.time()
    g = gs.games[gid]
    while _t.time() - t0 < 30 and g.phase != common.GAME_PHASE_ENDED and g.current_turn < 32:
        _t.sleep(0.05)
    for r in list(runners.values()):   # some thirty turns are enough: the match itself may take minutes of random play
        r.stop()
    for th in threads:
        th.join(20)
    conn.disconnect()
    assert len(res) == 2 and all(isinstance(v, int) for v in res.values()), res
    if g.phase == common.GAME_PHASE_ENDED:   # the agents are unseeded: a general may fall within a few turns
        assert g.slot == -1 and sorted(int(a) for a in g.final[0]["alive"][0]) == [0, 1], "one general fell"
        assert g.current_turn >= 1 and sum(res.values()) >= 1, (g.current_turn, res)
    else:
        assert all(v >= 10 for v in res.values()), res
        assert g.current_turn >= 20


def test_reference_experience_streaming_kats(server):
    """experience_streaming_test.go transliterated over the wire (the Go tests drive the service object with a mock
    stream; here the same requests go through gRPC and the experiences enter the way BufferManager buffers fill:
    through the store the collectors write to).
    :34-116 TestStreamExperienceBatches — a FOLLOWING stream started first, 100 experiences of one game added while it
    runs, cancelled: batches arrive, each with a stream id, a creation time and batch_id > 0 (ids count from 1,
    experience_service.go:342), at most batch_size experiences, metadata batch_size = its length (:345-347).
    :184-239 TestBatchProcessor — 12 experiences, batch size 5: at least two batches, none above 5, all 12 delivered.
    :241-343 TestStreamWithFilters — two games x three players x five turns; FilterByGame sees only game-A,
    FilterByPlayer only players 1 and 2."""
    import threading
    import time as _t
    import uuid as _uuid

    gs, _, xstub, _, _ = server

    def follow(req, seconds):
        got = []
        call = xstub.StreamExperienceBatches(req)

        def pump():
            try:
                for b in call:
                    got.append(b)
            except grpc.RpcError as exc:            # the cancelled stream ends the iteration
                assert exc.code() == grpc.StatusCode.CANCELLED
        th = threading.Thread(target=pump, daemon=True)
        th.start()
        return got, call, th

    def exp(game_id, player, turn, reward=0.0, done=False):
        return experience.Experience(experience_id=str(_uuid.uuid4()), game_id=game_id, player_id=player, turn=turn,
                                     reward=reward, done=done)

    # TestStreamExperienceBatches
    got, call, th = follow(experience.StreamExperiencesRequest(game_ids=["test-game-1"], batch_size=10, follow=True), 0)
    _t.sleep(0.1)
    for i in range(100):
        gs.store.add([exp("test-game-1", i % 2, i, float(i), i == 99)])
        if i % 10 == 0:
            _t.sleep(0.01)
    t0 = _t.time()
    while sum(len(b.experiences) for b in got) < 100 and _t.time() - t0 < 15:
        _t.sleep(0.02)
    call.cancel()
    th.join(3)
    assert not th.is_alive() and len(got) > 0
    assert len({b.stream_id for b in got}) == 1 and got[0].stream_id
    assert [b.batch_id for b in got] == list(range(1, len(got) + 1))
    for b in got:
        assert b.HasField("created_at") and 0 < len(b.experiences) <= 10
        assert b.metadata["batch_size"] == str(len(b.experiences)) and "compression" not in b.metadata
    assert [x.turn for b in got for x in b.experiences] == list(range(100))   # every experience, in order
    assert got[-1].experiences[-1].done

    # TestBatchProcessor: 12 experiences through batch size 5 (100 ms flush)
    gs.store.add([exp("test-game", i % 2, i) for i in range(12)])
    batches = list(xstub.StreamExperienceBatches(experience.StreamExperiencesRequest(game_ids=["test-game"], batch_size=5,
                                                                                     enable_compression=True)))
    assert len(batches) >= 2 and all(len(b.experiences) <= 5 for b in batches)
    assert sum(len(b.experiences) for b in batches) == 12
    assert all(b.metadata["compression"] == "none" for b in batches)            # experience_service.go:350-353

    # TestStreamWithFilters
    for g_id in ("game-A", "game-B"):
        for player in (1, 2, 3):
            gs.store.add([exp(g_id, player, turn, float(turn) * player) for turn in range(5)])
    def until(got, n, seconds=15.0):   # the Go test sleeps 500 ms; a loaded host gets as long as it needs
        t0 = _t.time()
        while sum(len(b.experiences) for b in got) < n and _t.time() - t0 < seconds:
            _t.sleep(0.02)
        _t.sleep(0.2)                  # anything beyond the expected count would arrive now
        return [x for b in got for x in b.experiences]

    got, call, th = follow(experience.StreamExperiencesRequest(game_ids=["game-A"], batch_size=5, follow=True), 0)
    xs = until(got, 15)
    call.cancel()
    th.join(3)
    assert len(xs) == 15 and all(x.game_id == "game-A" for x in xs)
    got, call, th = follow(experience.StreamExperiencesRequest(player_ids=[1, 2], batch_size=5, follow=True), 0)
    xs = until(got, 76)                # 20 of games A/B + player 1's 50 of test-game-1 and 6 of test-game
    call.cancel()
    th.join(3)
    assert xs and all(x.player_id in (1, 2) for x in xs)
    assert sum(1 for x in xs if x.game_id in ("game-A", "game-B")) == 20 and len(xs) == 76


def test_reference_mutex_kats(oracle_lib):
    """mutex_test.go transliterated (threads for goroutines).
    :13-87 TestNoDeadlock — a manager capped at ten games; ten concurrent creators whose games are then marked idle for
    twice the abandoned-game timeout; five cleanup passes run while another thread keeps creating and reading games:
    everything finishes within five seconds.
    :90-143 TestConcurrentGameAccess — ten concurrent action submissions (empty requests: the player passes) and ten
    concurrent state reads against one running game finish within two seconds."""
    import threading
    import time as _t

    srv, gs, port = serve("127.0.0.1:0", lib=oracle_lib, slots_per_pool=4, seed=9, max_games=10)
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = Stub(ch, GAME)
    cfg = game.GameConfig(width=10, height=10, max_players=2)
    try:
        def create_idle():
            try:
                gid = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
            except grpc.RpcError:
                return                                  # at capacity: skipped, as in the Go test
            g = gs.games[gid]
            with g.mu:
                g.last_activity = _t.time() - 2 * gs.abandoned_game_timeout

        ths = [threading.Thread(target=create_idle) for _ in range(10)]
        for th in ths:
            th.start()
        for th in ths:
            th.join(5)
        assert len(gs.games) == 10
        removed = []

        def cleaner():
            for _ in range(5):
                removed.append(gs.cleanup_games())
                _t.sleep(0.01)

        def creator():
            for _ in range(5):
                try:
                    gid = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
                    g = gs.games.get(gid)
                    if g is not None:
                        with g.mu:
                            _ = g.phase
                except grpc.RpcError:
                    pass
                _t.sleep(0.01)

        a, b = threading.Thread(target=cleaner), threading.Thread(target=creator)
        t0 = _t.time()
        a.start(), b.start()
        a.join(5), b.join(5)
        assert not a.is_alive() and not b.is_alive() and _t.time() - t0 < 5, "deadlock: cleanup did not complete"
        assert sum(removed) >= 10 and len(gs.games) <= 5         # the ten idle games went; the new ones are active

        # TestConcurrentGameAccess
        gid = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
        js = [stub.JoinGame(game.JoinGameRequest(game_id=gid, player_name=n)) for n in ("alice", "bob")]
        errors = []

        def submit(i):
            j = js[i % 2]
            try:
                stub.SubmitAction(game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token))
                stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token))
            except Exception as exc:  # noqa: BLE001
                errors.append(repr(exc))

        def read(i):
            g = gs.games[gid]
            with g.mu:
                _ = (g.current_turn, g.phase)

        ths = [threading.Thread(target=submit, args=(i,)) for i in range(10)] + [threading.Thread(target=read, args=(i,)) for i in range(10)]
        t0 = _t.time()
        for th in ths:
            th.start()
        for th in ths:
            th.join(2)
        assert not any(th.is_alive() for th in ths) and _t.time() - t0 < 2, "concurrent access took too long"
        assert not errors, errors
        g = gs.games[gid]
        assert g.phase == common.GAME_PHASE_RUNNING and 1 <= g.current_turn <= 5   # ten passes, two per turn at most
    finally:
        ch.close()
        srv.stop(0)
        gs.close()


def test_more_games_than_env_slots_through_one_server(oracle_lib):
    """A gym client creates a new game on every reset() (generals_env.py:167-177): a server must outlive its pool size."""
    srv, gs, port = serve("127.0.0.1:0", lib=oracle_lib, slots_per_pool=3, seed=5, max_games=8, finished_game_ttl=600.0)
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = Stub(ch, GAME)
    try:
        finished = []
        for k in range(8):                       # 8 games through 3 slots
            gid, js = _play_duel(gs, stub)
            g = gs.games[gid]
            assert g.phase == common.GAME_PHASE_ENDED and g.slot == -1, "the slot goes back when the game ends"
            st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=1, player_token=js[1].player_token)).state
            # the finished game keeps answering from its frozen final state
            assert (st.status, st.current_phase, st.winner_id, st.turn) == (common.GAME_STATUS_FINISHED, common.GAME_PHASE_ENDED, 0, 1)
            assert [p.status for p in st.players] == [common.PLAYER_STATUS_ACTIVE, common.PLAYER_STATUS_ELIMINATED]
            r = stub.SubmitAction(game.SubmitActionRequest(game_id=gid, player_id=0, player_token=js[0].player_token))
            assert r.error_code == common.ERROR_CODE_GAME_OVER
            finished.append(gid)
        assert len(gs.pools[(5, 5, 2)].free) == 3
        # max_games counts finished games until their TTL passes (cleanupGames), then they leave the map
        with pytest.raises(grpc.RpcError) as e:
            stub.CreateGame(game.CreateGameRequest(config=game.GameConfig(width=5, height=5, max_players=2)))
        assert e.value.code() == grpc.StatusCode.RESOURCE_EXHAUSTED
        import time as _t
        assert gs.cleanup_games(_t.time() + 601.0) == 8 and not gs.games
        gid, _ = _play_duel(gs, stub)
        assert gid == "game-9"
        # abandoned games (no activity for abandoned_game_timeout) release their slots too
        for _ in range(3):
            _start(stub, 5, 5)
        assert len(gs.pools[(5, 5, 2)].free) == 0
        assert gs.cleanup_games(_t.time() + 1801.0) == 4
        assert len(gs.pools[(5, 5, 2)].free) == 3 and not gs.games
        _play_duel(gs, stub)
    finally:
        ch.close()
        srv.stop(0)
        gs.close()


def test_idempotency_entries_expire_after_a_day(server):
    """IdempotencyManager.Check (idempotency.go:36-60): a cached response answers a repeated key for 24 hours; after that
    the request is processed again."""
    import time as _t

    gs, stub, _, _, _ = server
    gid, js = _start(stub, 5, 5)
    req = game.SubmitActionRequest(game_id=gid, player_id=0, player_token=js[0].player_token, idempotency_key="k-1")
    assert stub.SubmitAction(req).success
    assert stub.SubmitAction(game.SubmitActionRequest(game_id=gid, player_id=1, player_token=js[1].player_token)).success
    g = gs.games[gid]
    assert g.current_turn == 1
    req.action.type, req.action.turn_number = common.ACTION_TYPE_MOVE, 0      # a stale move under the same key
    getattr(req.action, "from").x = 0
    assert stub.SubmitAction(req).success, "within 24 h the cached response answers, whatever the request now says"
    resp, t0 = g.idempotency[(0, "k-1")]
    g.idempotency[(0, "k-1")] = (resp, t0 - 86401.0)
    r = stub.SubmitAction(req)
    assert not r.success and r.error_code == common.ERROR_CODE_INVALID_TURN, "an expired entry does not answer"


def test_more_running_games_than_slots_grow_the_pool(oracle_lib):
    """The reference gives every game its own Engine; only max_games bounds the running games (game_manager.go:104-110).
    A gym client abandons its game at every reset() and the server forgets it only after 30 minutes, so a pool whose
    slots are all held grows by another pool of the same shape instead of refusing the game: seven running games
    through pools of three slots, every one of them playable, and their slots return when they are swept."""
    import time as _t

    srv, gs, port = serve("127.0.0.1:0", lib=oracle_lib, slots_per_pool=3, seed=5)
    ch = grpc.insecure_channel(f"127.0.0.1:{port}")
    stub = Stub(ch, GAME)
    try:
        started = [_start(stub, 5, 5) for _ in range(7)]
        assert len(gs.more_pools[(5, 5, 2)]) == 2 and not gs.pools[(5, 5, 2)].free
        used = {(id(gs.games[gid].pool), gs.games[gid].slot) for gid, _ in started}
        assert len(used) == 7, "every running game has a slot of its own"
        for gid, js in started:                                   # both players pass: the turn runs in every game
            for j in js:
                r = stub.SubmitAction(game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token))
                assert r.success
            st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=0, player_token=js[0].player_token)).state
            assert st.turn == 1 and st.status == common.GAME_STATUS_IN_PROGRESS and len(st.board.tiles) == 25
        assert gs.cleanup_games(_t.time() + gs.abandoned_game_timeout + 1) == 7 and not gs.games
        assert len(gs.pools[(5, 5, 2)].free) == 3 and all(len(p.free) == 3 for p in gs.more_pools[(5, 5, 2)])
        _play_duel(gs, stub)                                      # and the first pool serves the next game
        assert len(gs.more_pools[(5, 5, 2)]) == 2
    finally:
        ch.close()
        srv.stop(0)
        gs.close()


def test_max_games_zero_means_unlimited(server):
    """max_games_test.go:48-68 (TestMaxGamesZeroMeansUnlimited): a server without a limit creates twenty lobby games."""
    gs, stub, _, _, _ = server
    assert gs.max_games == 0
    before = len(gs.games)
    cfg = game.GameConfig(width=10, height=10, max_players=2)
    ids = [stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id for _ in range(20)]
    assert all(ids) and len(set(ids)) == 20 and len(gs.games) == before + 20


def test_reference_cleanup_kats(server):
    """cleanup_test.go:14-103 (TestGameCleanup) and :105-149 (TestLastActivityUpdates): a lobby game without an engine is
    kept while it shows activity and removed once it has been idle past the abandoned-game timeout (35 min > 30 min);
    JoinGame refreshes lastActivity."""
    import time as _t

    gs, stub, _, _, _ = server
    cfg = game.GameConfig(width=10, height=10, max_players=2)
    gid = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
    g = gs.games[gid]
    g.last_activity = _t.time()
    gs.cleanup_games()
    assert gid in gs.games, "Active game should not be cleaned up"
    g.last_activity = _t.time() - 35 * 60
    gs.cleanup_games()
    assert gid not in gs.games, "Abandoned game without engine should be cleaned up"
    gid2 = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
    gs.games[gid2].last_activity = _t.time()
    gs.cleanup_games()
    assert gid2 in gs.games, "Active game should not be cleaned up"
    # TestLastActivityUpdates
    gid3 = stub.CreateGame(game.CreateGameRequest(config=cfg)).game_id
    initial = gs.games[gid3].last_activity
    _t.sleep(0.01)
    stub.JoinGame(game.JoinGameRequest(game_id=gid3, player_name="Player1"))
    assert gs.games[gid3].last_activity > initial, "Join should update last activity"


def test_reference_compat_emits_the_action_index_the_reference_server_emits(oracle_lib):
    """on the reference's gRPC path Experience.action is always 0 (turn_processor.go:194-199 reads
    MoveAction.From/To, converters.go:116-123 fills FromX/FromY/ToX/ToY).  Default: the real ActionToIndex."""
    for compat in (False, True):
        srv, gs, port = serve("127.0.0.1:0", lib=oracle_lib, slots_per_pool=2, seed=77, reference_compat=compat)
        ch = grpc.insecure_channel(f"127.0.0.1:{port}")
        stub, xstub = Stub(ch, GAME), Stub(ch, EXP)
        try:
            gid, js = _start(stub, 6, 6, collect=True)
            rng = np.random.default_rng(11)
            for turn in range(6):
                for j in js:
                    st = stub.GetGameState(game.GetGameStateRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)).state
                    req = game.SubmitActionRequest(game_id=gid, player_id=j.player_id, player_token=j.player_token)
                    req.action.CopyFrom(_legal_action(st, rng, st.turn))
                    assert stub.SubmitAction(req).success
            xs = list(xstub.StreamExperiences(experience.StreamExperiencesRequest(game_ids=[gid], follow=False)))
            assert len(xs) == 12
            if compat:
                assert all(x.action == 0 for x in xs)
            else:
                assert any(x.action != 0 for x in xs) and all(x.action_mask[x.action] for x in xs)
        finally:
            ch.close()
            srv.stop(0)
            gs.close()
