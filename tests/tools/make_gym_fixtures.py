#!/usr/bin/env python3
"""Pin the generals_gym contract against the RUNNING reference client.

What runs here is the reference's own, unmodified ``python/generals_gym/generals_env.py``
(imported from /root/reference behind a ten-line ``gymnasium`` stub: ``Env``, ``register``,
``spaces.Box/Discrete`` -- gymnasium is absent from the image) over the reference's own
protoc-generated ``generals_pb`` messages.  Its four contract functions

    GeneralsEnv._get_observation              generals_env.py:291-342
    GeneralsEnv._get_valid_actions_mask       generals_env.py:344-387
    GeneralsEnv._action_index_to_game_action  generals_env.py:389-441
    GeneralsEnv._calculate_reward             generals_env.py:499-561

and its ``reset()`` / ``step()`` (:150-289) are called as they stand; their outputs are committed
under tests/golden/gym_ref/ and the tests compare the oracle (CPU) and ``grl_gym_step`` /
``grl_gym_observe`` / ``grl_gym_encode`` (CUDA, -m gpu) with them bit for bit.

The Go game server cannot run here (no Go toolchain), so the client talks to ``FakeGameService``:
a stand-in for ``game_pb2_grpc.GameServiceStub`` whose games are single-env engines of the CPU
oracle and whose ``GetGameState`` builds the fog-filtered ``GameState`` message the way
``Server.convertGameStateToProto`` does (internal/grpc/gameserver/server.go:526-610, restated in
``state_proto`` below: that restatement, the engine trajectory and the map generator are pinned
elsewhere -- tests/kats.py, the mapgen goldens -- what THIS tool pins is the client side).

Two deliberate modelling decisions, both forced by the reference's wall-clock behaviour:
  * the turn is processed when the agent, having submitted, asks for the new state.  The real
    server processes it when every player has submitted or when its turn timer fires
    (game_manager.go:554-575, 643-690); the client simply sleeps 50 ms and reads.
  * ``Action.turn_number`` is not checked.  The client never sets it, and the real validator
    (action_validator.go:89-97) would answer INVALID_TURN to every action after turn 0, leaving the
    game to advance on its timer alone; the env contract mirrored by ``grl_gym_step`` is the
    client's evident one -- the action it submits is the action taken.
The server-side board validation IS kept (``ValidateCoreAction``, action_validator.go:113-137 ->
core/action.go:56-105): an action it refuses (the client's half move aimed at a mountain,
generals_env.py:421-428) is not buffered, the turn runs without it, and the client counts the step.

Gymnasium's ``TimeLimit`` (``gym.register(max_episode_steps=500)``, generals_env.py:607-611) is not
part of the class and is not emulated here; the fixtures record the class's own ``truncated``.

Usage: python tests/tools/make_gym_fixtures.py   (rewrites tests/golden/gym_ref/*.npz; this container only --
/root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF_PY = "/root/reference/python"
OUT_DIR = os.path.join(ROOT, "tests", "golden", "gym_ref")


# ---------------------------------------------------------------- importing the reference client
def install_gymnasium_stub():
    """The only gymnasium names generals_env.py touches at import/construction time."""
    gym = types.ModuleType("gymnasium")

    class Env:
        def reset(self, seed=None, options=None):
            return None

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class Discrete:
        def __init__(self, n):
            self.n = int(n)

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box, spaces.Discrete = Box, Discrete
    gym.Env, gym.spaces = Env, spaces
    gym.register = lambda **kw: None
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces


def load_reference():
    if not os.path.isdir(REF_PY):
        raise SystemExit(f"{REF_PY} not found: fixtures can only be regenerated where the reference is mounted")
    sys.dont_write_bytecode = True      # /root/reference is read-only
    install_gymnasium_stub()
    sys.path.insert(0, REF_PY)
    from generals_gym import generals_env
    from generals_pb.common.v1 import common_pb2
    from generals_pb.game.v1 import game_pb2

    generals_env.time.sleep = lambda s: None   # the client's 100 ms / 50 ms waits for the server
    return generals_env, game_pb2, common_pb2


# ---------------------------------------------------------------- the stand-in game server
class Game:
    """One server-side game: a single-env oracle engine + the proto view of server.go."""

    def __init__(self, lib, pb, W, H, P, fog, seed=None, boards=None):
        from helpers import new_engine

        self.game_pb2, self.common_pb2 = pb
        self.W, self.H, self.P, self.N = W, H, P, W * H
        self.e = new_engine(lib, W, H, P, 1, fog_of_war=1 if fog else 0, max_actions=max(2, P))
        if boards is not None:
            self.e.reset_boards(*boards)
        else:
            self.e.reset_seeded(np.array([seed], np.int64))
        self.fog = fog
        self.buffer = {}

    # core/action.go:56-105 MoveAction.Validate against the TRUE board (what ValidateCoreAction calls)
    def validate(self, player, fx, fy, tx, ty):
        st = self.e.get_state()
        W, H = self.W, self.H
        if not (0 <= fx < W and 0 <= fy < H) or not (0 <= tx < W and 0 <= ty < H):
            return False
        if (fx, fy) == (tx, ty) or abs(fx - tx) + abs(fy - ty) != 1:
            return False
        i, j = fy * W + fx, ty * W + tx
        if st["owner"][0, i] != player or st["army"][0, i] <= 1:
            return False
        return st["type"][0, j] != 3

    def submit(self, player, action):
        c = self.common_pb2
        if action.type != c.ACTION_TYPE_MOVE:
            self.buffer[player] = None
            return True
        fr, to = getattr(action, "from"), action.to
        if not self.validate(player, fr.x, fr.y, to.x, to.y):
            return False
        self.buffer[player] = (fr.x, fr.y, to.x, to.y, not action.half)   # converters.go:116-123
        return True

    def process_turn(self):
        from generalsreinforcementlearning_b200.engine import make_actions, set_action

        acts = make_actions(1, self.e.A)
        for player, a in sorted(self.buffer.items()):
            if a is not None:
                set_action(acts, 0, player, player, a[0], a[1], a[2], a[3], a[4])
        self.buffer = {}
        self.e.step(acts)

    def state_proto(self, player_id):
        """Server.convertGameStateToProto (server.go:526-610)."""
        g, c = self.game_pb2, self.common_pb2
        st = self.e.get_state()
        vis, fog = self.e.visibility()
        tile_type = {0: c.TILE_TYPE_NORMAL, 1: c.TILE_TYPE_GENERAL, 2: c.TILE_TYPE_CITY, 3: c.TILE_TYPE_MOUNTAIN}
        msg = g.GameState(game_id="fixture", turn=int(st["turn"][0]))
        over = bool(st["game_over"][0])
        # mapPhaseToStatus (converters.go:83-103): RUNNING -> IN_PROGRESS, ENDED -> FINISHED
        msg.status = c.GAME_STATUS_FINISHED if over else c.GAME_STATUS_IN_PROGRESS
        msg.current_phase = c.GAME_PHASE_ENDED if over else c.GAME_PHASE_RUNNING
        msg.winner_id = int(st["winner"][0]) if over else -1
        msg.board.width, msg.board.height = self.W, self.H
        for p in range(self.P):
            ps = msg.players.add()
            ps.id, ps.name = p, f"player{p}"
            alive = bool(st["alive"][0, p])
            ps.status = c.PLAYER_STATUS_ACTIVE if alive else c.PLAYER_STATUS_ELIMINATED
            ps.army_count = int(st["army_count"][0, p])
            ps.tile_count = int(st["owned"][0, p].sum())          # len(p.OwnedTiles): the CACHED list
            gi = int(st["general_idx"][0, p])
            if gi >= 0 and (not alive or st["owner"][0, gi] == player_id):
                ps.general_position.x, ps.general_position.y = gi % self.W, gi // self.W
        for i in range(self.N):
            t = msg.board.tiles.add()
            t.type = tile_type[int(st["type"][0, i])]
            t.owner_id = int(st["owner"][0, i])
            t.army_count = int(st["army"][0, i])
            t.visible = bool(vis[0, player_id, i])
            t.fog_of_war = bool(fog[0, player_id, i])
            if not t.visible and not t.fog_of_war:
                t.type, t.owner_id, t.army_count = c.TILE_TYPE_NORMAL, -1, 0
            elif t.fog_of_war and not t.visible:
                t.owner_id, t.army_count = -1, 0
        return msg


class FakeGameService:
    """Stands where game_pb2_grpc.GameServiceStub stands for ONE GeneralsEnv."""

    def __init__(self, lib, pb, scenario):
        self.lib, self.pb, self.scenario = lib, pb, scenario
        self.game = None
        self.tokens = {}
        self.creates = 0
        self.rejected = []       # (turn, player) of actions the server-side validation refused

    def CreateGame(self, req):
        self.creates += 1        # the first one is the constructor's connection probe (generals_env.py:134-137)
        self.config = req.config
        self.tokens = {}
        return self.pb[0].CreateGameResponse(game_id=f"g{self.creates}")

    def JoinGame(self, req):
        pid = len(self.tokens)
        token = f"tok{pid}"
        self.tokens[token] = pid
        if pid == 1:             # the game starts when the second client joins
            cfg, sc = self.config, self.scenario
            assert (cfg.width, cfg.height) == (sc["W"], sc["H"])
            self.game = Game(self.lib, self.pb, sc["W"], sc["H"], sc["P"], cfg.fog_of_war, seed=sc.get("seed"),
                             boards=sc.get("boards"))
        return self.pb[0].JoinGameResponse(player_id=pid, player_token=token)

    def SubmitAction(self, req):
        pid = self.tokens[req.player_token]
        ok = self.game.submit(pid, req.action)
        if not ok:
            self.rejected.append((int(self.game.e.get_state()["turn"][0]), pid))
        return self.pb[0].SubmitActionResponse(success=ok)

    def GetGameState(self, req):
        pid = self.tokens[req.player_token]
        if pid == 0 and self.pending_agent_turn:
            self.pending_agent_turn = False
            self.game.process_turn()
        return self.pb[0].GetGameStateResponse(state=self.game.state_proto(pid))

    pending_agent_turn = False


class ScriptedOpponent:
    """opponent_agent (generals_env.py:245-254): picks a Discrete(N*5) index for player 1 from ITS view with a
    seeded generator and decodes it with the reference's own _action_index_to_game_action on a shadow client."""

    def __init__(self, env_module, stub, W, H, rng, p_invalid, p_none, script=None):
        self.shadow = shadow_client(env_module, W, H, player_id=1)
        self.stub, self.rng, self.p_invalid, self.p_none, self.script = stub, rng, p_invalid, p_none, script
        self.log = []

    def select_action(self, _agent_state):
        sh = self.shadow
        sh.current_state = self.stub.game.state_proto(1)
        sh.valid_actions_mask = sh._get_valid_actions_mask()
        if self.script is not None:
            idx = self.script[len(self.log)] if len(self.log) < len(self.script) else -1
        else:
            idx = pick_index(sh.valid_actions_mask, self.rng, self.p_invalid, self.p_none)
        self.log.append(idx)
        if idx < 0 or idx >= len(sh.valid_actions_mask):
            return None
        return sh._action_index_to_game_action(idx)


def shadow_client(env_module, W, H, player_id, max_turns=500):
    """A GeneralsEnv without a connection: just the attributes its read-out methods use."""
    env = object.__new__(env_module.GeneralsEnv)
    env.board_width, env.board_height, env.board_size = W, H, W * H
    env.player_id, env.max_turns, env.turn_count = player_id, max_turns, 0
    env.action_space = sys.modules["gymnasium"].spaces.Discrete(W * H * 5)
    env.current_state, env.valid_actions_mask = None, None
    return env


def pick_index(mask, rng, p_invalid, p_none):
    r = rng.random()
    if r < p_none:
        return -1
    want = r >= p_none + p_invalid
    idx = np.flatnonzero(mask == want)
    if len(idx) == 0:
        return int(rng.integers(0, len(mask)))
    return int(idx[rng.integers(0, len(idx))])


def decode_table(env, N):
    """_action_index_to_game_action for EVERY index: [N*5][6] = valid, fx, fy, tx, ty, half."""
    out = np.zeros((N * 5, 6), np.int8)
    for a in range(N * 5):
        act = env._action_index_to_game_action(a)
        if act is not None:
            fr = getattr(act, "from")
            out[a] = (1, fr.x, fr.y, act.to.x, act.to.y, int(act.half))
    return out


# ---------------------------------------------------------------- fixture kind A: read-outs of given states
READOUT_CASES = [  # name, W, H, P, fog, turns-per-env (random policy), max_turns
    ("8x8x2p", 8, 8, 2, 1, (0, 1, 37, 90), 100),
    ("10x10x2p", 10, 10, 2, 1, (0, 25, 120, 260), 500),
    ("15x15x2p", 15, 15, 2, 1, (0, 50, 199, 420), 500),
    ("20x20x2p", 20, 20, 2, 1, (3, 75, 300, 499), 500),
    ("20x20x4p", 20, 20, 4, 1, (0, 60, 240), 500),
    ("9x7x3p", 9, 7, 3, 1, (0, 33, 140), 60),          # turn/max_turns saturates at 1
    ("10x10x2p_nofog", 10, 10, 2, 0, (0, 40, 180), 500),
    ("5x5x2p_ended", 5, 5, 2, 1, (400, 400, 400, 400, 400, 400), 500),   # small boards finish: dead players
    ("8x8x4p_late", 8, 8, 4, 1, (450, 450, 450, 450, 450, 450), 500),
]


def make_readouts(lib, ref, name, W, H, P, fog, turns, max_turns):
    from generalsreinforcementlearning_b200 import _abi

    env_module, game_pb2, common_pb2 = ref
    N = W * H
    states, obs, mask, decode, stats = [], [], [], [], []
    for k, T in enumerate(turns):
        g, seed = None, 4242 + 17 * k
        while g is None:       # a few seeds cannot seat every general on a small board (mapgen/generator.go:252)
            try:
                g = Game(lib, (game_pb2, common_pb2), W, H, P, fog, seed=seed)
            except RuntimeError:
                seed += 1000
        for _ in range(T):
            g.e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 99)
        st = g.e.get_state()
        states.append(st)
        o_env, m_env, d_env, s_env = [], [], [], []
        for p in range(P):
            cl = shadow_client(env_module, W, H, p, max_turns)
            cl.current_state = g.state_proto(p)
            cl.turn_count = int(st["turn"][0])       # the client counts the turns it has seen taken
            o_env.append(cl._get_observation())
            cl.valid_actions_mask = cl._get_valid_actions_mask()
            m_env.append(cl.valid_actions_mask.copy())
            d_env.append(decode_table(cl, N))
            me = [q for q in cl.current_state.players if q.id == p][0]
            s_env.append((me.army_count, me.tile_count, int(me.status == common_pb2.PLAYER_STATUS_ACTIVE)))
        obs.append(np.stack(o_env)), mask.append(np.stack(m_env)), decode.append(np.stack(d_env))
        stats.append(np.array(s_env, np.int32))
        g.e.close()
    res = {f"state_{k}": np.concatenate([s[k] for s in states]) for k in states[0]}
    res.update(obs=np.stack(obs), mask=np.stack(mask), decode=np.stack(decode), stats=np.stack(stats),
               meta=np.array([W, H, P, fog, max_turns], np.int32))
    alive = res["state_alive"]
    print(f"  readouts {name}: {len(turns)} states, dead players {int((alive == 0).sum())}, finished games "
          f"{int(res['state_game_over'].sum())}, valid actions {int(res['mask'].sum())}")
    return res


# ---------------------------------------------------------------- fixture kind B: whole episodes through step()
def elimination_board_4p():
    """8x8, four players.  Player 1 stands next to player 2's weak general and takes it on cue: the agent (player 0)
    sees an opponent leave (+50, generals_env.py:548-554) while the game goes on."""
    W = H = 8
    owner = np.full((1, W * H), -1, np.int32)
    army = np.zeros((1, W * H), np.int32)
    type_ = np.zeros((1, W * H), np.int32)

    def put(x, y, o, a, t=0):
        owner[0, y * W + x], army[0, y * W + x], type_[0, y * W + x] = o, a, t

    put(0, 0, 0, 5, 1), put(1, 0, 0, 4), put(0, 1, 0, 3)
    put(7, 7, 1, 6, 1), put(4, 3, 1, 30)
    put(4, 4, 2, 2, 1), put(5, 4, 2, 7), put(4, 5, 2, 1)
    put(7, 0, 3, 3, 1), put(6, 0, 3, 9)
    put(3, 6, -1, 40, 2), put(2, 2, -1, 0, 3), put(5, 1, -1, 0, 3)
    return (owner, army, type_)


def duel_board_2p():
    """5x5: the agent's stack stands next to the opponent's general -- it can end the game (+100) at once, or
    dither; the opponent's stack threatens the agent's general the same way (-100)."""
    W = H = 5
    owner = np.full((1, W * H), -1, np.int32)
    army = np.zeros((1, W * H), np.int32)
    type_ = np.zeros((1, W * H), np.int32)

    def put(x, y, o, a, t=0):
        owner[0, y * W + x], army[0, y * W + x], type_[0, y * W + x] = o, a, t

    put(0, 0, 0, 3, 1), put(3, 4, 0, 25)
    put(4, 4, 1, 4, 1), put(1, 0, 1, 25)
    put(2, 2, -1, 0, 3)
    return (owner, army, type_)


EPISODE_CASES = [  # name, scenario, #episodes, max steps, max_turns, p_invalid (agent), opponent mode
    ("5x5x2p", dict(W=5, H=5, P=2, fog=1), 6, 400, 500, 0.06, "scripted"),
    ("10x10x2p", dict(W=10, H=10, P=2, fog=1), 3, 70, 500, 0.08, "scripted"),
    ("10x10x2p_trunc", dict(W=10, H=10, P=2, fog=1), 2, 60, 24, 0.10, "scripted"),
    ("15x15x2p", dict(W=15, H=15, P=2, fog=1), 2, 60, 500, 0.08, "scripted"),
    ("20x20x2p", dict(W=20, H=20, P=2, fog=1), 2, 50, 500, 0.08, "scripted"),
    ("8x8x2p_nofog", dict(W=8, H=8, P=2, fog=0), 2, 60, 500, 0.08, "scripted"),
    ("8x8x4p_elim", dict(W=8, H=8, P=4, fog=1, boards=elimination_board_4p), 2, 12, 500, 0.0, "capture"),
    ("5x5x2p_duel", dict(W=5, H=5, P=2, fog=1, boards=duel_board_2p), 3, 6, 500, 0.0, "duel"),
]


def run_episode(lib, ref, sc, ep, max_steps, max_turns, p_invalid, mode):
    env_module, game_pb2, common_pb2 = ref
    W, H, P = sc["W"], sc["H"], sc["P"]
    N = W * H
    scenario = dict(sc)
    if "boards" in sc:
        scenario["boards"] = sc["boards"]()
    else:
        scenario["seed"] = 9000 + 31 * ep + 7 * W
    stub = FakeGameService(lib, (game_pb2, common_pb2), scenario)
    rng = np.random.default_rng(1000 * W + 10 * P + ep)

    # hand the client our stub instead of a channel (generals_env.py:127-142)
    env_module.grpc.insecure_channel = lambda addr: types.SimpleNamespace(close=lambda: None)
    env_module.game_pb2_grpc.GameServiceStub = lambda channel: stub
    script = None
    if mode == "capture":      # player 1 takes player 2's general at (4,4) from (4,3) on its `ep`-th move, else idles
        script = [-1] * ep + [(3 * W + 4) * 5 + 2]
    elif mode == "duel":       # ep 0: the agent wins; ep 1: the opponent wins first (player 0 dithers); ep 2: both strike
        script = {0: [-1, -1], 1: [(0 * W + 1) * 5 + 3], 2: [(0 * W + 1) * 5 + 3]}[ep]
    opp = ScriptedOpponent(env_module, stub, W, H, rng, 0.05, 0.05, script)
    env = env_module.GeneralsEnv(server_address="fake", board_width=W, board_height=H, max_players=P,
                                 fog_of_war=bool(sc["fog"]), opponent_agent=opp, max_turns=max_turns)
    obs0, info0 = env.reset()
    assert env.player_id == 0 and env.opponent_id == 1
    st0 = stub.game.e.get_state()
    rec = dict(obs=[obs0], mask=[info0["valid_actions_mask"].copy()], action=[], opp_action=[], reward=[], terminated=[],
               truncated=[], invalid=[], decoded=[], server_rejected=[])
    agent_script = None
    if mode == "duel":         # (3,4) -> right takes the general at (4,4)
        strike = (4 * W + 3) * 5 + 1
        agent_script = {0: [strike], 1: [(4 * W + 3) * 5 + 0, strike], 2: [strike]}[ep]
    for t in range(max_steps):
        if agent_script is not None:
            a = agent_script[t] if t < len(agent_script) else pick_index(env.valid_actions_mask, rng, 0, 0)
        else:
            a = pick_index(env.valid_actions_mask, rng, p_invalid, 0.0)
            if t % 11 == 5:
                a = int(rng.integers(0, N * 5))      # any index at all, half moves and masked-out ones included
        ga = env._action_index_to_game_action(a)
        dec = (0, 0, 0, 0, 0, 0)
        if ga is not None:
            fr = getattr(ga, "from")
            dec = (1, fr.x, fr.y, ga.to.x, ga.to.y, int(ga.half))
        n_opp, n_rej = len(opp.log), len(stub.rejected)
        stub.pending_agent_turn = ga is not None
        obs, reward, terminated, truncated, info = env.step(a)
        invalid = bool(info.get("invalid_action", False))
        assert "error" not in info
        assert invalid == (ga is None)
        rec["action"].append(a)
        rec["opp_action"].append(opp.log[n_opp] if len(opp.log) > n_opp else -1)
        rec["reward"].append(float(reward)), rec["terminated"].append(terminated), rec["truncated"].append(truncated)
        rec["invalid"].append(invalid), rec["decoded"].append(dec)
        rec["server_rejected"].append(sum(1 << p for (_, p) in stub.rejected[n_rej:]))
        rec["obs"].append(obs)
        # the client does not refresh its mask after a rejected action (:226-229); the mask it acts on next is this one
        rec["mask"].append(env.valid_actions_mask.copy())
        if terminated or truncated:
            break
    env.close()
    stub.game.e.close()
    T = len(rec["action"])
    out = dict(obs=np.stack(rec["obs"]).astype(np.float32), mask=np.stack(rec["mask"]),
               action=np.array(rec["action"], np.int64), opp_action=np.array(rec["opp_action"], np.int64),
               reward=np.array(rec["reward"], np.float64), terminated=np.array(rec["terminated"], np.uint8),
               truncated=np.array(rec["truncated"], np.uint8), invalid=np.array(rec["invalid"], np.uint8),
               decoded=np.array(rec["decoded"], np.int8), server_rejected=np.array(rec["server_rejected"], np.uint8),
               init_owner=st0["owner"], init_army=st0["army"], init_type=st0["type"])
    if "seed" in scenario:
        out["seed"] = np.array([scenario["seed"]], np.int64)
    return out, T


def main():
    from generalsreinforcementlearning_b200._abi import BoundLibrary

    lib = BoundLibrary(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    ref = load_reference()
    os.makedirs(OUT_DIR, exist_ok=True)
    for case in READOUT_CASES:
        res = make_readouts(lib, ref, *case)
        path = os.path.join(OUT_DIR, f"readouts_{case[0]}.npz")
        np.savez_compressed(path, **res)
        print("   ", path, os.path.getsize(path), "bytes")
    for name, sc, n_ep, max_steps, max_turns, p_invalid, mode in EPISODE_CASES:
        res = {"meta": np.array([sc["W"], sc["H"], sc["P"], sc["fog"], max_turns, n_ep], np.int32)}
        summary = []
        for ep in range(n_ep):
            out, T = run_episode(lib, ref, sc, ep, max_steps, max_turns, p_invalid, mode)
            for k, v in out.items():
                res[f"ep{ep}_{k}"] = v
            summary.append(f"{T} steps r={out['reward'].sum():+.2f} inv={int(out['invalid'].sum())} "
                           f"srvrej={int((out['server_rejected'] != 0).sum())} term={int(out['terminated'][-1])} "
                           f"trunc={int(out['truncated'][-1])}")
        path = os.path.join(OUT_DIR, f"episodes_{name}.npz")
        np.savez_compressed(path, **res)
        print(f"  episodes {name}: " + " | ".join(summary))
        print("   ", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
