"""gRPC ``GameService`` / ``ExperienceService`` front end over the batched engine.

The wire contract is the reference's (``proto/game/v1/game.proto:11-26``,
``proto/experience/v1/experience.proto:11-23``; schema in ``grpc_schema.py``); the behaviour follows
``internal/grpc/gameserver``:

  CreateGame / JoinGame     server.go:67-191, game_manager.go:110-165 (ids ``game-N``, tokens
                            ``token-<game>-<player>``, the engine starts when the last player joins)
  SubmitAction              server.go:193-294 + action_validator.go:35-139: game exists -> idempotency
                            cache -> phase RUNNING -> token -> turn number -> ``action.Validate`` against
                            the current board -> per-game action buffer -> when every player has
                            submitted (or the turn timer fires, game_manager.go:650-732) ONE turn
  GetGameState / StreamGame fog-filtered proto view (server.go:526-610), delta updates when fewer than
                            N/5 tiles changed (server.go:638-777)
  experiences               one record per player that submitted a move, state/mask from before the
                            turn (collector.go:30-98), streamed singly or in batches of ``batch_size``
                            (default 32) flushed after ``max_batch_wait_ms`` (default 100 ms)
                            (experience_service.go:287-378,510-514)

Each game occupies one env SLOT of a ``BatchedEngine`` pool keyed by (width, height, players).  A
turn is one fused kernel launch in which every other slot carries ``GRL_ACTION_FLAG_SKIP_ENV`` —
the per-game turn barrier of game_manager.go:559-600 expressed in the batched engine.  Observation
and mask planes stay in device memory; only the slot being served is copied to the host.

This is the servicer a Python deployment runs; the Go binding of the same calls is in INTEGRATION.md.
"""
from __future__ import annotations

import threading
import time
import uuid
from collections import deque
from concurrent import futures
from typing import Dict, List, Optional

import grpc
import numpy as np

from . import _abi
from .engine import BatchedEngine, make_actions, make_config, set_action
from .grpc_schema import SERVICES, common, experience, game

_TILE_TYPE = {0: common.TILE_TYPE_NORMAL, 1: common.TILE_TYPE_GENERAL, 2: common.TILE_TYPE_CITY, 3: common.TILE_TYPE_MOUNTAIN}
_VALIDATE_MSG = {  # core/errors.go:8-17
    _abi.STEP_INVALID_COORDINATES: "invalid coordinates", _abi.STEP_NOT_ADJACENT: "tiles are not adjacent",
    _abi.STEP_NOT_OWNED: "tile not owned by player", _abi.STEP_INSUFFICIENT_ARMY: "insufficient army to move",
    _abi.STEP_MOVE_TO_SELF: "cannot move to the same tile", _abi.STEP_TARGET_IS_MOUNTAIN: "target tile is a mountain",
}
_PHASE_STATUS = {  # converters.go:84-102
    common.GAME_PHASE_LOBBY: common.GAME_STATUS_WAITING, common.GAME_PHASE_RUNNING: common.GAME_STATUS_IN_PROGRESS,
    common.GAME_PHASE_ENDED: common.GAME_STATUS_FINISHED, common.GAME_PHASE_ERROR: common.GAME_STATUS_CANCELLED,
    common.GAME_PHASE_STARTING: common.GAME_STATUS_IN_PROGRESS, common.GAME_PHASE_PAUSED: common.GAME_STATUS_IN_PROGRESS,
    common.GAME_PHASE_ENDING: common.GAME_STATUS_IN_PROGRESS,
}


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b7 = n & 0x7F
        n >>= 7
        if n:
            out.append(b7 | 0x80)
        else:
            out.append(b7)
            return bytes(out)


def _tag(field: int, wire_type: int) -> bytes:
    return _varint((field << 3) | wire_type)


def _packed_field(field: int, payload: bytes) -> bytes:
    """A length-delimited field: strings, sub-messages and packed repeated scalars."""
    return _tag(field, 2) + _varint(len(payload)) + payload


def _now(ts) -> None:
    t = time.time()
    ts.seconds, ts.nanos = int(t), int((t % 1) * 1e9)


def validate_move(owner, army, type_, W, H, player, fx, fy, tx, ty) -> int:
    """core.MoveAction.Validate (core/action.go:56-105), same precedence; 0 when legal."""
    if not (0 <= fx < W and 0 <= fy < H):
        return _abi.STEP_INVALID_COORDINATES
    if not (0 <= tx < W and 0 <= ty < H):
        return _abi.STEP_INVALID_COORDINATES
    if fx == tx and fy == ty:
        return _abi.STEP_MOVE_TO_SELF
    if abs(fx - tx) + abs(fy - ty) != 1:
        return _abi.STEP_NOT_ADJACENT
    if owner[fy * W + fx] != player:
        return _abi.STEP_NOT_OWNED
    if army[fy * W + fx] <= 1:
        return _abi.STEP_INSUFFICIENT_ARMY
    if type_[ty * W + tx] == _abi.TILE_MOUNTAIN:
        return _abi.STEP_TARGET_IS_MOUNTAIN
    return 0


class EnginePool:
    """S env slots of one board shape on one device, with resident read-out planes."""

    def __init__(self, lib, W, H, P, slots, device=0):
        import torch

        self.torch = torch
        self.W, self.H, self.P, self.N, self.S = W, H, P, W * H, slots
        self.engine = BatchedEngine(lib, make_config(lib, num_envs=slots, width=W, height=H, num_players=P,
                                                     max_actions=P, device=device, host_threads=1))
        dev = torch.device("cuda", device) if lib.prefix == "grl_" else torch.device("cpu")
        self.engine.use_torch_stream()
        S, N = slots, self.N
        self.obs = torch.zeros((S, P, 9, H, W), dtype=torch.float32, device=dev)       # StateToTensor of every slot
        self.ser_mask = torch.zeros((S, P, N * 4), dtype=torch.uint8, device=dev)     # serializer mask (U,D,L,R)
        self.eng_mask = torch.zeros((S, P, N * 4), dtype=torch.uint8, device=dev)     # engine mask (U,R,D,L)
        self.reward = torch.zeros((S, P), dtype=torch.float32, device=dev)
        self.done = torch.zeros(S, dtype=torch.uint8, device=dev)
        self.winner = torch.zeros(S, dtype=torch.int8, device=dev)
        self.err = torch.zeros(S, dtype=torch.uint8, device=dev)
        self.aidx = torch.zeros((S, P), dtype=torch.int32, device=dev)
        self.free = deque(range(slots))
        self.lock = threading.RLock()  # the handle is not internally locked (grlcuda.h)

    def outputs(self):
        e = self.engine
        return e.outputs(obs=self.obs, reward=self.reward, done=self.done, winner=self.winner, step_error=self.err,
                         action_index=self.aidx)

    def refresh(self):
        self.engine.observe(self.engine.outputs(obs=self.obs, done=self.done, winner=self.winner))
        self.engine.mask(_abi.MASK_SERIALIZER_UDLR, self.ser_mask)
        self.engine.mask(_abi.MASK_ENGINE_URDL, self.eng_mask)


class _Player:
    def __init__(self, pid, name, token):
        self.id, self.name, self.token = pid, name, token


class GameInstance:
    def __init__(self, gid, cfg):
        self.id, self.config = gid, cfg
        self.players: List[_Player] = []
        self.phase = common.GAME_PHASE_LOBBY
        self.pool: Optional[EnginePool] = None
        self.slot = -1
        self.current_turn = 0
        self.actions: Dict[int, Optional[tuple]] = {}
        self.idempotency: Dict[tuple, object] = {}
        self.streams: Dict[int, deque] = {}
        self.stream_cv = threading.Condition()
        self.mu = threading.RLock()
        self.timer: Optional[threading.Timer] = None
        self.started_at = 0.0
        self.prev_alive: List[bool] = []
        self.last_activity = time.time()          # gameInstance.lastActivity (game_manager.go:160,562)
        self.final = None                          # (state planes, engine masks) kept once the env slot is released


class ExperienceStore:
    """Everything the collectors emitted, with blocking readers (BufferManager + StreamMerger)."""

    def __init__(self, capacity=100000):
        self.items: deque = deque(maxlen=capacity)
        self.total = 0
        self.cv = threading.Condition()

    def add(self, exps):
        with self.cv:
            for x in exps:
                self.items.append((self.total, x))
                self.total += 1
            self.cv.notify_all()

    def read_from(self, cursor, timeout):
        """Items with sequence >= cursor; waits up to ``timeout`` seconds for the first one."""
        with self.cv:
            if self.total <= cursor:
                self.cv.wait(timeout)
            first = self.items[0][0] if self.items else self.total
            start = max(cursor, first)
            out = [x for (i, x) in self.items if i >= start]
            return out, max(self.total, cursor)


class GameServer:
    """Implements both services; register with ``add_to_server`` or use ``serve``."""

    def __init__(self, lib=None, slots_per_pool: int = 256, device: int = 0, max_games: int = 0, seed: Optional[int] = None,
                 finished_game_ttl: float = 600.0, abandoned_game_timeout: float = 1800.0, cleanup_interval: float = 300.0,
                 reference_compat: bool = False):
        if lib is None:
            from . import load_library

            lib = load_library()
        self.lib, self.slots_per_pool, self.device, self.max_games = lib, slots_per_pool, device, max_games
        self.games: Dict[str, GameInstance] = {}
        self.pools: Dict[tuple, EnginePool] = {}
        # The reference gives every game its own Engine: the only bound on running games is max_games
        # (game_manager.go:104-110).  A gym client creates a new game on every reset() and abandons the old one, which
        # the server only forgets after abandoned_game_timeout (30 min) — so when every slot of a shape's pool is held,
        # a further pool of the same shape is created instead of refusing the game.
        self.more_pools: Dict[tuple, List[EnginePool]] = {}
        self.next_id = 0
        self.mu = threading.Lock()
        self.store = ExperienceStore()
        self.seed = seed  # None: time-seeded maps like the reference (engine_initializer.go:91-94)
        self.submitted_ids = set()
        # server.go:42-44: cleanupInterval 5 min, finishedGameTTL 10 min, abandonedGameTimeout 30 min
        self.finished_game_ttl, self.abandoned_game_timeout = finished_game_ttl, abandoned_game_timeout
        self.cleanup_interval = cleanup_interval
        self._last_cleanup = time.time()
        # collectExperiences copies MoveAction.From/To, the Coordinate fields
        # (internal/game/turn_processor.go:194-199), but the server's convertProtoAction fills only FromX/FromY/ToX/ToY
        # (internal/grpc/gameserver/converters.go:116-123): From == To == (0,0) and Serializer.ActionToIndex
        # (experience/serializer.go:179-198) yields 0 for every move of a served game.  reference_compat=True reproduces
        # that wire value; the default emits the index ActionToIndex defines for the move that was made, which is what
        # a learner needs.
        self.reference_compat = reference_compat

    # ------------------------------------------------------------------ GameService
    def CreateGame(self, req, ctx):
        now = time.time()
        if now - self._last_cleanup >= self.cleanup_interval or (self.max_games > 0 and len(self.games) >= self.max_games):
            self.cleanup_games(now)
        with self.mu:
            if self.max_games > 0 and len(self.games) >= self.max_games:
                ctx.abort(grpc.StatusCode.RESOURCE_EXHAUSTED,
                          f"failed to create game: server at capacity: {len(self.games)}/{self.max_games} games active")
            self.next_id += 1
            gid = f"game-{self.next_id}"
            cfg = game.GameConfig()
            if req.HasField("config"):
                cfg.CopyFrom(req.config)
            else:  # game_manager.go:131-140
                cfg.width, cfg.height, cfg.max_players, cfg.fog_of_war, cfg.turn_time_ms = 20, 20, 2, True, 0
            self.games[gid] = GameInstance(gid, cfg)
        return game.CreateGameResponse(game_id=gid, config=cfg)

    def _pool_for(self, cfg) -> EnginePool:
        """A pool of the config's shape with a free slot: the shape's first pool, one of its further pools, or — after
        finished and abandoned games have been swept — a new one."""
        key = (cfg.width, cfg.height, cfg.max_players)

        def pick():
            with self.mu:
                if key not in self.pools:
                    self.pools[key] = EnginePool(self.lib, cfg.width, cfg.height, cfg.max_players, self.slots_per_pool, self.device)
                for p in [self.pools[key]] + self.more_pools.get(key, []):
                    if p.free:
                        return p
            return None

        pool = pick()
        if pool is None:
            self.cleanup_games()
            pool = pick()
        if pool is None:
            with self.mu:
                pool = EnginePool(self.lib, cfg.width, cfg.height, cfg.max_players, self.slots_per_pool, self.device)
                self.more_pools.setdefault(key, []).append(pool)
        return pool

    def _start_engine(self, g: GameInstance, ctx):
        cfg = g.config
        if not (1 <= cfg.width <= _abi.GRL_MAX_DIM and 1 <= cfg.height <= _abi.GRL_MAX_DIM
                and 1 <= cfg.max_players <= _abi.GRL_MAX_PLAYERS):
            ctx.abort(grpc.StatusCode.INTERNAL, f"failed to start game engine for game {g.id}: unsupported board")
        while True:
            pool = self._pool_for(cfg)
            with pool.lock:
                if not pool.free:     # another game took the last slot between the look-up and the lock
                    continue
                g.slot = pool.free.popleft()
                seed = (int(time.time_ns()) if self.seed is None else self.seed + self.next_id) & 0x7FFFFFFFFFFFFFFF
                try:
                    pool.engine.reset_seeded([seed], [g.slot])
                except RuntimeError as exc:
                    pool.free.append(g.slot)
                    g.slot = -1
                    ctx.abort(grpc.StatusCode.INTERNAL, f"failed to start game engine for game {g.id}: {exc}")
                pool.refresh()
            break
        g.pool, g.phase, g.current_turn, g.actions = pool, common.GAME_PHASE_RUNNING, 0, {}
        g.started_at = g.last_activity = time.time()
        g.prev_alive = [True] * cfg.max_players
        ev = game.GameUpdate()
        _now(ev.event.game_started.started_at)
        self._broadcast(g, lambda pid: ev)

    def JoinGame(self, req, ctx):
        g = self.games.get(req.game_id)
        if g is None:
            ctx.abort(grpc.StatusCode.NOT_FOUND, f"game {req.game_id} not found: request from player {req.player_name}")
        start_timer = False
        with g.mu:
            for p in g.players:
                if p.name == req.player_name:
                    return game.JoinGameResponse(player_id=p.id, player_token=p.token, initial_state=self._game_state(g, p.id))
            if g.phase != common.GAME_PHASE_LOBBY:
                ctx.abort(grpc.StatusCode.FAILED_PRECONDITION,
                          f"cannot join game {g.id}: game is in {common.GamePhase.names[g.phase]} phase")
            if len(g.players) >= g.config.max_players:
                ctx.abort(grpc.StatusCode.RESOURCE_EXHAUSTED, f"game {g.id} is full: {len(g.players)}/{g.config.max_players} players")
            pid = len(g.players)
            p = _Player(pid, req.player_name, f"token-{g.id}-{pid}")
            g.players.append(p)
            g.last_activity = time.time()
            if len(g.players) == g.config.max_players:
                self._start_engine(g, ctx)
                start_timer = True
        if start_timer:
            self._start_turn_timer(g)
        return game.JoinGameResponse(player_id=p.id, player_token=p.token, initial_state=self._game_state(g, p.id))

    def _auth(self, g, pid, token) -> bool:
        return any(p.id == pid and p.token == token for p in g.players)

    def SubmitAction(self, req, ctx):
        g = self.games.get(req.game_id)
        if g is None:
            return game.SubmitActionResponse(success=False, error_code=common.ERROR_CODE_GAME_NOT_FOUND,
                                             error_message=f"game {req.game_id} not found")
        key = (req.player_id, req.idempotency_key)

        def done(resp):
            if req.idempotency_key:                     # IdempotencyManager.Store (idempotency.go:62-85)
                now = time.time()
                g.idempotency[key] = (resp, now)
                if len(g.idempotency) > 1000:           # :81-84: entries older than 24 h leave once the cache is large
                    for k in [k for k, (_, t0) in g.idempotency.items() if now - t0 > 86400.0]:
                        del g.idempotency[k]
            return resp

        if req.idempotency_key and key in g.idempotency:   # IdempotencyManager.Check (:36-60): valid for 24 hours
            cached, t0 = g.idempotency[key]
            if time.time() - t0 <= 86400.0:
                return cached
        if g.phase != common.GAME_PHASE_RUNNING:
            code = common.ERROR_CODE_GAME_OVER if g.phase == common.GAME_PHASE_ENDED else common.ERROR_CODE_INVALID_PHASE
            return done(game.SubmitActionResponse(
                success=False, error_code=code,
                error_message=f"game {g.id} cannot accept actions in {common.GamePhase.names[g.phase]} phase"))
        if not self._auth(g, req.player_id, req.player_token):
            return done(game.SubmitActionResponse(
                success=False, error_code=common.ERROR_CODE_INVALID_PLAYER,
                error_message=f"invalid player credentials for game {g.id}: player {req.player_id}"))
        with g.mu:
            current_turn = g.current_turn
        has_action = req.HasField("action")
        if has_action and req.action.turn_number != current_turn:
            return done(game.SubmitActionResponse(
                success=False, error_code=common.ERROR_CODE_INVALID_TURN,
                error_message=f"invalid turn number for game {g.id}: expected {current_turn}, got {req.action.turn_number}"))
        move = None
        if has_action and req.action.type == common.ACTION_TYPE_MOVE:  # converters.go:105-132
            a = req.action
            if not a.HasField("from") or not a.HasField("to"):
                return done(game.SubmitActionResponse(
                    success=False, error_code=common.ERROR_CODE_INVALID_TURN,
                    error_message=f"invalid action for game {g.id} player {req.player_id}: move action for player "
                                  f"{req.player_id} requires from and to coordinates"))
            f = getattr(a, "from")
            move = (f.x, f.y, a.to.x, a.to.y, not a.half)
        elif has_action and req.action.type != common.ACTION_TYPE_UNSPECIFIED:
            return done(game.SubmitActionResponse(
                success=False, error_code=common.ERROR_CODE_INVALID_TURN,
                error_message=f"invalid action for game {g.id} player {req.player_id}: unsupported action type"))
        if move is not None:  # ValidateCoreAction against the CURRENT board (action_validator.go:114-139)
            with g.pool.lock:
                st = g.pool.engine.get_state(g.slot, 1)
            code = validate_move(st["owner"][0], st["army"][0], st["type"][0], g.pool.W, g.pool.H, req.player_id, *move[:4])
            if code:
                return done(game.SubmitActionResponse(
                    success=False, error_code=common.ERROR_CODE_INVALID_TURN,
                    error_message=f"action validation failed for game {g.id} player {req.player_id} turn {current_turn}: "
                                  f"player {req.player_id}: move from ({move[0]},{move[1]}) to ({move[2]},{move[3]}): "
                                  f"{_VALIDATE_MSG[code]}"))
        with g.mu:
            g.actions[req.player_id] = move
            g.last_activity = time.time()
            all_in = len(g.actions) >= len(g.players)  # game_manager.go:554-573
        if all_in:
            if not self._process_turn(g):
                return done(game.SubmitActionResponse(
                    success=False, error_code=common.ERROR_CODE_UNSPECIFIED,
                    error_message=f"failed to process turn {current_turn} for game {g.id}"))
            if g.phase == common.GAME_PHASE_RUNNING:
                self._start_turn_timer(g)
        return done(game.SubmitActionResponse(success=True, next_turn_number=current_turn + 1))

    def GetGameState(self, req, ctx):
        g = self.games.get(req.game_id)
        if g is None:
            ctx.abort(grpc.StatusCode.NOT_FOUND, f"game {req.game_id} not found: requested by player {req.player_id}")
        if not self._auth(g, req.player_id, req.player_token):
            ctx.abort(grpc.StatusCode.PERMISSION_DENIED, f"invalid player credentials for game {g.id}: player {req.player_id}")
        return game.GetGameStateResponse(state=self._game_state(g, req.player_id))

    def StreamGame(self, req, ctx):
        g = self.games.get(req.game_id)
        if g is None:
            ctx.abort(grpc.StatusCode.NOT_FOUND, f"game {req.game_id} not found")
        if not self._auth(g, req.player_id, req.player_token):
            ctx.abort(grpc.StatusCode.PERMISSION_DENIED, f"invalid player credentials for game {g.id}: player {req.player_id}")
        q: deque = deque(maxlen=10)  # the reference's buffered channel of 10
        with g.stream_cv:
            g.streams[req.player_id] = q
        first = game.GameUpdate(full_state=self._game_state(g, req.player_id))
        _now(first.timestamp)
        yield first
        try:
            while ctx.is_active():
                with g.stream_cv:
                    if not q:
                        g.stream_cv.wait(0.25)
                    items = list(q)
                    q.clear()
                for u in items:
                    yield u
                if g.phase == common.GAME_PHASE_ENDED and not items:
                    return
        finally:
            with g.stream_cv:
                if g.streams.get(req.player_id) is q:
                    del g.streams[req.player_id]

    # ------------------------------------------------------------------ turn processing
    def _start_turn_timer(self, g: GameInstance):
        ms = g.config.turn_time_ms
        if ms <= 0:
            return
        with g.mu:
            if g.timer:
                g.timer.cancel()
            turn = g.current_turn

            def fire():  # processTurnTimeout (game_manager.go:693-732): play the turn with what was collected
                if g.phase == common.GAME_PHASE_RUNNING and g.current_turn == turn:
                    self._process_turn(g)
                    if g.phase == common.GAME_PHASE_RUNNING:
                        self._start_turn_timer(g)

            g.timer = threading.Timer(ms / 1000.0, fire)
            g.timer.daemon = True
            g.timer.start()

    def _process_turn(self, g: GameInstance) -> bool:
        """gameInstance.processTurn (game_manager.go:576-647): ONE turn of this game's slot."""
        pool = g.pool
        P = pool.P
        with g.mu, pool.lock:
            moves, g.actions = g.actions, {}
            acts = make_actions(pool.S, pool.engine.A)
            acts["flags"][:, 0] = _abi.ACTION_FLAG_SKIP_ENV       # every other game waits at its own barrier
            acts["flags"][g.slot, 0] = 0
            k = 0
            for pid in sorted(moves):
                m = moves[pid]
                if m is not None and k < pool.engine.A:
                    set_action(acts, g.slot, k, pid, m[0], m[1], m[2], m[3], m[4])
                    k += 1
            collect = g.config.collect_experiences
            if collect:
                prev_obs = pool.obs[g.slot].cpu().numpy().copy()
                prev_mask = pool.ser_mask[g.slot].cpu().numpy().copy()
            pool.engine.step_fused(acts, pool.outputs())
            pool.engine.mask(_abi.MASK_SERIALIZER_UDLR, pool.ser_mask)
            pool.engine.mask(_abi.MASK_ENGINE_URDL, pool.eng_mask)
            err = int(pool.err[g.slot])
            st = pool.engine.get_state(g.slot, 1)
            if err not in (0, _abi.STEP_ARMY_OVERFLOW):
                # Engine.Step returned the first validation error AFTER the half-applied turn (Q5);
                # processTurn returns before it updates currentTurn (game_manager.go:602-605)
                return False
            g.current_turn = int(st["turn"][0])
            over = bool(st["game_over"][0])
            if collect:
                self._collect(g, prev_obs, prev_mask, over)
            alive = [bool(v) for v in st["alive"][0]]
            for pid in range(P):
                if g.prev_alive[pid] and not alive[pid]:
                    ev = game.GameUpdate()
                    ev.event.player_eliminated.player_id, ev.event.player_eliminated.eliminated_by = pid, -1
                    self._broadcast(g, lambda _pid, ev=ev: ev)
            g.prev_alive = alive
            if sum(alive) <= 1:
                if g.timer:
                    g.timer.cancel()
                g.phase = common.GAME_PHASE_ENDED
                ev = game.GameUpdate()
                ev.event.game_ended.winner_id = int(st["winner"][0])
                _now(ev.event.game_ended.ended_at)
                self._broadcast(g, lambda _pid, ev=ev: ev)
            if g.streams:
                self._broadcast(g, lambda pid: self._stream_update(g, st, pid))
            if g.phase == common.GAME_PHASE_ENDED:
                self._release_slot(g, st)
        return True

    def _release_slot(self, g: GameInstance, st=None):
        """Hand the game's env slot back to its pool.  The reference keeps a finished game's engine until
        cleanupGames drops the whole game (game_manager.go:257-343); here the engine is a slot of a shared batch, so
        the final state is frozen on the host (GetGameState keeps answering from it) and the slot is reused at once —
        a gym client creates a new game on every reset() (generals_env.py:167-177)."""
        pool = g.pool
        if pool is None or g.slot < 0:
            return
        with pool.lock:
            if st is None:
                st = pool.engine.get_state(g.slot, 1)
            g.final = (st, pool.eng_mask[g.slot].cpu().numpy().copy())
            pool.free.append(g.slot)
            g.slot = -1

    def cleanup_games(self, now: Optional[float] = None) -> int:
        """GameManager.cleanupGames (game_manager.go:257-343): finished games leave after finished_game_ttl,
        games without activity after abandoned_game_timeout; timers stopped, streams closed, slots released."""
        now = time.time() if now is None else now
        self._last_cleanup = now
        with self.mu:
            refs = list(self.games.items())
        drop = []
        # A sweep can run on a request thread that already holds ITS game's lock (a JoinGame that found every slot
        # taken): a game whose lock is busy is in use — it is left for the next sweep instead of being waited for, so two
        # such threads cannot wait for each other.
        for gid, g in refs:
            if not g.mu.acquire(timeout=0.05):
                continue
            try:
                idle = now - g.last_activity
                if (g.phase == common.GAME_PHASE_ENDED and idle > self.finished_game_ttl) or \
                        (g.phase != common.GAME_PHASE_ENDED and idle > self.abandoned_game_timeout):
                    drop.append((gid, g))
            finally:
                g.mu.release()
        dropped = []
        for gid, g in drop:
            if not g.mu.acquire(timeout=0.05):
                continue
            try:
                if g.timer:
                    g.timer.cancel()
                if g.phase != common.GAME_PHASE_ENDED:
                    g.phase = common.GAME_PHASE_ENDED     # an abandoned game stops accepting actions
                self._release_slot(g)
                g.idempotency.clear()
                dropped.append((gid, g))
            finally:
                g.mu.release()
            with g.stream_cv:                              # StreamManager.CloseAll
                g.streams.clear()
                g.stream_cv.notify_all()
        with self.mu:
            for gid, _ in dropped:
                self.games.pop(gid, None)
        return len(dropped)

    def _collect(self, g, prev_obs, prev_mask, over):
        """SimpleCollector.OnStateTransition (collector.go:30-98)."""
        pool = g.pool
        aidx = pool.aidx[g.slot].cpu().numpy()
        reward = pool.reward[g.slot].cpu().numpy()
        nxt = pool.obs[g.slot].cpu().numpy()
        out = []
        for pid in range(pool.P):
            if aidx[pid] < 0:
                continue
            x = experience.Experience(experience_id=str(uuid.uuid4()), game_id=g.id, player_id=pid, turn=g.current_turn,
                                      action=0 if self.reference_compat else int(aidx[pid]), reward=float(reward[pid]),
                                      done=over)
            x.state.shape.extend([9, pool.H, pool.W])
            x.state.data.extend(prev_obs[pid].reshape(-1).tolist())
            x.next_state.shape.extend([9, pool.H, pool.W])
            x.next_state.data.extend(nxt[pid].reshape(-1).tolist())
            x.action_mask.extend(prev_mask[pid].astype(bool).tolist())
            _now(x.collected_at)
            out.append(x)
        self.store.add(out)

    # ------------------------------------------------------------------ state conversion
    def _player_states(self, g, st, viewer=None):
        W = g.pool.W
        out = []
        for i, p in enumerate(g.players):
            ps = game.PlayerState(id=p.id, name=p.name, status=common.PLAYER_STATUS_ACTIVE,
                                  army_count=int(st["army_count"][0, i]), tile_count=int(st["owned"][0, i].sum()),
                                  color="#%06X" % (i * 0x333333))
            alive = bool(st["alive"][0, i])
            if not alive:
                ps.status = common.PLAYER_STATUS_ELIMINATED
            gi = int(st["general_idx"][0, i])
            show = (not alive) if viewer is None else (not alive or int(st["owner"][0, gi]) == viewer)
            if gi >= 0 and show:
                ps.general_position.x, ps.general_position.y = gi % W, gi // W
            out.append(ps)
        return out

    @staticmethod
    def _view_tile(st, i, viewer, fog_on):
        vis = (not fog_on) or bool((int(st["visible"][0, i]) >> viewer) & 1)
        typ = int(st["type"][0, i])
        fog = fog_on and not vis and typ != _abi.TILE_NORMAL
        t = game.Tile(type=_TILE_TYPE[typ], owner_id=int(st["owner"][0, i]), army_count=int(st["army"][0, i]), visible=vis,
                      fog_of_war=fog)
        if not vis and not fog:
            t.type, t.owner_id, t.army_count = common.TILE_TYPE_NORMAL, -1, 0
        elif fog and not vis:
            t.owner_id, t.army_count = -1, 0
        return t

    def _game_state(self, g: GameInstance, viewer: int):
        """createGameState / convertGameStateToProto (server.go:462-610)."""
        cfg = g.config
        s = game.GameState(game_id=g.id, status=_PHASE_STATUS.get(g.phase, common.GAME_STATUS_WAITING), winner_id=-1,
                           current_phase=g.phase)
        s.board.width, s.board.height = cfg.width, cfg.height
        if g.pool is None:  # lobby placeholder (server.go:472-515)
            for i, p in enumerate(g.players):
                s.players.add(id=p.id, name=p.name, status=common.PLAYER_STATUS_ACTIVE, army_count=1, tile_count=1,
                              color="#%06X" % (i * 0x333333))
            for _ in range(cfg.width * cfg.height):
                s.board.tiles.add(type=common.TILE_TYPE_NORMAL, owner_id=-1, army_count=0, visible=True, fog_of_war=False)
            return s
        pool = g.pool
        with pool.lock:
            if g.slot < 0:      # finished: the slot went back to the pool, the final state is frozen on the host
                st, masks = g.final
            else:
                st, masks = pool.engine.get_state(g.slot, 1), None
            if not 0 <= viewer < pool.P:
                mask = np.zeros(pool.N * 4, np.uint8)
            else:
                mask = masks[viewer] if masks is not None else pool.eng_mask[g.slot, viewer].cpu().numpy()
        s.turn = int(st["turn"][0])
        s.players.extend(self._player_states(g, st, viewer))
        fog_on = bool(pool.engine.cfg.fog_of_war)
        s.board.tiles.extend(self._view_tile(st, i, viewer, fog_on) for i in range(pool.N))
        if st["game_over"][0]:
            s.winner_id = int(st["winner"][0])
        s.action_mask.extend(mask.astype(bool).tolist())
        if g.started_at:
            s.started_at.seconds = int(g.started_at)
        return s

    def _stream_update(self, g, st, pid):
        """createStreamUpdate (server.go:638-777): delta when 0 < changes < N/5, else the full view."""
        pool = g.pool
        chg = np.nonzero(st["changed"][0])[0]
        vch = np.nonzero(st["vis_changed"][0])[0]
        total = len(chg) + len(vch)
        u = game.GameUpdate()
        _now(u.timestamp)
        if 0 < total < pool.N // 5:
            u.delta.turn = int(st["turn"][0])
            fog_on = bool(pool.engine.cfg.fog_of_war)
            seen = set()
            for i in list(chg) + list(vch):
                if int(i) in seen:
                    continue
                seen.add(int(i))
                tu = u.delta.tile_updates.add()
                tu.position.x, tu.position.y = int(i) % pool.W, int(i) // pool.W
                tu.tile.CopyFrom(self._view_tile(st, int(i), pid, fog_on))
            for ps in self._player_states(g, st, None):
                u.delta.player_updates.add(player_id=ps.id, state=ps)
        else:
            u.full_state.CopyFrom(self._game_state(g, pid))
        return u

    def _broadcast(self, g, make):
        with g.stream_cv:
            for pid, q in g.streams.items():
                q.append(make(pid))
            g.stream_cv.notify_all()

    # ------------------------------------------------------------------ ExperienceService
    @staticmethod
    def _matches(req, x) -> bool:
        if req.game_ids and x.game_id not in req.game_ids:
            return False
        if req.player_ids and x.player_id not in req.player_ids:
            return False
        return x.turn >= req.min_turn

    @staticmethod
    def _validate_stream_request(req, ctx):
        """validateStreamRequest (experience_service.go:510-520), reported as :161-163 / :292-294 do."""
        if req.batch_size > 1000:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "invalid request: batch size too large (max 1000)")

    def StreamExperiences(self, req, ctx):
        self._validate_stream_request(req, ctx)
        cursor = 0
        while ctx.is_active():
            items, cursor = self.store.read_from(cursor, 0.1)
            for x in items:
                if self._matches(req, x):
                    yield x
            if not req.follow and not items:
                return

    def StreamExperienceBatches(self, req, ctx):
        self._validate_stream_request(req, ctx)
        size = req.batch_size if req.batch_size > 0 else 32          # experience_service.go:510-514
        wait = (req.max_batch_wait_ms if req.max_batch_wait_ms > 0 else 100) / 1000.0
        stream_id, batch_id, cursor = str(uuid.uuid4()), 0, 0
        pending: List = []
        oldest = None
        while ctx.is_active():
            items, cursor = self.store.read_from(cursor, wait / 2)
            for x in items:
                if self._matches(req, x):
                    pending.append(x)
                    oldest = oldest or time.time()
            flush = len(pending) >= size or (pending and time.time() - oldest >= wait) or (pending and not req.follow and not items)
            while flush and pending:
                batch_id += 1                                        # experience_service.go:342: atomic.AddInt32 -> ids from 1
                b = experience.ExperienceBatch(batch_id=batch_id, stream_id=stream_id)
                b.experiences.extend(pending[:size])
                _now(b.created_at)
                b.metadata["batch_size"] = str(len(b.experiences))    # :345-347
                if req.enable_compression:                           # :350-353: requested, not implemented by the reference
                    b.metadata["compression"] = "none"
                pending = pending[size:]
                oldest = time.time() if pending else None
                yield b
                flush = len(pending) >= size or (pending and not req.follow and not items)
            if not req.follow and not items and not pending:
                return

    def SubmitExperiences(self, req, ctx):
        # experience_service.go:381-448: the request-level checks abort, an invalid experience (validateExperience,
        # :522-540) is counted as rejected.  Beyond the reference, an experience_id already submitted is rejected as the
        # duplicate the proto's `rejected` field names (experience.proto:103).
        if len(req.experiences) == 0:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "no experiences provided")
        if len(req.experiences) > 1000:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "too many experiences (max 1000)")
        if not req.experiences[0].game_id:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "game ID required")
        accepted = rejected = 0
        fresh = []
        for x in req.experiences:
            if not x.game_id or not x.HasField("state") or not x.HasField("next_state") or \
                    len(x.state.data) == 0 or len(x.next_state.data) == 0:
                rejected += 1
                continue
            if x.experience_id and x.experience_id in self.submitted_ids:
                rejected += 1
                continue
            if x.experience_id:
                self.submitted_ids.add(x.experience_id)
            fresh.append(x)
            accepted += 1
        self.store.add(fresh)
        return experience.SubmitExperiencesResponse(accepted=accepted, rejected=rejected)

    def GetExperienceStats(self, req, ctx):
        with self.store.cv:
            xs = [x for (_, x) in self.store.items if not req.game_ids or x.game_id in req.game_ids]
        r = experience.GetExperienceStatsResponse(total_experiences=len(xs), total_games=len({x.game_id for x in xs}))
        for x in xs:
            r.experiences_per_game[x.game_id] += 1
            r.experiences_per_player[x.player_id] += 1
        if xs:
            rewards = [x.reward for x in xs]
            r.average_reward, r.min_reward, r.max_reward = float(np.mean(rewards)), min(rewards), max(rewards)
            r.oldest_experience.CopyFrom(xs[0].collected_at)
            r.newest_experience.CopyFrom(xs[-1].collected_at)
        return r

    def ingest_records(self, records, width: int, height: int, game_prefix: str = "env", players: int = 2) -> int:
        """Feed device-gathered experience records (``sharding.gather_experience`` dicts of tensors) into the stream
        store, as the per-game collectors do for served games.  Either float32 records (``sharding.pack_experience``:
        ``state`` / ``next_state`` tensors, ``mask_bits`` = the serializer mask bytes [R, 4*W*H],
        GRL_MASK_SERIALIZER_UDLR) or compact ones (``sharding.pack_experience_packed``), which are expanded here on the
        host (grl_expand_obs + the serializer mask derived from the packed record).  Messages are framed straight in
        protobuf wire format from the arrays' bytes (a packed repeated float IS its little-endian bytes) instead of
        element by element.  Returns the number of records added."""
        if "state_packed" in records:
            from . import sharding

            host = sharding.expand_experience(records, self.lib, width, height, players)
        else:
            host = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in records.items()
                    if k not in ("dropped", "counts")}
        n = int(host["action"].shape[0])
        shape = _packed_field(1, b"".join(_varint(d) for d in (9, height, width)))
        state = np.ascontiguousarray(host["state"], dtype="<f4").reshape(n, -1)
        nxt = np.ascontiguousarray(host["next_state"], dtype="<f4").reshape(n, -1)
        mask = np.ascontiguousarray(host["mask_bits"]).reshape(n, -1).astype(np.uint8)
        reward = np.ascontiguousarray(host["reward"], dtype="<f4")
        t = time.time()
        stamp = _packed_field(11, _tag(1, 0) + _varint(int(t)) + _tag(2, 0) + _varint(int((t % 1) * 1e9)))
        out = []
        for i in range(n):
            gid = f"{game_prefix}-{int(host['env_id'][i])}".encode()
            wire = b"".join((
                _packed_field(1, str(uuid.uuid4()).encode()), _packed_field(2, gid),
                _tag(3, 0), _varint(int(host["player"][i])), _tag(4, 0), _varint(int(host["turn"][i])),
                _packed_field(5, shape + _packed_field(2, state[i].tobytes())),
                _tag(6, 0), _varint(int(host["action"][i]) & 0xFFFFFFFFFFFFFFFF),
                _tag(7, 5), reward[i].tobytes(),
                _packed_field(8, shape + _packed_field(2, nxt[i].tobytes())),
                (_tag(9, 0) + b"\x01") if host["done"][i] else b"",
                _packed_field(10, (mask[i] != 0).astype(np.uint8).tobytes()), stamp))
            out.append(experience.Experience.FromString(wire))
        self.store.add(out)
        return n

    # ------------------------------------------------------------------ plumbing
    def add_to_server(self, server: grpc.Server) -> None:
        for sname, methods in SERVICES.items():
            ns = game if "game" in sname else experience
            handlers = {}
            for mname, (req, resp, streaming) in methods.items():
                fn = getattr(self, mname)
                make = grpc.unary_stream_rpc_method_handler if streaming else grpc.unary_unary_rpc_method_handler
                handlers[mname] = make(fn, request_deserializer=getattr(ns, req).FromString,
                                       response_serializer=getattr(ns, resp).SerializeToString)
            server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(sname, handlers),))

    def close(self):
        for g in self.games.values():
            if g.timer:
                g.timer.cancel()
        for p in list(self.pools.values()) + [q for ps in self.more_pools.values() for q in ps]:
            p.engine.close()


def serve(address: str = "127.0.0.1:50051", max_workers: int = 16, **kwargs):
    """Start a gRPC server; returns (grpc.Server, GameServer, bound_port)."""
    gs = GameServer(**kwargs)
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    gs.add_to_server(server)
    port = server.add_insecure_port(address)
    server.start()
    return server, gs, port


class Stub:
    """Client stub for either service over a channel (what ``*_pb2_grpc.*Stub`` would give)."""

    def __init__(self, channel: grpc.Channel, service: str):
        ns = game if "game" in service else experience
        for mname, (req, resp, streaming) in SERVICES[service].items():
            make = channel.unary_stream if streaming else channel.unary_unary
            setattr(self, mname, make(f"/{service}/{mname}", request_serializer=getattr(ns, req).SerializeToString,
                                      response_deserializer=getattr(ns, resp).FromString))
