"""The reference's demo policy ``GenerateRandomActions`` (internal/game/demo_helpers.go:12-62)
and the Go ``math/rand`` (v1) generator it draws from.

Host-side and per game, like the reference: it exists for demos, replays of the reference's
UI games and as a baseline agent; the batched synthetic policy of the benchmark is the
counter-based one inside the CUDA library (``GRL_STEP_FLAG_RANDOM_POLICY``).

``GoRand`` restates go1.24 ``math/rand`` (rng.go: additive lagged Fibonacci 607/273 seeded by
``seedrand``; rand.go: Int63, Int31, Int31n, Intn, Float64, Float32).  Its 607-word ``rngCooked``
table is the reconstructed one shared with the C++ mapgen (csrc/go_rng_cooked.inc, SURVEY
Appendix B); tests pin it against the canonical ``Seed(1)`` outputs and the oracle's generator.
"""
from __future__ import annotations

import os
import re
import struct
from typing import List, Optional

_LEN, _TAP, _MAX31, _MASK63 = 607, 273, (1 << 31) - 1, (1 << 63) - 1
_COOKED: Optional[List[int]] = None


def _cooked() -> List[int]:
    global _COOKED
    if _COOKED is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "go_rng_cooked.inc")
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        vals = [int(tok.rstrip("LlUu"), 0) for tok in re.findall(r"-?(?:0x[0-9a-fA-F]+|\d+)[LlUu]*", text)]
        assert len(vals) == _LEN, f"go_rng_cooked.inc holds {len(vals)} values"
        _COOKED = [v & 0xFFFFFFFFFFFFFFFF for v in vals]
    return _COOKED


def _seedrand(x: int) -> int:  # rng.go: x[n+1] = 48271 * x[n] mod (2**31 - 1)
    hi, lo = divmod(x, 44488)
    x = 48271 * lo - 3399 * hi
    return x + _MAX31 if x < 0 else x


class GoRand:
    """``rand.New(rand.NewSource(seed))``."""

    def __init__(self, seed: int):
        self.seed(seed)

    def seed(self, seed: int) -> None:
        self.tap, self.feed = 0, _LEN - _TAP
        seed %= _MAX31          # Go's % keeps the sign of the dividend; Python's is already non-negative
        if seed == 0:
            seed = 89482311
        x = seed
        cooked = _cooked()
        self.vec = [0] * _LEN
        for i in range(-20, _LEN):
            x = _seedrand(x)
            if i >= 0:
                u = (x << 40) & 0xFFFFFFFFFFFFFFFF
                x = _seedrand(x)
                u ^= (x << 20) & 0xFFFFFFFFFFFFFFFF
                x = _seedrand(x)
                u ^= x
                self.vec[i] = u ^ cooked[i]

    def uint64(self) -> int:
        self.tap = self.tap - 1 if self.tap > 0 else _LEN - 1
        self.feed = self.feed - 1 if self.feed > 0 else _LEN - 1
        x = (self.vec[self.feed] + self.vec[self.tap]) & 0xFFFFFFFFFFFFFFFF
        self.vec[self.feed] = x
        return x

    def int63(self) -> int:
        return self.uint64() & _MASK63

    def int31(self) -> int:
        return self.int63() >> 32

    def int31n(self, n: int) -> int:
        if n & (n - 1) == 0:
            return self.int31() & (n - 1)
        mx = (1 << 31) - 1 - (1 << 31) % n
        v = self.int31()
        while v > mx:
            v = self.int31()
        return v % n

    def intn(self, n: int) -> int:
        if n <= 0:
            raise ValueError("invalid argument to Intn")
        if n <= _MAX31:
            return self.int31n(n)
        mx = (1 << 63) - 1 - (1 << 63) % n
        v = self.int63()
        while v > mx:
            v = self.int63()
        return v % n

    def float64(self) -> float:
        while True:  # Go 1 value stream: float64(Int63()) / (1 << 63), re-drawn when it rounds to 1
            f = float(self.int63()) / float(1 << 63)
            if f != 1.0:
                return f

    def float32(self) -> float:
        while True:
            f = struct.unpack("f", struct.pack("f", self.float64()))[0]
            if f != 1.0:
                return f


def generate_random_actions(owner, army, type_, alive, width: int, height: int, rng: GoRand):
    """GenerateRandomActions for one game: ``[(player, fx, fy, tx, ty, move_all), ...]``.

    Per alive player: with probability 0.7 no action; otherwise every legal (tile, direction)
    pair in row-major tile order and direction order down, up, right, left, each with its own
    MoveAll draw (``Float32() < 0.7``), then one is chosen uniformly."""
    out = []
    for pid, is_alive in enumerate(alive):
        if not is_alive:
            continue
        if rng.float32() > 0.3:
            continue
        moves = []
        for y in range(height):
            for x in range(width):
                i = y * width + x
                if owner[i] != pid or army[i] <= 1:
                    continue
                for dx, dy in ((0, 1), (0, -1), (1, 0), (-1, 0)):
                    tx, ty = x + dx, y + dy
                    if tx < 0 or tx >= width or ty < 0 or ty >= height:
                        continue
                    if type_[ty * width + tx] == 3:
                        continue
                    moves.append((pid, x, y, tx, ty, rng.float32() < 0.7))
        if moves:
            out.append(moves[rng.intn(len(moves))])
    return out


def demo_actions_for(engine, env: int, rng: GoRand):
    """The demo policy for env slot ``env`` of a BatchedEngine, as a filled grl_action row."""
    from .engine import make_actions, set_action

    st = engine.get_state(env, 1)
    moves = generate_random_actions(st["owner"][0], st["army"][0], st["type"][0], st["alive"][0], engine.W, engine.H, rng)
    acts = make_actions(1, engine.A)
    for k, (pid, fx, fy, tx, ty, move_all) in enumerate(moves[: engine.A]):
        set_action(acts, 0, k, pid, fx, fy, tx, ty, move_all)
    return acts
