"""ASCII board renderer: ``Engine.Board(playerID)`` (internal/game/rendering.go:34-143).

Produces the reference's string byte for byte — header row, one line per board row with the
ANSI colour codes, the legend — from the planar state of one env (``grl_get_state``).
Host-side debugging aid; nothing here touches the turn path.
"""
from __future__ import annotations

COLOR_RESET = "\033[0m"
COLOR_WHITE = "\033[37m"
COLOR_GRAY = "\033[90m"
# rendering.go:31: red, blue, green, yellow, purple, cyan
PLAYER_COLORS = ["\033[31m", "\033[34m", "\033[32m", "\033[33m", "\033[35m", "\033[36m"]
EMPTY, CITY, GENERAL, MOUNTAIN = "·", "⬢", "♔", "▲"
PLAYER_SYMBOLS = "ABCDEFGH"
T_NORMAL, T_GENERAL, T_CITY, T_MOUNTAIN = 0, 1, 2, 3


def _player_color(pid: int) -> str:  # rendering.go:212-217
    return PLAYER_COLORS[pid] if 0 <= pid < len(PLAYER_COLORS) else COLOR_WHITE


def _tile(owner: int, army: int, type_: int, visible: bool) -> str:
    """getTileDisplayDirect (rendering.go:86-143)."""
    out = []
    if not visible:
        out += [COLOR_GRAY, " "]
    elif type_ == T_MOUNTAIN:
        out += [COLOR_GRAY, " ", MOUNTAIN]
    elif type_ == T_GENERAL:
        out += [_player_color(owner), PLAYER_SYMBOLS[owner % len(PLAYER_SYMBOLS)], GENERAL]
    elif type_ == T_CITY and owner < 0:
        out += [COLOR_WHITE, " ", CITY]
    elif type_ == T_CITY:
        out += [_player_color(owner), PLAYER_SYMBOLS[owner % len(PLAYER_SYMBOLS)], CITY]
    elif owner < 0 and type_ == T_NORMAL:
        out.append(COLOR_GRAY)
        if army == 0:
            out += [" ", EMPTY]
        elif army >= 100:
            out.append("++")
        elif army >= 10:
            out.append("%2d" % army)
        else:
            out += [" ", "%1d" % army]
    elif type_ == T_NORMAL:
        out += [_player_color(owner), PLAYER_SYMBOLS[owner % len(PLAYER_SYMBOLS)]]
        if army >= 100:
            out.append("+")
        elif army >= 10:
            out.append("%1d" % army)  # fmt "%*d" with width 1 prints every digit (rendering.go:132)
        else:
            out.append(" ")
    out += [COLOR_RESET, " "]
    return "".join(out)


def render_state(owner, army, type_, visible_bits, width: int, height: int, player_id: int, fog: bool = True) -> str:
    """Engine.Board for one game given its planar state (row-major lists/arrays of length W*H)."""
    rows = ["    " + "".join("%2d" % x for x in range(width)) + "\n"]
    for y in range(height):
        line = ["%2d" % y, " "]
        for x in range(width):
            i = y * width + x
            vis = player_id < 0 or not fog or bool((int(visible_bits[i]) >> player_id) & 1)
            line.append(_tile(int(owner[i]), int(army[i]), int(type_[i]), vis))
        rows.append("".join(line) + "\n")
    rows.append("\n" + EMPTY + "=empty " + CITY + "=city " + GENERAL + "=general " + MOUNTAIN + "=mountain A-H=players\n")
    return "".join(rows)


def render_board(engine, env: int = 0, player_id: int = 0) -> str:
    """``Engine.Board(playerID)`` of env slot ``env`` of a BatchedEngine (player_id < 0: no fog)."""
    st = engine.get_state(env, 1)
    return render_state(st["owner"][0], st["army"][0], st["type"][0], st["visible"][0], engine.W, engine.H, player_id,
                        fog=bool(engine.cfg.fog_of_war))
