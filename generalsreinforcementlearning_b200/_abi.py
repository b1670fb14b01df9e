"""ctypes mirror of include/grlcuda.h (types + prototypes).

The same binder serves the CUDA product (prefix ``grl_``) and, in tests only,
the CPU oracle (prefix ``grlo_``): both export the identical C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

GRL_MAX_DIM = 32
GRL_MAX_PLAYERS = 8
GRL_MAX_ACTIONS = 12
GRL_OBS_CHANNELS = 9

GRL_OK = 0
STEP_OK = 0
STEP_INVALID_COORDINATES = 1
STEP_NOT_ADJACENT = 2
STEP_NOT_OWNED = 3
STEP_INSUFFICIENT_ARMY = 4
STEP_GAME_OVER = 5
STEP_MOVE_TO_SELF = 7
STEP_TARGET_IS_MOUNTAIN = 8
STEP_ARMY_OVERFLOW = 100

TILE_NORMAL, TILE_GENERAL, TILE_CITY, TILE_MOUNTAIN = 0, 1, 2, 3
NEUTRAL = -1

MASK_ENGINE_URDL = 0
MASK_SERIALIZER_UDLR = 1
MASK_ENGINE_URDL_BITS = 2
MASK_ENGINE_HALF_BITS = 3

STEP_FLAG_RANDOM_POLICY = 1
ACTION_FLAG_SKIP_ENV = 1
GRL_GYM_CHANNELS = 9

ACTION_DTYPE = np.dtype(
    [
        ("player_id", "i1"),
        ("from_x", "i1"),
        ("from_y", "i1"),
        ("to_x", "i1"),
        ("to_y", "i1"),
        ("move_all", "u1"),
        ("present", "u1"),
        ("flags", "u1"),
    ]
)
assert ACTION_DTYPE.itemsize == 8


class RewardConfig(C.Structure):
    _fields_ = [
        (n, C.c_float)
        for n in (
            "win_game",
            "lose_game",
            "capture_city",
            "lose_city",
            "capture_general",
            "lose_general",
            "territory_gained",
            "territory_lost",
            "army_gained",
            "army_lost",
            "army_advantage",
        )
    ]


class Config(C.Structure):
    _fields_ = [
        (n, C.c_int32)
        for n in (
            "num_envs",
            "width",
            "height",
            "num_players",
            "device",
            "max_actions",
            "fog_of_war",
            "env_id_base",
            "city_ratio",
            "city_start_army",
            "min_general_spacing",
            "production_general",
            "production_city",
            "production_normal",
            "normal_growth_interval",
            "host_threads",
        )
    ] + [("reward", RewardConfig)]


class StepOutputs(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p),
        ("mask_bits", C.c_void_p),
        ("reward", C.c_void_p),
        ("done", C.c_void_p),
        ("winner", C.c_void_p),
        ("step_error", C.c_void_p),
        ("action_index", C.c_void_p),
        ("obs_packed", C.c_void_p),
    ]


class GymOutputs(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("mask", C.c_void_p), ("stats", C.c_void_p)]


class GymStepIO(C.Structure):
    _fields_ = [("action", C.c_void_p), ("opponent_action", C.c_void_p), ("out", GymOutputs), ("actions", C.c_void_p),
                ("prev_stats", C.c_void_p), ("turns", C.c_void_p), ("calls", C.c_void_p), ("reward", C.c_void_p),
                ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("valid", C.c_void_p), ("done", C.c_void_p),
                ("winner", C.c_void_p), ("step_error", C.c_void_p), ("n_finished", C.c_void_p),
                ("agent_seed", C.c_uint64), ("sampled_action", C.c_void_p)]


class GymAutoresetIO(C.Structure):
    _fields_ = [("terminated", C.c_void_p), ("truncated", C.c_void_p), ("episode", C.c_void_p), ("turns", C.c_void_p),
                ("calls", C.c_void_p), ("out", GymOutputs), ("final_obs", C.c_void_p), ("n_reset", C.c_void_p)]


class ReplayRowsIO(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("final_obs", C.c_void_p), ("done", C.c_void_p), ("next_states", C.c_void_p),
                ("states", C.c_void_p), ("capacity", C.c_int64), ("next_row0", C.c_int64), ("state_row0", C.c_int64),
                ("views", C.c_int32), ("view", C.c_int32), ("obs_floats", C.c_int32), ("reserved", C.c_int32)]


STATE_FIELDS = (
    ("owner", np.int32, "N"),
    ("army", np.int32, "N"),
    ("type", np.int32, "N"),
    ("visible", np.uint32, "N"),
    ("owned", np.uint8, "PN"),
    ("changed", np.uint8, "N"),
    ("vis_changed", np.uint8, "N"),
    ("turn", np.int32, ""),
    ("game_over", np.int32, ""),
    ("winner", np.int32, ""),
    ("alive", np.int32, "P"),
    ("army_count", np.int32, "P"),
    ("general_idx", np.int32, "P"),
    ("step_error", np.int32, ""),
)


class StatePlanes(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in STATE_FIELDS]


# symbols every implementation of the ABI must export (tests check the .so against this
# list AND against a parse of include/grlcuda.h)
ABI_FUNCTIONS = {
    "abi_version": (C.c_int, []),
    "status_string": (C.c_char_p, [C.c_int]),
    "last_error": (C.c_char_p, []),
    "default_config": (C.c_int, [C.POINTER(Config)]),
    "create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "destroy": (C.c_int, [C.c_void_p]),
    "sync": (C.c_int, [C.c_void_p]),
    "set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "get_config": (C.c_int, [C.c_void_p, C.POINTER(Config)]),
    "reset_seeded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "reset_boards": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mapgen": (C.c_int, [C.POINTER(Config), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]),
    "step_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(StepOutputs)]),
    "observe": (C.c_int, [C.c_void_p, C.POINTER(StepOutputs)]),
    "mask": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "visibility": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gym_observe": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(GymOutputs)]),
    "gym_observe_envs": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(GymOutputs)]),
    "gym_autoreset": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(GymAutoresetIO)]),
    "gym_sample": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p]),
    "replay_push_rows": (C.c_int, [C.c_void_p, C.POINTER(ReplayRowsIO)]),
    "gym_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "gym_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(GymStepIO)]),
    "sample_actions": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "get_state": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(StatePlanes)]),
    "set_state": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(StatePlanes)]),
    "state_hash": (C.c_int, [C.c_void_p, C.c_void_p]),
    "buffer_hash": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]),
    "stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "launch_count": (C.c_int, [C.c_void_p, C.c_void_p]),
    "obs_packed_words": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "expand_obs": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
}


class BoundLibrary:
    """A loaded shared object exporting the ABI under ``prefix``."""

    def __init__(self, path: str, prefix: str):
        self.path = path
        self.prefix = prefix
        self.cdll = C.CDLL(path)
        for name, (restype, argtypes) in ABI_FUNCTIONS.items():
            fn = getattr(self.cdll, prefix + name)
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)

    def check(self, status: int, what: str) -> None:
        if status != GRL_OK:
            detail = self.last_error()
            msg = self.status_string(status)
            raise RuntimeError(
                f"{self.prefix}{what} failed: {msg.decode() if msg else status}"
                + (f" ({detail.decode()})" if detail else "")
            )
