"""Wire schema of the reference's gRPC API, built at import time without protoc.

The image has ``grpcio`` + ``protobuf`` but no ``grpc_tools``, and the reference's generated stubs
(``python/generals_pb``) do not travel to the GPU box.  The wire contract is what matters, so the
three files of ``proto/`` (common/v1/common.proto, game/v1/game.proto,
experience/v1/experience.proto) are restated here as descriptor tables — same packages, message
names, field names, numbers and types — and turned into message classes through a private
descriptor pool.  ``tests/test_grpc_service.py`` checks byte-level compatibility against the
reference's own generated classes when ``/root/reference`` is present.

Usage:  ``from .grpc_schema import common, game, experience`` then ``game.GameState(...)``;
service/method tables are in ``SERVICES``.
"""
from __future__ import annotations

from types import SimpleNamespace

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory, timestamp_pb2

F = descriptor_pb2.FieldDescriptorProto
_SCALAR = {"int32": F.TYPE_INT32, "int64": F.TYPE_INT64, "bool": F.TYPE_BOOL, "string": F.TYPE_STRING,
           "float": F.TYPE_FLOAT}

COMMON_ENUMS = {
    "TileType": ["TILE_TYPE_UNSPECIFIED", "TILE_TYPE_NORMAL", "TILE_TYPE_GENERAL", "TILE_TYPE_CITY", "TILE_TYPE_MOUNTAIN"],
    "PlayerStatus": ["PLAYER_STATUS_UNSPECIFIED", "PLAYER_STATUS_ACTIVE", "PLAYER_STATUS_ELIMINATED",
                     "PLAYER_STATUS_DISCONNECTED"],
    "ActionType": ["ACTION_TYPE_UNSPECIFIED", "ACTION_TYPE_MOVE"],
    "ErrorCode": ["ERROR_CODE_UNSPECIFIED", "ERROR_CODE_INVALID_COORDINATES", "ERROR_CODE_NOT_ADJACENT",
                  "ERROR_CODE_NOT_OWNED", "ERROR_CODE_INSUFFICIENT_ARMY", "ERROR_CODE_GAME_OVER", "ERROR_CODE_INVALID_PLAYER",
                  "ERROR_CODE_MOVE_TO_SELF", "ERROR_CODE_TARGET_IS_MOUNTAIN", "ERROR_CODE_INVALID_TURN",
                  "ERROR_CODE_GAME_NOT_FOUND", "ERROR_CODE_GAME_FULL", "ERROR_CODE_ALREADY_JOINED",
                  "ERROR_CODE_INVALID_PHASE", "ERROR_CODE_PHASE_TRANSITION"],
    "GameStatus": ["GAME_STATUS_UNSPECIFIED", "GAME_STATUS_WAITING", "GAME_STATUS_IN_PROGRESS", "GAME_STATUS_FINISHED",
                   "GAME_STATUS_CANCELLED"],
    "GamePhase": ["GAME_PHASE_UNSPECIFIED", "GAME_PHASE_INITIALIZING", "GAME_PHASE_LOBBY", "GAME_PHASE_STARTING",
                  "GAME_PHASE_RUNNING", "GAME_PHASE_PAUSED", "GAME_PHASE_ENDING", "GAME_PHASE_ENDED", "GAME_PHASE_ERROR",
                  "GAME_PHASE_RESET"],
}
C = ".generals.common.v1."
G = ".generals.game.v1."
E = ".generals.experience.v1."
TS = ".google.protobuf.Timestamp"

# message -> [(name, number, type, label)]; type is a scalar name, an enum/message path; label "" | "repeated"
# | "oneof:<name>" | "map:<key>,<value>"
COMMON_MESSAGES = {"Coordinate": [("x", 1, "int32", ""), ("y", 2, "int32", "")]}

GAME_MESSAGES = {
    "CreateGameRequest": [("config", 1, G + "GameConfig", "")],
    "CreateGameResponse": [("game_id", 1, "string", ""), ("config", 2, G + "GameConfig", "")],
    "GameConfig": [("width", 1, "int32", ""), ("height", 2, "int32", ""), ("max_players", 3, "int32", ""),
                   ("fog_of_war", 4, "bool", ""), ("turn_time_ms", 5, "int32", ""), ("collect_experiences", 6, "bool", "")],
    "JoinGameRequest": [("game_id", 1, "string", ""), ("player_name", 2, "string", ""), ("player_token", 3, "string", "")],
    "JoinGameResponse": [("player_id", 1, "int32", ""), ("player_token", 2, "string", ""),
                         ("initial_state", 3, G + "GameState", "")],
    "SubmitActionRequest": [("game_id", 1, "string", ""), ("player_id", 2, "int32", ""), ("player_token", 3, "string", ""),
                            ("action", 4, G + "Action", ""), ("idempotency_key", 5, "string", "")],
    "SubmitActionResponse": [("success", 1, "bool", ""), ("error_code", 2, "enum:" + C + "ErrorCode", ""),
                             ("error_message", 3, "string", ""), ("next_turn_number", 4, "int32", "")],
    "Action": [("type", 1, "enum:" + C + "ActionType", ""), ("from", 2, C + "Coordinate", ""), ("to", 3, C + "Coordinate", ""),
               ("turn_number", 4, "int32", ""), ("half", 5, "bool", "")],
    "GetGameStateRequest": [("game_id", 1, "string", ""), ("player_id", 2, "int32", ""), ("player_token", 3, "string", "")],
    "GetGameStateResponse": [("state", 1, G + "GameState", "")],
    "StreamGameRequest": [("game_id", 1, "string", ""), ("player_id", 2, "int32", ""), ("player_token", 3, "string", "")],
    "GameUpdate": [("full_state", 1, G + "GameState", "oneof:update"), ("delta", 2, G + "GameStateDelta", "oneof:update"),
                   ("event", 3, G + "GameEvent", "oneof:update"), ("timestamp", 4, TS, "")],
    "GameState": [("game_id", 1, "string", ""), ("status", 2, "enum:" + C + "GameStatus", ""), ("turn", 3, "int32", ""),
                  ("board", 4, G + "Board", ""), ("players", 5, G + "PlayerState", "repeated"), ("winner_id", 6, "int32", ""),
                  ("started_at", 7, TS, ""), ("updated_at", 8, TS, ""), ("action_mask", 9, "bool", "repeated"),
                  ("current_phase", 10, "enum:" + C + "GamePhase", "")],
    "Board": [("width", 1, "int32", ""), ("height", 2, "int32", ""), ("tiles", 3, G + "Tile", "repeated")],
    "Tile": [("type", 1, "enum:" + C + "TileType", ""), ("owner_id", 2, "int32", ""), ("army_count", 3, "int32", ""),
             ("visible", 4, "bool", ""), ("fog_of_war", 5, "bool", "")],
    "PlayerState": [("id", 1, "int32", ""), ("name", 2, "string", ""), ("status", 3, "enum:" + C + "PlayerStatus", ""),
                    ("army_count", 4, "int32", ""), ("tile_count", 5, "int32", ""),
                    ("general_position", 6, C + "Coordinate", ""), ("color", 7, "string", "")],
    "GameStateDelta": [("turn", 1, "int32", ""), ("tile_updates", 2, G + "TileUpdate", "repeated"),
                       ("player_updates", 3, G + "PlayerUpdate", "repeated")],
    "TileUpdate": [("position", 1, C + "Coordinate", ""), ("tile", 2, G + "Tile", "")],
    "PlayerUpdate": [("player_id", 1, "int32", ""), ("state", 2, G + "PlayerState", "")],
    "GameEvent": [("player_joined", 1, G + "PlayerJoinedEvent", "oneof:event"),
                  ("player_eliminated", 2, G + "PlayerEliminatedEvent", "oneof:event"),
                  ("game_started", 3, G + "GameStartedEvent", "oneof:event"),
                  ("game_ended", 4, G + "GameEndedEvent", "oneof:event"),
                  ("player_disconnected", 5, G + "PlayerDisconnectedEvent", "oneof:event"),
                  ("player_reconnected", 6, G + "PlayerReconnectedEvent", "oneof:event"),
                  ("phase_changed", 7, G + "PhaseChangedEvent", "oneof:event")],
    "PlayerJoinedEvent": [("player_id", 1, "int32", ""), ("player_name", 2, "string", "")],
    "PlayerEliminatedEvent": [("player_id", 1, "int32", ""), ("eliminated_by", 2, "int32", "")],
    "GameStartedEvent": [("started_at", 1, TS, "")],
    "GameEndedEvent": [("winner_id", 1, "int32", ""), ("ended_at", 2, TS, "")],
    "PlayerDisconnectedEvent": [("player_id", 1, "int32", "")],
    "PlayerReconnectedEvent": [("player_id", 1, "int32", "")],
    "PhaseChangedEvent": [("previous_phase", 1, "enum:" + C + "GamePhase", ""), ("new_phase", 2, "enum:" + C + "GamePhase", ""),
                          ("reason", 3, "string", "")],
}

EXPERIENCE_MESSAGES = {
    "Experience": [("experience_id", 1, "string", ""), ("game_id", 2, "string", ""), ("player_id", 3, "int32", ""),
                   ("turn", 4, "int32", ""), ("state", 5, E + "TensorState", ""), ("action", 6, "int32", ""),
                   ("reward", 7, "float", ""), ("next_state", 8, E + "TensorState", ""), ("done", 9, "bool", ""),
                   ("action_mask", 10, "bool", "repeated"), ("collected_at", 11, TS, ""),
                   ("metadata", 12, "", "map:string,string")],
    "TensorState": [("shape", 1, "int32", "repeated"), ("data", 2, "float", "repeated")],
    "ExperienceBatch": [("experiences", 1, E + "Experience", "repeated"), ("batch_id", 2, "int32", ""),
                        ("stream_id", 3, "string", ""), ("created_at", 4, TS, ""), ("metadata", 5, "", "map:string,string")],
    "StreamExperiencesRequest": [("game_ids", 1, "string", "repeated"), ("player_ids", 2, "int32", "repeated"),
                                 ("min_turn", 3, "int64", ""), ("follow", 4, "bool", ""), ("batch_size", 5, "int32", ""),
                                 ("enable_compression", 6, "bool", ""), ("max_batch_wait_ms", 7, "int32", "")],
    "SubmitExperiencesRequest": [("experiences", 1, E + "Experience", "repeated")],
    "SubmitExperiencesResponse": [("accepted", 1, "int32", ""), ("rejected", 2, "int32", ""), ("errors", 3, "string", "repeated")],
    "GetExperienceStatsRequest": [("game_ids", 1, "string", "repeated")],
    "GetExperienceStatsResponse": [("total_experiences", 1, "int64", ""), ("total_games", 2, "int64", ""),
                                   ("experiences_per_game", 3, "", "map:string,int64"),
                                   ("experiences_per_player", 4, "", "map:int32,int64"), ("average_reward", 5, "float", ""),
                                   ("min_reward", 6, "float", ""), ("max_reward", 7, "float", ""),
                                   ("oldest_experience", 8, TS, ""), ("newest_experience", 9, TS, "")],
    "RewardConfig": [(n, i + 1, "float", "") for i, n in enumerate(
        ["territory_gained", "territory_lost", "army_gained", "army_lost", "enemy_general_captured", "own_general_lost",
         "win_game", "lose_game", "city_captured", "city_lost"])],
}

# service -> method -> (request, response, server_streaming)
SERVICES = {
    "generals.game.v1.GameService": {
        "CreateGame": ("CreateGameRequest", "CreateGameResponse", False),
        "JoinGame": ("JoinGameRequest", "JoinGameResponse", False),
        "SubmitAction": ("SubmitActionRequest", "SubmitActionResponse", False),
        "GetGameState": ("GetGameStateRequest", "GetGameStateResponse", False),
        "StreamGame": ("StreamGameRequest", "GameUpdate", True),
    },
    "generals.experience.v1.ExperienceService": {
        "StreamExperiences": ("StreamExperiencesRequest", "Experience", True),
        "StreamExperienceBatches": ("StreamExperiencesRequest", "ExperienceBatch", True),
        "SubmitExperiences": ("SubmitExperiencesRequest", "SubmitExperiencesResponse", False),
        "GetExperienceStats": ("GetExperienceStatsRequest", "GetExperienceStatsResponse", False),
    },
}


def _camel(name: str) -> str:
    return "".join(p.capitalize() for p in name.split("_"))


def _add_message(fd, pkg_path, name, fields):
    m = fd.message_type.add()
    m.name = name
    oneofs = {}
    for fname, num, ftype, label in fields:
        f = m.field.add()
        f.name, f.number = fname, num
        f.label = F.LABEL_REPEATED if label == "repeated" or label.startswith("map:") else F.LABEL_OPTIONAL
        if label.startswith("map:"):
            k, v = label[4:].split(",")
            entry = m.nested_type.add()
            entry.name = _camel(fname) + "Entry"
            entry.options.map_entry = True
            for en, enum_, et in (("key", 1, k), ("value", 2, v)):
                ef = entry.field.add()
                ef.name, ef.number, ef.label, ef.type = en, enum_, F.LABEL_OPTIONAL, _SCALAR[et]
            f.type, f.type_name = F.TYPE_MESSAGE, f"{pkg_path}{name}.{entry.name}"
        elif ftype in _SCALAR:
            f.type = _SCALAR[ftype]
        elif ftype.startswith("enum:"):
            f.type, f.type_name = F.TYPE_ENUM, ftype[5:]
        else:
            f.type, f.type_name = F.TYPE_MESSAGE, ftype
        if label.startswith("oneof:"):
            on = label[6:]
            if on not in oneofs:
                oneofs[on] = len(m.oneof_decl)
                m.oneof_decl.add().name = on
            f.oneof_index = oneofs[on]


def _build():
    pool = descriptor_pool.DescriptorPool()
    pool.AddSerializedFile(timestamp_pb2.DESCRIPTOR.serialized_pb)
    files = [
        ("common/v1/common.proto", "generals.common.v1", C, COMMON_MESSAGES, COMMON_ENUMS, []),
        ("game/v1/game.proto", "generals.game.v1", G, GAME_MESSAGES, {}, ["common/v1/common.proto", "google/protobuf/timestamp.proto"]),
        ("experience/v1/experience.proto", "generals.experience.v1", E, EXPERIENCE_MESSAGES, {},
         ["common/v1/common.proto", "google/protobuf/timestamp.proto"]),
    ]
    spaces = {}
    for fname, pkg, path, messages, enums, deps in files:
        fd = descriptor_pb2.FileDescriptorProto()
        fd.name, fd.package, fd.syntax = fname, pkg, "proto3"
        fd.dependency.extend(deps)
        for ename, values in enums.items():
            e = fd.enum_type.add()
            e.name = ename
            for i, v in enumerate(values):
                ev = e.value.add()
                ev.name, ev.number = v, i
        for mname, fields in messages.items():
            _add_message(fd, path, mname, fields)
        for sname, methods in SERVICES.items():
            if sname.rsplit(".", 1)[0] != pkg:
                continue
            sv = fd.service.add()
            sv.name = sname.rsplit(".", 1)[1]
            for mname, (req, resp, streaming) in methods.items():
                md = sv.method.add()
                md.name, md.input_type, md.output_type, md.server_streaming = mname, path + req, path + resp, streaming
        pool.AddSerializedFile(fd.SerializeToString())
        ns = SimpleNamespace()
        for mname in messages:
            setattr(ns, mname, message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{pkg}.{mname}")))
        for ename, values in enums.items():
            for i, v in enumerate(values):
                setattr(ns, v, i)
            setattr(ns, ename, SimpleNamespace(**{v: i for i, v in enumerate(values)}, names=list(values)))
        spaces[pkg] = ns
    return pool, spaces["generals.common.v1"], spaces["generals.game.v1"], spaces["generals.experience.v1"]


POOL, common, game, experience = _build()


def message_class(service: str, name: str):
    return getattr(game if "game" in service else experience, name)
