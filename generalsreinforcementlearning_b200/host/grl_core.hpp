// grl_core.hpp — C++ mirror of the reference's `internal/game/core` package (types only).
//
// The reference is Go; this image has no Go toolchain, so the compiled host layer that sits
// above the C ABI (include/grlcuda.h) is C++ and keeps the reference's names, argument
// meaning and error behaviour:
//
//   core::Tile, core::Board        internal/game/core/board.go:7-18, 20-26, 108-126
//   core::Coordinate, Direction    internal/game/core/coordinate.go:5-154
//   core::MoveAction + Validate    internal/game/core/action.go:23-35, 56-105
//   core::Err* sentinels, Wrap*, GameError   internal/game/core/errors.go:8-79
//
// Nothing here steps a game: the turn runs on the GPU (grl_engine.hpp).  Validate is the
// client-side pre-check the reference's server runs on a state copy before it queues an action
// (internal/grpc/gameserver/action_validator.go:126-127).
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace grl {
namespace core {

// ---- errors (errors.go) ------------------------------------------------------------------
// A Go sentinel is an identity; `code` is the number the C ABI's step_error plane and the proto
// ErrorCode enum use for it (proto/common/v1/common.proto:39-55), 0 when it has none.
struct Sentinel {
  const char *text;
  int code;
};

inline const Sentinel ErrInvalidCoordinates{"invalid coordinates", 1};
inline const Sentinel ErrNotAdjacent{"tiles are not adjacent", 2};
inline const Sentinel ErrNotOwned{"tile not owned by player", 3};
inline const Sentinel ErrInsufficientArmy{"insufficient army to move", 4};
inline const Sentinel ErrGameOver{"game is over", 5};
inline const Sentinel ErrInvalidPlayer{"invalid player ID", 0};
inline const Sentinel ErrMoveToSelf{"cannot move to the same tile", 7};
inline const Sentinel ErrTargetIsMountain{"target tile is a mountain", 8};

// Go `error`: nil when empty; a message chain that ends in a sentinel (fmt.Errorf("...: %w")).
class Error {
 public:
  Error() = default;
  Error(const Sentinel &s) : msg_(std::make_shared<std::string>(s.text)), root_(&s) {}
  static Error New(std::string text) {
    Error e;
    e.msg_ = std::make_shared<std::string>(std::move(text));
    return e;
  }
  // fmt.Errorf(prefix + ": %w", inner)
  static Error Wrap(const std::string &prefix, const Error &inner) {
    if (!inner) return Error();
    Error e;
    e.msg_ = std::make_shared<std::string>(prefix + ": " + inner.String());
    e.root_ = inner.root_;
    return e;
  }
  explicit operator bool() const { return msg_ != nullptr; }
  bool IsNil() const { return msg_ == nullptr; }
  std::string String() const { return msg_ ? *msg_ : std::string("<nil>"); }
  bool Is(const Sentinel &s) const { return root_ == &s; }  // errors.Is
  const Sentinel *Root() const { return root_; }

 private:
  std::shared_ptr<std::string> msg_;
  const Sentinel *root_ = nullptr;
};

// The sentinel a GRL_STEP_* / proto ErrorCode number stands for (nullptr: none).
inline const Sentinel *SentinelForCode(int code) {
  static const Sentinel *all[] = {&ErrInvalidCoordinates, &ErrNotAdjacent,  &ErrNotOwned,        &ErrInsufficientArmy,
                                  &ErrGameOver,           &ErrMoveToSelf,   &ErrTargetIsMountain};
  for (const Sentinel *s : all)
    if (s->code == code) return s;
  return nullptr;
}

// errors.go:35-41
inline Error WrapGameStateError(int turn, const std::string &phase, const Error &err) {
  return Error::Wrap("game turn " + std::to_string(turn) + " [" + phase + "]", err);
}
// errors.go:43-49
inline Error WrapPlayerError(int playerID, const std::string &operation, const Error &err) {
  return Error::Wrap("player " + std::to_string(playerID) + " " + operation, err);
}

// errors.go:51-79: a structured error with game context; Unwrap() is `Err`, so errors.Is reaches the sentinel
struct GameError {
  int Turn = 0, PlayerID = 0;
  std::string Operation;
  Error Err;
  std::string String() const {  // Error() in Go
    if (PlayerID != 0)
      return "turn " + std::to_string(Turn) + ": player " + std::to_string(PlayerID) + " " + Operation + ": " + Err.String();
    return "turn " + std::to_string(Turn) + ": " + Operation + ": " + Err.String();
  }
  const Error &Unwrap() const { return Err; }
  bool Is(const Sentinel &s) const { return Err.Is(s); }
};
inline GameError NewGameError(int turn, int playerID, const std::string &operation, const Error &err) {
  return GameError{turn, playerID, operation, err};
}

// ---- coordinate.go -----------------------------------------------------------------------
enum Direction : int { North = 0, East = 1, South = 2, West = 3 };  // coordinate.go:106-113

struct Coordinate {
  int X = 0, Y = 0;
  bool IsValid(int width, int height) const { return X >= 0 && X < width && Y >= 0 && Y < height; }
  int ToIndex(int width) const { return Y * width + X; }
  int DistanceTo(Coordinate o) const {
    int dx = X - o.X, dy = Y - o.Y;
    return (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
  }
  bool IsAdjacentTo(Coordinate o) const {
    int dx = X - o.X, dy = Y - o.Y;
    return (dx == 0 && (dy == 1 || dy == -1)) || (dy == 0 && (dx == 1 || dx == -1));
  }
  // coordinate.go:57-78: north, east, south, west — the order the gym client and the legal-move mask enumerate
  std::vector<Coordinate> Neighbors() const { return {{X, Y - 1}, {X + 1, Y}, {X, Y + 1}, {X - 1, Y}}; }
  std::vector<Coordinate> ValidNeighbors(int width, int height) const {
    std::vector<Coordinate> valid;
    valid.reserve(4);
    for (const Coordinate &n : Neighbors())
      if (n.IsValid(width, height)) valid.push_back(n);
    return valid;
  }
  Coordinate Add(Coordinate o) const { return {X + o.X, Y + o.Y}; }
  Coordinate Sub(Coordinate o) const { return {X - o.X, Y - o.Y}; }
  bool Equal(Coordinate o) const { return X == o.X && Y == o.Y; }
  std::string String() const { return "(" + std::to_string(X) + "," + std::to_string(Y) + ")"; }
  // coordinate.go:123-129: one step in a direction; a value that is not a Direction leaves the coordinate unchanged
  Coordinate Move(int direction) const;
  // coordinate.go:131-154: the direction of an adjacent coordinate, -1 when `o` is not adjacent
  int DirectionTo(Coordinate o) const {
    if (!IsAdjacentTo(o)) return -1;
    const int dx = o.X - X, dy = o.Y - Y;
    if (dy == -1) return North;
    if (dx == 1) return East;
    if (dy == 1) return South;
    if (dx == -1) return West;
    return -1;
  }
  bool operator==(Coordinate o) const { return X == o.X && Y == o.Y; }
  bool operator!=(Coordinate o) const { return !(*this == o); }
  bool operator<(Coordinate o) const { return Y != o.Y ? Y < o.Y : X < o.X; }  // Go structs are map keys; std::map needs an order
};
inline Coordinate NewCoordinate(int x, int y) { return Coordinate{x, y}; }
inline Coordinate FromIndex(int idx, int width) { return Coordinate{idx % width, idx / width}; }
// coordinate.go:115-121 (a map in Go; indexed by Direction here)
inline const Coordinate DirectionVectors[4] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
inline Coordinate Coordinate::Move(int direction) const {
  return (direction >= North && direction <= West) ? Add(DirectionVectors[direction]) : *this;
}

// ---- board.go ----------------------------------------------------------------------------
enum : int { TileNormal = 0, TileGeneral = 1, TileCity = 2, TileMountain = 3 };
constexpr int NeutralID = -1;

struct Tile {
  int Owner = NeutralID;
  int Army = 0;
  int Type = TileNormal;
  uint32_t VisibleBitfield = 0;  // bit i: player i sees this tile

  bool IsNeutral() const { return Owner == NeutralID; }
  bool IsCity() const { return Type == TileCity; }
  bool IsGeneral() const { return Type == TileGeneral; }
  bool IsMountain() const { return Type == TileMountain; }
  bool IsEmpty() const { return IsNeutral() && Type == TileNormal && Army == 0; }
  bool IsVisibleTo(int playerID) const {
    if (playerID < 0 || playerID >= 32) return false;
    return (VisibleBitfield & (1u << unsigned(playerID))) != 0;
  }
};

struct Board {
  int W = 0, H = 0;
  std::vector<Tile> T;  // row-major, length W*H

  int Idx(int x, int y) const { return y * W + x; }
  std::pair<int, int> XY(int idx) const { return {idx % W, idx / W}; }
  bool InBounds(int x, int y) const { return x >= 0 && x < W && y >= 0 && y < H; }
  Tile *GetTile(int x, int y) { return InBounds(x, y) ? &T[Idx(x, y)] : nullptr; }
  const Tile *GetTile(int x, int y) const { return InBounds(x, y) ? &T[Idx(x, y)] : nullptr; }
  int Distance(int x1, int y1, int x2, int y2) const { return Coordinate{x1, y1}.DistanceTo({x2, y2}); }
  // board.go:136-159, the coordinate-based forms
  bool InBoundsCoord(Coordinate c) const { return InBounds(c.X, c.Y); }
  Tile *GetTileCoord(Coordinate c) { return GetTile(c.X, c.Y); }
  const Tile *GetTileCoord(Coordinate c) const { return GetTile(c.X, c.Y); }
  void SetTile(Coordinate c, const Tile &tile) {
    if (InBoundsCoord(c)) T[Idx(c.X, c.Y)] = tile;
  }
  int IdxCoord(Coordinate c) const { return Idx(c.X, c.Y); }
  std::shared_ptr<Board> Clone() const { return std::make_shared<Board>(*this); }
};
inline std::shared_ptr<Board> NewBoard(int w, int h) {
  auto b = std::make_shared<Board>();
  b->W = w;
  b->H = h;
  b->T.assign(size_t(w) * h, Tile{});
  return b;
}

// ---- action.go ---------------------------------------------------------------------------
struct MoveAction {
  int PlayerID = 0;
  int FromX = 0, FromY = 0, ToX = 0, ToY = 0;
  bool MoveAll = false;  // true: leave one behind; false: move half (min 1)
  // action.go:30-32: the coordinate fields.  As in the reference they only feed GetFrom/GetTo (and through them the
  // adjacency test of Validate); bounds, ownership, armies and the move itself read FromX/FromY/ToX/ToY, and so does the
  // device.  Declared after MoveAll so that the six-value aggregate form used throughout keeps its meaning.
  Coordinate From{}, To{};

  int GetPlayerID() const { return PlayerID; }
  // action.go:40-54: the coordinate field when it is set (non-zero), the legacy pair otherwise
  Coordinate GetFrom() const { return From != Coordinate{} ? From : Coordinate{FromX, FromY}; }
  Coordinate GetTo() const { return To != Coordinate{} ? To : Coordinate{ToX, ToY}; }

  // action.go:56-105, checks in the reference's order.
  Error Validate(const Board &b, int playerID) const {
    const std::string who = "player " + std::to_string(PlayerID);
    auto at = [](int x, int y) { return "(" + std::to_string(x) + "," + std::to_string(y) + ")"; };
    if (!b.InBounds(FromX, FromY))
      return Error::Wrap(who + ": move from " + at(FromX, FromY) + " out of bounds", ErrInvalidCoordinates);
    if (!b.InBounds(ToX, ToY))
      return Error::Wrap(who + ": move to " + at(ToX, ToY) + " out of bounds", ErrInvalidCoordinates);
    if (FromX == ToX && FromY == ToY)
      return Error::Wrap(who + ": move from/to same tile " + at(FromX, FromY), ErrMoveToSelf);
    if (!GetFrom().IsAdjacentTo(GetTo()))
      return Error::Wrap(who + ": move from " + at(FromX, FromY) + " to " + at(ToX, ToY) + " not adjacent",
                         ErrNotAdjacent);
    const Tile &from = b.T[b.Idx(FromX, FromY)];
    if (from.Owner != playerID)
      return Error::Wrap(who + ": tile at " + at(FromX, FromY) + " owned by player " + std::to_string(from.Owner),
                         ErrNotOwned);
    if (from.Army <= 1)
      return Error::Wrap(who + ": tile at " + at(FromX, FromY) + " has only " + std::to_string(from.Army) + " army",
                         ErrInsufficientArmy);
    if (b.T[b.Idx(ToX, ToY)].IsMountain())
      return Error::Wrap(who + ": cannot move to mountain at " + at(ToX, ToY), ErrTargetIsMountain);
    return Error();
  }
};
// The reference's `[]core.Action` holds only *MoveAction today (action.go:9-14).
using Action = MoveAction;

// errors.go:20-33
inline Error WrapActionError(const MoveAction &a, const Error &err) {
  auto at = [](int x, int y) { return "(" + std::to_string(x) + "," + std::to_string(y) + ")"; };
  return Error::Wrap("player " + std::to_string(a.PlayerID) + ": move from " + at(a.FromX, a.FromY) + " to " +
                         at(a.ToX, a.ToY),
                     err);
}

// utils.go:10-12: fmt.Sprintf("%*d", width, num) — left-padded with spaces, never truncated
inline std::string IntToStringFixedWidth(int num, int width) {
  std::string d = std::to_string(num);
  const size_t w = width < 0 ? size_t(-(long long)width) : size_t(width);
  if (d.size() >= w) return d;
  return width < 0 ? d + std::string(w - d.size(), ' ') : std::string(w - d.size(), ' ') + d;  // a negative width left-justifies
}
// utils.go:14-19: fmt.Sprintf("%T", action); the only Action is *MoveAction
inline std::string GetActionType(const MoveAction *action) { return action ? "*core.MoveAction" : "nil"; }

// errors.go:20-33 with an action of unknown type (Go: a nil Action interface): the generic prefix
inline Error WrapActionError(const MoveAction *a, const Error &err) {
  return a ? WrapActionError(*a, err) : Error::Wrap("player action", err);
}

}  // namespace core

namespace errors {
inline bool Is(const core::Error &err, const core::Sentinel &target) { return err.Is(target); }
}  // namespace errors
}  // namespace grl
