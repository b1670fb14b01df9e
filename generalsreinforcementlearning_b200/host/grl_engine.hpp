// grl_engine.hpp — C++ mirror of the reference's `internal/game` package surface over libgrlcuda.so.
//
// The reference's seam for the turn path is the public method set of *game.Engine
// (internal/game/engine.go:45-71, 75, 197-198, 248, 271-300; visibility.go:153; rendering.go:34).
// This header presents the same names with the same argument meaning and error behaviour, in C++
// because the reference is compiled code and its own toolchain (Go) is absent from this image:
//
//   game::GameConfig / NewEngine          engine.go:45-71, engine_initializer.go:34-87
//   game::Engine::Step                    engine.go:75 (TurnProcessor.ProcessTurn, turn_processor.go:29-77)
//   game::Engine::GameState / IsGameOver / GetWinner / GetLegalActionMask /
//         ComputePlayerVisibility / GetChangedTiles / GetVisibilityChangedTiles / Board
//   game::Player, game::GameState         state.go:7-100
//   game::ExperienceCollector             experience_collector.go:4-10, turn_processor.go:182-217
//   experience::Transition                what SimpleCollector.OnStateTransition derives per player
//                                         (internal/experience/collector.go:30-98), read from the device
//
// Every game lives in HBM as one env slot of a batched grl_env; an EnginePool owns the slots of one
// device and steps all of them with ONE fused launch (StepAll).  Engine is the per-game view the
// reference's callers hold (gameInstance, internal/grpc/gameserver/game_manager.go:510,602).
// There is no CPU engine behind this: without libgrlcuda.so construction throws.
#pragma once

#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "grl_core.hpp"

struct grl_env;  // include/grlcuda.h
struct grl_action;
struct grl_step_outputs;

namespace grl {

// ---- context.Context (what Step reads of it: turn_processor.go:31-36, 79-93) -----------------
namespace context {
inline const core::Sentinel Canceled{"context canceled", 0};
class Context {
 public:
  core::Error Err() const { return flag_ && flag_->load() ? core::Error(Canceled) : core::Error(); }
  void Cancel() const {
    if (flag_) flag_->store(true);
  }
  static Context WithCancel() {
    Context c;
    c.flag_ = std::make_shared<std::atomic<bool>>(false);
    return c;
  }

 private:
  std::shared_ptr<std::atomic<bool>> flag_;
};
inline Context Background() { return Context(); }
}  // namespace context

// ---- math/rand (go1.24, the v1 value stream): rand.New(rand.NewSource(seed)) -----------------------------------
//      NewEngine hands a FRESH generator's seed to the device map generator (engine_initializer.go:90-94 builds the
//      generator from cfg.Rng; a fresh one is fully described by its seed); GenerateRandomActions draws from it on the
//      host exactly as demo_helpers.go does.
namespace rand {
struct Source {
  int64_t seed;
};
class Rand {  // rng.go (*rngSource) + rand.go
 public:
  explicit Rand(int64_t seed) { Seed(seed); }
  void Seed(int64_t seed);
  int64_t Int63();
  uint64_t Uint64();
  int32_t Int31() { return int32_t(Int63() >> 32); }
  int32_t Int31n(int32_t n);
  int Intn(int n);
  double Float64();
  float Float32();
  int64_t seed;            // what it was seeded with
  uint64_t draws = 0;      // values drawn since (NewEngine needs draws == 0)

 private:
  uint64_t vec_[607];
  int tap_ = 0, feed_ = 0;
};
inline Source NewSource(int64_t seed) { return Source{seed}; }
inline std::shared_ptr<Rand> New(Source s) { return std::make_shared<Rand>(s.seed); }
}  // namespace rand

// ---- the bound C ABI ---------------------------------------------------------------------------
// A table of the include/grlcuda.h entry points resolved with dlsym.  Default() binds
// libgrlcuda.so (env GRLCUDA_LIB, else next to libgrlhost.so, else the loader path) and throws when it
// is missing.  Open(path, prefix) exists for the tests, which bind the CPU oracle's copy of the same
// ABI (prefix "grlo_") where no GPU is present.
class Library {
 public:
  static std::shared_ptr<Library> Default();
  static std::shared_ptr<Library> Open(const std::string &path, const std::string &prefix);
  ~Library();
  struct Fns;  // function pointers, defined in grl_engine.cpp
  const Fns &fn() const { return *fns_; }
  const std::string &path() const { return path_; }

 private:
  Library() = default;
  void *handle_ = nullptr;
  std::unique_ptr<Fns> fns_;
  std::string path_;
};

namespace game {

// state.go:7-24
struct Player {
  int ID = 0;
  bool Alive = false;
  int ArmyCount = 0;
  int GeneralIdx = -1;
  std::vector<int> OwnedTiles;  // ascending tile order (the reference's order is unspecified)
  int GetID() const { return ID; }
  bool IsAlive() const { return Alive; }
};

// state.go:26-100.  Board is shared between copies exactly as in Go (GameState() is a shallow copy).
struct GameState {
  int Turn = 0;
  std::shared_ptr<core::Board> Board;
  std::vector<Player> Players;
  std::map<int, bool> ChangedTiles;
  bool FogOfWarEnabled = true;
  std::map<int, bool> VisibilityChangedTiles;

  std::shared_ptr<GameState> Clone() const;
  bool IsGameOver() const;
  int GetWinner() const;
};

// visibility.go:146-151
struct PlayerVisibility {
  std::vector<bool> VisibleTiles, FogTiles;
};

// internal/game/action.go:5-18
enum ActionType { ActionTypeMove = 0, ActionTypeNoOp = 1 };
struct Action {
  ActionType Type = ActionTypeMove;
  core::Coordinate From, To;
};

// experience_collector.go:4-10
class ExperienceCollector {
 public:
  virtual ~ExperienceCollector() = default;
  virtual void OnStateTransition(const GameState *prevState, const GameState *currState,
                                 const std::map<int, Action> &actions) = 0;
  virtual void OnGameEnd(const GameState *finalState) = 0;
};

// engine.go:45-59 (Logger/EventBus/StateMachine belong to the control plane and are not mirrored)
struct GameConfig {
  int Width = 0, Height = 0, Players = 0;
  std::shared_ptr<rand::Rand> Rng;  // nil: seeded from the clock, as engine_initializer.go:91-94
  std::string GameID;
  ::grl::game::ExperienceCollector *ExperienceCollector = nullptr;
};

class Engine;
std::unique_ptr<Engine> NewEngine(const context::Context &ctx, const GameConfig &cfg, std::shared_ptr<Library> lib,
                                  int device);

// One CUDA device's worth of game slots: a grl_env of `numEnvs` games of one shape.
class EnginePool {
 public:
  EnginePool(std::shared_ptr<Library> lib, int numEnvs, int width, int height, int players, int device = 0,
             int maxActionsPerPlayer = 1);
  ~EnginePool();
  EnginePool(const EnginePool &) = delete;
  EnginePool &operator=(const EnginePool &) = delete;

  // NewEngine in a free slot (map from cfg.Rng's seed, turn-0 stats + fog; engine_initializer.go:34-87).
  // Returns nullptr when the pool is full or cfg's shape is not the pool's, as NewEngine returns nil on
  // an initialisation error (engine.go:62-71).
  std::unique_ptr<Engine> NewEngine(const context::Context &ctx, const GameConfig &cfg);
  // A game built from a caller's board (the reference's tests assemble engines by hand,
  // action_mask_test.go:16-55): players start alive with empty lists; nothing is recomputed.
  std::unique_ptr<Engine> NewEngineFromBoard(const core::Board &board, int players);

  // Engine.Step for every slot in one fused launch.  perSlot[i] are slot i's actions this turn; slots
  // absent from the map take NO turn (the per-game turn barrier of game_manager.go:554-600).
  // Returns one error per stepped slot (nil entries included).
  std::map<int, core::Error> StepAll(const context::Context &ctx,
                                     const std::map<int, std::vector<core::Action>> &perSlot);

  // The bulk form of the same call for callers that own whole batches (a self-play driver, a trainer): `actions` is
  // the C ABI's [numEnvs][max_actions] array (host or device memory; nullptr with `randomPolicy` = every alive player
  // plays the synthetic uniformly random legal move), `out` the C ABI's output planes (any member may be null).  One
  // fused launch; per-slot views are refreshed lazily afterwards.
  void StepBatch(const ::grl_action *actions, const ::grl_step_outputs *out, bool randomPolicy = false,
                 uint64_t policySeed = 0);

  int NumEnvs() const { return B_; }
  int Width() const { return W_; }
  int Height() const { return H_; }
  int NumPlayers() const { return P_; }
  uint64_t LaunchCount() const;  // CUDA kernels launched by this pool (0 on the oracle binding)
  const Library &lib() const { return *lib_; }
  grl_env *handle() const { return env_; }

 private:
  friend class Engine;
  int Acquire();
  void Release(int slot);
  void Check(int status, const char *what) const;

  std::shared_ptr<Library> lib_;
  grl_env *env_ = nullptr;
  int B_, W_, H_, P_, A_;
  std::vector<int> free_;
  std::vector<Engine *> engines_;  // slot -> live view (nullptr: free)
  // per-pool result planes of the last launch
  std::vector<uint8_t> done_, stepErr_;
  std::vector<int8_t> winner_;
  std::vector<float> reward_;
  std::vector<int32_t> actionIndex_;
};

}  // namespace game

namespace experience {
// What SimpleCollector.OnStateTransition computes for one acting player (collector.go:30-98), taken
// from the device's read-outs instead of six host-side board scans.
struct Transition {
  int PlayerID = 0;
  int Turn = 0;                    // currState.Turn
  std::vector<float> State;        // StateToTensor(prevState, p), [9][H][W]  (serializer.go:37-109)
  int Action = -1;                 // ActionToIndex, U,D,L,R order            (serializer.go:179-198)
  float Reward = 0;                // CalculateReward(prev, curr, p)          (rewards.go:45-175)
  std::vector<float> NextState;    // StateToTensor(currState, p)
  bool Done = false;               // currState.IsGameOver()
  std::vector<bool> ActionMask;    // GenerateActionMask(prevState, p)        (serializer.go:112-176)
};
}  // namespace experience

namespace game {

class Engine {
 public:
  ~Engine();
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;

  // engine.go:75.  Error kinds, in the reference's order (turn_processor.go:29-77, 95-113): context
  // cancelled; ErrGameOver when the game has ended; the first action-validation error of the turn —
  // returned after the turn's valid moves were applied and before production.
  core::Error Step(const context::Context &ctx, const std::vector<core::Action> &actions);

  game::GameState GameState();  // engine.go:197 — shallow copy; Board is shared with the engine
  bool IsGameOver();            // engine.go:198
  int GetWinner();              // engine.go:248-263
  std::vector<bool> GetLegalActionMask(int playerID);           // engine.go:271-280, dirs U,R,D,L
  PlayerVisibility ComputePlayerVisibility(int playerID);        // visibility.go:153-190
  std::map<int, bool> GetChangedTiles();                         // engine.go:283-289
  std::map<int, bool> GetVisibilityChangedTiles();               // engine.go:292-298
  std::string Board(int playerID);                               // rendering.go:34-143
  ExperienceCollector *GetExperienceCollector() const { return collector_; }

  // The reference's own tests reach into `engine.gs` and `engine.gameOver` (same package).  gs() is that
  // pointer: the host copy of this slot's state, refreshed from HBM when stale.  After editing it, call
  // Upload() — the point where a Go test would go on using the edited struct.
  game::GameState *gs();
  void Upload();
  void SetGameOver(bool over);

  // Read-outs the experience path takes per player, straight from the device.
  std::vector<float> StateTensor(int playerID);                  // Serializer.StateToTensor
  std::vector<bool> SerializerActionMask(int playerID);          // Serializer.GenerateActionMask
  float LastReward(int playerID) const;                          // reward of the last Step
  // The last Step's transition for `playerID` (valid when a collector is attached).
  const experience::Transition *LastTransition(int playerID) const;

  int Slot() const { return slot_; }
  EnginePool &Pool() { return *pool_; }

 private:
  friend class EnginePool;
  friend std::unique_ptr<Engine> NewEngine(const context::Context &, const GameConfig &, std::shared_ptr<Library>, int);
  Engine(EnginePool *pool, int slot, std::shared_ptr<EnginePool> owned);
  void Refresh();

  EnginePool *pool_;
  std::shared_ptr<EnginePool> owned_;  // standalone NewEngine owns a one-slot pool
  int slot_;
  bool stale_ = true;
  bool gameOver_ = false;
  int winner_ = -1;
  game::GameState gs_;
  ExperienceCollector *collector_ = nullptr;
  std::vector<experience::Transition> last_;
};

// demo_helpers.go:12-62: per alive player, with probability 0.3, one uniformly chosen legal move (every (tile, direction)
// pair in row-major order, directions down, up, right, left, each with its own MoveAll draw) — the reference's demo
// policy, drawing from Go's generator so that a seed replays the reference's choices.
std::vector<core::Action> GenerateRandomActions(Engine &g, rand::Rand &rng);

// engine.go:62-71: a private one-game engine (its own one-slot pool on `device`).  Batched callers use
// EnginePool::NewEngine instead.
std::unique_ptr<Engine> NewEngine(const context::Context &ctx, const GameConfig &cfg,
                                  std::shared_ptr<Library> lib = nullptr, int device = 0);

}  // namespace game
}  // namespace grl
