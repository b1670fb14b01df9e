// grl_engine.cpp — the host mirror of `internal/game` over the C ABI (see grl_engine.hpp).
//
// Only plumbing lives here: action packing, plane slicing, error synthesis, the ASCII renderer.  Every
// rule of the game (moves, captures, eliminations, production, stats, fog, masks, tensors, rewards) is
// evaluated by the library this file binds with dlsym.
#include "grl_engine.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "../../include/grlcuda.h"

namespace grl {

// ---- Library -------------------------------------------------------------------------------
struct Library::Fns {
  int (*abi_version)(void);
  const char *(*status_string)(int);
  const char *(*last_error)(void);
  int (*default_config)(grl_config *);
  int (*create)(const grl_config *, grl_env **);
  int (*destroy)(grl_env *);
  int (*sync)(grl_env *);
  int (*reset_seeded)(grl_env *, const int32_t *, int32_t, const int64_t *);
  int (*reset_boards)(grl_env *, const int32_t *, int32_t, const int32_t *, const int32_t *, const int32_t *);
  int (*step_fused)(grl_env *, const grl_action *, uint32_t, uint64_t, const grl_step_outputs *);
  int (*observe)(grl_env *, const grl_step_outputs *);
  int (*mask)(grl_env *, int, void *);
  int (*visibility)(grl_env *, uint8_t *, uint8_t *);
  int (*get_state)(grl_env *, int32_t, int32_t, const grl_state_planes *);
  int (*set_state)(grl_env *, int32_t, int32_t, const grl_state_planes *);
  int (*launch_count)(grl_env *, uint64_t *);
};

Library::~Library() {
  if (handle_) dlclose(handle_);
}

std::shared_ptr<Library> Library::Open(const std::string &path, const std::string &prefix) {
  void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) throw std::runtime_error("grl: cannot load " + path + ": " + dlerror());
  std::shared_ptr<Library> lib(new Library());
  lib->handle_ = h;
  lib->path_ = path;
  lib->fns_.reset(new Fns());
  auto sym = [&](const char *name) -> void * {
    void *p = dlsym(h, (prefix + name).c_str());
    if (!p) throw std::runtime_error("grl: " + path + " does not export " + prefix + name);
    return p;
  };
  Fns &f = *lib->fns_;
#define GRL_BIND(field, name) f.field = reinterpret_cast<decltype(f.field)>(sym(name))
  GRL_BIND(abi_version, "abi_version");
  GRL_BIND(status_string, "status_string");
  GRL_BIND(last_error, "last_error");
  GRL_BIND(default_config, "default_config");
  GRL_BIND(create, "create");
  GRL_BIND(destroy, "destroy");
  GRL_BIND(sync, "sync");
  GRL_BIND(reset_seeded, "reset_seeded");
  GRL_BIND(reset_boards, "reset_boards");
  GRL_BIND(step_fused, "step_fused");
  GRL_BIND(observe, "observe");
  GRL_BIND(mask, "mask");
  GRL_BIND(visibility, "visibility");
  GRL_BIND(get_state, "get_state");
  GRL_BIND(set_state, "set_state");
  GRL_BIND(launch_count, "launch_count");
#undef GRL_BIND
  if (f.abi_version() != GRL_ABI_VERSION)
    throw std::runtime_error("grl: " + path + " has ABI version " + std::to_string(f.abi_version()) + ", header has " +
                             std::to_string(GRL_ABI_VERSION));
  return lib;
}

std::shared_ptr<Library> Library::Default() {
  static std::weak_ptr<Library> cached;
  if (auto l = cached.lock()) return l;
  std::vector<std::string> candidates;
  if (const char *env = std::getenv("GRLCUDA_LIB")) candidates.push_back(env);
  Dl_info info;
  if (dladdr(reinterpret_cast<void *>(&Library::Default), &info) && info.dli_fname) {
    std::string self(info.dli_fname);
    size_t slash = self.rfind('/');
    std::string dir = slash == std::string::npos ? "." : self.substr(0, slash);
    candidates.push_back(dir + "/libgrlcuda.so");
    candidates.push_back(dir + "/../csrc/libgrlcuda.so");
  }
  candidates.push_back("libgrlcuda.so");
  std::string why;
  for (const std::string &c : candidates) {
    try {
      auto l = Open(c, "grl_");
      cached = l;
      return l;
    } catch (const std::exception &e) {
      why += std::string("\n  ") + e.what();
    }
  }
  // No CPU engine stands behind this layer: fail loudly.
  throw std::runtime_error("grl: libgrlcuda.so not found (set GRLCUDA_LIB); tried:" + why);
}

// ---- math/rand ------------------------------------------------------------------------------
namespace rand {
namespace {
constexpr int kLen = 607, kTap = 273;
constexpr int64_t kInt32Max = 2147483647;
const int64_t kCooked[kLen] = {
#include "../csrc/go_rng_cooked.inc"
};
int32_t seedrand(int32_t x) {  // rng.go: x = 48271 * x mod (2^31 - 1), Schrage
  const int32_t hi = x / 44488, lo = x % 44488;
  x = 48271 * lo - 3399 * hi;
  return x < 0 ? x + int32_t(kInt32Max) : x;
}
}  // namespace

void Rand::Seed(int64_t s) {
  seed = s;
  draws = 0;
  tap_ = 0;
  feed_ = kLen - kTap;
  s %= kInt32Max;
  if (s < 0) s += kInt32Max;
  if (s == 0) s = 89482311;
  int32_t x = int32_t(s);
  for (int i = -20; i < kLen; i++) {
    x = seedrand(x);
    if (i < 0) continue;
    uint64_t u = uint64_t(x) << 40;
    x = seedrand(x);
    u ^= uint64_t(x) << 20;
    x = seedrand(x);
    u ^= uint64_t(x);
    vec_[i] = u ^ uint64_t(kCooked[i]);
  }
}
uint64_t Rand::Uint64() {
  draws++;
  if (--tap_ < 0) tap_ += kLen;
  if (--feed_ < 0) feed_ += kLen;
  const uint64_t x = vec_[feed_] + vec_[tap_];
  vec_[feed_] = x;
  return x;
}
int64_t Rand::Int63() { return int64_t(Uint64() & 0x7fffffffffffffffULL); }
int32_t Rand::Int31n(int32_t n) {  // rand.go Int31n
  if ((n & (n - 1)) == 0) return Int31() & (n - 1);
  const int32_t mx = int32_t((1u << 31) - 1 - (1u << 31) % uint32_t(n));
  int32_t v = Int31();
  while (v > mx) v = Int31();
  return v % n;
}
int Rand::Intn(int n) {
  if (n <= 0) throw std::invalid_argument("invalid argument to Intn");
  return Int31n(int32_t(n));
}
double Rand::Float64() {  // Go 1 value stream: float64(Int63()) / (1 << 63), re-drawn when it rounds to 1
  for (;;) {
    const double f = double(Int63()) / 9223372036854775808.0;
    if (f != 1.0) return f;
  }
}
float Rand::Float32() {
  for (;;) {
    const float f = float(Float64());
    if (f != 1.0f) return f;
  }
}
}  // namespace rand

namespace game {

std::vector<core::Action> GenerateRandomActions(Engine &g, rand::Rand &rng) {
  std::vector<core::Action> actions;
  const GameState state = g.GameState();
  const core::Board &board = *state.Board;
  static const int dirs[4][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}};
  for (const Player &player : state.Players) {
    if (!player.Alive) continue;
    if (rng.Float32() > 0.3f) continue;
    std::vector<core::MoveAction> validMoves;
    for (int y = 0; y < board.H; y++) {
      for (int x = 0; x < board.W; x++) {
        const core::Tile &tile = board.T[board.Idx(x, y)];
        if (tile.Owner != player.ID || tile.Army <= 1) continue;
        for (const auto &dir : dirs) {
          const int toX = x + dir[0], toY = y + dir[1];
          if (toX < 0 || toX >= board.W || toY < 0 || toY >= board.H) continue;
          if (board.T[board.Idx(toX, toY)].IsMountain()) continue;
          core::MoveAction move;
          move.PlayerID = player.ID;
          move.FromX = x, move.FromY = y, move.ToX = toX, move.ToY = toY;
          move.MoveAll = rng.Float32() < 0.7f;
          validMoves.push_back(move);
        }
      }
    }
    if (!validMoves.empty()) actions.push_back(validMoves[rng.Intn(int(validMoves.size()))]);
  }
  return actions;
}

// ---- GameState (state.go) ------------------------------------------------------------------
std::shared_ptr<GameState> GameState::Clone() const {
  auto c = std::make_shared<GameState>(*this);
  c->Board = Board ? Board->Clone() : nullptr;
  return c;
}
bool GameState::IsGameOver() const {
  int alive = 0;
  for (const Player &p : Players) alive += p.Alive ? 1 : 0;
  return alive <= 1;
}
int GameState::GetWinner() const {
  int alive = 0, who = -1;
  for (const Player &p : Players)
    if (p.Alive) alive++, who = p.ID;
  return alive == 1 ? who : -1;
}

// ---- EnginePool ----------------------------------------------------------------------------
void EnginePool::Check(int status, const char *what) const {
  if (status == GRL_OK) return;
  const auto &f = lib_->fn();
  throw std::runtime_error(std::string("grl: ") + what + ": " + f.status_string(status) + ": " + f.last_error());
}

EnginePool::EnginePool(std::shared_ptr<Library> lib, int numEnvs, int width, int height, int players, int device,
                       int maxActionsPerPlayer)
    : lib_(lib ? std::move(lib) : Library::Default()), B_(numEnvs), W_(width), H_(height), P_(players) {
  const auto &f = lib_->fn();
  grl_config cfg;
  Check(f.default_config(&cfg), "grl_default_config");  // config.go:198-209, rewards.go:23-37
  cfg.num_envs = numEnvs;
  cfg.width = width;
  cfg.height = height;
  cfg.num_players = players;
  cfg.device = device;
  A_ = std::max(1, std::min(GRL_MAX_ACTIONS, players * std::max(1, maxActionsPerPlayer)));
  cfg.max_actions = A_;
  Check(f.create(&cfg, &env_), "grl_create");
  for (int i = numEnvs - 1; i >= 0; i--) free_.push_back(i);
  engines_.assign(numEnvs, nullptr);
  done_.assign(numEnvs, 0);
  stepErr_.assign(numEnvs, 0);
  winner_.assign(numEnvs, -1);
  reward_.assign(size_t(numEnvs) * players, 0.f);
  actionIndex_.assign(size_t(numEnvs) * players, -1);
}

EnginePool::~EnginePool() {
  for (Engine *e : engines_)
    if (e) e->pool_ = nullptr;
  if (env_) lib_->fn().destroy(env_);
}

uint64_t EnginePool::LaunchCount() const {
  uint64_t n = 0;
  Check(lib_->fn().launch_count(env_, &n), "grl_launch_count");
  return n;
}

int EnginePool::Acquire() {
  if (free_.empty()) return -1;
  int s = free_.back();
  free_.pop_back();
  return s;
}
void EnginePool::Release(int slot) {
  engines_[slot] = nullptr;
  free_.push_back(slot);
}

std::unique_ptr<Engine> EnginePool::NewEngine(const context::Context &ctx, const GameConfig &cfg) {
  (void)ctx;
  if (cfg.Width != W_ || cfg.Height != H_ || cfg.Players != P_) return nullptr;
  if (cfg.Rng && cfg.Rng->draws != 0)  // the device generator is seeded, not handed a state: only a fresh generator maps onto it
    throw std::invalid_argument("grl: NewEngine needs a fresh rand.Rand (values were already drawn from cfg.Rng)");
  int slot = Acquire();
  if (slot < 0) return nullptr;
  int64_t seed = cfg.Rng ? cfg.Rng->seed
                         : std::chrono::duration_cast<std::chrono::nanoseconds>(
                               std::chrono::system_clock::now().time_since_epoch())
                               .count();  // engine_initializer.go:91-94
  int32_t id = slot;
  int st = lib_->fn().reset_seeded(env_, &id, 1, &seed);
  if (st != GRL_OK) {  // engine.go:65-69: initialisation error -> nil engine
    free_.push_back(slot);
    return nullptr;
  }
  std::unique_ptr<Engine> e(new Engine(this, slot, nullptr));
  e->collector_ = cfg.ExperienceCollector;
  engines_[slot] = e.get();
  done_[slot] = 0;
  winner_[slot] = -1;
  e->Refresh();
  // engine_initializer.go:218-225 runs checkGameOver at turn 0 (a 1-player game is never over,
  // win_conditions.go:27-31)
  return e;
}

std::unique_ptr<Engine> EnginePool::NewEngineFromBoard(const core::Board &board, int players) {
  if (board.W != W_ || board.H != H_ || players != P_) return nullptr;
  int slot = Acquire();
  if (slot < 0) return nullptr;
  std::unique_ptr<Engine> e(new Engine(this, slot, nullptr));
  engines_[slot] = e.get();
  e->gs_.Turn = 0;
  e->gs_.Board = board.Clone();
  e->gs_.Players.assign(players, Player{});
  for (int p = 0; p < players; p++) {
    e->gs_.Players[p].ID = p;
    e->gs_.Players[p].Alive = true;
  }
  e->gs_.FogOfWarEnabled = true;
  e->gameOver_ = false;
  e->stale_ = false;
  e->Upload();
  return e;
}

std::map<int, core::Error> EnginePool::StepAll(const context::Context &ctx,
                                               const std::map<int, std::vector<core::Action>> &perSlot) {
  std::map<int, core::Error> errs;
  if (core::Error c = ctx.Err()) {  // turn_processor.go:31-36
    for (const auto &kv : perSlot) errs[kv.first] = c;
    return errs;
  }
  const auto &f = lib_->fn();
  const int N = W_ * H_;
  std::vector<grl_action> acts(size_t(B_) * A_);
  std::memset(acts.data(), 0, acts.size() * sizeof(grl_action));
  for (int b = 0; b < B_; b++) acts[size_t(b) * A_].flags = GRL_ACTION_FLAG_SKIP_ENV;
  bool collect = false;
  for (const auto &kv : perSlot) {
    const int slot = kv.first;
    if (slot < 0 || slot >= B_) throw std::out_of_range("grl: StepAll: slot out of range");
    grl_action *row = &acts[size_t(slot) * A_];
    row[0].flags = 0;
    if (kv.second.size() > size_t(A_))
      throw std::length_error("grl: StepAll: more actions than the pool's max_actions (GRL_MAX_ACTIONS caps it at 12)");
    int k = 0;
    for (const core::Action &m : kv.second) {
      grl_action &a = row[k++];
      // int8 fields: anything outside the board is out of bounds either way (action.go:58-64)
      auto clamp8 = [](int v) { return int8_t(v < -128 ? -128 : (v > 127 ? 127 : v)); };
      a.player_id = clamp8(m.PlayerID);
      a.from_x = clamp8(m.FromX);
      a.from_y = clamp8(m.FromY);
      a.to_x = clamp8(m.ToX);
      a.to_y = clamp8(m.ToY);
      a.move_all = m.MoveAll ? 1 : 0;
      a.present = 1;
    }
    if (engines_[slot] && engines_[slot]->collector_) collect = true;
  }

  // Experience collection needs s and mask(s) of the acting players before the turn
  // (turn_processor.go:124-129 clones the state; collector.go:44-55).
  std::vector<float> prevObs, nextObs;
  std::vector<uint8_t> prevMask;
  std::map<int, std::shared_ptr<game::GameState>> prevStates;
  if (collect) {
    prevObs.resize(size_t(B_) * P_ * GRL_OBS_CHANNELS * N);
    prevMask.resize(size_t(B_) * P_ * N * 4);
    nextObs.resize(prevObs.size());
    grl_step_outputs ro{};
    ro.obs = prevObs.data();
    Check(f.observe(env_, &ro), "grl_observe");
    Check(f.mask(env_, GRL_MASK_SERIALIZER_UDLR, prevMask.data()), "grl_mask");
    for (const auto &kv : perSlot)
      if (Engine *e = engines_[kv.first])
        if (e->collector_ && !e->gameOver_) prevStates[kv.first] = e->gs()->Clone();
  }

  grl_step_outputs out{};
  out.reward = reward_.data();
  out.done = done_.data();
  out.winner = winner_.data();
  out.step_error = stepErr_.data();
  out.action_index = actionIndex_.data();
  if (collect) out.obs = nextObs.data();
  Check(f.step_fused(env_, acts.data(), GRL_STEP_FLAG_NONE, 0, &out), "grl_step_fused");
  Check(f.sync(env_), "grl_sync");

  for (const auto &kv : perSlot) {
    const int slot = kv.first;
    Engine *e = engines_[slot];
    const int code = stepErr_[slot];
    int turn = -1;
    if (e) {
      e->stale_ = true;
      const bool wasOver = e->gameOver_;
      e->gameOver_ = done_[slot] != 0;
      e->winner_ = winner_[slot];
      turn = e->gs()->Turn;
      e->last_.clear();
      auto ps = prevStates.find(slot);
      if (e->collector_ && ps != prevStates.end() && code == GRL_STEP_OK && !wasOver) {
        // collectExperiences (turn_processor.go:182-217): one record per player that submitted a
        // MoveAction; a turn that returned an error never reaches it.
        std::map<int, Action> actionMap;
        for (const core::Action &m : kv.second)
          actionMap[m.PlayerID] = Action{ActionTypeMove, m.GetFrom(), m.GetTo()};
        for (const auto &pa : actionMap) {
          const int p = pa.first;
          if (p < 0 || p >= P_) continue;
          experience::Transition t;
          t.PlayerID = p;
          t.Turn = e->gs_.Turn;
          const size_t o = (size_t(slot) * P_ + p) * GRL_OBS_CHANNELS * N;
          t.State.assign(prevObs.begin() + o, prevObs.begin() + o + size_t(GRL_OBS_CHANNELS) * N);
          t.NextState.assign(nextObs.begin() + o, nextObs.begin() + o + size_t(GRL_OBS_CHANNELS) * N);
          t.Action = actionIndex_[size_t(slot) * P_ + p];
          t.Reward = reward_[size_t(slot) * P_ + p];
          t.Done = e->gs_.IsGameOver();
          const size_t mo = (size_t(slot) * P_ + p) * N * 4;
          t.ActionMask.resize(size_t(N) * 4);
          for (int i = 0; i < N * 4; i++) t.ActionMask[i] = prevMask[mo + i] != 0;
          e->last_.push_back(std::move(t));
        }
        e->collector_->OnStateTransition(ps->second.get(), &e->gs_, actionMap);
        if (e->gameOver_) e->collector_->OnGameEnd(&e->gs_);
      }
    }
    if (code == GRL_STEP_OK) {
      errs[slot] = core::Error();
    } else if (code == GRL_STEP_GAME_OVER) {
      errs[slot] = core::WrapGameStateError(turn, "step", core::ErrGameOver);  // turn_processor.go:105-110
    } else if (const core::Sentinel *s = core::SentinelForCode(code)) {
      // engine.go:110-113 wraps with "processing actions", turn_processor.go:148-155 with "action processing"
      errs[slot] = core::WrapGameStateError(
          turn, "action processing", core::WrapGameStateError(turn, "processing actions", core::Error(*s)));
    } else {
      errs[slot] = core::Error::New("grlcuda: step error " + std::to_string(code));
    }
  }
  return errs;
}

void EnginePool::StepBatch(const grl_action *actions, const grl_step_outputs *out, bool randomPolicy, uint64_t policySeed) {
  const auto &f = lib_->fn();
  Check(f.step_fused(env_, actions, randomPolicy ? GRL_STEP_FLAG_RANDOM_POLICY : GRL_STEP_FLAG_NONE, policySeed, out),
        "grl_step_fused");
  Check(f.sync(env_), "grl_sync");
  for (Engine *e : engines_)
    if (e) e->stale_ = true;  // gs(), IsGameOver(), GetWinner() re-read the slot from HBM on demand
}

// ---- Engine --------------------------------------------------------------------------------
Engine::Engine(EnginePool *pool, int slot, std::shared_ptr<EnginePool> owned)
    : pool_(pool), owned_(std::move(owned)), slot_(slot) {}

Engine::~Engine() {
  if (pool_) pool_->Release(slot_);
}

std::unique_ptr<Engine> NewEngine(const context::Context &ctx, const GameConfig &cfg, std::shared_ptr<Library> lib,
                                  int device) {
  std::shared_ptr<EnginePool> pool;
  try {
    pool = std::make_shared<EnginePool>(std::move(lib), 1, cfg.Width, cfg.Height, cfg.Players, device);
  } catch (const std::runtime_error &) {
    if (!cfg.Width || !cfg.Height || !cfg.Players) return nullptr;  // engine.go:65-69
    throw;
  }
  std::unique_ptr<Engine> e = pool->NewEngine(ctx, cfg);
  if (e) e->owned_ = pool;
  return e;
}

void Engine::Refresh() {
  if (!pool_) throw std::runtime_error("grl: engine outlived its pool");
  const int N = pool_->W_ * pool_->H_, P = pool_->P_;
  std::vector<int32_t> owner(N), army(N), type(N), alive(P), armyCount(P), generalIdx(P);
  std::vector<uint32_t> visible(N);
  std::vector<uint8_t> owned(size_t(P) * N), changed(N), vchg(N);
  int32_t turn = 0, over = 0, winner = -1;
  grl_state_planes pl{};
  pl.owner = owner.data();
  pl.army = army.data();
  pl.type = type.data();
  pl.visible = visible.data();
  pl.owned = owned.data();
  pl.changed = changed.data();
  pl.vis_changed = vchg.data();
  pl.turn = &turn;
  pl.game_over = &over;
  pl.winner = &winner;
  pl.alive = alive.data();
  pl.army_count = armyCount.data();
  pl.general_idx = generalIdx.data();
  pool_->Check(pool_->lib_->fn().get_state(pool_->env_, slot_, 1, &pl), "grl_get_state");
  if (!gs_.Board || gs_.Board->W != pool_->W_ || gs_.Board->H != pool_->H_) gs_.Board = core::NewBoard(pool_->W_, pool_->H_);
  for (int i = 0; i < N; i++) {
    core::Tile &t = gs_.Board->T[i];
    t.Owner = owner[i];
    t.Army = army[i];
    t.Type = type[i];
    t.VisibleBitfield = visible[i];
  }
  gs_.Turn = turn;
  gs_.Players.resize(P);
  gs_.ChangedTiles.clear();
  gs_.VisibilityChangedTiles.clear();
  for (int i = 0; i < N; i++) {
    if (changed[i]) gs_.ChangedTiles[i] = true;
    if (vchg[i]) gs_.VisibilityChangedTiles[i] = true;
  }
  for (int p = 0; p < P; p++) {
    Player &pl2 = gs_.Players[p];
    pl2.ID = p;
    pl2.Alive = alive[p] != 0;
    pl2.ArmyCount = armyCount[p];
    pl2.GeneralIdx = generalIdx[p];
    pl2.OwnedTiles.clear();
    for (int i = 0; i < N; i++)
      if (owned[size_t(p) * N + i]) pl2.OwnedTiles.push_back(i);
  }
  gameOver_ = over != 0;
  winner_ = winner;
  stale_ = false;
}

game::GameState *Engine::gs() {
  if (stale_) Refresh();
  return &gs_;
}

void Engine::Upload() {
  if (!pool_) throw std::runtime_error("grl: engine outlived its pool");
  const int N = pool_->W_ * pool_->H_, P = pool_->P_;
  if (!gs_.Board || int(gs_.Board->T.size()) != N || int(gs_.Players.size()) != P)
    throw std::invalid_argument("grl: Upload: state does not have the pool's shape");
  std::vector<int32_t> owner(N), army(N), type(N), alive(P), armyCount(P), generalIdx(P);
  std::vector<uint32_t> visible(N);
  std::vector<uint8_t> owned(size_t(P) * N, 0), changed(N, 0), vchg(N, 0);
  for (int i = 0; i < N; i++) {
    const core::Tile &t = gs_.Board->T[i];
    owner[i] = t.Owner;
    army[i] = t.Army;
    type[i] = t.Type;
    visible[i] = t.VisibleBitfield;
  }
  for (const auto &kv : gs_.ChangedTiles)
    if (kv.second && kv.first >= 0 && kv.first < N) changed[kv.first] = 1;
  for (const auto &kv : gs_.VisibilityChangedTiles)
    if (kv.second && kv.first >= 0 && kv.first < N) vchg[kv.first] = 1;
  for (int p = 0; p < P; p++) {
    const Player &pl2 = gs_.Players[p];
    alive[p] = pl2.Alive ? 1 : 0;
    armyCount[p] = pl2.ArmyCount;
    generalIdx[p] = pl2.GeneralIdx;
    for (int i : pl2.OwnedTiles)
      if (i >= 0 && i < N) owned[size_t(p) * N + i] = 1;
  }
  int32_t turn = gs_.Turn, over = gameOver_ ? 1 : 0, err = 0;
  grl_state_planes pl{};
  pl.owner = owner.data();
  pl.army = army.data();
  pl.type = type.data();
  pl.visible = visible.data();
  pl.owned = owned.data();
  pl.changed = changed.data();
  pl.vis_changed = vchg.data();
  pl.turn = &turn;
  pl.game_over = &over;
  pl.alive = alive.data();
  pl.army_count = armyCount.data();
  pl.general_idx = generalIdx.data();
  pl.step_error = &err;
  pool_->Check(pool_->lib_->fn().set_state(pool_->env_, slot_, 1, &pl), "grl_set_state");
  pool_->done_[slot_] = uint8_t(over);
  stale_ = false;
}

void Engine::SetGameOver(bool over) {
  gs();
  gameOver_ = over;
  Upload();
}

core::Error Engine::Step(const context::Context &ctx, const std::vector<core::Action> &actions) {
  if (!pool_) throw std::runtime_error("grl: engine outlived its pool");
  std::map<int, std::vector<core::Action>> one;
  one[slot_] = actions;
  return pool_->StepAll(ctx, one)[slot_];
}

game::GameState Engine::GameState() { return *gs(); }
bool Engine::IsGameOver() {
  gs();
  return gameOver_;
}
int Engine::GetWinner() {
  gs();
  return gameOver_ ? winner_ : -1;  // engine.go:250-253
}

std::vector<bool> Engine::GetLegalActionMask(int playerID) {
  if (!pool_) throw std::runtime_error("grl: engine outlived its pool");
  const int N = pool_->W_ * pool_->H_, P = pool_->P_, B = pool_->B_;
  std::vector<bool> mask(size_t(N) * 4, false);
  if (playerID < 0 || playerID >= P) return mask;  // engine.go:273-276
  std::vector<uint8_t> all(size_t(B) * P * N * 4);
  pool_->Check(pool_->lib_->fn().mask(pool_->env_, GRL_MASK_ENGINE_URDL, all.data()), "grl_mask");
  const size_t base = (size_t(slot_) * P + playerID) * N * 4;
  for (int i = 0; i < N * 4; i++) mask[i] = all[base + i] != 0;
  return mask;
}

std::vector<bool> Engine::SerializerActionMask(int playerID) {
  if (!pool_) throw std::runtime_error("grl: engine outlived its pool");
  const int N = pool_->W_ * pool_->H_, P = pool_->P_, B = pool_->B_;
  std::vector<bool> mask(size_t(N) * 4, false);
  if (playerID < 0 || playerID >= P) return mask;
  std::vector<uint8_t> all(size_t(B) * P * N * 4);
  pool_->Check(pool_->lib_->fn().mask(pool_->env_, GRL_MASK_SERIALIZER_UDLR, all.data()), "grl_mask");
  const size_t base = (size_t(slot_) * P + playerID) * N * 4;
  for (int i = 0; i < N * 4; i++) mask[i] = all[base + i] != 0;
  return mask;
}

PlayerVisibility Engine::ComputePlayerVisibility(int playerID) {
  if (!pool_) throw std::runtime_error("grl: engine outlived its pool");
  const int N = pool_->W_ * pool_->H_, P = pool_->P_, B = pool_->B_;
  PlayerVisibility vis;
  vis.VisibleTiles.assign(N, false);
  vis.FogTiles.assign(N, false);
  if (playerID < 0 || playerID >= P) {
    // Tile.IsVisibleTo is false for an id without a bit; fog still marks the special tiles
    // (visibility_optimized.go:166-195)
    const game::GameState *s = gs();
    for (int i = 0; i < N; i++) vis.FogTiles[i] = s->Board->T[i].Type != core::TileNormal;
    return vis;
  }
  std::vector<uint8_t> v(size_t(B) * P * N), g(size_t(B) * P * N);
  pool_->Check(pool_->lib_->fn().visibility(pool_->env_, v.data(), g.data()), "grl_visibility");
  const size_t base = (size_t(slot_) * P + playerID) * N;
  for (int i = 0; i < N; i++) {
    vis.VisibleTiles[i] = v[base + i] != 0;
    vis.FogTiles[i] = g[base + i] != 0;
  }
  return vis;
}

std::map<int, bool> Engine::GetChangedTiles() { return gs()->ChangedTiles; }
std::map<int, bool> Engine::GetVisibilityChangedTiles() { return gs()->VisibilityChangedTiles; }

std::vector<float> Engine::StateTensor(int playerID) {
  if (!pool_) throw std::runtime_error("grl: engine outlived its pool");
  const int N = pool_->W_ * pool_->H_, P = pool_->P_, B = pool_->B_;
  if (playerID < 0 || playerID >= P) throw std::out_of_range("grl: StateTensor: player out of range");
  std::vector<float> all(size_t(B) * P * GRL_OBS_CHANNELS * N);
  grl_step_outputs ro{};
  ro.obs = all.data();
  pool_->Check(pool_->lib_->fn().observe(pool_->env_, &ro), "grl_observe");
  const size_t o = (size_t(slot_) * P + playerID) * GRL_OBS_CHANNELS * N;
  return std::vector<float>(all.begin() + o, all.begin() + o + size_t(GRL_OBS_CHANNELS) * N);
}

float Engine::LastReward(int playerID) const {
  if (!pool_ || playerID < 0 || playerID >= pool_->P_) return 0.f;
  return pool_->reward_[size_t(slot_) * pool_->P_ + playerID];
}

const experience::Transition *Engine::LastTransition(int playerID) const {
  for (const experience::Transition &t : last_)
    if (t.PlayerID == playerID) return &t;
  return nullptr;
}

// ---- rendering.go:34-143 (ANSI colours :10-30) ------------------------------------------------
namespace {
const char *kReset = "\033[0m", *kGray = "\033[90m", *kWhite = "\033[37m";
const char *kPlayerColors[] = {"\033[31m", "\033[34m", "\033[32m", "\033[33m", "\033[35m", "\033[36m"};
const char *PlayerColor(int owner) {  // rendering.go:212-217
  if (owner < 0 || owner >= 6) return kWhite;
  return kPlayerColors[owner];
}
std::string FixedWidth(int v, int width) {  // core/utils.go IntToStringFixedWidth: right-aligned, space padded
  std::string s = std::to_string(v);
  if (int(s.size()) < width) s.insert(0, size_t(width) - s.size(), ' ');
  return s;
}
}  // namespace

std::string Engine::Board(int playerID) {
  const game::GameState *s = gs();
  const core::Board &b = *s->Board;
  static const char *Empty = "\xC2\xB7", *City = "\xE2\xAC\xA2", *General = "\xE2\x99\x94", *Mountain = "\xE2\x96\xB2";
  static const char Symbols[] = "ABCDEFGH";
  std::string sb = "    ";
  for (int x = 0; x < b.W; x++) sb += FixedWidth(x, 2);
  sb += "\n";
  for (int y = 0; y < b.H; y++) {
    sb += FixedWidth(y, 2);
    sb += " ";
    for (int x = 0; x < b.W; x++) {
      const core::Tile &t = b.T[b.Idx(x, y)];
      const bool visible = playerID < 0 || !s->FogOfWarEnabled || t.IsVisibleTo(playerID);
      if (!visible) {
        sb += kGray;
        sb += " ";
      } else if (t.IsMountain()) {
        sb += kGray;
        sb += " ";
        sb += Mountain;
      } else if (t.IsGeneral()) {
        sb += PlayerColor(t.Owner);
        sb += Symbols[((t.Owner % 8) + 8) % 8];
        sb += General;
      } else if (t.IsCity() && t.IsNeutral()) {
        sb += kWhite;
        sb += " ";
        sb += City;
      } else if (t.IsCity()) {
        sb += PlayerColor(t.Owner);
        sb += Symbols[((t.Owner % 8) + 8) % 8];
        sb += City;
      } else if (t.IsNeutral() && t.Type == core::TileNormal) {
        sb += kGray;
        if (t.Army == 0) {
          sb += " ";
          sb += Empty;
        } else if (t.Army >= 100) {
          sb += "++";
        } else if (t.Army >= 10) {
          sb += FixedWidth(t.Army, 2);
        } else {
          sb += " ";
          sb += FixedWidth(t.Army, 1);
        }
      } else if (t.Type == core::TileNormal) {
        sb += PlayerColor(t.Owner);
        sb += Symbols[((t.Owner % 8) + 8) % 8];
        if (t.Army >= 100)
          sb += "+";
        else if (t.Army >= 10)
          sb += FixedWidth(t.Army, 1);
        else
          sb += " ";
      }
      sb += kReset;
      sb += " ";
    }
    sb += "\n";
  }
  sb += "\n";
  sb += Empty;
  sb += "=empty ";
  sb += City;
  sb += "=city ";
  sb += General;
  sb += "=general ";
  sb += Mountain;
  sb += "=mountain A-H=players\n";
  return sb;
}

}  // namespace game
}  // namespace grl
