// host_test.cpp — the reference's Go tests for the turn path, transliterated against the C++ host mirror
// (generalsreinforcementlearning_b200/host).  Each TEST names the Go test it follows; the asserted values
// are the reference's own.  The binary takes the library to bind:
//
//   host_test --lib <path/libgrlcuda.so> --prefix grl_      (GPU box: the product, `-m gpu`)
//   host_test --lib <path/libgrloracle.so> --prefix grlo_   (no GPU: the CPU oracle's copy of the ABI —
//                                                            checks the host layer's own logic)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../generalsreinforcementlearning_b200/host/grl_engine.hpp"
#include "../../generalsreinforcementlearning_b200/host/grl_experience.hpp"

using namespace grl;
using core::MoveAction;
using core::Tile;

static std::shared_ptr<Library> g_lib;
static int g_failures = 0, g_checks = 0;
static const char *g_current = "";

#define EXPECT(cond, ...)                                                          \
  do {                                                                             \
    g_checks++;                                                                    \
    if (!(cond)) {                                                                 \
      g_failures++;                                                                \
      std::printf("  FAIL %s:%d [%s] %s — ", __FILE__, __LINE__, g_current, #cond); \
      std::printf(__VA_ARGS__);                                                    \
      std::printf("\n");                                                           \
    }                                                                              \
  } while (0)
#define EXPECT_EQ(want, got, ...) EXPECT((want) == (got), __VA_ARGS__)
#define REQUIRE(cond, ...)          \
  do {                              \
    const bool require_ok_ = bool(cond); \
    EXPECT(require_ok_, __VA_ARGS__);    \
    if (!require_ok_) return;       \
  } while (0)

struct TestCase {
  const char *name;
  std::function<void()> fn;
};
static std::vector<TestCase> &registry() {
  static std::vector<TestCase> r;
  return r;
}
struct Registrar {
  Registrar(const char *n, std::function<void()> f) { registry().push_back({n, std::move(f)}); }
};
#define TEST(name)                          \
  static void name();                       \
  static Registrar reg_##name(#name, name); \
  static void name()

// ---- helpers in the style of the Go tests -------------------------------------------------------
static std::shared_ptr<rand::Rand> newTestRNG() { return rand::New(rand::NewSource(12345)); }  // engine_test.go:15-17

static std::unique_ptr<game::Engine> newEngine(int w, int h, int players, game::ExperienceCollector *c = nullptr) {
  game::GameConfig cfg;
  cfg.Width = w;
  cfg.Height = h;
  cfg.Players = players;
  cfg.Rng = newTestRNG();
  cfg.ExperienceCollector = c;
  return game::NewEngine(context::Background(), cfg, g_lib);
}

// createTestEngineForActionMask (action_mask_test.go:16-55): a blank board, players alive with empty lists.
static std::shared_ptr<game::EnginePool> g_keep;  // keeps hand-built engines' pools alive for a test
static std::unique_ptr<game::Engine> createTestEngineForActionMask(int w, int h, int players) {
  g_keep = std::make_shared<game::EnginePool>(g_lib, 1, w, h, players);
  return g_keep->NewEngineFromBoard(*core::NewBoard(w, h), players);
}

// engine.updatePlayerStats() as the Go tests call it at turn 0 after editing the board by hand: the full
// rebuild of stats.go:33-63.  Test-side only (the product computes stats on the device every turn); the
// edited state is then uploaded to HBM.
static void updatePlayerStats(game::Engine &e) {
  game::GameState *gs = e.gs();
  for (game::Player &p : gs->Players) {
    p.OwnedTiles.clear();
    p.ArmyCount = 0;
    p.GeneralIdx = -1;
  }
  for (int i = 0; i < int(gs->Board->T.size()); i++) {
    const Tile &t = gs->Board->T[i];
    if (t.Owner < 0 || t.Owner >= int(gs->Players.size())) continue;
    game::Player &p = gs->Players[t.Owner];
    p.OwnedTiles.push_back(i);
    p.ArmyCount += t.Army;
    if (t.IsGeneral()) p.GeneralIdx = i;
  }
  for (game::Player &p : gs->Players) p.Alive = p.GeneralIdx != -1;
  e.Upload();
}

static Tile mk(int owner, int army, int type) {
  Tile t;
  t.Owner = owner;
  t.Army = army;
  t.Type = type;
  return t;
}

// ---- internal/game/engine_test.go -----------------------------------------------------------------
TEST(TestNewEngine) {  // engine_test.go:25-63
  int width = 8, height = 8, numPlayers = 2;
  auto engine = newEngine(width, height, numPlayers);
  REQUIRE(engine != nullptr, "Engine should not be nil");
  game::GameState *gs = engine->gs();
  REQUIRE(gs->Board != nullptr, "Board should not be nil");
  EXPECT_EQ(width, gs->Board->W, "Board width mismatch");
  EXPECT_EQ(height, gs->Board->H, "Board height mismatch");
  REQUIRE(int(gs->Players.size()) == numPlayers, "Incorrect number of players");
  EXPECT(!engine->IsGameOver(), "Game should not be over at start");
  EXPECT_EQ(0, gs->Turn, "Initial turn should be 0");
  int generalsFound = 0;
  for (int i = 0; i < numPlayers; i++) {
    const game::Player &player = gs->Players[i];
    EXPECT(player.Alive, "Player %d should be alive", i);
    EXPECT(player.GeneralIdx != -1, "Player %d should have a general assigned", i);
    if (player.GeneralIdx != -1) {
      const Tile &g = gs->Board->T[player.GeneralIdx];
      EXPECT_EQ(i, g.Owner, "General tile owner mismatch for player %d", i);
      EXPECT(g.IsGeneral(), "General tile type mismatch for player %d", i);
      generalsFound++;
    }
    EXPECT(player.ArmyCount >= 1, "Player %d should have at least 1 army", i);
  }
  EXPECT_EQ(numPlayers, generalsFound, "All players should have a general on the board");
}

TEST(TestEngine_Step_BasicTurn) {  // engine_test.go:65-86
  auto engine = newEngine(5, 5, 1);
  REQUIRE(engine != nullptr, "Engine should not be nil");
  int initialTurn = engine->gs()->Turn;
  int initialArmy = engine->gs()->Players[0].ArmyCount;
  core::Error err = engine->Step(context::Background(), {});
  REQUIRE(!err, "Step: %s", err.String().c_str());
  EXPECT_EQ(initialTurn + 1, engine->gs()->Turn, "Turn should increment");
  EXPECT_EQ(initialArmy + 1, engine->gs()->Players[0].ArmyCount, "general production");
  EXPECT(!engine->IsGameOver(), "Game should not be over with 1 player and no actions");
}

TEST(TestEngine_Step_GameOverReturnError) {  // engine_test.go:88-102
  auto engine = newEngine(5, 5, 1);
  REQUIRE(engine != nullptr, "nil engine");
  engine->SetGameOver(true);  // engine.gameOver = true
  core::Error err = engine->Step(context::Background(), {});
  EXPECT(errors::Is(err, core::ErrGameOver), "Step should return ErrGameOver, got %s", err.String().c_str());
  EXPECT(err.String().find("game turn 0 [step]: game is over") != std::string::npos, "message: %s", err.String().c_str());
}

TEST(TestEngine_Step_ContextCancelled) {  // turn_processor.go:31-36, 79-93
  auto engine = newEngine(5, 5, 1);
  REQUIRE(engine != nullptr, "nil engine");
  context::Context ctx = context::Context::WithCancel();
  ctx.Cancel();
  core::Error err = engine->Step(ctx, {});
  EXPECT(errors::Is(err, context::Canceled), "expected context.Canceled");
  EXPECT_EQ(0, engine->gs()->Turn, "a cancelled Step takes no turn");
}

TEST(TestEngine_ProcessTurnProduction) {  // engine_test.go:104-182, through Step on turn 25 and turn 24
  for (int scenario = 0; scenario < 2; scenario++) {
    auto engine = newEngine(5, 5, 1);
    REQUIRE(engine != nullptr, "nil engine");
    game::GameState *gs = engine->gs();
    int playerID = 0;
    int generalIdx = gs->Players[0].GeneralIdx;
    REQUIRE(generalIdx != -1, "Player should have a general");
    int cityIdx = -1, landIdx = -1;
    for (int i = 0; i < int(gs->Board->T.size()); i++) {
      Tile &t = gs->Board->T[i];
      if (t.Type == core::TileNormal && t.IsNeutral()) {
        t = mk(playerID, 5, core::TileCity);
        cityIdx = i;
        break;
      }
    }
    REQUIRE(cityIdx != -1, "Could not place a test city");
    for (int i = 0; i < int(gs->Board->T.size()); i++) {
      Tile &t = gs->Board->T[i];
      if (t.Type == core::TileNormal && t.IsNeutral() && i != cityIdx && i != generalIdx) {
        t.Owner = playerID;
        t.Army = 2;
        landIdx = i;
        break;
      }
    }
    REQUIRE(landIdx != -1, "Could not place test land");
    // Step increments the turn first, so start one before the turn under test.
    gs->Turn = scenario == 0 ? 24 : 23;
    updatePlayerStats(*engine);
    int g0 = gs->Board->T[generalIdx].Army, c0 = gs->Board->T[cityIdx].Army, l0 = gs->Board->T[landIdx].Army;
    core::Error err = engine->Step(context::Background(), {});
    REQUIRE(!err, "Step: %s", err.String().c_str());
    gs = engine->gs();
    EXPECT_EQ(g0 + 1, gs->Board->T[generalIdx].Army, "General production mismatch (scenario %d)", scenario);
    EXPECT_EQ(c0 + 1, gs->Board->T[cityIdx].Army, "City production mismatch (scenario %d)", scenario);
    EXPECT_EQ(l0 + (scenario == 0 ? 1 : 0), gs->Board->T[landIdx].Army, "Normal land production (scenario %d)", scenario);
  }
}

TEST(TestEngine_PlayerEliminationAndTileTurnover) {  // engine_test.go:184-248
  auto engine = newEngine(5, 5, 2);
  REQUIRE(engine != nullptr, "nil engine");
  game::GameState *gs = engine->gs();
  core::Board &board = *gs->Board;
  int p0AttackerIdx = board.Idx(0, 0);
  board.T[p0AttackerIdx] = mk(0, 20, core::TileNormal);
  int p1GeneralOriginalIdx = gs->Players[1].GeneralIdx;
  REQUIRE(p1GeneralOriginalIdx != -1, "Player 1 should have a general from NewEngine");
  int p1NewGeneralIdx = board.Idx(0, 1);
  if (p1GeneralOriginalIdx != p1NewGeneralIdx) board.T[p1GeneralOriginalIdx] = mk(core::NeutralID, 0, core::TileNormal);
  board.T[p1NewGeneralIdx] = mk(1, 1, core::TileGeneral);
  gs->Players[1].GeneralIdx = p1NewGeneralIdx;
  int p1CityIdx = board.Idx(1, 1);
  board.T[p1CityIdx] = mk(1, 5, core::TileCity);
  int p1LandIdx = board.Idx(2, 2);
  board.T[p1LandIdx] = mk(1, 3, core::TileNormal);
  REQUIRE(gs->Players[0].GeneralIdx != p1NewGeneralIdx, "Test setup conflict");
  updatePlayerStats(*engine);

  MoveAction action;
  action.PlayerID = 0;
  action.FromX = 0, action.FromY = 0, action.ToX = 0, action.ToY = 1;
  action.MoveAll = true;
  core::Error err = engine->Step(context::Background(), {action});
  REQUIRE(!err, "Step should not error during capture: %s", err.String().c_str());

  gs = engine->gs();
  EXPECT(!gs->Players[1].Alive, "Player 1 should be eliminated");
  EXPECT_EQ(-1, gs->Players[1].GeneralIdx, "Player 1 should have no general index");
  EXPECT(gs->Players[0].Alive, "Player 0 should still be alive");
  EXPECT_EQ(0, gs->Board->T[p1NewGeneralIdx].Owner, "Captured general tile should be owned by Player 0");
  EXPECT_EQ(19, gs->Board->T[p1NewGeneralIdx].Army, "20 - 1 left - 1 defender + 1 production");
  EXPECT_EQ(0, gs->Board->T[p1CityIdx].Owner, "Player 1's city should now be owned by Player 0");
  EXPECT_EQ(6, gs->Board->T[p1CityIdx].Army, "5 + 1 production");
  EXPECT_EQ(0, gs->Board->T[p1LandIdx].Owner, "Player 1's land should now be owned by Player 0");
  EXPECT_EQ(3, gs->Board->T[p1LandIdx].Army, "no production for normal tiles on turn 1");
  EXPECT(engine->IsGameOver(), "Game should be over after elimination");
  EXPECT_EQ(0, engine->GetWinner(), "Player 0 should be the winner");
}

TEST(TestEngine_Step_ActionFromDeadPlayer) {  // engine_test.go:250-303
  auto engine = newEngine(5, 5, 2);
  REQUIRE(engine != nullptr, "nil engine");
  game::GameState *gs = engine->gs();
  core::Board &board = *gs->Board;
  if (gs->Players[1].GeneralIdx != -1) board.T[gs->Players[1].GeneralIdx] = mk(core::NeutralID, 0, core::TileNormal);
  gs->Players[1].Alive = false;
  gs->Players[1].GeneralIdx = -1;
  int p0TileIdx = board.Idx(0, 0);
  board.T[p0TileIdx] = mk(0, 10, core::TileNormal);
  int p1OwnedTileIdx = board.Idx(1, 1);
  board.T[p1OwnedTileIdx] = mk(1, 5, core::TileNormal);
  // the Go test edits tiles without refreshing the cached lists; P0's move validates against ownership only
  engine->Upload();

  MoveAction dead, live;
  dead.PlayerID = 1, dead.FromX = 1, dead.FromY = 1, dead.ToX = 1, dead.ToY = 2, dead.MoveAll = true;
  live.PlayerID = 0, live.FromX = 0, live.FromY = 0, live.ToX = 0, live.ToY = 1, live.MoveAll = true;
  core::Error err = engine->Step(context::Background(), {dead, live});
  REQUIRE(!err, "Step: %s", err.String().c_str());
  gs = engine->gs();
  EXPECT_EQ(1, gs->Board->T[p1OwnedTileIdx].Owner, "P1's tile ownership should not change");
  EXPECT_EQ(5, gs->Board->T[p1OwnedTileIdx].Army, "P1's tile army should not change");
  EXPECT_EQ(1, gs->Board->T[p0TileIdx].Army, "P0's original tile should have 1 army left");
  const Tile &target = gs->Board->T[gs->Board->Idx(0, 1)];
  EXPECT_EQ(0, target.Owner, "P0 should own target tile (0,1)");
  EXPECT_EQ(9, target.Army, "P0's target tile army count is wrong");
  EXPECT(gs->Players[0].Alive, "Player 0 should be alive");
  EXPECT(!gs->Players[1].Alive, "Player 1 should remain dead");
}

// ---- internal/game/action_mask_test.go -------------------------------------------------------------
static int countTrue(const std::vector<bool> &m) {
  int n = 0;
  for (bool b : m) n += b ? 1 : 0;
  return n;
}

TEST(TestGetLegalActionMask_BasicScenario) {  // action_mask_test.go:57-103
  auto engine = createTestEngineForActionMask(3, 3, 2);
  REQUIRE(engine != nullptr, "nil engine");
  int centerIdx = engine->gs()->Board->Idx(1, 1);
  engine->gs()->Board->T[centerIdx].Owner = 0;
  engine->gs()->Board->T[centerIdx].Army = 5;
  engine->gs()->Players[0].OwnedTiles = {centerIdx};
  engine->Upload();
  std::vector<bool> mask = engine->GetLegalActionMask(0);
  EXPECT_EQ(36, int(mask.size()), "mask length");
  int base = (1 * 3 + 1) * 4;
  EXPECT(mask[base + 0], "Should be able to move up");
  EXPECT(mask[base + 1], "Should be able to move right");
  EXPECT(mask[base + 2], "Should be able to move down");
  EXPECT(mask[base + 3], "Should be able to move left");
  EXPECT_EQ(4, countTrue(mask), "Should have exactly 4 legal moves");
}

TEST(TestGetLegalActionMask_EdgeTiles) {  // action_mask_test.go:105-127
  auto engine = createTestEngineForActionMask(3, 3, 1);
  REQUIRE(engine != nullptr, "nil engine");
  int cornerIdx = engine->gs()->Board->Idx(0, 0);
  engine->gs()->Board->T[cornerIdx].Owner = 0;
  engine->gs()->Board->T[cornerIdx].Army = 3;
  engine->gs()->Players[0].OwnedTiles = {cornerIdx};
  engine->Upload();
  std::vector<bool> mask = engine->GetLegalActionMask(0);
  EXPECT(!mask[0], "Cannot move up from top edge");
  EXPECT(mask[1], "Should be able to move right");
  EXPECT(mask[2], "Should be able to move down");
  EXPECT(!mask[3], "Cannot move left from left edge");
}

TEST(TestGetLegalActionMask_InsufficientArmy) {  // action_mask_test.go:129-167
  auto engine = createTestEngineForActionMask(3, 3, 1);
  REQUIRE(engine != nullptr, "nil engine");
  core::Board &b = *engine->gs()->Board;
  int t1 = b.Idx(0, 0), t2 = b.Idx(1, 1);
  b.T[t1].Owner = 0, b.T[t1].Army = 1;
  b.T[t2].Owner = 0, b.T[t2].Army = 2;
  engine->gs()->Players[0].OwnedTiles = {t1, t2};
  engine->Upload();
  std::vector<bool> mask = engine->GetLegalActionMask(0);
  for (int d = 0; d < 4; d++) EXPECT(!mask[t1 * 4 + d], "Should not be able to move from tile with 1 army");
  bool any = false;
  for (int d = 0; d < 4; d++) any = any || mask[t2 * 4 + d];
  EXPECT(any, "Should have at least one legal move from tile with 2 armies");
}

TEST(TestGetLegalActionMask_Mountains) {  // action_mask_test.go:169-202
  auto engine = createTestEngineForActionMask(3, 3, 1);
  REQUIRE(engine != nullptr, "nil engine");
  core::Board &b = *engine->gs()->Board;
  int c = b.Idx(1, 1);
  b.T[c].Owner = 0, b.T[c].Army = 5;
  engine->gs()->Players[0].OwnedTiles = {c};
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++)
      if (dx || dy) b.T[b.Idx(1 + dx, 1 + dy)].Type = core::TileMountain;
  engine->Upload();
  std::vector<bool> mask = engine->GetLegalActionMask(0);
  for (int d = 0; d < 4; d++) EXPECT(!mask[c * 4 + d], "Should not be able to move to mountain tile");
}

TEST(TestGetLegalActionMask_DeadPlayer) {  // action_mask_test.go:204-222
  auto engine = createTestEngineForActionMask(3, 3, 2);
  REQUIRE(engine != nullptr, "nil engine");
  core::Board &b = *engine->gs()->Board;
  int t = b.Idx(1, 1);
  b.T[t].Owner = 0, b.T[t].Army = 10;
  engine->gs()->Players[0].OwnedTiles = {t};
  engine->gs()->Players[0].Alive = false;
  engine->Upload();
  EXPECT_EQ(0, countTrue(engine->GetLegalActionMask(0)), "Dead player should have no legal moves");
}

TEST(TestGetLegalActionMask_InvalidPlayer) {  // action_mask_test.go:224-243
  auto engine = createTestEngineForActionMask(3, 3, 2);
  REQUIRE(engine != nullptr, "nil engine");
  std::vector<bool> mask = engine->GetLegalActionMask(-1);
  EXPECT_EQ(36, int(mask.size()), "len");
  EXPECT_EQ(0, countTrue(mask), "Invalid player should have no legal moves");
  mask = engine->GetLegalActionMask(5);
  EXPECT_EQ(36, int(mask.size()), "len");
  EXPECT_EQ(0, countTrue(mask), "Invalid player should have no legal moves");
}

TEST(TestGetLegalActionMask_ComplexScenario) {  // action_mask_test.go:245-292
  auto engine = createTestEngineForActionMask(5, 5, 2);
  REQUIRE(engine != nullptr, "nil engine");
  core::Board &b = *engine->gs()->Board;
  const int tiles[4][3] = {{1, 1, 5}, {2, 1, 1}, {3, 3, 3}, {0, 0, 2}};
  std::vector<int> owned;
  for (auto &t : tiles) {
    int idx = b.Idx(t[0], t[1]);
    b.T[idx].Owner = 0, b.T[idx].Army = t[2];
    owned.push_back(idx);
  }
  engine->gs()->Players[0].OwnedTiles = owned;
  b.T[b.Idx(1, 2)].Type = core::TileMountain;
  engine->Upload();
  std::vector<bool> mask = engine->GetLegalActionMask(0);
  int legal = countTrue(mask);
  EXPECT(legal > 0, "Should have some legal moves");
  EXPECT(legal < 20, "Should not have too many legal moves");
  EXPECT_EQ(9, legal, "(1,1): U,R,L = 3; (3,3): 4; (0,0): R,D = 2");
  EXPECT(!mask[(1 * 5 + 1) * 4 + 2], "Should not be able to move into mountain");
}

// ---- internal/game/core/action_test.go:23-186: every case through the host-side Validate AND through
//      Step on the device (same sentinel either way) ----------------------------------------------------
struct ValidateCase {
  const char *name;
  MoveAction action;
  int fromOwner, fromArmy;  // tile (FromX,FromY) when in bounds; owner -2: leave the board untouched
  bool mountainTarget;
  const core::Sentinel *want;  // nullptr: valid
};

static MoveAction mv(int fx, int fy, int tx, int ty, bool all = false) {
  MoveAction m;
  m.PlayerID = 0, m.FromX = fx, m.FromY = fy, m.ToX = tx, m.ToY = ty, m.MoveAll = all;
  return m;
}

TEST(TestMoveAction_Validate) {
  const int W = 5, H = 5;
  const std::vector<ValidateCase> cases = {
      {"ValidMove", mv(1, 1, 1, 2, true), 0, 5, false, nullptr},
      {"FromXNegative", mv(-1, 1, 0, 1), 0, 5, false, &core::ErrInvalidCoordinates},
      {"FromXTooLarge", mv(W, 1, 0, 1), 0, 5, false, &core::ErrInvalidCoordinates},
      {"FromYNegative", mv(1, -1, 1, 0), 0, 5, false, &core::ErrInvalidCoordinates},
      {"FromYTooLarge", mv(1, H, 1, 0), 0, 5, false, &core::ErrInvalidCoordinates},
      {"ToXNegative", mv(1, 1, -1, 1), 0, 5, false, &core::ErrInvalidCoordinates},
      {"ToXTooLarge", mv(1, 1, W, 1), 0, 5, false, &core::ErrInvalidCoordinates},
      {"ToYNegative", mv(1, 1, 1, -1), 0, 5, false, &core::ErrInvalidCoordinates},
      {"ToYTooLarge", mv(1, 1, 1, H), 0, 5, false, &core::ErrInvalidCoordinates},
      {"MoveToSelf", mv(1, 1, 1, 1), 0, 10, false, &core::ErrMoveToSelf},
      {"NotAdjacent", mv(1, 1, 3, 3), 0, 10, false, &core::ErrNotAdjacent},
      {"NotAdjacentButSameRowFar", mv(1, 1, 3, 1), 0, 10, false, &core::ErrNotAdjacent},
      {"NotOwned", mv(1, 1, 1, 2), 1, 10, false, &core::ErrNotOwned},
      {"InsufficientArmy_ArmyIs1", mv(1, 1, 1, 2), 0, 1, false, &core::ErrInsufficientArmy},
      {"InsufficientArmy_ArmyIs0", mv(1, 1, 1, 2), 0, 0, false, &core::ErrInsufficientArmy},
      {"TargetIsMountain", mv(1, 1, 1, 2), 0, 10, true, &core::ErrTargetIsMountain},
  };
  for (const ValidateCase &tc : cases) {
    auto board = core::NewBoard(W, H);
    if (board->InBounds(tc.action.FromX, tc.action.FromY)) {
      Tile &f = board->T[board->Idx(tc.action.FromX, tc.action.FromY)];
      f.Owner = tc.fromOwner;
      f.Army = tc.fromArmy;
    }
    if (tc.mountainTarget) board->T[board->Idx(tc.action.ToX, tc.action.ToY)].Type = core::TileMountain;
    core::Error err = tc.action.Validate(*board, 0);
    if (tc.want)
      EXPECT(errors::Is(err, *tc.want), "%s: Validate gave %s", tc.name, err.String().c_str());
    else
      EXPECT(!err, "%s: Validate gave %s", tc.name, err.String().c_str());

    // the same case through Engine.Step: the device reports the same sentinel (step_error plane)
    game::EnginePool pool(g_lib, 1, W, H, 2);
    auto engine = pool.NewEngineFromBoard(*board, 2);
    REQUIRE(engine != nullptr, "nil engine");
    core::Error serr = engine->Step(context::Background(), {tc.action});
    if (tc.want) {
      EXPECT(errors::Is(serr, *tc.want), "%s: Step gave %s", tc.name, serr.String().c_str());
      EXPECT(serr.String().find("game turn 1 [action processing]: game turn 1 [processing actions]: ") == 0,
             "%s: wrapping %s", tc.name, serr.String().c_str());
    } else {
      EXPECT(!serr, "%s: Step gave %s", tc.name, serr.String().c_str());
    }
    EXPECT_EQ(1, engine->gs()->Turn, "%s: the turn counter advances even when the action fails (SURVEY Q5)", tc.name);
  }
}

// ---- internal/game/core/movement_test.go:71-235 TestApplyMoveAction_BasicMovement, at engine level ------
TEST(TestApplyMoveAction_BasicMovement) {
  struct Row {
    const char *name;
    int fo, fa, to, ta;
    bool moveAll;
    int expFrom, expTo, expOwner;
  };
  const Row rows[] = {
      {"move all to own tile", 0, 10, 0, 5, true, 1, 14, 0},
      {"move half to own tile", 0, 10, 0, 5, false, 5, 10, 0},
      {"capture neutral tile", 0, 10, -1, 3, true, 1, 6, 0},
      {"failed attack on enemy tile", 0, 5, 1, 10, true, 1, 6, 1},
      {"exact army match (no capture)", 0, 6, 1, 5, true, 1, 0, 1},
      {"move half with odd number", 0, 3, 0, 0, false, 2, 1, 0},
  };
  for (const Row &r : rows) {
    auto board = core::NewBoard(3, 3);
    board->T[0] = mk(r.fo, r.fa, core::TileNormal);
    board->T[1] = mk(r.to, r.ta, core::TileNormal);
    game::EnginePool pool(g_lib, 1, 3, 3, 2);
    auto engine = pool.NewEngineFromBoard(*board, 2);
    REQUIRE(engine != nullptr, "nil engine");
    core::Error err = engine->Step(context::Background(), {mv(0, 0, 1, 0, r.moveAll)});
    EXPECT(!err, "%s: %s", r.name, err.String().c_str());
    game::GameState gs = engine->GameState();
    EXPECT_EQ(r.expFrom, gs.Board->T[0].Army, "%s: from army", r.name);
    EXPECT_EQ(r.expTo, gs.Board->T[1].Army, "%s: to army", r.name);
    EXPECT_EQ(r.expOwner, gs.Board->T[1].Owner, "%s: to owner", r.name);
    std::map<int, bool> changed = engine->GetChangedTiles();
    EXPECT(changed.count(0) && changed.count(1), "%s: both tiles enter ChangedTiles (movement.go:57-60)", r.name);
    bool captured = r.expOwner == 0 && r.to != 0;
    EXPECT_EQ(captured ? size_t(1) : size_t(0), engine->GetVisibilityChangedTiles().count(1), "%s: vis-changed", r.name);
  }
}

// ---- visibility.go:153-190 / visibility_optimized.go:166-195 ------------------------------------------
TEST(TestComputePlayerVisibility) {
  auto engine = newEngine(8, 8, 2);
  REQUIRE(engine != nullptr, "nil engine");
  game::GameState gs = engine->GameState();
  for (int p = 0; p < 2; p++) {
    game::PlayerVisibility vis = engine->ComputePlayerVisibility(p);
    REQUIRE(int(vis.VisibleTiles.size()) == 64 && int(vis.FogTiles.size()) == 64, "sizes");
    auto gxy = gs.Board->XY(gs.Players[p].GeneralIdx);
    for (int i = 0; i < 64; i++) {
      auto xy = gs.Board->XY(i);
      bool near = std::abs(xy.first - gxy.first) <= 1 && std::abs(xy.second - gxy.second) <= 1;
      EXPECT_EQ(near, bool(vis.VisibleTiles[i]), "player %d tile %d: 3x3 around the general at turn 0", p, i);
      EXPECT_EQ(gs.Board->T[i].IsVisibleTo(p), bool(vis.VisibleTiles[i]), "bitfield agrees");
      EXPECT_EQ(!near && gs.Board->T[i].Type != core::TileNormal, bool(vis.FogTiles[i]), "fog = special tiles out of sight");
    }
  }
}

// ---- experience_collector.go:4-10 + turn_processor.go:182-217 + experience/collector.go:30-98 ----------
struct RecordingCollector : game::ExperienceCollector {
  int transitions = 0, ends = 0;
  int prevTurn = -1, currTurn = -1;
  std::map<int, game::Action> lastActions;
  int prevArmyAtFrom = -1;
  void OnStateTransition(const game::GameState *prev, const game::GameState *curr,
                         const std::map<int, game::Action> &actions) override {
    transitions++;
    prevTurn = prev->Turn;
    currTurn = curr->Turn;
    lastActions = actions;
    for (const auto &kv : actions) prevArmyAtFrom = prev->Board->T[kv.second.From.ToIndex(prev->Board->W)].Army;
  }
  void OnGameEnd(const game::GameState *) override { ends++; }
};

TEST(TestExperienceCollection) {
  RecordingCollector rec;
  auto engine = newEngine(5, 5, 2, &rec);
  REQUIRE(engine != nullptr, "nil engine");
  EXPECT(engine->GetExperienceCollector() == &rec, "collector attached");
  // let the general grow, then move it out
  for (int i = 0; i < 2; i++) REQUIRE(!engine->Step(context::Background(), {}), "idle step");
  EXPECT_EQ(2, rec.transitions, "OnStateTransition runs every turn, with an empty action map when nobody moved (got %d)", rec.transitions);
  game::GameState gs = engine->GameState();
  int g = gs.Players[0].GeneralIdx;
  auto xy = gs.Board->XY(g);
  std::vector<bool> mask = engine->GetLegalActionMask(0);
  const int dxs[4] = {0, 1, 0, -1}, dys[4] = {-1, 0, 1, 0};
  int dir = -1;
  for (int d = 0; d < 4; d++)
    if (mask[g * 4 + d]) {
      dir = d;
      break;
    }
  REQUIRE(dir >= 0, "the general (army 4) has a legal move");
  MoveAction m = mv(xy.first, xy.second, xy.first + dxs[dir], xy.second + dys[dir], true);
  std::vector<float> before = engine->StateTensor(0);
  std::vector<bool> smask = engine->SerializerActionMask(0);
  int armyBefore = gs.Board->T[g].Army;
  REQUIRE(!engine->Step(context::Background(), {m}), "move step");
  EXPECT_EQ(3, rec.transitions, "one more transition");
  EXPECT_EQ(2, rec.prevTurn, "prev state is the clone taken before Turn++");
  EXPECT_EQ(3, rec.currTurn, "curr state is the engine's state after the turn");
  EXPECT_EQ(armyBefore, rec.prevArmyAtFrom, "prev state is a deep copy (state.go:37-70)");
  REQUIRE(rec.lastActions.count(0) == 1, "player 0's action is in the map");
  EXPECT(rec.lastActions[0].From == m.GetFrom() && rec.lastActions[0].To == m.GetTo(), "action coordinates");
  const experience::Transition *t = engine->LastTransition(0);
  REQUIRE(t != nullptr, "transition for the acting player");
  EXPECT(engine->LastTransition(1) == nullptr, "no record for a player that did not act (collector.go:33-36)");
  EXPECT_EQ(3, t->Turn, "Turn = currState.Turn");
  EXPECT(t->State == before, "State = StateToTensor(prevState)");
  EXPECT(t->NextState == engine->StateTensor(0), "NextState = StateToTensor(currState)");
  EXPECT(t->ActionMask == smask, "ActionMask = GenerateActionMask(prevState)");
  // ActionToIndex uses the serializer's direction order up,down,left,right (serializer.go:179-198)
  const int udlr[4] = {0, 3, 1, 2};  // engine dir U,R,D,L -> serializer dir
  EXPECT_EQ(g * 4 + udlr[dir], t->Action, "action index");
  EXPECT(t->ActionMask[t->Action], "the move taken was legal in the serializer's mask");
  EXPECT(!t->Done, "game not over");
  EXPECT(std::fabs(t->Reward - engine->LastReward(0)) == 0.f, "reward plane");
  // territory +1 (0.01) + army: general 4 -> 1 stays +1 production, 3 move to a neutral tile ...: just sign
  EXPECT(t->Reward > 0.f, "capturing a neutral tile is rewarded, got %g", double(t->Reward));
  EXPECT_EQ(0, rec.ends, "OnGameEnd only when the game ends");
}

// ---- internal/experience/buffer_test.go ------------------------------------------------------------------
static experience::Experience createTestExperience(const std::string &id, int32_t turn) {  // buffer_test.go:16-34
  experience::Experience e;
  e.ExperienceId = id;
  e.GameId = "test-game";
  e.Turn = turn;
  e.State.Shape = {9, 10, 10};
  e.State.Data.assign(900, 0.f);
  e.NextState = e.State;
  e.Action = 42;
  e.Reward = 1.0f;
  return e;
}

TEST(TestBuffer_Creation) {  // buffer_test.go:36-44
  experience::Buffer buffer(100);
  EXPECT_EQ(100, buffer.Capacity(), "capacity");
  EXPECT_EQ(0, buffer.Size(), "size");
  EXPECT(!buffer.IsFull(), "not full");
  EXPECT_EQ(10000, experience::Buffer(0).Capacity(), "default capacity (buffer.go:44-46)");
}

TEST(TestBuffer_AddAndGet) {  // buffer_test.go:46-68
  experience::Buffer buffer(10);
  for (int i = 0; i < 5; i++) EXPECT(!buffer.Add(createTestExperience(std::string(1, char('a' + i)), i)), "Add");
  EXPECT_EQ(5, buffer.Size(), "size");
  auto got = buffer.Get(3);
  EXPECT_EQ(size_t(3), got.size(), "len");
  EXPECT_EQ(2, buffer.Size(), "size after Get");
  EXPECT(got[0].ExperienceId == "a" && got[1].ExperienceId == "b" && got[2].ExperienceId == "c", "FIFO order");
}

TEST(TestBuffer_CircularBehavior) {  // buffer_test.go:70-97
  experience::Buffer buffer(3);
  for (int i = 0; i < 5; i++) EXPECT(!buffer.Add(createTestExperience(std::string(1, char('a' + i)), i)), "Add");
  EXPECT_EQ(3, buffer.Size(), "only the last 3");
  EXPECT(buffer.IsFull(), "full");
  experience::BufferStats stats = buffer.Stats();
  EXPECT_EQ(int64_t(5), stats.TotalAdded, "TotalAdded");
  EXPECT_EQ(int64_t(2), stats.TotalDropped, "TotalDropped");
  auto all = buffer.GetAll();
  EXPECT_EQ(size_t(3), all.size(), "len");
  EXPECT(all[0].ExperienceId == "c" && all[1].ExperienceId == "d" && all[2].ExperienceId == "e", "a and b were dropped");
}

TEST(TestBuffer_AddBatch_Sample_Clear_Close) {  // buffer_test.go:99-157, 234-252
  experience::Buffer buffer(10);
  std::vector<experience::Experience> batch;
  for (int i = 0; i < 7; i++) batch.push_back(createTestExperience(std::string(1, char('a' + i)), i));
  EXPECT(!buffer.AddBatch(batch), "AddBatch");
  EXPECT_EQ(7, buffer.Size(), "size");
  EXPECT_EQ(size_t(3), buffer.Sample(3).size(), "sample fewer than available");
  EXPECT_EQ(7, buffer.Size(), "sampling does not remove");
  EXPECT_EQ(size_t(7), buffer.Sample(10).size(), "sample more than available");
  auto latest = buffer.GetLatest(2);
  EXPECT(latest.size() == 2 && latest[0].ExperienceId == "f" && latest[1].ExperienceId == "g", "GetLatest");
  buffer.Clear();
  EXPECT_EQ(0, buffer.Size(), "cleared");
  EXPECT(!buffer.Add(createTestExperience("new", 0)), "can add after clear");
  EXPECT(!buffer.Close(), "Close");
  EXPECT(errors::Is(buffer.Add(createTestExperience("x", 0)), experience::ErrBufferClosed), "Add on a closed buffer");
  EXPECT(errors::Is(buffer.Close(), experience::ErrBufferClosed), "double close");
}

// ---- internal/experience/collector_test.go: through a real engine, tensors from the device ----------------
TEST(TestSimpleCollector) {
  experience::SimpleCollector collector(100, "test-game-123");  // collector_test.go:12-20
  EXPECT_EQ(100, collector.GetBuffer().Capacity(), "capacity");
  EXPECT(collector.GameID() == "test-game-123", "game id");
  EXPECT_EQ(0, collector.GetExperienceCount(), "empty");

  auto engine = newEngine(5, 5, 2, &collector);
  REQUIRE(engine != nullptr, "nil engine");
  collector.Attach(engine.get());
  for (int i = 0; i < 2; i++) REQUIRE(!engine->Step(context::Background(), {}), "idle step");
  EXPECT_EQ(0, collector.GetExperienceCount(), "no record for a turn without actions (collector.go:33)");
  // collector_test.go:60-95 TestSimpleCollector_MultipleActions: both players move, one record each
  std::vector<core::Action> acts;
  const int dxs[4] = {0, 1, 0, -1}, dys[4] = {-1, 0, 1, 0};
  for (int p = 0; p < 2; p++) {
    game::GameState gs = engine->GameState();
    int g = gs.Players[p].GeneralIdx;
    std::vector<bool> mask = engine->GetLegalActionMask(p);
    for (int d = 0; d < 4; d++)
      if (mask[g * 4 + d]) {
        MoveAction m = mv(g % 5, g / 5, g % 5 + dxs[d], g / 5 + dys[d], true);
        m.PlayerID = p;
        acts.push_back(m);
        break;
      }
  }
  REQUIRE(acts.size() == 2, "both generals can move");
  REQUIRE(!engine->Step(context::Background(), acts), "move step");
  EXPECT_EQ(2, collector.GetExperienceCount(), "one experience per acting player");
  auto exps = collector.GetExperiences();
  REQUIRE(exps.size() == 2, "len");
  for (const experience::Experience &exp : exps) {  // collector_test.go:22-58
    EXPECT(exp.GameId == "test-game-123", "GameId");
    EXPECT(exp.PlayerId == 0 || exp.PlayerId == 1, "PlayerId");
    EXPECT_EQ(3, exp.Turn, "Turn = currState.Turn");
    EXPECT(exp.State.Shape == (std::vector<int32_t>{9, 5, 5}) && exp.State.Data.size() == 225, "State tensor");
    EXPECT(exp.NextState.Shape == (std::vector<int32_t>{9, 5, 5}) && exp.NextState.Data.size() == 225, "NextState tensor");
    EXPECT_EQ(size_t(100), exp.ActionMask.size(), "ActionMask");
    EXPECT(exp.Action >= 0 && exp.Action < 100 && exp.ActionMask[exp.Action], "the action taken is legal in the mask of the previous state");
    EXPECT(!exp.Done, "not done");
    EXPECT(exp.ExperienceId.size() == 36, "an id");
    EXPECT(exp.Metadata.at("collector_version") == "1.0.0", "metadata");
    // channel 7 (visible) of the previous state marks the acting general's 3x3 neighbourhood
    float seen = 0;
    for (int i = 0; i < 25; i++) seen += exp.State.Data[7 * 25 + i];
    EXPECT(seen >= 4.f, "visible channel populated");
  }
  EXPECT(exps[0].PlayerId != exps[1].PlayerId, "one per player");
  EXPECT(exps[0].ExperienceId != exps[1].ExperienceId, "unique ids");
  EXPECT_EQ(0, collector.GetExperienceCount(), "GetExperiences drains the buffer");
}

// collector_test.go:98-123 BufferOverflow, :124-141 OnGameEnd, :142-174 GameEndingExperience, :175-196 Clear — the
// transitions come from real turns here (the records' tensors are the device's), the asserted counts and values are the
// reference's
static bool firstLegalMove(game::Engine &engine, int p, int W, MoveAction *out) {
  const int dxs[4] = {0, 1, 0, -1}, dys[4] = {-1, 0, 1, 0};
  std::vector<bool> mask = engine.GetLegalActionMask(p);
  for (int a = 0; a < int(mask.size()); a++)
    if (mask[a]) {
      const int t = a / 4, d = a % 4;
      *out = mv(t % W, t / W, t % W + dxs[d], t / W + dys[d], false);
      out->PlayerID = p;
      return true;
    }
  return false;
}

TEST(TestSimpleCollector_BufferOverflowAndClear) {
  experience::SimpleCollector collector(2, "test-game-123");  // a small buffer
  auto engine = newEngine(5, 5, 2, &collector);
  REQUIRE(engine != nullptr, "nil engine");
  collector.Attach(engine.get());
  for (int i = 0; i < 3; i++) {  // the general needs two armies before it can move
    REQUIRE(!engine->Step(context::Background(), {}), "idle step");
  }
  for (int i = 0; i < 3; i++) {
    MoveAction m;
    REQUIRE(firstLegalMove(*engine, 0, 5, &m), "player 0 has a move on turn %d", i);
    REQUIRE(!engine->Step(context::Background(), {m}), "move step");
    EXPECT_EQ(i < 2 ? i + 1 : 2, collector.GetExperienceCount(), "the third record finds the buffer full: still 2");
  }
  collector.Clear();  // collector_test.go:175-196
  EXPECT_EQ(0, collector.GetExperienceCount(), "Clear empties the buffer");
  MoveAction m;
  if (firstLegalMove(*engine, 0, 5, &m)) {
    REQUIRE(!engine->Step(context::Background(), {m}), "move step");
    EXPECT_EQ(1, collector.GetExperienceCount(), "and it fills again");
  }
}

TEST(TestSimpleCollector_GameEndingExperience) {
  experience::SimpleCollector collector(100, "test-game-123");
  auto engine = newEngine(5, 5, 2, &collector);
  REQUIRE(engine != nullptr, "nil engine");
  collector.Attach(engine.get());
  // the board of TestEngine_PlayerEliminationAndTileTurnover: player 0 takes player 1's general in one move
  game::GameState *gs = engine->gs();
  core::Board &board = *gs->Board;
  board.T[board.Idx(0, 0)] = mk(0, 20, core::TileNormal);
  const int old = gs->Players[1].GeneralIdx, gen = board.Idx(0, 1);
  if (old != gen) board.T[old] = mk(core::NeutralID, 0, core::TileNormal);
  board.T[gen] = mk(1, 1, core::TileGeneral);
  gs->Players[1].GeneralIdx = gen;
  REQUIRE(gs->Players[0].GeneralIdx != gen, "test setup conflict");
  updatePlayerStats(*engine);
  MoveAction action = mv(0, 0, 0, 1, true);
  action.PlayerID = 0;
  REQUIRE(!engine->Step(context::Background(), {action}), "capture step");
  EXPECT(engine->IsGameOver(), "the game ends with the capture");
  auto exps = collector.GetExperiences();
  REQUIRE(exps.size() == 1, "one record, for the player that acted");
  EXPECT(exps[0].Done, "Done");
  EXPECT_EQ(1.0f, exps[0].Reward, "the winner gets +1.0 (rewards.go:48-52)");
  EXPECT_EQ(0, exps[0].PlayerId, "player 0");
  // collector_test.go:124-141: OnGameEnd with a finished state does not throw; the engine reports the end once
  game::GameState fin = engine->GameState();
  collector.OnGameEnd(&fin);
  EXPECT(collector.GamesEnded() >= 1, "game end seen");
}

// ---- the pool: many games, one launch per turn --------------------------------------------------------
TEST(TestEnginePool_SlotsAreIndependentGames) {
  const int B = 6, W = 10, H = 10, P = 2;
  game::EnginePool pool(g_lib, B, W, H, P);
  std::vector<std::unique_ptr<game::Engine>> slots, solo;
  for (int i = 0; i < B; i++) {
    game::GameConfig cfg;
    cfg.Width = W, cfg.Height = H, cfg.Players = P;
    cfg.Rng = rand::New(rand::NewSource(12345 + i));
    slots.push_back(pool.NewEngine(context::Background(), cfg));
    solo.push_back(game::NewEngine(context::Background(), cfg, g_lib));
    REQUIRE(slots.back() && solo.back(), "engines");
  }
  game::GameConfig extra;
  extra.Width = W, extra.Height = H, extra.Players = P;
  EXPECT(pool.NewEngine(context::Background(), extra) == nullptr, "a full pool hands out nil");
  extra.Width = 7;
  EXPECT(pool.NewEngine(context::Background(), extra) == nullptr, "so does a shape mismatch");

  uint64_t launches0 = pool.LaunchCount();
  for (int turn = 0; turn < 30; turn++) {
    // every other slot sits out every third turn (its players have not all submitted yet)
    std::map<int, std::vector<core::Action>> perSlot;
    for (int i = 0; i < B; i++) {
      if (i % 2 == 1 && turn % 3 == 2) continue;
      std::vector<core::Action> acts;
      for (int p = 0; p < P; p++) {
        std::vector<bool> mask = slots[i]->GetLegalActionMask(p);
        int n = countTrue(mask);
        if (!n) continue;
        int k = (turn * 7 + i * 3 + p) % n;
        for (int a = 0; a < int(mask.size()); a++)
          if (mask[a] && k-- == 0) {
            const int dxs[4] = {0, 1, 0, -1}, dys[4] = {-1, 0, 1, 0};
            int tile = a / 4, d = a % 4;
            MoveAction m = mv(tile % W, tile / W, tile % W + dxs[d], tile / W + dys[d], (turn + p) % 2 == 0);
            m.PlayerID = p;
            acts.push_back(m);
            break;
          }
      }
      perSlot[i] = acts;
    }
    std::map<int, core::Error> errs = pool.StepAll(context::Background(), perSlot);
    EXPECT_EQ(perSlot.size(), errs.size(), "one error entry per stepped slot");
    for (const auto &kv : perSlot) {
      EXPECT(!errs[kv.first], "slot %d turn %d: %s", kv.first, turn, errs[kv.first].String().c_str());
      core::Error e2 = solo[kv.first]->Step(context::Background(), kv.second);
      EXPECT(!e2, "solo %d: %s", kv.first, e2.String().c_str());
    }
    for (int i = 0; i < B; i++) {
      game::GameState a = slots[i]->GameState(), b = solo[i]->GameState();
      bool same = a.Turn == b.Turn && a.ChangedTiles == b.ChangedTiles;
      for (int t = 0; same && t < W * H; t++) {
        const Tile &x = a.Board->T[t], &y = b.Board->T[t];
        same = x.Owner == y.Owner && x.Army == y.Army && x.Type == y.Type && x.VisibleBitfield == y.VisibleBitfield;
      }
      for (int p = 0; same && p < P; p++)
        same = a.Players[p].OwnedTiles == b.Players[p].OwnedTiles && a.Players[p].ArmyCount == b.Players[p].ArmyCount;
      EXPECT(same, "slot %d equals its private engine after turn %d", i, turn);
    }
  }
  if (g_lib->path().find("libgrlcuda") != std::string::npos) {
    // masks and state reads launch read-out kernels too; the turns themselves are one launch per StepAll
    EXPECT(pool.LaunchCount() - launches0 >= 30, "the CUDA library launched kernels (%llu)",
           (unsigned long long)(pool.LaunchCount() - launches0));
  }
  slots[2].reset();  // a finished game frees its slot
  game::GameConfig again;
  again.Width = W, again.Height = H, again.Players = P;
  again.Rng = newTestRNG();
  auto reused = pool.NewEngine(context::Background(), again);
  REQUIRE(reused != nullptr, "freed slot is reusable");
  EXPECT_EQ(2, reused->Slot(), "same slot");
  EXPECT_EQ(0, reused->gs()->Turn, "fresh game");
}

TEST(TestEnginePool_StepBatch) {
  // bulk callers hand the C ABI's arrays straight through; per-slot views follow
  const int B = 512, W = 10, H = 10, P = 2;
  game::EnginePool pool(g_lib, B, W, H, P);
  std::vector<std::unique_ptr<game::Engine>> slots;
  for (int i = 0; i < B; i++) {
    game::GameConfig cfg;
    cfg.Width = W, cfg.Height = H, cfg.Players = P;
    cfg.Rng = rand::New(rand::NewSource(500 + i));
    slots.push_back(pool.NewEngine(context::Background(), cfg));
    REQUIRE(slots.back() != nullptr, "engine %d", i);
  }
  for (int t = 0; t < 40; t++) pool.StepBatch(nullptr, nullptr, /*randomPolicy=*/true, /*policySeed=*/9);
  int moved = 0;
  for (int i = 0; i < B; i += 37) {
    game::GameState gs = slots[i]->GameState();
    EXPECT_EQ(40, gs.Turn, "slot %d took 40 turns", i);
    int tiles = int(gs.Players[0].OwnedTiles.size()) + int(gs.Players[1].OwnedTiles.size());
    moved += tiles > 2;
    EXPECT(!slots[i]->IsGameOver() || slots[i]->GetWinner() >= -1, "flags readable after a bulk step");
  }
  EXPECT(moved > 0, "the random policy expands territory");
}

// ---- math/rand + demo_helpers.go:12-62 ----------------------------------------------------------------------
TEST(TestGoRandAndGenerateRandomActions) {
  rand::Rand r(1);  // the canonical first values of rand.New(rand.NewSource(1))
  EXPECT(r.Int63() == 5577006791947779410LL && r.Int63() == 8674665223082153551LL && r.Int63() == 6129484611666145821LL, "Int63");
  r.Seed(1);
  const int want[10] = {81, 87, 47, 59, 81, 18, 25, 40, 56, 0};
  for (int i = 0; i < 10; i++) EXPECT_EQ(want[i], r.Intn(100), "Intn(100) #%d", i);
  r.Seed(1);
  EXPECT(std::fabs(r.Float64() - 0.6046602879796196) < 1e-16, "Float64");
  r.Seed(1);
  EXPECT(std::fabs(r.Float32() - 0.6046603f) < 1e-7f, "Float32");

  // the demo policy only ever proposes moves the engine accepts, and moves the game along
  auto engine = newEngine(10, 10, 2);
  REQUIRE(engine != nullptr, "nil engine");
  rand::Rand rng(99);
  int played = 0;
  for (int t = 0; t < 200 && !engine->IsGameOver(); t++) {
    std::vector<core::Action> acts = game::GenerateRandomActions(*engine, rng);
    for (const core::Action &a : acts) EXPECT(!a.Validate(*engine->gs()->Board, a.PlayerID), "proposed moves validate");
    played += int(acts.size());
    core::Error err = engine->Step(context::Background(), acts);
    EXPECT(!err, "turn %d: %s", t, err.String().c_str());
  }
  EXPECT(played > 20, "about 0.3 moves per player and turn (%d)", played);
  int tiles = 0;
  for (const game::Player &p : engine->gs()->Players) tiles += int(p.OwnedTiles.size());
  EXPECT(tiles > 2, "territory grew");
  // NewEngine takes the seed of a FRESH generator
  game::GameConfig cfg;
  cfg.Width = 8, cfg.Height = 8, cfg.Players = 2;
  cfg.Rng = rand::New(rand::NewSource(12345));
  auto a = game::NewEngine(context::Background(), cfg, g_lib), b = newEngine(8, 8, 2);
  REQUIRE(a && b, "engines");
  EXPECT(a->gs()->Players[0].GeneralIdx == b->gs()->Players[0].GeneralIdx, "same seed, same map");
}

// ---- rendering.go:34-143 ----------------------------------------------------------------------------
TEST(TestBoardRendering) {
  auto engine = createTestEngineForActionMask(3, 2, 2);
  REQUIRE(engine != nullptr, "nil engine");
  core::Board &b = *engine->gs()->Board;
  b.T[0] = mk(0, 7, core::TileGeneral);
  b.T[1] = mk(0, 12, core::TileNormal);
  b.T[2] = mk(-1, 0, core::TileMountain);
  b.T[3] = mk(-1, 40, core::TileCity);
  b.T[4] = mk(-1, 0, core::TileNormal);
  b.T[5] = mk(1, 150, core::TileNormal);
  for (Tile &t : b.T) t.VisibleBitfield = 1;  // player 0 sees everything
  b.T[5].VisibleBitfield = 2;                 // except the far corner
  engine->Upload();
  const std::string R = "\033[0m", G = "\033[90m", Wh = "\033[37m", Red = "\033[31m", Blue = "\033[34m";
  std::string want = "     0 1 2\n";
  want += " 0 " + Red + "A\xE2\x99\x94" + R + " " + Red + "A12" + R + " " + G + " \xE2\x96\xB2" + R + " \n";
  want += " 1 " + Wh + " \xE2\xAC\xA2" + R + " " + G + " \xC2\xB7" + R + " " + G + " " + R + " \n";
  want += "\n\xC2\xB7=empty \xE2\xAC\xA2=city \xE2\x99\x94=general \xE2\x96\xB2=mountain A-H=players\n";
  EXPECT(engine->Board(0) == want, "Board(0):\n%s\nwant:\n%s", engine->Board(0).c_str(), want.c_str());
  std::string all = engine->Board(-1);  // spectator: no fog (rendering.go:96)
  EXPECT(all.find(Blue + "B+" + R) != std::string::npos, "spectator sees player 1's 150-army tile as B+");
}

// ---- core.Coordinate / Direction (coordinate_test.go), the coordinate forms of Board and MoveAction
// (coordinate_integration_test.go:10-88), core/utils (utils_test.go) ---------------------------------
using core::Coordinate;
static Coordinate C(int x, int y) { return Coordinate{x, y}; }

TEST(TestCoordinate_IndexRoundTrip) {  // coordinate_test.go:9-69
  Coordinate c = core::NewCoordinate(3, 5);
  EXPECT(c.X == 3 && c.Y == 5, "NewCoordinate");
  struct { int index, width; Coordinate want; } cases[] = {{0, 10, {0, 0}}, {9, 10, {9, 0}}, {10, 10, {0, 1}}, {55, 10, {5, 5}},
                                                          {99, 10, {9, 9}}, {7, 4, {3, 1}}};
  for (auto &t : cases) {
    EXPECT(core::FromIndex(t.index, t.width) == t.want, "FromIndex(%d, %d)", t.index, t.width);
    EXPECT_EQ(t.index, t.want.ToIndex(t.width), "ToIndex of (%d,%d)", t.want.X, t.want.Y);
  }
  for (int i = 0; i < 100; i++) EXPECT_EQ(i, core::FromIndex(i, 10).ToIndex(10), "round trip of %d", i);
}

TEST(TestCoordinate_IsValid) {  // coordinate_test.go:71-96
  struct { Coordinate c; bool valid; } cases[] = {{{0, 0}, true},  {{5, 5}, true},   {{9, 9}, true},   {{-1, 5}, false}, {{5, -1}, false},
                                                  {{10, 5}, false}, {{5, 10}, false}, {{-1, -1}, false}, {{10, 10}, false}};
  for (auto &t : cases) EXPECT_EQ(t.valid, t.c.IsValid(10, 10), "IsValid of (%d,%d)", t.c.X, t.c.Y);
}

TEST(TestCoordinate_DistanceAndAdjacency) {  // coordinate_test.go:98-153
  struct { Coordinate a, b; int want; } dist[] = {{{5, 5}, {5, 5}, 0}, {{5, 5}, {6, 5}, 1},  {{5, 5}, {5, 6}, 1},
                                                  {{0, 0}, {1, 1}, 2}, {{0, 0}, {5, 7}, 12}, {{-2, -3}, {2, 3}, 10}};
  for (auto &t : dist) {
    EXPECT_EQ(t.want, t.a.DistanceTo(t.b), "distance");
    EXPECT_EQ(t.want, t.b.DistanceTo(t.a), "distance is symmetric");
  }
  const Coordinate center{5, 5};
  struct { Coordinate o; bool adj; } adj[] = {{{5, 4}, true},  {{6, 5}, true},  {{5, 6}, true},  {{4, 5}, true},  {{6, 4}, false}, {{6, 6}, false},
                                              {{4, 6}, false}, {{4, 4}, false}, {{5, 5}, false}, {{7, 5}, false}, {{0, 0}, false}};
  for (auto &t : adj) {
    EXPECT_EQ(t.adj, center.IsAdjacentTo(t.o), "adjacency of (%d,%d)", t.o.X, t.o.Y);
    EXPECT_EQ(t.adj, t.o.IsAdjacentTo(center), "adjacency is symmetric");
  }
}

TEST(TestCoordinate_Neighbors) {  // coordinate_test.go:155-197
  auto has = [](const std::vector<Coordinate> &v, Coordinate c) {
    for (const Coordinate &x : v)
      if (x == c) return true;
    return false;
  };
  auto nb = C(5, 5).Neighbors();
  EXPECT_EQ(size_t(4), nb.size(), "four neighbours");
  EXPECT(has(nb, {5, 4}) && has(nb, {6, 5}) && has(nb, {5, 6}) && has(nb, {4, 5}), "north, east, south, west");
  EXPECT(nb[0] == C(5, 4) && nb[1] == C(6, 5) && nb[2] == C(5, 6) && nb[3] == C(4, 5),
         "in the order of coordinate.go:59-64");
  struct { Coordinate c; int w, h, count; } cases[] = {{{5, 5}, 10, 10, 4}, {{0, 0}, 10, 10, 2}, {{9, 0}, 10, 10, 2}, {{0, 9}, 10, 10, 2},
                                                       {{9, 9}, 10, 10, 2}, {{5, 0}, 10, 10, 3}, {{5, 9}, 10, 10, 3}, {{0, 5}, 10, 10, 3},
                                                       {{9, 5}, 10, 10, 3}, {{0, 0}, 1, 1, 0}};
  for (auto &t : cases) {
    auto valid = t.c.ValidNeighbors(t.w, t.h);
    EXPECT_EQ(size_t(t.count), valid.size(), "valid neighbours of (%d,%d)", t.c.X, t.c.Y);
    for (const Coordinate &n : valid) EXPECT(n.IsValid(t.w, t.h) && n.IsAdjacentTo(t.c), "a valid neighbour is in bounds and adjacent");
  }
}

TEST(TestCoordinate_Arithmetic) {  // coordinate_test.go:199-258
  Coordinate c1{3, 4}, c2{2, -1};
  EXPECT(c1.Add(c2) == C(5, 3), "Add");
  EXPECT(c1 == C(3, 4) && c2 == C(2, -1), "operands unchanged");
  EXPECT(C(5, 3).Sub(c2) == C(3, 4), "Sub");
  struct { Coordinate a, b; bool eq; } eq[] = {{{0, 0}, {0, 0}, true},  {{5, 7}, {5, 7}, true},  {{-1, -1}, {-1, -1}, true},
                                               {{5, 7}, {7, 5}, false}, {{0, 0}, {0, 1}, false}, {{0, 0}, {1, 0}, false}};
  for (auto &t : eq) EXPECT(t.a.Equal(t.b) == t.eq && t.b.Equal(t.a) == t.eq, "Equal, both ways");
  EXPECT_EQ(std::string("(0,0)"), (C(0, 0).String()), "String");
  EXPECT_EQ(std::string("(5,7)"), (C(5, 7).String()), "String");
  EXPECT_EQ(std::string("(-1,-2)"), (C(-1, -2).String()), "String");
  EXPECT_EQ(std::string("(100,200)"), (C(100, 200).String()), "String");
}

TEST(TestCoordinate_Directions) {  // coordinate_test.go:260-322
  const Coordinate start{5, 5};
  EXPECT(start.Move(core::North) == C(5, 4), "north");
  EXPECT(start.Move(core::East) == C(6, 5), "east");
  EXPECT(start.Move(core::South) == C(5, 6), "south");
  EXPECT(start.Move(core::West) == C(4, 5), "west");
  EXPECT(start.Move(7) == start, "an unknown direction moves nowhere (coordinate.go:124-128)");
  EXPECT_EQ(int(core::North), start.DirectionTo({5, 4}), "direction north");
  EXPECT_EQ(int(core::East), start.DirectionTo({6, 5}), "direction east");
  EXPECT_EQ(int(core::South), start.DirectionTo({5, 6}), "direction south");
  EXPECT_EQ(int(core::West), start.DirectionTo({4, 5}), "direction west");
  EXPECT_EQ(-1, start.DirectionTo({6, 6}), "diagonal");
  EXPECT_EQ(-1, start.DirectionTo({5, 5}), "same");
  EXPECT_EQ(-1, start.DirectionTo({10, 10}), "far");
  EXPECT(core::DirectionVectors[core::North] == C(0, -1) && core::DirectionVectors[core::East] == C(1, 0) &&
             core::DirectionVectors[core::South] == C(0, 1) && core::DirectionVectors[core::West] == C(-1, 0),
         "direction vectors");
  for (const Coordinate &v : core::DirectionVectors) EXPECT_EQ(1, v.DistanceTo({0, 0}), "unit length");
}

TEST(TestCoordinate_ComparableAsMapKey) {  // coordinate_test.go:324-340
  std::map<Coordinate, std::string> m;
  m[C(5, 5)] = "first";
  m[C(6, 5)] = "third";
  EXPECT_EQ(std::string("first"), m[(C(5, 5))], "equal coordinates are one key");
  EXPECT_EQ(std::string("third"), m[(C(6, 5))], "third");
  EXPECT_EQ(size_t(2), m.size(), "two keys");
}

TEST(TestCoordinateIntegration_Board) {  // coordinate_integration_test.go:12-41
  auto board = core::NewBoard(10, 10);
  const Coordinate coord = core::NewCoordinate(5, 5);
  EXPECT(board->InBoundsCoord(coord), "in bounds");
  EXPECT(!board->InBoundsCoord(core::NewCoordinate(-1, 5)) && !board->InBoundsCoord(core::NewCoordinate(10, 5)), "out of bounds");
  Tile *tile = board->GetTileCoord(coord);
  REQUIRE(tile != nullptr, "tile");
  EXPECT_EQ(core::NeutralID, tile->Owner, "a fresh tile is neutral");
  Tile nt;
  nt.Owner = 1, nt.Army = 10, nt.Type = core::TileNormal;
  board->SetTile(coord, nt);
  EXPECT(board->GetTileCoord(coord)->Owner == 1 && board->GetTileCoord(coord)->Army == 10, "SetTile");
  EXPECT_EQ(55, board->IdxCoord(coord), "IdxCoord = 5 * 10 + 5");
  board->SetTile(core::NewCoordinate(-1, -1), nt);  // out of bounds: ignored
  EXPECT(board->GetTileCoord(core::NewCoordinate(-1, -1)) == nullptr, "no tile out of bounds");
}

TEST(TestCoordinateIntegration_MoveAction) {  // coordinate_integration_test.go:44-88
  MoveAction a1{1, 3, 4, 3, 5, true};
  EXPECT(a1.GetFrom() == C(3, 4) && a1.GetTo() == C(3, 5), "legacy fields");
  MoveAction a2;
  a2.PlayerID = 1, a2.From = core::NewCoordinate(5, 6), a2.To = core::NewCoordinate(5, 7);
  EXPECT(a2.GetFrom() == C(5, 6) && a2.GetTo() == C(5, 7), "coordinate fields win when set");
  MoveAction a3;
  a3.PlayerID = 1, a3.From = core::NewCoordinate(2, 3), a3.ToX = 2, a3.ToY = 4, a3.MoveAll = true;
  EXPECT(a3.GetFrom() == C(2, 3) && a3.GetTo() == C(2, 4), "mixed: To falls back to ToX, ToY");
}

TEST(TestIntToStringFixedWidth) {  // utils_test.go:10-93
  struct { int num, width; const char *want; } cases[] = {{5, 3, "  5"},   {42, 3, " 42"}, {123, 3, "123"},           {1234, 3, "1234"}, {-5, 3, " -5"},
                                                          {0, 3, "  0"},   {7, 1, "7"},    {99, 10, "        99"},   {123, 0, "123"},   {123, -5, "123  "},
                                                          {999999999, 5, "999999999"}};
  for (auto &t : cases) EXPECT_EQ(std::string(t.want), core::IntToStringFixedWidth(t.num, t.width), "%d in width %d", t.num, t.width);
}

TEST(TestGetActionType) {  // utils_test.go:95-141
  EXPECT_EQ(std::string("nil"), core::GetActionType(nullptr), "nil action");
  MoveAction mv{0, 0, 0, 1, 0, true};
  EXPECT_EQ(std::string("*core.MoveAction"), core::GetActionType(&mv), "a move");
}

// ---- core.Board / core.Tile (board_test.go:10-281) -----------------------------------------------------
TEST(TestBoard_Basics) {
  struct { int w, h; } sizes[] = {{5, 5}, {10, 20}, {100, 100}, {1, 1}};  // TestNewBoard
  for (auto &sz : sizes) {
    auto b = core::NewBoard(sz.w, sz.h);
    EXPECT(b->W == sz.w && b->H == sz.h && int(b->T.size()) == sz.w * sz.h, "%dx%d", sz.w, sz.h);
    bool fresh = true;
    for (const Tile &t : b->T) fresh = fresh && t.Owner == core::NeutralID && t.Type == core::TileNormal && t.Army == 0 && t.VisibleBitfield == 0;
    EXPECT(fresh, "every tile neutral, normal, empty, unseen");
  }
  auto board = core::NewBoard(5, 5);
  struct { int x, y, idx; } at[] = {{0, 0, 0}, {4, 0, 4}, {0, 1, 5}, {2, 2, 12}, {4, 4, 24}};  // TestBoard_Idx, TestBoard_XY
  for (auto &t : at) {
    EXPECT_EQ(t.idx, board->Idx(t.x, t.y), "Idx(%d,%d)", t.x, t.y);
    EXPECT(board->XY(t.idx) == std::make_pair(t.x, t.y), "XY(%d)", t.idx);
  }
  struct { int x, y; bool in; } bounds[] = {{0, 0, true},  {4, 0, true},  {0, 4, true},  {4, 4, true},   {2, 2, true},  {-1, 2, false},
                                            {2, -1, false}, {5, 2, false}, {2, 5, false}, {-1, -1, false}, {10, 10, false}};  // TestBoard_InBounds
  for (auto &t : bounds) EXPECT_EQ(t.in, board->InBounds(t.x, t.y), "InBounds(%d,%d)", t.x, t.y);
  board->T[0] = mk(0, 10, core::TileGeneral);  // TestBoard_GetTile
  board->T[12] = mk(1, 5, core::TileCity);
  Tile *g = board->GetTile(0, 0), *c = board->GetTile(2, 2);
  REQUIRE(g && c, "tiles in bounds");
  EXPECT(g->Owner == 0 && g->Army == 10 && g->Type == core::TileGeneral, "general tile");
  EXPECT(c->Owner == 1 && c->Army == 5 && c->Type == core::TileCity, "city tile");
  EXPECT(board->GetTile(-1, 0) == nullptr && board->GetTile(10, 10) == nullptr, "nil out of bounds");
  auto b10 = core::NewBoard(10, 10);  // TestBoard_Distance
  struct { int x1, y1, x2, y2, d; } dist[] = {{5, 5, 5, 5, 0}, {0, 0, 5, 0, 5}, {0, 0, 0, 5, 5}, {0, 0, 3, 4, 7}, {2, 2, -1, -1, 6}, {0, 0, 9, 9, 18}};
  for (auto &t : dist) EXPECT(b10->Distance(t.x1, t.y1, t.x2, t.y2) == t.d && b10->Distance(t.x2, t.y2, t.x1, t.y1) == t.d, "distance %d", t.d);
  auto one = core::NewBoard(1, 1);  // TestBoard_EdgeCases
  EXPECT(one->T.size() == 1 && one->InBounds(0, 0) && !one->InBounds(1, 0) && !one->InBounds(0, 1), "1x1");
  auto big = core::NewBoard(1000, 1000);
  EXPECT(big->T.size() == 1000000 && big->InBounds(999, 999) && !big->InBounds(1000, 1000), "1000x1000");
  auto m = core::NewBoard(5, 5);  // TestBoard_ModifyTile
  Tile *t = m->GetTile(2, 2);
  REQUIRE(t != nullptr, "tile");
  t->Owner = 1, t->Army = 50, t->Type = core::TileCity;
  Tile *same = m->GetTile(2, 2);
  EXPECT(same->Owner == 1 && same->Army == 50 && same->Type == core::TileCity, "modifications persist");
}

TEST(TestTile_Predicates) {  // board_test.go:183-246
  EXPECT(mk(core::NeutralID, 0, core::TileNormal).IsNeutral() && !mk(0, 0, core::TileNormal).IsNeutral() && !mk(1, 0, core::TileNormal).IsNeutral(), "IsNeutral");
  struct { int type; bool city, general, mountain; } types[] = {{core::TileNormal, false, false, false}, {core::TileCity, true, false, false},
                                                                {core::TileGeneral, false, true, false}, {core::TileMountain, false, false, true}};
  for (auto &t : types) {
    Tile tile = mk(core::NeutralID, 0, t.type);
    EXPECT(tile.IsCity() == t.city && tile.IsGeneral() == t.general && tile.IsMountain() == t.mountain, "type %d", t.type);
  }
  EXPECT(mk(core::NeutralID, 0, core::TileNormal).IsEmpty(), "empty neutral tile");
  EXPECT(!mk(core::NeutralID, 5, core::TileNormal).IsEmpty(), "neutral with army");
  EXPECT(!mk(0, 0, core::TileNormal).IsEmpty(), "owned empty tile");
  EXPECT(!mk(core::NeutralID, 0, core::TileCity).IsEmpty() && !mk(core::NeutralID, 0, core::TileGeneral).IsEmpty() &&
             !mk(core::NeutralID, 0, core::TileMountain).IsEmpty(), "neutral city / general / mountain");
}

// ---- core/errors (errors_test.go) ---------------------------------------------------------------------
TEST(TestWrapActionError) {  // errors_test.go:12-60
  MoveAction a0{1, 0, 0, 1, 0, false};
  EXPECT(core::WrapActionError(a0, core::Error()).IsNil(), "a nil error stays nil");
  MoveAction a1{1, 5, 3, 5, 4, false};
  core::Error w1 = core::WrapActionError(a1, core::ErrInvalidCoordinates);
  EXPECT_EQ(std::string("player 1: move from (5,3) to (5,4): invalid coordinates"), w1.String(), "message");
  EXPECT(errors::Is(w1, core::ErrInvalidCoordinates), "errors.Is through the wrap");
  MoveAction a2{2, 10, 10, 11, 10, false};
  core::Error w2 = core::WrapActionError(a2, core::ErrNotOwned);
  EXPECT_EQ(std::string("player 2: move from (10,10) to (11,10): tile not owned by player"), w2.String(), "message");
  EXPECT(errors::Is(w2, core::ErrNotOwned), "errors.Is through the wrap");
  core::Error w3 = core::WrapActionError(static_cast<const MoveAction *>(nullptr), core::ErrGameOver);
  EXPECT_EQ(std::string("player action: game is over"), w3.String(), "generic fallback");
  EXPECT(errors::Is(w3, core::ErrGameOver), "errors.Is through the fallback");
}

TEST(TestWrapGameStateAndPlayerError) {  // errors_test.go:62-158
  EXPECT(core::WrapGameStateError(50, "action", core::Error()).IsNil(), "a nil error stays nil");
  core::Error g1 = core::WrapGameStateError(100, "action", core::ErrGameOver);
  EXPECT_EQ(std::string("game turn 100 [action]: game is over"), g1.String(), "message");
  EXPECT(errors::Is(g1, core::ErrGameOver), "errors.Is");
  core::Error g2 = core::WrapGameStateError(25, "production", core::Error::New("failed to apply production"));
  EXPECT_EQ(std::string("game turn 25 [production]: failed to apply production"), g2.String(), "a plain error wraps too");
  EXPECT(!errors::Is(g2, core::ErrGameOver), "and carries no sentinel");
  EXPECT(core::WrapPlayerError(1, "move validation", core::Error()).IsNil(), "a nil error stays nil");
  core::Error p1 = core::WrapPlayerError(3, "move validation", core::ErrInsufficientArmy);
  EXPECT_EQ(std::string("player 3 move validation: insufficient army to move"), p1.String(), "message");
  EXPECT(errors::Is(p1, core::ErrInsufficientArmy), "errors.Is");
  core::Error p2 = core::WrapPlayerError(0, "action processing", core::ErrInvalidPlayer);
  EXPECT_EQ(std::string("player 0 action processing: invalid player ID"), p2.String(), "message");
  EXPECT(errors::Is(p2, core::ErrInvalidPlayer), "errors.Is");
}

TEST(TestGameError) {  // errors_test.go:160-186
  core::GameError e1 = core::NewGameError(150, 2, "capture general", core::ErrNotOwned);
  EXPECT_EQ(std::string("turn 150: player 2 capture general: tile not owned by player"), e1.String(), "with a player");
  EXPECT(e1.Is(core::ErrNotOwned), "unwraps to the sentinel");
  core::GameError e2 = core::NewGameError(200, 0, "win condition check", core::ErrGameOver);
  EXPECT_EQ(std::string("turn 200: win condition check: game is over"), e2.String(), "player 0 is left out (errors.go:61-66)");
  EXPECT(e2.Is(core::ErrGameOver), "unwraps to the sentinel");
  core::GameError e3 = core::NewGameError(50, 1, "network sync", core::Error::New("network timeout"));
  EXPECT(e3.Turn == 50 && e3.PlayerID == 1 && e3.Operation == "network sync", "fields");
  EXPECT_EQ(std::string("network timeout"), e3.Unwrap().String(), "Unwrap");
}

TEST(TestErrorUtilitiesUsageExample) {  // errors_test.go:188-216: a move from a neutral tile on an empty board
  auto board = core::NewBoard(10, 10);
  MoveAction move{1, 5, 3, 5, 4, false};
  core::Error err = move.Validate(*board, 1);
  REQUIRE(!err.IsNil(), "the tile is not the player's");
  EXPECT(errors::Is(err, core::ErrNotOwned), "not owned");
  core::Error wrapped = core::WrapActionError(move, err);
  EXPECT(wrapped.String().rfind("player 1: move from (5,3) to (5,4): ", 0) == 0, "the action context comes first: %s", wrapped.String().c_str());
  EXPECT(errors::Is(wrapped, core::ErrNotOwned), "the chain keeps the sentinel");
  EXPECT(errors::Is(core::WrapPlayerError(1, "move", core::ErrInsufficientArmy), core::ErrInsufficientArmy), "example 4");
}

// ---- the product binding must not fall back to anything ---------------------------------------------
TEST(TestLibraryMissingFailsLoudly) {
  bool threw = false;
  try {
    Library::Open("/nonexistent/libgrlcuda.so", "grl_");
  } catch (const std::runtime_error &) {
    threw = true;
  }
  EXPECT(threw, "a missing library is an error, not a fallback");
}

int main(int argc, char **argv) {
  std::string lib, prefix = "grl_", only;
  for (int i = 1; i < argc; i++) {
    if (!std::strcmp(argv[i], "--lib") && i + 1 < argc) lib = argv[++i];
    else if (!std::strcmp(argv[i], "--prefix") && i + 1 < argc) prefix = argv[++i];
    else if (!std::strcmp(argv[i], "--only") && i + 1 < argc) only = argv[++i];
  }
  try {
    g_lib = lib.empty() ? Library::Default() : Library::Open(lib, prefix);
  } catch (const std::exception &e) {
    std::printf("cannot bind the engine library: %s\n", e.what());
    return 2;
  }
  std::printf("bound %s (prefix %s)\n", g_lib->path().c_str(), prefix.c_str());
  int ran = 0;
  for (const TestCase &t : registry()) {
    if (!only.empty() && only != t.name) continue;
    g_current = t.name;
    int before = g_failures;
    try {
      t.fn();
    } catch (const std::exception &e) {
      g_failures++;
      std::printf("  EXCEPTION [%s] %s\n", t.name, e.what());
    }
    g_keep.reset();
    std::printf("%s %s\n", g_failures == before ? "ok  " : "FAIL", t.name);
    ran++;
  }
  std::printf("%d tests, %d checks, %d failures\n", ran, g_checks, g_failures);
  return g_failures ? 1 : 0;
}
