/*
 * grl_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference Go engine's turn-processing path
 * (mitchelldurbincs/GeneralsReinforcementLearning), written to mirror the Go
 * source literally: ordered OwnedTiles slices, tile sets, per-tile structs,
 * the same loops in the same order.  It deliberately shares NO code with the
 * CUDA product (generalsreinforcementlearning_b200/csrc), which uses packed
 * bitmasks instead of lists; agreement between the two is the parity claim.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product never does.
 *
 * Parity pinning: the Go toolchain is absent from the build image, so the
 * reference cannot run here.  This restatement is pinned by transliterations of
 * the reference's own known-answer tests (tests/test_oracle_kat.py; SURVEY.md
 * Appendix C) and by its seeded mapgen golden counts
 * (internal/game/mapgen/generator_test.go:84,124,148,175,451-454).  Multi-turn
 * trajectories have no reference goldens (the reference has none): for those,
 * parity is "restatement of the source", as DESIGN.md states.
 *
 * Exports the ABI of include/grlcuda.h with the prefix grlo_ (host pointers only).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../include/grlcuda.h"

/* ======================================================================== */
/* Go math/rand (v1) — go1.24 src/math/rand/{rng.go,rand.go}                 */
/* Call sites: internal/game/mapgen/generator.go:84,106,117,135,152,190      */
/* ======================================================================== */

#define RNG_LEN 607
#define RNG_TAP 273
#define INT32_MAX_GO 2147483647

static const int64_t rng_cooked[RNG_LEN] = {
#include "go_rng_cooked.inc"
};

typedef struct {
  int tap, feed;
  int64_t vec[RNG_LEN];
} go_rand;

/* rng.go seedrand: x[n+1] = 48271 * x[n] mod (2**31 - 1) */
static int32_t go_seedrand(int32_t x) {
  const int32_t A = 48271, Q = 44488, R = 3399;
  int32_t hi = x / Q, lo = x % Q;
  x = A * lo - R * hi;
  if (x < 0) x += INT32_MAX_GO;
  return x;
}

/* rng.go (*rngSource).Seed */
static void go_rand_seed(go_rand *r, int64_t seed) {
  r->tap = 0;
  r->feed = RNG_LEN - RNG_TAP;
  seed = seed % INT32_MAX_GO;
  if (seed < 0) seed += INT32_MAX_GO;
  if (seed == 0) seed = 89482311;
  int32_t x = (int32_t)seed;
  for (int i = -20; i < RNG_LEN; i++) {
    x = go_seedrand(x);
    if (i >= 0) {
      int64_t u = (int64_t)((uint64_t)x << 40);
      x = go_seedrand(x);
      u ^= (int64_t)((uint64_t)x << 20);
      x = go_seedrand(x);
      u ^= (int64_t)x;
      u ^= rng_cooked[i];
      r->vec[i] = u;
    }
  }
}

/* rng.go (*rngSource).Uint64 */
static uint64_t go_rand_uint64(go_rand *r) {
  r->tap--;
  if (r->tap < 0) r->tap += RNG_LEN;
  r->feed--;
  if (r->feed < 0) r->feed += RNG_LEN;
  int64_t x = (int64_t)((uint64_t)r->vec[r->feed] + (uint64_t)r->vec[r->tap]);
  r->vec[r->feed] = x;
  return (uint64_t)x;
}

static int64_t go_rand_int63(go_rand *r) { return (int64_t)(go_rand_uint64(r) & 0x7fffffffffffffffULL); }
static int32_t go_rand_int31(go_rand *r) { return (int32_t)(go_rand_int63(r) >> 32); }
static uint32_t go_rand_uint32(go_rand *r) { return (uint32_t)(go_rand_int63(r) >> 31); }

/* rand.go (*Rand).Int31n / Intn (n <= 1<<31-1 always here) */
static int go_rand_intn(go_rand *r, int n) {
  if ((n & (n - 1)) == 0) return go_rand_int31(r) & (n - 1);
  int32_t max = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
  int32_t v = go_rand_int31(r);
  while (v > max) v = go_rand_int31(r);
  return v % n;
}

/* rand.go (*Rand).int31n — Lemire's method, used by Shuffle */
static int32_t go_rand_int31n_lemire(go_rand *r, int32_t n) {
  uint32_t v = go_rand_uint32(r);
  uint64_t prod = (uint64_t)v * (uint64_t)n;
  uint32_t low = (uint32_t)prod;
  if (low < (uint32_t)n) {
    uint32_t thresh = (uint32_t)(-n) % (uint32_t)n;
    while (low < thresh) {
      v = go_rand_uint32(r);
      prod = (uint64_t)v * (uint64_t)n;
      low = (uint32_t)prod;
    }
  }
  return (int32_t)(prod >> 32);
}

/* ======================================================================== */
/* Game model (internal/game/core/board.go:7-18, internal/game/state.go:7-34) */
/* ======================================================================== */

typedef struct {
  int owner, army, type;
  uint32_t vis; /* VisibleBitfield */
} tile_t;

typedef struct {
  int alive, army_count, general_idx;
  int *owned; /* OwnedTiles, ordered like the Go slice */
  int n_owned;
} player_t;

typedef struct {
  int W, H, N, P;
  tile_t *T;
  player_t *pl;
  uint8_t *changed; /* ChangedTiles as a flag set; iterated in ascending index order */
  int n_changed;
  uint8_t *vchg; /* VisibilityChangedTiles */
  int n_vchg;
  int turn, game_over, fog;
  int step_error;
  /* snapshot of the pre-step state = GameState.Clone() (turn_processor.go:116-121) */
  int *prev_owner, *prev_army;
  int prev_valid;
  float reward[GRL_MAX_PLAYERS];
  int action_index[GRL_MAX_PLAYERS];
} game_t;

struct grlo_env {
  grl_config cfg;
  int N;
  game_t *g;
  uint64_t stats[4];
  int nthreads;
};
typedef struct grlo_env grlo_env;

static __thread char g_err[256];

static void set_add(uint8_t *set, int *n, int idx) {
  if (!set[idx]) {
    set[idx] = 1;
    (*n)++;
  }
}
static void set_clear(uint8_t *set, int *n, int N) {
  memset(set, 0, (size_t)(unsigned)N);
  *n = 0;
}

/* ---- mapgen (internal/game/mapgen/generator.go) ------------------------- */

typedef struct {
  int W, H, players, city_ratio, city_start_army, spacing, veins, min_vein, max_vein;
} mapcfg_t;

/* generator.go:25-47 DefaultMapConfig */
static mapcfg_t default_map_config(const grl_config *c) {
  mapcfg_t m;
  int w = c->width, h = c->height;
  int spacing = c->min_general_spacing;
  int max_feasible = w / 2 + h / 2;
  if (spacing > max_feasible) spacing = max_feasible;
  m.W = w;
  m.H = h;
  m.players = c->num_players;
  m.city_ratio = c->city_ratio;
  m.city_start_army = c->city_start_army;
  m.spacing = spacing;
  m.veins = (w * h) / 50;
  m.min_vein = 3;
  m.max_vein = w / 4;
  return m;
}

static int iabs(int v) { return v < 0 ? -v : v; }

/* generator.go:77-142 placeMountains */
static void place_mountains(const mapcfg_t *m, go_rand *rng, tile_t *T) {
  int W = m->W, H = m->H;
  for (int v = 0; v < m->veins; v++) {
    int sx0 = -1, sy0 = -1, found = 0;
    for (int a = 0; a < 100; a++) {
      int sx = go_rand_intn(rng, W);
      int sy = go_rand_intn(rng, H);
      int s = sy * W + sx;
      if (T[s].type == GRL_TILE_NORMAL && T[s].owner == GRL_NEUTRAL) {
        sx0 = sx;
        sy0 = sy;
        found = 1;
        break;
      }
    }
    if (!found) continue;
    int cx = sx0, cy = sy0;
    T[cy * W + cx].type = GRL_TILE_MOUNTAIN;
    T[cy * W + cx].army = 0;
    int len = m->min_vein;
    if (m->max_vein > m->min_vein) len += go_rand_intn(rng, m->max_vein - m->min_vein + 1);
    for (int i = 1; i < len; i++) {
      int dx[4] = {0, 1, 0, -1};
      int dy[4] = {-1, 0, 1, 0};
      /* rand.Shuffle(4, swap) */
      for (int k = 3; k > 0; k--) {
        int j = (int)go_rand_int31n_lemire(rng, (int32_t)(k + 1));
        int t = dx[k];
        dx[k] = dx[j];
        dx[j] = t;
        t = dy[k];
        dy[k] = dy[j];
        dy[j] = t;
      }
      int cand_x[4], cand_y[4], nc = 0;
      for (int j = 0; j < 4; j++) {
        int nx = cx + dx[j], ny = cy + dy[j];
        if (nx >= 0 && nx < W && ny >= 0 && ny < H) {
          int n = ny * W + nx;
          if (T[n].type == GRL_TILE_NORMAL && T[n].owner == GRL_NEUTRAL) {
            cand_x[nc] = nx;
            cand_y[nc] = ny;
            nc++;
          }
        }
      }
      if (nc == 0) break;
      int pick = go_rand_intn(rng, nc);
      cx = cand_x[pick];
      cy = cand_y[pick];
      T[cy * W + cx].type = GRL_TILE_MOUNTAIN;
      T[cy * W + cx].army = 0;
    }
  }
}

/* generator.go:144-164 placeCities */
static void place_cities(const mapcfg_t *m, go_rand *rng, tile_t *T) {
  int want = (m->W * m->H) / m->city_ratio;
  int placed = 0, attempts = 0, max_attempts = want * 20;
  while (placed < want && attempts < max_attempts) {
    int x = go_rand_intn(rng, m->W);
    int y = go_rand_intn(rng, m->H);
    tile_t *t = &T[y * m->W + x];
    if (t->owner == GRL_NEUTRAL && t->type == GRL_TILE_NORMAL) {
      t->type = GRL_TILE_CITY;
      t->army = m->city_start_army;
      placed++;
    }
    attempts++;
  }
}

/* generator.go:186-253 findGeneralLocation; returns index or -1 */
static int find_general_location(const mapcfg_t *m, go_rand *rng, const tile_t *T, const int *existing,
                                 int n_existing) {
  int W = m->W, H = m->H;
  int max_attempts = W * H;
  for (int a = 0; a < max_attempts; a++) {
    int x = go_rand_intn(rng, W);
    int y = go_rand_intn(rng, H);
    int idx = y * W + x;
    if (T[idx].owner != GRL_NEUTRAL || T[idx].type != GRL_TILE_NORMAL) continue;
    int ok = 1;
    for (int e = 0; e < n_existing; e++) {
      int ox = existing[e] % W, oy = existing[e] / W;
      if (iabs(x - ox) + iabs(y - oy) < m->spacing) {
        ok = 0;
        break;
      }
    }
    if (ok) return idx;
  }
  for (int idx = 0; idx < W * H; idx++) {
    if (T[idx].owner == GRL_NEUTRAL && T[idx].type == GRL_TILE_NORMAL) {
      int x = idx % W, y = idx / W, ok = 1;
      for (int e = 0; e < n_existing; e++) {
        int ox = existing[e] % W, oy = existing[e] / W;
        if (iabs(x - ox) + iabs(y - oy) < m->spacing) {
          ok = 0;
          break;
        }
      }
      if (ok) return idx;
    }
  }
  return -1;
}

/* generator.go:64-75 GenerateMap */
static int generate_map(const mapcfg_t *m, go_rand *rng, tile_t *T) {
  int N = m->W * m->H;
  for (int i = 0; i < N; i++) { /* core.NewBoard, board.go:96-105 */
    T[i].owner = GRL_NEUTRAL;
    T[i].army = 0;
    T[i].type = GRL_TILE_NORMAL;
    T[i].vis = 0;
  }
  place_mountains(m, rng, T);
  place_cities(m, rng, T);
  int placed[GRL_MAX_PLAYERS];
  for (int pid = 0; pid < m->players; pid++) { /* generator.go:166-184 placeGenerals */
    int idx = find_general_location(m, rng, T, placed, pid);
    if (idx < 0) return GRL_ERR_MAPGEN;
    T[idx].owner = pid;
    T[idx].army = 2;
    T[idx].type = GRL_TILE_GENERAL;
    placed[pid] = idx;
  }
  return GRL_OK;
}

/* ---- stats (internal/game/stats.go) -------------------------------------- */

/* stats.go:33-63 performFullStatsUpdate */
static void stats_full(game_t *g) {
  for (int p = 0; p < g->P; p++) {
    g->pl[p].army_count = 0;
    g->pl[p].general_idx = -1;
    g->pl[p].n_owned = 0;
  }
  for (int idx = 0; idx < g->N; idx++) {
    tile_t *t = &g->T[idx];
    if (t->owner >= 0 && t->owner < g->P) {
      player_t *p = &g->pl[t->owner];
      p->army_count += t->army;
      p->owned[p->n_owned++] = idx;
      if (t->type == GRL_TILE_GENERAL) p->general_idx = idx;
    }
  }
  for (int p = 0; p < g->P; p++) g->pl[p].alive = g->pl[p].general_idx != -1;
}

/* stats.go:66-144 performIncrementalStatsUpdate.  Go iterates the
 * tempTileOwnership map in runtime-random order; ascending order is used here, which
 * fixes list ORDER only.  GeneralIdx is "the last general-type tile seen", which for a
 * player holding two or more general-type tiles depends on that random order in Go
 * (parity unpinned): the canonical choice here is the HIGHEST index, which
 * is also what the full rebuild's ascending scan yields. */
static void stats_incremental(game_t *g) {
  for (int p = 0; p < g->P; p++) {
    player_t *pl = &g->pl[p];
    pl->army_count = 0;
    pl->general_idx = -1;
    int n_new = 0;
    for (int k = 0; k < pl->n_owned; k++) {
      int idx = pl->owned[k];
      if (g->T[idx].owner == p) {
        pl->army_count += g->T[idx].army;
        pl->owned[n_new++] = idx; /* newOwnedTiles reuses the backing array */
        if (g->T[idx].type == GRL_TILE_GENERAL && idx > pl->general_idx) pl->general_idx = idx;
      }
    }
    for (int idx = 0; idx < g->N; idx++) {
      if (!g->changed[idx] || g->T[idx].owner != p) continue;
      int found = 0;
      for (int k = 0; k < n_new; k++)
        if (pl->owned[k] == idx) {
          found = 1;
          break;
        }
      if (!found) {
        pl->army_count += g->T[idx].army;
        pl->owned[n_new++] = idx;
        if (g->T[idx].type == GRL_TILE_GENERAL && idx > pl->general_idx) pl->general_idx = idx;
      }
    }
    pl->n_owned = n_new;
  }
  for (int p = 0; p < g->P; p++) g->pl[p].alive = g->pl[p].general_idx != -1;
}

/* stats.go:8-30 updatePlayerStats */
static void update_player_stats(game_t *g) {
  if (g->n_changed == 0 && g->turn > 0) return;
  int threshold = g->N / 5;
  if (g->turn == 0 || g->n_changed > threshold) {
    stats_full(g);
    return;
  }
  stats_incremental(g);
}

/* ---- fog of war (internal/game/visibility_optimized.go) ------------------ */

static void set_visibility_around(game_t *g, int idx, uint32_t bit) { /* :118-128 */
  int x = idx % g->W, y = idx / g->W;
  for (int dy = -1; dy <= 1; dy++)
    for (int dx = -1; dx <= 1; dx++) {
      int nx = x + dx, ny = y + dy;
      if (nx >= 0 && nx < g->W && ny >= 0 && ny < g->H) g->T[ny * g->W + nx].vis |= bit;
    }
}

static void fog_full(game_t *g) { /* :33-53 */
  for (int i = 0; i < g->N; i++) g->T[i].vis = 0;
  for (int p = 0; p < g->P; p++) {
    if (!g->pl[p].alive) continue;
    for (int k = 0; k < g->pl[p].n_owned; k++) set_visibility_around(g, g->pl[p].owned[k], 1u << p);
  }
}

static void fog_incremental(game_t *g) { /* :56-97 */
  int affected[GRL_MAX_PLAYERS] = {0};
  for (int idx = 0; idx < g->N; idx++) { /* collectAffectedPlayersOptimized :100-115 */
    if (!g->vchg[idx]) continue;
    int x = idx % g->W, y = idx / g->W;
    for (int dx = -2; dx <= 2; dx++)
      for (int dy = -2; dy <= 2; dy++) {
        int nx = x + dx, ny = y + dy;
        if (nx >= 0 && nx < g->W && ny >= 0 && ny < g->H) {
          int owner = g->T[ny * g->W + nx].owner;
          if (owner >= 0 && owner < g->P) affected[owner] = 1;
        }
      }
  }
  uint32_t all = 0; /* clearVisibilityAroundOptimized :131-149 */
  for (int p = 0; p < g->P; p++) all |= 1u << p;
  for (int idx = 0; idx < g->N; idx++) {
    if (!g->vchg[idx]) continue;
    int x = idx % g->W, y = idx / g->W;
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        int nx = x + dx, ny = y + dy;
        if (nx >= 0 && nx < g->W && ny >= 0 && ny < g->H) g->T[ny * g->W + nx].vis &= ~all;
      }
  }
  for (int p = 0; p < g->P; p++) { /* :84-94 */
    if (!affected[p] || !g->pl[p].alive) continue;
    for (int k = 0; k < g->pl[p].n_owned; k++) set_visibility_around(g, g->pl[p].owned[k], 1u << p);
  }
}

static void update_fog_of_war(game_t *g) { /* :16-30 */
  if (!g->fog) return;
  int threshold = g->N / 10;
  if (g->turn == 0 || g->n_vchg > threshold) {
    fog_full(g);
    return;
  }
  fog_incremental(g);
}

/* ---- win condition (internal/game/rules/win_conditions.go:21-57) --------- */

static void check_game_over_ex(const game_t *g, int *over, int *winner) {
  int alive = 0, last = 0;
  for (int p = 0; p < g->P; p++)
    if (g->pl[p].alive) {
      alive++;
      last = p;
    }
  int go = g->P > 1 ? (alive <= 1) : (alive == 0);
  *over = go;
  *winner = (go && alive == 1) ? last : -1;
}

/* engine.go:248-263 GetWinner */
static int engine_get_winner(const game_t *g) {
  if (!g->game_over) return -1;
  int over, winner;
  check_game_over_ex(g, &over, &winner);
  return winner;
}

/* ---- validation and movement (core/action.go:56-105, core/movement.go) --- */

static int validate_move(const game_t *g, int player, int fx, int fy, int tx, int ty) {
  if (!(fx >= 0 && fx < g->W && fy >= 0 && fy < g->H)) return GRL_STEP_INVALID_COORDINATES;
  if (!(tx >= 0 && tx < g->W && ty >= 0 && ty < g->H)) return GRL_STEP_INVALID_COORDINATES;
  if (fx == tx && fy == ty) return GRL_STEP_MOVE_TO_SELF;
  int dx = fx - tx, dy = fy - ty; /* coordinate.go:46-53 IsAdjacentTo */
  if (!((dx == 0 && (dy == 1 || dy == -1)) || (dy == 0 && (dx == 1 || dx == -1)))) return GRL_STEP_NOT_ADJACENT;
  const tile_t *from = &g->T[fy * g->W + fx];
  if (from->owner != player) return GRL_STEP_NOT_OWNED;
  if (from->army <= 1) return GRL_STEP_INSUFFICIENT_ARMY;
  if (g->T[ty * g->W + tx].type == GRL_TILE_MOUNTAIN) return GRL_STEP_TARGET_IS_MOUNTAIN;
  return GRL_STEP_OK;
}

typedef struct {
  int tile_type, capturer, prev_owner, idx;
} capture_t;

/* movement.go:23-89 ApplyMoveAction; returns error code; *cap_valid set on capture */
static int apply_move(game_t *g, const grl_action *a, capture_t *cap, int *cap_valid) {
  *cap_valid = 0;
  int err = validate_move(g, a->player_id, a->from_x, a->from_y, a->to_x, a->to_y);
  if (err) return err;
  int fi = a->from_y * g->W + a->from_x, ti = a->to_y * g->W + a->to_x;
  tile_t *from = &g->T[fi], *to = &g->T[ti];
  int orig_owner = to->owner;
  int moved;
  if (a->move_all)
    moved = from->army - 1;
  else {
    moved = from->army / 2;
    if (moved == 0) moved = 1;
  }
  from->army -= moved;
  set_add(g->changed, &g->n_changed, fi);
  set_add(g->changed, &g->n_changed, ti);
  if (to->owner == a->player_id) {
    to->army += moved;
    return GRL_STEP_OK;
  }
  if (moved > to->army) {
    to->owner = a->player_id;
    to->army = moved - to->army;
    cap->tile_type = to->type;
    cap->capturer = a->player_id;
    cap->prev_owner = orig_owner;
    cap->idx = ti;
    *cap_valid = 1;
  } else {
    to->army -= moved;
  }
  return GRL_STEP_OK;
}

/* ---- production (internal/game/production_manager.go:26-101) ------------- */

static void process_production(game_t *g, const grl_config *c) {
  int grow = (g->turn % c->normal_growth_interval) == 0;
  for (int p = 0; p < g->P; p++) {
    if (!g->pl[p].alive) continue;
    for (int k = 0; k < g->pl[p].n_owned; k++) {
      int idx = g->pl[p].owned[k];
      tile_t *t = &g->T[idx];
      int prod = 0;
      switch (t->type) {
        case GRL_TILE_GENERAL:
          prod = c->production_general;
          t->army += prod;
          break;
        case GRL_TILE_CITY:
          prod = c->production_city;
          t->army += prod;
          break;
        case GRL_TILE_NORMAL:
          if (grow) {
            prod = c->production_normal;
            t->army += prod;
          }
          break;
        default:
          break;
      }
      if (prod > 0) set_add(g->changed, &g->n_changed, idx);
    }
  }
}

/* ---- reward (internal/experience/rewards.go:45-175) ----------------------- */

static float calc_reward(const game_t *g, const grl_reward_config *rc, int p) {
  volatile float reward = 0.0f; /* volatile: one rounding per Go statement, no contraction */
  /* state.go:73-100 GameState.IsGameOver / GetWinner on the current state */
  int alive = 0, alive_id = -1;
  for (int q = 0; q < g->P; q++)
    if (g->pl[q].alive) {
      alive++;
      alive_id = q;
    }
  if (alive <= 1) {
    int winner = (alive == 1) ? alive_id : -1;
    if (winner == p) return rc->win_game;
    if (winner != -1) return rc->lose_game;
  }
  int prev_terr = 0, curr_terr = 0, prev_army = 0, curr_army = 0;
  int cg = 0, cl = 0, gg = 0, gl = 0, own = 0, enemy = 0;
  for (int i = 0; i < g->N; i++) {
    int po = g->prev_owner[i], co = g->T[i].owner;
    if (po == p) {
      prev_terr++;
      prev_army += g->prev_army[i];
    }
    if (co == p) {
      curr_terr++;
      curr_army += g->T[i].army;
    }
    if (g->T[i].type == GRL_TILE_CITY) {
      if (po != p && co == p) cg++;
      if (po == p && co != p) cl++;
    }
    if (g->T[i].type == GRL_TILE_GENERAL) {
      if (po != p && po >= 0 && co == p) gg++;
      if (po == p && co != p) gl++;
    }
    if (co == p)
      own += g->T[i].army;
    else if (co >= 0)
      enemy += g->T[i].army;
  }
  volatile float term;
  term = (float)(curr_terr - prev_terr) * rc->territory_gained;
  reward = reward + term;
  term = (float)(curr_army - prev_army) * rc->army_gained;
  reward = reward + term;
  term = (float)cg * rc->capture_city;
  reward = reward + term;
  term = (float)cl * rc->lose_city;
  reward = reward + term;
  term = (float)gg * rc->capture_general;
  reward = reward + term;
  term = (float)gl * rc->lose_general;
  reward = reward + term;
  volatile float adv = 0.0f;
  int total = own + enemy;
  if (total != 0) adv = (float)(own - enemy) / (float)total;
  term = adv * rc->army_advantage;
  reward = reward + term;
  return reward;
}

/* serializer.go:179-198 ActionToIndex (dirs Up=0, Down=1, Left=2, Right=3) */
static int action_to_index(const grl_action *a, int W) {
  int dir = 0;
  int dx = a->to_x - a->from_x, dy = a->to_y - a->from_y;
  if (dy == -1 && dx == 0)
    dir = 0;
  else if (dy == 1 && dx == 0)
    dir = 1;
  else if (dy == 0 && dx == -1)
    dir = 2;
  else if (dy == 0 && dx == 1)
    dir = 3;
  return (a->from_y * W + a->from_x) * 4 + dir;
}

/* ---- the turn (internal/game/turn_processor.go:29-77) --------------------- */

/* counters: [0] steps executed, [1] error turns, [2] games finished, [3] rejected */
static void process_turn(game_t *g, const grl_config *c, const grl_action *acts, int n_slots, uint64_t cnt[4]) {
  for (int p = 0; p < g->P; p++) {
    g->reward[p] = 0.0f;
    g->action_index[p] = -1;
  }
  /* validateGameState :95-113 */
  if (g->game_over) {
    g->step_error = GRL_STEP_GAME_OVER;
    cnt[3]++;
    return;
  }
  g->step_error = GRL_STEP_OK;
  cnt[0]++;
  /* captureStateForExperience :116-121 */
  for (int i = 0; i < g->N; i++) {
    g->prev_owner[i] = g->T[i].owner;
    g->prev_army[i] = g->T[i].army;
  }
  /* initializeTurn :124-135 */
  g->turn++;
  update_fog_of_war(g);
  set_clear(g->changed, &g->n_changed, g->N);
  set_clear(g->vchg, &g->n_vchg, g->N);

  /* processActions engine.go:80-115 -> ActionProcessor.ProcessActions action_processor.go:36-99 */
  grl_action sorted[GRL_MAX_ACTIONS];
  int n = 0;
  for (int s = 0; s < n_slots; s++)
    if (acts && acts[s].present) sorted[n++] = acts[s];
  for (int i = 1; i < n; i++) { /* sort.Slice, n <= 12: insertion sort (stable) */
    grl_action key = sorted[i];
    int j = i - 1;
    while (j >= 0 && sorted[j].player_id > key.player_id) {
      sorted[j + 1] = sorted[j];
      j--;
    }
    sorted[j + 1] = key;
  }
  int first_error = GRL_STEP_OK;
  capture_t caps[GRL_MAX_ACTIONS];
  int n_caps = 0;
  int alive_at_start[GRL_MAX_PLAYERS];
  for (int p = 0; p < g->P; p++) alive_at_start[p] = g->pl[p].alive;
  for (int i = 0; i < n; i++) {
    int pid = sorted[i].player_id;
    if (pid < 0 || pid >= g->P || !alive_at_start[pid]) continue; /* :56-60 */
    capture_t cap;
    int cap_valid;
    int err = apply_move(g, &sorted[i], &cap, &cap_valid);
    if (err) {
      if (first_error == GRL_STEP_OK) first_error = err; /* :66-77 */
      continue;
    }
    if (cap_valid) {
      caps[n_caps++] = cap;
      set_add(g->vchg, &g->n_vchg, cap.idx); /* :78-87, engine.go:96-98 */
    }
  }
  /* core.ProcessCaptures movement.go:100-118 + handleEliminationsAndTileTurnover engine.go:118-152 */
  if (n_caps > 0) {
    int processed[GRL_MAX_PLAYERS] = {0};
    int n_orders = 0;
    int ord_el[GRL_MAX_ACTIONS], ord_new[GRL_MAX_ACTIONS];
    for (int i = 0; i < n_caps; i++) {
      if (caps[i].tile_type == GRL_TILE_GENERAL && caps[i].prev_owner != GRL_NEUTRAL &&
          caps[i].prev_owner != caps[i].capturer && !processed[caps[i].prev_owner]) {
        ord_el[n_orders] = caps[i].prev_owner;
        ord_new[n_orders] = caps[i].capturer;
        n_orders++;
        processed[caps[i].prev_owner] = 1;
      }
    }
    if (n_orders > 0) {
      for (int o = 0; o < n_orders; o++) {
        player_t *el = &g->pl[ord_el[o]];
        for (int k = 0; k < el->n_owned; k++) {
          int idx = el->owned[k];
          if (g->T[idx].owner == ord_el[o]) {
            g->T[idx].owner = ord_new[o];
            set_add(g->changed, &g->n_changed, idx);
            set_add(g->vchg, &g->n_vchg, idx);
          }
        }
        el->alive = 0;
        el->general_idx = -1;
      }
      update_player_stats(g); /* engine.go:107 */
    }
  }
  if (first_error != GRL_STEP_OK) { /* engine.go:111-113, turn_processor.go:55-57 */
    g->step_error = first_error;
    cnt[1]++;
    for (int p = 0; p < g->P; p++) g->reward[p] = calc_reward(g, &c->reward, p);
    return;
  }
  process_production(g, c); /* :150-158 */
  update_player_stats(g);   /* :170-179 */
  {
    int over, winner;
    check_game_over_ex(g, &over, &winner); /* engine.go:160-194 */
    if (over && !g->game_over) cnt[2]++;
    g->game_over = over;
  }
  /* collectExperiences :182-217 -> SimpleCollector.OnStateTransition collector.go:30-98 */
  for (int p = 0; p < g->P; p++) g->reward[p] = calc_reward(g, &c->reward, p);
  for (int s = 0; s < n_slots; s++) {
    if (!acts || !acts[s].present) continue;
    int pid = acts[s].player_id;
    if (pid < 0 || pid >= g->P) continue;
    g->action_index[pid] = action_to_index(&acts[s], g->W); /* last one wins, like the Go map */
  }
}

/* ---- read-outs ------------------------------------------------------------ */

/* experience/serializer.go:37-109 StateToTensor */
static void state_to_tensor(const game_t *g, int p, float *out) {
  int N = g->N;
  memset(out, 0, sizeof(float) * (size_t)GRL_OBS_CHANNELS * (size_t)N);
  for (int i = 0; i < N; i++) {
    const tile_t *t = &g->T[i];
    int visible = !g->fog || ((t->vis >> p) & 1u);
    if (visible) out[7 * N + i] = 1.0f;
    if (!visible) {
      out[8 * N + i] = 1.0f;
      continue;
    }
    if (t->type == GRL_TILE_MOUNTAIN) {
      out[6 * N + i] = 1.0f;
      continue;
    }
    if (t->type == GRL_TILE_CITY || t->type == GRL_TILE_GENERAL) out[5 * N + i] = 1.0f;
    if (t->owner == p) {
      if (t->army > 0) {
        volatile float v = (float)t->army / 1000.0f;
        if (v > 1.0f) v = 1.0f;
        out[0 * N + i] = v;
      }
      out[2 * N + i] = 1.0f;
    } else if (t->owner >= 0) {
      if (t->army > 0) {
        volatile float v = (float)t->army / 1000.0f;
        if (v > 1.0f) v = 1.0f;
        out[1 * N + i] = v;
      }
      out[3 * N + i] = 1.0f;
    } else {
      out[4 * N + i] = 1.0f;
    }
  }
}

/* rules/legal_moves.go:19-73 via engine.go:271-280; dirs up,right,down,left */
static void engine_mask(const game_t *g, int p, uint8_t *mask) {
  memset(mask, 0, (size_t)g->N * 4);
  if (p < 0 || p >= g->P) return;
  if (!g->pl[p].alive) return;
  static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0};
  for (int k = 0; k < g->pl[p].n_owned; k++) {
    int idx = g->pl[p].owned[k];
    const tile_t *t = &g->T[idx];
    if (t->owner != p || t->army <= 1) continue;
    int x = idx % g->W, y = idx / g->W;
    for (int d = 0; d < 4; d++)
      if (validate_move(g, p, x, y, x + DX[d], y + DY[d]) == GRL_STEP_OK) mask[(y * g->W + x) * 4 + d] = 1;
  }
}

/* experience/serializer.go:112-176 GenerateActionMask; dirs up,down,left,right */
static void serializer_mask(const game_t *g, int p, uint8_t *mask) {
  int W = g->W, H = g->H;
  memset(mask, 0, (size_t)g->N * 4);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const tile_t *t = &g->T[y * W + x];
      if (t->owner != p || t->army < 2) continue;
      int base = (y * W + x) * 4;
      if (y > 0 && g->T[(y - 1) * W + x].type != GRL_TILE_MOUNTAIN) mask[base + 0] = 1;
      if (y < H - 1 && g->T[(y + 1) * W + x].type != GRL_TILE_MOUNTAIN) mask[base + 1] = 1;
      if (x > 0 && g->T[y * W + x - 1].type != GRL_TILE_MOUNTAIN) mask[base + 2] = 1;
      if (x < W - 1 && g->T[y * W + x + 1].type != GRL_TILE_MOUNTAIN) mask[base + 3] = 1;
    }
}

/* ---- synthetic policy: counter-based, replayable --------------- */

static uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

static uint64_t policy_draw(uint64_t seed, uint64_t env, uint64_t turn, uint64_t player) {
  uint64_t x = mix64(seed ^ (env * 0xD6E8FEB86659FD93ULL));
  x = mix64(x ^ (turn * 0xA0761D6478BD642FULL) ^ (player << 56));
  return x;
}

static void sample_actions(const game_t *g, uint64_t seed, uint64_t env_global, grl_action *slots, int n_slots,
                           uint8_t *scratch) {
  for (int s = 0; s < n_slots; s++) memset(&slots[s], 0, sizeof(grl_action));
  if (g->game_over) return;
  for (int p = 0; p < g->P && p < n_slots; p++) {
    engine_mask(g, p, scratch);
    int cnt = 0;
    for (int i = 0; i < g->N * 4; i++) cnt += scratch[i];
    if (cnt == 0) continue;
    uint64_t r = policy_draw(seed, env_global, (uint64_t)g->turn, (uint64_t)p);
    int k = (int)((uint32_t)r % (uint32_t)cnt);
    int pick = -1;
    for (int i = 0; i < g->N * 4; i++)
      if (scratch[i] && k-- == 0) {
        pick = i;
        break;
      }
    static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0};
    int tile = pick / 4, d = pick % 4;
    grl_action *a = &slots[p];
    a->player_id = (int8_t)p;
    a->from_x = (int8_t)(tile % g->W);
    a->from_y = (int8_t)(tile / g->W);
    a->to_x = (int8_t)(a->from_x + DX[d]);
    a->to_y = (int8_t)(a->from_y + DY[d]);
    a->move_all = (uint8_t)((r >> 32) & 1u);
    a->present = 1;
  }
}

/* ---- digests (same definition as the CUDA library; DESIGN.md "digests") ---- */

static uint64_t state_hash(const game_t *g) {
  uint64_t h = 0;
  for (int i = 0; i < g->N; i++) {
    uint64_t lists = 0;
    for (int p = 0; p < g->P; p++)
      for (int k = 0; k < g->pl[p].n_owned; k++)
        if (g->pl[p].owned[k] == i) lists |= 1ULL << p;
    const tile_t *t = &g->T[i];
    uint64_t pack = (uint64_t)(t->owner + 1) | ((uint64_t)t->type << 4) | ((uint64_t)g->changed[i] << 6) |
                    ((uint64_t)g->vchg[i] << 7) | ((uint64_t)(t->vis & 0xFFu) << 8) | (lists << 16) |
                    ((uint64_t)(uint32_t)t->army << 24);
    h += mix64(pack ^ ((uint64_t)(i + 1) * 0xD6E8FEB86659FD93ULL));
  }
  uint64_t alive = 0;
  for (int p = 0; p < g->P; p++) alive |= (uint64_t)(g->pl[p].alive != 0) << p;
  h += mix64(0x1000000000ULL + (uint64_t)(uint32_t)g->turn);
  h += mix64(0x2000000000ULL + ((uint64_t)(g->game_over != 0)) + (alive << 8));
  for (int p = 0; p < g->P; p++)
    h += mix64(0x3000000000ULL + ((uint64_t)p << 40) + (uint64_t)(uint32_t)g->pl[p].army_count);
  return h;
}

static uint64_t row_hash(const uint32_t *w, size_t n) {
  uint64_t h = 0;
  for (size_t i = 0; i < n; i++) h += mix64((uint64_t)w[i] ^ ((uint64_t)(i + 1) * 0xD6E8FEB86659FD93ULL));
  return h;
}

/* ======================================================================== */
/* Exported ABI (grlo_ prefix)                                               */
/* ======================================================================== */

int grlo_abi_version(void) { return GRL_ABI_VERSION; }

const char *grlo_status_string(int s) {
  switch (s) {
    case GRL_OK: return "ok";
    case GRL_ERR_INVALID_ARG: return "invalid argument";
    case GRL_ERR_CUDA: return "cuda error";
    case GRL_ERR_NOMEM: return "out of memory";
    case GRL_ERR_MAPGEN: return "map generation failed";
    case GRL_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown";
  }
}
const char *grlo_last_error(void) { return g_err; }

int grlo_default_config(grl_config *c) {
  if (!c) return GRL_ERR_INVALID_ARG;
  memset(c, 0, sizeof(*c));
  c->num_envs = 1;
  c->width = 20;
  c->height = 20;
  c->num_players = 2;
  c->max_actions = 2;
  c->fog_of_war = 1;
  c->city_ratio = 20; /* internal/config/config.go:198-209 */
  c->city_start_army = 40;
  c->min_general_spacing = 5;
  c->production_general = 1;
  c->production_city = 1;
  c->production_normal = 1;
  c->normal_growth_interval = 25;
  c->reward.win_game = 1.0f; /* experience/rewards.go:23-37 */
  c->reward.lose_game = -1.0f;
  c->reward.capture_city = 0.1f;
  c->reward.lose_city = -0.1f;
  c->reward.capture_general = 0.5f;
  c->reward.lose_general = -0.5f;
  c->reward.territory_gained = 0.01f;
  c->reward.territory_lost = -0.01f;
  c->reward.army_gained = 0.001f;
  c->reward.army_lost = -0.001f;
  c->reward.army_advantage = 0.05f;
  return GRL_OK;
}

static int check_config(const grl_config *c) {
  if (!c || c->num_envs < 1 || c->width < 1 || c->width > GRL_MAX_DIM || c->height < 1 || c->height > GRL_MAX_DIM ||
      c->num_players < 1 || c->num_players > GRL_MAX_PLAYERS || c->max_actions < 1 ||
      c->max_actions > GRL_MAX_ACTIONS || c->city_ratio < 1 || c->normal_growth_interval < 1) {
    snprintf(g_err, sizeof g_err, "bad config");
    return GRL_ERR_INVALID_ARG;
  }
  return GRL_OK;
}

static void game_free(game_t *g) {
  if (g->pl)
    for (int p = 0; p < g->P; p++) free(g->pl[p].owned);
  free(g->T);
  free(g->pl);
  free(g->changed);
  free(g->vchg);
  free(g->prev_owner);
  free(g->prev_army);
}

int grlo_create(const grl_config *c, grlo_env **out) {
  int st = check_config(c);
  if (st) return st;
  if (!out) return GRL_ERR_INVALID_ARG;
  grlo_env *e = (grlo_env *)calloc(1, sizeof(*e));
  if (!e) return GRL_ERR_NOMEM;
  e->cfg = *c;
  e->N = c->width * c->height;
  e->nthreads = c->host_threads > 0 ? c->host_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (e->nthreads < 1) e->nthreads = 1;
  e->g = (game_t *)calloc((size_t)c->num_envs, sizeof(game_t));
  if (!e->g) {
    free(e);
    return GRL_ERR_NOMEM;
  }
  for (int b = 0; b < c->num_envs; b++) {
    game_t *g = &e->g[b];
    g->W = c->width;
    g->H = c->height;
    g->N = e->N;
    g->P = c->num_players;
    g->fog = c->fog_of_war != 0;
    g->T = (tile_t *)calloc((size_t)e->N, sizeof(tile_t));
    g->pl = (player_t *)calloc((size_t)g->P, sizeof(player_t));
    g->changed = (uint8_t *)calloc((size_t)e->N, 1);
    g->vchg = (uint8_t *)calloc((size_t)e->N, 1);
    g->prev_owner = (int *)calloc((size_t)e->N, sizeof(int));
    g->prev_army = (int *)calloc((size_t)e->N, sizeof(int));
    for (int p = 0; p < g->P; p++) g->pl[p].owned = (int *)calloc((size_t)e->N, sizeof(int));
    for (int i = 0; i < e->N; i++) g->T[i].owner = GRL_NEUTRAL;
    for (int p = 0; p < g->P; p++) {
      g->pl[p].general_idx = -1;
      g->action_index[p] = -1;
    }
    g->game_over = 1; /* not reset yet: stepping is rejected */
  }
  *out = e;
  return GRL_OK;
}

int grlo_destroy(grlo_env *e) {
  if (!e) return GRL_OK;
  for (int b = 0; b < e->cfg.num_envs; b++) game_free(&e->g[b]);
  free(e->g);
  free(e);
  return GRL_OK;
}

int grlo_sync(grlo_env *e) {
  (void)e;
  return GRL_OK;
}

int grlo_set_stream(grlo_env *e, void *stream) {
  (void)e;
  (void)stream;
  return GRL_OK;
}

int grlo_get_config(const grlo_env *e, grl_config *out) {
  if (!e || !out) return GRL_ERR_INVALID_ARG;
  *out = e->cfg;
  return GRL_OK;
}

/* engine_initializer.go:113-143,218-225: players alive, turn 0, full stats, full fog, game-over */
static void initial_setup(game_t *g) {
  g->turn = 0;
  g->game_over = 0;
  g->step_error = 0;
  set_clear(g->changed, &g->n_changed, g->N);
  set_clear(g->vchg, &g->n_vchg, g->N);
  for (int p = 0; p < g->P; p++) {
    g->pl[p].alive = 1;
    g->pl[p].n_owned = 0;
    g->reward[p] = 0.0f;
    g->action_index[p] = -1;
  }
  update_player_stats(g);
  update_fog_of_war(g);
  int over, winner;
  check_game_over_ex(g, &over, &winner);
  g->game_over = over;
  for (int i = 0; i < g->N; i++) {
    g->prev_owner[i] = g->T[i].owner;
    g->prev_army[i] = g->T[i].army;
  }
}

int grlo_mapgen(const grl_config *c, int64_t seed, int32_t *owner, int32_t *army, int32_t *type) {
  int st = check_config(c);
  if (st) return st;
  int N = c->width * c->height;
  tile_t *T = (tile_t *)calloc((size_t)N, sizeof(tile_t));
  go_rand *rng = (go_rand *)malloc(sizeof(go_rand));
  go_rand_seed(rng, seed);
  mapcfg_t m = default_map_config(c);
  st = generate_map(&m, rng, T);
  for (int i = 0; i < N; i++) {
    if (owner) owner[i] = T[i].owner;
    if (army) army[i] = T[i].army;
    if (type) type[i] = T[i].type;
  }
  free(rng);
  free(T);
  return st;
}

int grlo_reset_seeded(grlo_env *e, const int32_t *env_ids, int32_t n, const int64_t *seeds) {
  if (!e || !seeds || n < 0) return GRL_ERR_INVALID_ARG;
  go_rand *rng = (go_rand *)malloc(sizeof(go_rand));
  mapcfg_t m = default_map_config(&e->cfg);
  int rc = GRL_OK;
  for (int i = 0; i < n; i++) {
    int b = env_ids ? env_ids[i] : i;
    if (b < 0 || b >= e->cfg.num_envs) {
      rc = GRL_ERR_INVALID_ARG;
      break;
    }
    go_rand_seed(rng, seeds[i]);
    int st = generate_map(&m, rng, e->g[b].T);
    if (st) {
      rc = st;
      break;
    }
    initial_setup(&e->g[b]);
  }
  free(rng);
  return rc;
}

int grlo_reset_boards(grlo_env *e, const int32_t *env_ids, int32_t n, const int32_t *owner, const int32_t *army,
                      const int32_t *type) {
  if (!e || !owner || !army || !type || n < 0) return GRL_ERR_INVALID_ARG;
  for (int i = 0; i < n; i++) {
    int b = env_ids ? env_ids[i] : i;
    if (b < 0 || b >= e->cfg.num_envs) return GRL_ERR_INVALID_ARG;
    game_t *g = &e->g[b];
    for (int t = 0; t < e->N; t++) {
      g->T[t].owner = owner[(size_t)i * e->N + t];
      g->T[t].army = army[(size_t)i * e->N + t];
      g->T[t].type = type[(size_t)i * e->N + t];
      g->T[t].vis = 0;
    }
    initial_setup(g);
  }
  return GRL_OK;
}

/* ---- threaded batch drivers ---- */
typedef struct {
  grlo_env *e;
  const grl_action *actions;
  uint32_t flags;
  uint64_t seed;
  const grl_step_outputs *out;
  int do_step;
  int b0, b1;
  uint64_t cnt[4];
} job_t;

int32_t grlo_obs_packed_words(int32_t width, int32_t height, int32_t num_players) {
  if (width < 1 || width > GRL_MAX_DIM || height < 1 || height > GRL_MAX_DIM || num_players < 1 || num_players > GRL_MAX_PLAYERS)
    return 0;
  int N = width * height, NW = (N + 31) / 32, NA = (N + 7) & ~7;
  return ((2 * num_players + 2) * NW + NA / 2 + 3) & ~3;
}

/* the checker's own expansion of packed records: serializer.go:37-109 tile by tile from the record's planes */
int grlo_expand_obs(int32_t width, int32_t height, int32_t num_players, const uint32_t *packed, int32_t count, float *obs,
                    int32_t threads) {
  (void)threads;
  int RW = grlo_obs_packed_words(width, height, num_players);
  if (!packed || !obs || count < 0 || RW == 0) return GRL_ERR_INVALID_ARG;
  int N = width * height, P = num_players, NW = (N + 31) / 32;
  for (int c = 0; c < count; c++) {
    const uint32_t *rec = packed + (size_t)c * RW;
    const uint16_t *army = (const uint16_t *)(rec + (2 * P + 2) * NW);
    for (int p = 0; p < P; p++) {
      float *out = obs + ((size_t)c * P + p) * GRL_OBS_CHANNELS * N;
      memset(out, 0, sizeof(float) * (size_t)GRL_OBS_CHANNELS * N);
      for (int i = 0; i < N; i++) {
        int w = i >> 5;
        uint32_t bit = 1u << (i & 31);
        int owner = -1;
        for (int q = 0; q < P; q++)
          if (rec[q * NW + w] & bit) owner = q;
        if (!(rec[(P + p) * NW + w] & bit)) {
          out[8 * N + i] = 1.0f;
          continue;
        }
        out[7 * N + i] = 1.0f;
        if (rec[2 * P * NW + w] & bit) {
          out[6 * N + i] = 1.0f;
          continue;
        }
        if (rec[(2 * P + 1) * NW + w] & bit) out[5 * N + i] = 1.0f;
        volatile float v = (float)army[i] / 1000.0f;
        if (v > 1.0f) v = 1.0f;
        if (owner == p) {
          if (army[i] > 0) out[0 * N + i] = v;
          out[2 * N + i] = 1.0f;
        } else if (owner >= 0) {
          if (army[i] > 0) out[1 * N + i] = v;
          out[3 * N + i] = 1.0f;
        } else {
          out[4 * N + i] = 1.0f;
        }
      }
    }
  }
  return GRL_OK;
}

static void write_outputs(grlo_env *e, int b, const grl_step_outputs *out, uint8_t *scratch) {
  game_t *g = &e->g[b];
  int N = e->N, P = g->P;
  if (!out) return;
  if (out->obs)
    for (int p = 0; p < P; p++) state_to_tensor(g, p, out->obs + ((size_t)b * P + p) * GRL_OBS_CHANNELS * N);
  if (out->mask_bits) {
    int words = (4 * N + 31) / 32;
    for (int p = 0; p < P; p++) {
      uint32_t *dst = out->mask_bits + ((size_t)b * P + p) * words;
      memset(dst, 0, sizeof(uint32_t) * (size_t)words);
      engine_mask(g, p, scratch);
      for (int i = 0; i < 4 * N; i++)
        if (scratch[i]) dst[i >> 5] |= 1u << (i & 31);
    }
  }
  if (out->reward)
    for (int p = 0; p < P; p++) out->reward[(size_t)b * P + p] = g->reward[p];
  if (out->done) out->done[b] = (uint8_t)g->game_over;
  if (out->winner) out->winner[b] = (int8_t)engine_get_winner(g);
  if (out->step_error) out->step_error[b] = (uint8_t)g->step_error;
  if (out->action_index)
    for (int p = 0; p < P; p++) out->action_index[(size_t)b * P + p] = g->action_index[p];
  if (out->obs_packed) { /* include/grlcuda.h: own[P][NW] vis[P][NW] mountain[NW] city|general[NW] army u16[NA] */
    int NW = (N + 31) / 32, RW = grlo_obs_packed_words(e->cfg.width, e->cfg.height, P);
    uint32_t *rec = out->obs_packed + (size_t)b * RW;
    memset(rec, 0, sizeof(uint32_t) * (size_t)RW);
    uint16_t *army = (uint16_t *)(rec + (2 * P + 2) * NW);
    for (int i = 0; i < N; i++) {
      const tile_t *t = &g->T[i];
      uint32_t bit = 1u << (i & 31);
      if (t->owner >= 0) rec[t->owner * NW + (i >> 5)] |= bit;
      for (int p = 0; p < P; p++)
        if (!g->fog || ((t->vis >> p) & 1u)) rec[(P + p) * NW + (i >> 5)] |= bit;
      if (t->type == GRL_TILE_MOUNTAIN) rec[2 * P * NW + (i >> 5)] |= bit;
      if (t->type == GRL_TILE_CITY || t->type == GRL_TILE_GENERAL) rec[(2 * P + 1) * NW + (i >> 5)] |= bit;
      army[i] = (uint16_t)t->army;
    }
  }
}

static void *job_run(void *arg) {
  job_t *j = (job_t *)arg;
  grlo_env *e = j->e;
  int A = e->cfg.max_actions;
  uint8_t *scratch = (uint8_t *)malloc((size_t)e->N * 4);
  grl_action slots[GRL_MAX_ACTIONS];
  for (int b = j->b0; b < j->b1; b++) {
    game_t *g = &e->g[b];
    if (j->do_step) {
      const grl_action *acts = j->actions ? j->actions + (size_t)b * A : NULL;
      if (j->flags & GRL_STEP_FLAG_RANDOM_POLICY) {
        sample_actions(g, j->seed, (uint64_t)(e->cfg.env_id_base + b), slots, A, scratch);
        acts = slots;
      }
      /* GRL_ACTION_FLAG_SKIP_ENV: the game is still waiting at its turn barrier
         (game_manager.go:559-600) — Step is simply not called for it */
      if (acts && !(j->flags & GRL_STEP_FLAG_RANDOM_POLICY) && (acts[0].flags & GRL_ACTION_FLAG_SKIP_ENV)) {
        write_outputs(e, b, j->out, scratch);
        continue;
      }
      process_turn(g, &e->cfg, acts, A, j->cnt);
    }
    write_outputs(e, b, j->out, scratch);
  }
  free(scratch);
  return NULL;
}

static int run_jobs(grlo_env *e, const grl_action *actions, uint32_t flags, uint64_t seed,
                    const grl_step_outputs *out, int do_step) {
  int B = e->cfg.num_envs;
  int nt = e->nthreads;
  if (nt > B) nt = B;
  if (nt > 256) nt = 256;
  job_t jobs[256];
  pthread_t th[256];
  int per = (B + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    job_t *j = &jobs[t];
    memset(j, 0, sizeof(*j));
    j->e = e;
    j->actions = actions;
    j->flags = flags;
    j->seed = seed;
    j->out = out;
    j->do_step = do_step;
    j->b0 = t * per;
    j->b1 = j->b0 + per > B ? B : j->b0 + per;
    if (j->b0 > B) j->b0 = B;
  }
  if (nt == 1) {
    job_run(&jobs[0]);
  } else {
    for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  }
  for (int t = 0; t < nt; t++)
    for (int k = 0; k < 4; k++) e->stats[k] += jobs[t].cnt[k];
  return GRL_OK;
}

int grlo_step(grlo_env *e, const grl_action *actions, uint32_t flags, uint64_t policy_seed) {
  if (!e) return GRL_ERR_INVALID_ARG;
  return run_jobs(e, actions, flags, policy_seed, NULL, 1);
}

int grlo_step_fused(grlo_env *e, const grl_action *actions, uint32_t flags, uint64_t policy_seed,
                    const grl_step_outputs *out) {
  if (!e) return GRL_ERR_INVALID_ARG;
  return run_jobs(e, actions, flags, policy_seed, out, 1);
}

int grlo_observe(grlo_env *e, const grl_step_outputs *out) {
  if (!e || !out) return GRL_ERR_INVALID_ARG;
  return run_jobs(e, NULL, 0, 0, out, 0);
}

int grlo_mask(grlo_env *e, int variant, void *out) {
  if (!e || !out) return GRL_ERR_INVALID_ARG;
  int N = e->N, P = e->cfg.num_players, B = e->cfg.num_envs;
  int words = (4 * N + 31) / 32;
  uint8_t *scratch = (uint8_t *)malloc((size_t)N * 4);
  for (int b = 0; b < B; b++)
    for (int p = 0; p < P; p++) {
      game_t *g = &e->g[b];
      size_t bp = (size_t)b * P + p;
      switch (variant) {
        case GRL_MASK_ENGINE_URDL:
          engine_mask(g, p, (uint8_t *)out + bp * (size_t)N * 4);
          break;
        case GRL_MASK_SERIALIZER_UDLR:
          serializer_mask(g, p, (uint8_t *)out + bp * (size_t)N * 4);
          break;
        case GRL_MASK_ENGINE_URDL_BITS:
        case GRL_MASK_ENGINE_HALF_BITS: {
          int rep = variant == GRL_MASK_ENGINE_HALF_BITS ? 2 : 1;
          uint32_t *dst = (uint32_t *)out + bp * (size_t)words * rep;
          memset(dst, 0, sizeof(uint32_t) * (size_t)words * rep);
          engine_mask(g, p, scratch);
          for (int r = 0; r < rep; r++)
            for (int i = 0; i < 4 * N; i++)
              if (scratch[i]) dst[(size_t)r * words + (i >> 5)] |= 1u << (i & 31);
          break;
        }
        default:
          free(scratch);
          return GRL_ERR_INVALID_ARG;
      }
    }
  free(scratch);
  return GRL_OK;
}

/* visibility_optimized.go:166-195 ComputePlayerVisibilityOptimized */
int grlo_visibility(grlo_env *e, uint8_t *visible, uint8_t *fog) {
  if (!e) return GRL_ERR_INVALID_ARG;
  int N = e->N, P = e->cfg.num_players;
  for (int b = 0; b < e->cfg.num_envs; b++)
    for (int p = 0; p < P; p++) {
      game_t *g = &e->g[b];
      size_t base = ((size_t)b * P + p) * N;
      for (int i = 0; i < N; i++) {
        int vis, fg = 0;
        if (!g->fog) {
          vis = 1;
        } else {
          vis = (g->T[i].vis >> p) & 1u;
          if (!vis && g->T[i].type != GRL_TILE_NORMAL) fg = 1;
        }
        if (visible) visible[base + i] = (uint8_t)vis;
        if (fog) fog[base + i] = (uint8_t)fg;
      }
    }
  return GRL_OK;
}

/* generic parallel range over [0,n): used by the batch read-outs below (test-speed only) */
typedef struct {
  void (*fn)(void *ctx, int i0, int i1);
  void *ctx;
  int i0, i1;
} range_job_t;

static void *range_run(void *arg) {
  range_job_t *j = (range_job_t *)arg;
  j->fn(j->ctx, j->i0, j->i1);
  return NULL;
}

static void par_range(int nthreads, int n, void (*fn)(void *, int, int), void *ctx) {
  int nt = nthreads;
  if (nt > 256) nt = 256;
  if (nt > n / 64) nt = n / 64; /* not worth a thread below ~64 items each */
  if (nt <= 1) {
    fn(ctx, 0, n);
    return;
  }
  range_job_t jobs[256];
  pthread_t th[256];
  int per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    jobs[t].fn = fn;
    jobs[t].ctx = ctx;
    jobs[t].i0 = t * per > n ? n : t * per;
    jobs[t].i1 = (t + 1) * per > n ? n : (t + 1) * per;
    pthread_create(&th[t], NULL, range_run, &jobs[t]);
  }
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
}

typedef struct {
  grlo_env *e;
  uint64_t seed;
  grl_action *actions;
  uint64_t *out;
  const uint32_t *buf;
  size_t row_words;
} batch_ctx_t;

static void sample_range(void *c, int i0, int i1) {
  batch_ctx_t *x = (batch_ctx_t *)c;
  grlo_env *e = x->e;
  uint8_t *scratch = (uint8_t *)malloc((size_t)e->N * 4);
  int A = e->cfg.max_actions;
  for (int b = i0; b < i1; b++)
    sample_actions(&e->g[b], x->seed, (uint64_t)(e->cfg.env_id_base + b), x->actions + (size_t)b * A, A, scratch);
  free(scratch);
}

static void hash_range(void *c, int i0, int i1) {
  batch_ctx_t *x = (batch_ctx_t *)c;
  for (int b = i0; b < i1; b++) x->out[b] = state_hash(&x->e->g[b]);
}

static void row_hash_range(void *c, int i0, int i1) {
  batch_ctx_t *x = (batch_ctx_t *)c;
  for (int r = i0; r < i1; r++) x->out[r] = row_hash(x->buf + (size_t)r * x->row_words, x->row_words);
}

int grlo_sample_actions(grlo_env *e, uint64_t policy_seed, grl_action *actions) {
  if (!e || !actions) return GRL_ERR_INVALID_ARG;
  batch_ctx_t x = {e, policy_seed, actions, NULL, NULL, 0};
  par_range(e->nthreads, e->cfg.num_envs, sample_range, &x);
  return GRL_OK;
}

/* ---- generals_gym read-outs ------------------------------------------------------------
 * The proto view of one player (server.go:556-582 convertGameStateToProto), then
 * GeneralsEnv._get_observation / _get_valid_actions_mask (generals_env.py:291-387). */
typedef struct {
  int type, owner, army, visible;
} view_tile_t;

static void proto_view(const game_t *g, int p, view_tile_t *v) {
  for (int i = 0; i < g->N; i++) {
    int vis = g->fog ? (int)((g->T[i].vis >> p) & 1u) : 1;
    int fogt = g->fog && !vis && g->T[i].type != GRL_TILE_NORMAL;
    v[i].type = g->T[i].type;
    v[i].owner = g->T[i].owner;
    v[i].army = g->T[i].army;
    v[i].visible = vis;
    if (!vis && !fogt) { /* completely hidden */
      v[i].type = GRL_TILE_NORMAL;
      v[i].owner = -1;
      v[i].army = 0;
    } else if (fogt && !vis) { /* in fog: type shown, state hidden */
      v[i].owner = -1;
      v[i].army = 0;
    }
  }
}

static void gym_range(void *c, int i0, int i1);

typedef struct {
  grlo_env *e;
  int max_turns;
  const grl_gym_outputs *out;
} gym_ctx_t;

static void gym_range(void *c, int i0, int i1) {
  gym_ctx_t *x = (gym_ctx_t *)c;
  grlo_env *e = x->e;
  int N = e->N, P = e->cfg.num_players, W = e->cfg.width, H = e->cfg.height;
  view_tile_t *v = (view_tile_t *)malloc(sizeof(view_tile_t) * (size_t)N);
  static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0}; /* generals_env.py:369 */
  for (int b = i0; b < i1; b++) {
    const game_t *g = &e->g[b];
    for (int p = 0; p < P; p++) {
      proto_view(g, p, v);
      size_t bp = (size_t)b * P + p;
      if (x->out->obs) {
        float *o = x->out->obs + bp * GRL_GYM_CHANNELS * N;
        memset(o, 0, sizeof(float) * (size_t)GRL_GYM_CHANNELS * N);
        double tf = (double)g->turn / (double)x->max_turns;
        if (tf > 1.0) tf = 1.0;
        for (int i = 0; i < N; i++) {
          if (v[i].visible) o[0 * N + i] = 1.0f;
          if (v[i].owner == p)
            o[1 * N + i] = 0.5f;
          else if (v[i].owner >= 0)
            o[1 * N + i] = 1.0f;
          if (v[i].army > 0) o[2 * N + i] = (float)(log((double)v[i].army + 1.0) / 10.0);
          if (v[i].type == GRL_TILE_NORMAL)
            o[3 * N + i] = 1.0f;
          else if (v[i].type == GRL_TILE_MOUNTAIN)
            o[4 * N + i] = 1.0f;
          else if (v[i].type == GRL_TILE_CITY)
            o[5 * N + i] = 1.0f;
          else if (v[i].type == GRL_TILE_GENERAL)
            o[6 * N + i] = 1.0f;
          o[7 * N + i] = (float)tf;
        }
      }
      if (x->out->mask) {
        uint8_t *m = x->out->mask + bp * (size_t)N * 5;
        memset(m, 0, (size_t)N * 5);
        for (int y = 0; y < H; y++)
          for (int xx = 0; xx < W; xx++) {
            int idx = y * W + xx;
            if (v[idx].owner != p || v[idx].army <= 1) continue;
            for (int d = 0; d < 4; d++) {
              int nx = xx + DX[d], ny = y + DY[d];
              if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
              if (v[ny * W + nx].type == GRL_TILE_MOUNTAIN) continue;
              m[idx * 5 + d] = 1;
              m[idx * 5 + 4] = 1;
            }
          }
      }
      if (x->out->stats) {
        int32_t *st = x->out->stats + bp * 4;
        st[0] = g->pl[p].army_count;
        st[1] = g->pl[p].n_owned;
        st[2] = g->pl[p].alive;
        st[3] = g->pl[p].general_idx;
      }
    }
  }
  free(v);
}

/* GeneralsEnv._action_index_to_game_action (generals_env.py:389-441) */
int grlo_gym_encode(grlo_env *e, const int64_t *action_idx, int32_t player, int32_t slot, const uint8_t *mask,
                    int32_t skip_invalid, grl_action *actions, uint8_t *valid) {
  if (!e || !action_idx || !mask || !actions) return GRL_ERR_INVALID_ARG;
  int N = e->N, P = e->cfg.num_players, W = e->cfg.width, H = e->cfg.height, A = e->cfg.max_actions;
  if (player < 0 || player >= P || slot < 0 || slot >= A) return GRL_ERR_INVALID_ARG;
  static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0};
  for (int b = 0; b < e->cfg.num_envs; b++) {
    int64_t a = action_idx[b];
    int ok = a >= 0 && a < (int64_t)N * 5 && mask[((size_t)b * P + player) * N * 5 + a];
    grl_action *rec = &actions[(size_t)b * A + slot];
    memset(rec, 0, sizeof(*rec));
    if (ok) {
      int from_idx = (int)(a / 5), info = (int)(a % 5);
      int fx = from_idx % W, fy = from_idx / W, tx = fx, ty = fy;
      if (info < 4) {
        tx = fx + DX[info];
        ty = fy + DY[info];
      } else { /* half move: the first in-bounds direction */
        for (int d = 0; d < 4; d++) {
          tx = fx + DX[d];
          ty = fy + DY[d];
          if (tx >= 0 && tx < W && ty >= 0 && ty < H) break;
        }
      }
      rec->player_id = (int8_t)player;
      rec->from_x = (int8_t)fx;
      rec->from_y = (int8_t)fy;
      rec->to_x = (int8_t)tx;
      rec->to_y = (int8_t)ty;
      rec->move_all = info != 4; /* Action.half == false -> MoveAll (converters.go:123) */
      rec->present = 1;
    } else if (skip_invalid) {
      actions[(size_t)b * A].flags |= GRL_ACTION_FLAG_SKIP_ENV;
    }
    if (valid) valid[b] = (uint8_t)ok;
  }
  return GRL_OK;
}

/* grl_replay_push_rows (include/grlcuda.h): what vector_env.py:170-176 pushes env by env, as plain loops */
int grlo_replay_push_rows(grlo_env *e, const grl_replay_rows_io *io) {
  if (!e || !io || !io->obs || io->capacity < 1 || io->obs_floats < 1 || io->views < 1 || io->view < 0 || io->view >= io->views)
    return GRL_ERR_INVALID_ARG;
  const int B = e->cfg.num_envs;
  if (io->capacity < B || io->next_row0 < 0 || io->state_row0 < 0) return GRL_ERR_INVALID_ARG;
  const size_t F = (size_t)io->obs_floats;
  for (int b = 0; b < B; b++) {
    const float *row = io->obs + ((size_t)b * (size_t)io->views + (size_t)io->view) * F;
    if (io->next_states) {
      const float *src = (io->done && io->final_obs && io->done[b]) ? io->final_obs + (size_t)b * F : row;
      memcpy(io->next_states + (size_t)((io->next_row0 + b) % io->capacity) * F, src, F * sizeof(float));
    }
    if (io->states) memcpy(io->states + (size_t)((io->state_row0 + b) % io->capacity) * F, row, F * sizeof(float));
  }
  return GRL_OK;
}

/* a uniformly random valid gym action per env: the k-th set mask entry in index order (grlcuda.h) */
int grlo_gym_sample(grlo_env *e, uint64_t seed, const uint8_t *mask, int32_t player, int64_t *action) {
  if (!e || !mask || !action) return GRL_ERR_INVALID_ARG;
  int P = e->cfg.num_players, M = e->N * 5;
  if (player < 0 || player >= P) return GRL_ERR_INVALID_ARG;
  for (int b = 0; b < e->cfg.num_envs; b++) {
    const uint8_t *row = mask + ((size_t)b * P + player) * M;
    int total = 0;
    for (int i = 0; i < M; i++) total += row[i] != 0;
    int64_t pick = 0;
    if (total > 0) {
      uint64_t r = policy_draw(seed, (uint64_t)(e->cfg.env_id_base + b), 0, (uint64_t)player);
      int k = (int)(r % (uint64_t)total);
      for (int i = 0; i < M; i++)
        if (row[i] && k-- == 0) {
          pick = i;
          break;
        }
    }
    action[b] = pick;
  }
  return GRL_OK;
}

int grlo_gym_observe(grlo_env *e, int32_t max_turns, const grl_gym_outputs *out);

/* GeneralsEnv.step (generals_env.py:210-289) for every env; reward :499-561 in float64 like the client */
int grlo_gym_step(grlo_env *e, int32_t max_turns, uint64_t opponent_seed, const grl_gym_step_io *io) {
  if (!e || !io || (!io->action && !io->sampled_action) || !io->out.mask || !io->out.stats || !io->actions || !io->prev_stats || !io->turns ||
      !io->calls || !io->reward || !io->terminated || !io->truncated || !io->valid || !io->done || !io->winner ||
      !io->step_error || max_turns < 1)
    return GRL_ERR_INVALID_ARG;
  int B = e->cfg.num_envs, P = e->cfg.num_players, A = e->cfg.max_actions;
  if (P < 2 || A < 2) return GRL_ERR_INVALID_ARG;
  memcpy(io->prev_stats, io->out.stats, sizeof(int32_t) * (size_t)B * P * 4);
  int st;
  if (io->opponent_action) {
    memset(io->actions, 0, sizeof(grl_action) * (size_t)B * A);
    if ((st = grlo_gym_encode(e, io->opponent_action, 1, 1, io->out.mask, 0, io->actions, NULL))) return st;
  } else { /* the random opponent: a uniformly random legal FULL move (:443-497) */
    if ((st = grlo_sample_actions(e, opponent_seed, io->actions))) return st;
    for (int b = 0; b < B; b++) io->actions[(size_t)b * A + 1].move_all = 1;
  }
  const int64_t *agent = io->action;
  if (!agent) { /* the random agent (python/generals_agent/random_agent.py): a uniformly random entry of the CURRENT mask */
    if ((st = grlo_gym_sample(e, io->agent_seed, io->out.mask, 0, io->sampled_action))) return st;
    agent = io->sampled_action;
  }
  if ((st = grlo_gym_encode(e, agent, 0, 0, io->out.mask, 1, io->actions, io->valid))) return st;
  /* Server.SubmitAction -> ActionValidator.ValidateCoreAction (internal/grpc/gameserver/server.go:241,
   * action_validator.go:113-137): MoveAction.Validate against the board AT SUBMISSION; a refused action is never
   * buffered and the turn runs without it.  The client does not look at the response, so it still counts the
   * step.  What the client's own mask lets through and the server refuses: a half move, which the client aims
   * at the first in-bounds direction whatever stands there (generals_env.py:421-428) -- at a mountain. */
  for (int b = 0; b < B; b++)
    for (int s = 0; s < 2; s++) {
      grl_action *a = &io->actions[(size_t)b * A + s];
      if (a->present && validate_move(&e->g[b], a->player_id, a->from_x, a->from_y, a->to_x, a->to_y) != GRL_STEP_OK)
        a->present = 0;
    }
  grl_step_outputs so;
  memset(&so, 0, sizeof(so));
  so.done = io->done;
  so.winner = io->winner;
  so.step_error = io->step_error;
  if ((st = grlo_step_fused(e, io->actions, 0, 0, &so))) return st;
  if ((st = grlo_gym_observe(e, max_turns, &io->out))) return st;
  int finished = 0;
  for (int b = 0; b < B; b++) {
    int valid = io->valid[b];
    io->turns[b] += valid;
    io->calls[b] += 1;
    int terminated = io->done[b] && valid;
    int truncated = (io->turns[b] >= max_turns && valid) || io->calls[b] >= max_turns;
    const int32_t *cur = io->out.stats + (size_t)b * P * 4, *prev = io->prev_stats + (size_t)b * P * 4;
    double r = 0.0;
    if (!valid) {
      r = -0.1;
    } else if (terminated) {
      r = io->winner[b] == 0 ? 100.0 : -100.0;
    } else {
      r += (double)(cur[1] - prev[1]) * 1.0;
      r += (double)(cur[0] - prev[0]) * 0.01;
      for (int q = 1; q < P; q++)
        if (prev[q * 4 + 2] == 1 && cur[q * 4 + 2] == 0) r += 50.0;
    }
    io->reward[b] = r;
    io->terminated[b] = (uint8_t)terminated;
    io->truncated[b] = (uint8_t)truncated;
    finished += terminated || truncated;
  }
  if (io->n_finished) *io->n_finished = finished;
  return GRL_OK;
}

int grlo_gym_observe(grlo_env *e, int32_t max_turns, const grl_gym_outputs *out) {
  if (!e || !out || max_turns < 1) return GRL_ERR_INVALID_ARG;
  gym_ctx_t x = {e, max_turns, out};
  par_range(e->nthreads, e->cfg.num_envs, gym_range, &x);
  return GRL_OK;
}

int grlo_gym_observe_envs(grlo_env *e, int32_t max_turns, const int32_t *env_ids, int32_t n, const grl_gym_outputs *out) {
  if (!e || !out || !env_ids || max_turns < 1 || n < 0) return GRL_ERR_INVALID_ARG;
  gym_ctx_t x = {e, max_turns, out};
  for (int i = 0; i < n; i++) {
    if (env_ids[i] < 0 || env_ids[i] >= e->cfg.num_envs) return GRL_ERR_INVALID_ARG;
    gym_range(&x, env_ids[i], env_ids[i] + 1);
  }
  return GRL_OK;
}

/* the auto-reset of a vector env (grlcuda.h): re-seed every env whose episode ended, keep player 0's last view */
int grlo_gym_autoreset(grlo_env *e, int32_t max_turns, int64_t base_seed, const grl_gym_autoreset_io *io) {
  if (!e || !io || max_turns < 1 || !io->terminated || !io->truncated || !io->episode || !io->turns || !io->calls)
    return GRL_ERR_INVALID_ARG;
  if (io->final_obs && !io->out.obs) return GRL_ERR_INVALID_ARG;
  int B = e->cfg.num_envs, P = e->cfg.num_players, N = e->N, n = 0;
  gym_ctx_t x = {e, max_turns, &io->out};
  for (int b = 0; b < B; b++) {
    if (!(io->terminated[b] | io->truncated[b])) continue;
    io->episode[b] += 1;
    int64_t seed = base_seed + b + io->episode[b] * (int64_t)B;
    if (io->final_obs)
      memcpy(io->final_obs + (size_t)b * GRL_GYM_CHANNELS * N, io->out.obs + (size_t)b * P * GRL_GYM_CHANNELS * N,
             sizeof(float) * GRL_GYM_CHANNELS * (size_t)N);
    int32_t id = b;
    int st = grlo_reset_seeded(e, &id, 1, &seed);
    if (st) return st;
    io->turns[b] = 0;
    io->calls[b] = 0;
    gym_range(&x, b, b + 1);
    n++;
  }
  if (io->n_reset) *io->n_reset = n;
  return GRL_OK;
}

int grlo_get_state(grlo_env *e, int32_t first, int32_t count, const grl_state_planes *o) {
  if (!e || !o || first < 0 || count < 0 || first + count > e->cfg.num_envs) return GRL_ERR_INVALID_ARG;
  int N = e->N, P = e->cfg.num_players;
  for (int c = 0; c < count; c++) {
    game_t *g = &e->g[first + c];
    for (int i = 0; i < N; i++) {
      size_t k = (size_t)c * N + i;
      if (o->owner) o->owner[k] = g->T[i].owner;
      if (o->army) o->army[k] = g->T[i].army;
      if (o->type) o->type[k] = g->T[i].type;
      if (o->visible) o->visible[k] = g->T[i].vis;
      if (o->changed) o->changed[k] = g->changed[i];
      if (o->vis_changed) o->vis_changed[k] = g->vchg[i];
    }
    if (o->owned) {
      memset(o->owned + (size_t)c * P * N, 0, (size_t)P * N);
      for (int p = 0; p < P; p++)
        for (int k = 0; k < g->pl[p].n_owned; k++) o->owned[((size_t)c * P + p) * N + g->pl[p].owned[k]] = 1;
    }
    if (o->turn) o->turn[c] = g->turn;
    if (o->game_over) o->game_over[c] = g->game_over;
    if (o->winner) o->winner[c] = engine_get_winner(g);
    if (o->step_error) o->step_error[c] = g->step_error;
    for (int p = 0; p < P; p++) {
      if (o->alive) o->alive[(size_t)c * P + p] = g->pl[p].alive;
      if (o->army_count) o->army_count[(size_t)c * P + p] = g->pl[p].army_count;
      if (o->general_idx) o->general_idx[(size_t)c * P + p] = g->pl[p].general_idx;
    }
  }
  return GRL_OK;
}

int grlo_set_state(grlo_env *e, int32_t first, int32_t count, const grl_state_planes *in) {
  if (!e || !in || first < 0 || count < 0 || first + count > e->cfg.num_envs) return GRL_ERR_INVALID_ARG;
  int N = e->N, P = e->cfg.num_players;
  for (int c = 0; c < count; c++) {
    game_t *g = &e->g[first + c];
    for (int i = 0; i < N; i++) {
      size_t k = (size_t)c * N + i;
      if (in->owner) g->T[i].owner = in->owner[k];
      if (in->army) g->T[i].army = in->army[k];
      if (in->type) g->T[i].type = in->type[k];
      if (in->visible) g->T[i].vis = in->visible[k] & ((1u << P) - 1u); /* seats beyond P are not addressable (grlcuda.h) */
    }
    if (in->changed) {
      set_clear(g->changed, &g->n_changed, N);
      for (int i = 0; i < N; i++)
        if (in->changed[(size_t)c * N + i]) set_add(g->changed, &g->n_changed, i);
    }
    if (in->vis_changed) {
      set_clear(g->vchg, &g->n_vchg, N);
      for (int i = 0; i < N; i++)
        if (in->vis_changed[(size_t)c * N + i]) set_add(g->vchg, &g->n_vchg, i);
    }
    if (in->owned)
      for (int p = 0; p < P; p++) {
        g->pl[p].n_owned = 0;
        for (int i = 0; i < N; i++)
          if (in->owned[((size_t)c * P + p) * N + i]) g->pl[p].owned[g->pl[p].n_owned++] = i;
      }
    if (in->turn) g->turn = in->turn[c];
    if (in->game_over) g->game_over = in->game_over[c];
    if (in->step_error) g->step_error = in->step_error[c];
    for (int p = 0; p < P; p++) {
      if (in->alive) g->pl[p].alive = in->alive[(size_t)c * P + p];
      if (in->army_count) g->pl[p].army_count = in->army_count[(size_t)c * P + p];
      if (in->general_idx) g->pl[p].general_idx = in->general_idx[(size_t)c * P + p];
      g->reward[p] = 0.0f;
      g->action_index[p] = -1;
    }
    for (int i = 0; i < N; i++) {
      g->prev_owner[i] = g->T[i].owner;
      g->prev_army[i] = g->T[i].army;
    }
  }
  return GRL_OK;
}

int grlo_state_hash(grlo_env *e, uint64_t *out) {
  if (!e || !out) return GRL_ERR_INVALID_ARG;
  batch_ctx_t x = {e, 0, NULL, out, NULL, 0};
  par_range(e->nthreads, e->cfg.num_envs, hash_range, &x);
  return GRL_OK;
}

int grlo_buffer_hash(grlo_env *e, const void *buf, size_t row_words, int32_t rows, uint64_t *out) {
  if (!buf || !out || rows < 0) return GRL_ERR_INVALID_ARG;
  batch_ctx_t x = {e, 0, NULL, out, (const uint32_t *)buf, row_words};
  par_range(e ? e->nthreads : 1, rows, row_hash_range, &x);
  return GRL_OK;
}

int grlo_stats(grlo_env *e, uint64_t out[4]) {
  if (!e || !out) return GRL_ERR_INVALID_ARG;
  memcpy(out, e->stats, sizeof(e->stats));
  return GRL_OK;
}

/* ---- test hooks for transliterated reference unit tests (not in the product ABI) ---- */

/* mapgen stages with explicit parameters, as generator_test.go drives them */
int grlo_test_place_mountains(int W, int H, int veins, int min_len, int max_len, int64_t seed, int32_t *type) {
  int N = W * H;
  tile_t *T = (tile_t *)calloc((size_t)N, sizeof(tile_t));
  for (int i = 0; i < N; i++) T[i].owner = GRL_NEUTRAL;
  go_rand *rng = (go_rand *)malloc(sizeof(go_rand));
  go_rand_seed(rng, seed);
  mapcfg_t m = {W, H, 0, 20, 40, 5, veins, min_len, max_len};
  place_mountains(&m, rng, T);
  for (int i = 0; i < N; i++) type[i] = T[i].type;
  free(rng);
  free(T);
  return GRL_OK;
}

int grlo_test_place_cities(int W, int H, int city_ratio, int city_army, int64_t seed, int32_t *type, int32_t *army) {
  int N = W * H;
  tile_t *T = (tile_t *)calloc((size_t)N, sizeof(tile_t));
  for (int i = 0; i < N; i++) {
    T[i].owner = GRL_NEUTRAL;
    T[i].type = type[i]; /* in/out: a caller-prepared board (generator_test.go:196-244 pre-fills mountains) */
  }
  go_rand *rng = (go_rand *)malloc(sizeof(go_rand));
  go_rand_seed(rng, seed);
  mapcfg_t m = {W, H, 0, city_ratio, city_army, 5, 0, 3, 3};
  place_cities(&m, rng, T);
  for (int i = 0; i < N; i++) {
    type[i] = T[i].type;
    army[i] = T[i].army;
  }
  free(rng);
  free(T);
  return GRL_OK;
}

/* raw generator draws: kind 0 Int63, 1 Intn(n), 2 Uint32, 3 Shuffle-int31n(n) */
int grlo_test_gorand(int64_t seed, int kind, int n, int count, int64_t *out) {
  go_rand *rng = (go_rand *)malloc(sizeof(go_rand));
  go_rand_seed(rng, seed);
  for (int i = 0; i < count; i++) {
    switch (kind) {
      case 0: out[i] = go_rand_int63(rng); break;
      case 1: out[i] = go_rand_intn(rng, n); break;
      case 2: out[i] = go_rand_uint32(rng); break;
      default: out[i] = go_rand_int31n_lemire(rng, n); break;
    }
  }
  free(rng);
  return GRL_OK;
}

/* core.MoveAction.Validate against env b's current board (core/action.go:56-105) */
int grlo_test_validate(grlo_env *e, int b, int player, int fx, int fy, int tx, int ty) {
  return validate_move(&e->g[b], player, fx, fy, tx, ty);
}

/* experience.CalculateRewardWithConfig(prev, curr, p) with curr = env b's state and
 * prev given as planes (rewards_test.go builds arbitrary prev/curr pairs by hand) */
float grlo_test_reward(grlo_env *e, int b, int p, const int32_t *prev_owner, const int32_t *prev_army) {
  game_t *g = &e->g[b];
  for (int i = 0; i < g->N; i++) {
    g->prev_owner[i] = prev_owner[i];
    g->prev_army[i] = prev_army[i];
  }
  return calc_reward(g, &e->cfg.reward, p);
}

/* Generator.GenerateMap with an explicit MapConfig (generator_test.go:394-455 overrides fields) */
int grlo_test_generate_map(int W, int H, int players, int city_ratio, int city_army, int spacing, int veins,
                           int min_len, int max_len, int64_t seed, int32_t *owner, int32_t *army, int32_t *type) {
  int N = W * H;
  tile_t *T = (tile_t *)calloc((size_t)N, sizeof(tile_t));
  go_rand *rng = (go_rand *)malloc(sizeof(go_rand));
  go_rand_seed(rng, seed);
  mapcfg_t m = {W, H, players, city_ratio, city_army, spacing, veins, min_len, max_len};
  int st = generate_map(&m, rng, T);
  for (int i = 0; i < N; i++) {
    owner[i] = T[i].owner;
    army[i] = T[i].army;
    type[i] = T[i].type;
  }
  free(rng);
  free(T);
  return st;
}

/* Generator.placeGenerals on a caller-prepared board (generator_test.go:336-392) */
int grlo_test_place_generals(int W, int H, int players, int spacing, int64_t seed, int32_t *type, int32_t *owner) {
  int N = W * H;
  tile_t *T = (tile_t *)calloc((size_t)N, sizeof(tile_t));
  for (int i = 0; i < N; i++) {
    T[i].owner = GRL_NEUTRAL;
    T[i].type = type[i];
  }
  go_rand *rng = (go_rand *)malloc(sizeof(go_rand));
  go_rand_seed(rng, seed);
  mapcfg_t m = {W, H, players, 20, 40, spacing, 0, 3, 3};
  int st = GRL_OK;
  int placed[GRL_MAX_PLAYERS];
  for (int pid = 0; pid < players; pid++) {
    int idx = find_general_location(&m, rng, T, placed, pid);
    if (idx < 0) {
      st = GRL_ERR_MAPGEN;
      break;
    }
    T[idx].owner = pid;
    T[idx].army = 2;
    T[idx].type = GRL_TILE_GENERAL;
    placed[pid] = idx;
  }
  for (int i = 0; i < N; i++) {
    owner[i] = T[i].owner;
    type[i] = T[i].type;
  }
  free(rng);
  free(T);
  return st;
}

int grlo_launch_count(grlo_env *e, uint64_t *out) {
  if (!e || !out) return GRL_ERR_INVALID_ARG;
  *out = 0; /* no GPU work happens in the oracle */
  return GRL_OK;
}
