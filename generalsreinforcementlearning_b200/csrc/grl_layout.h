// grl_layout.h — device-resident state layout shared by the host ABI code and the kernels.
//
// One game ("env") is one contiguous, 16-byte aligned SLAB of 32-bit words in HBM, so a
// warp (or one TMA bulk copy) moves it with fully coalesced 128-bit transactions:
//
//   header   8 + 5*P words (rounded up to 4)
//     [0] turn            [1] flags: bits0-7 alive, bit8 gameOver, bits16-23 last step_error
//     [2] steps executed  [3] error turns  [4] games finished  [5] steps rejected (game over)
//     [6] sticky army-overflow flag        [7] reserved
//     per player p at 8+5p: armyCount, generalIdx, trueArmy (sum over true ownership),
//                           reward bits (fp32 of the last step), actionIndex of the last step
//   own [P][NW]   ownership as one N-bit linear bitmask per player (tile t = bit t&31 of word t>>5)
//   list[P][NW]   the reference's cached Player.OwnedTiles, as a set
//   vis [P][NW]   Tile.VisibleBitfield transposed: bit p of tile t = bit t of vis[p]
//   changed[NW], vchg[NW]   GameState.ChangedTiles / VisibilityChangedTiles
//   army u16[NA]  (NA = N rounded up to 8)
//
// Terrain never changes within an episode (core/movement.go:38), so it lives in a separate
// read-only STATIC slab: mountain[NW], city[NW], general[NW] (rounded up to 4 words).
//
// NW = ceil(N/32) <= 32 because W,H <= 32: lane j of a warp owns word j of every mask.
#pragma once
#include <stdint.h>

#define GRL_HDR_TURN 0
#define GRL_HDR_FLAGS 1
#define GRL_HDR_STEPS 2
#define GRL_HDR_ERRORS 3
#define GRL_HDR_FINISHED 4
#define GRL_HDR_REJECTED 5
#define GRL_HDR_OVERFLOW 6
#define GRL_HDR_PLAYER0 8
#define GRL_HDR_PER_PLAYER 5
#define GRL_PL_ARMY_COUNT 0
#define GRL_PL_GENERAL_IDX 1
#define GRL_PL_TRUE_ARMY 2
#define GRL_PL_REWARD 3
#define GRL_PL_ACTION_INDEX 4

#define GRL_FLAG_OVER (1u << 8)
#define GRL_FLAG_ERR_SHIFT 16

struct GrlLayout {
  int N, P, NW, NA;
  int hdr_words;
  int off_own, off_list, off_vis, off_changed, off_vchg, off_army;
  int slab_words;    // multiple of 8: a slab is a whole number of 32-byte sectors
  int static_words;  // multiple of 4: [M][C][G]
};

static inline GrlLayout grl_make_layout(int W, int H, int P) {
  GrlLayout L;
  L.N = W * H;
  L.P = P;
  L.NW = (L.N + 31) / 32;
  L.NA = (L.N + 7) & ~7;
  L.hdr_words = (GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * P + 3) & ~3;
  L.off_own = L.hdr_words;
  L.off_list = L.off_own + P * L.NW;
  L.off_vis = L.off_list + P * L.NW;
  L.off_changed = L.off_vis + P * L.NW;
  L.off_vchg = L.off_changed + L.NW;
  L.off_army = (L.off_vchg + L.NW + 3) & ~3;
  L.slab_words = (L.off_army + L.NA / 2 + 7) & ~7;
  L.static_words = (3 * L.NW + 3) & ~3;
  return L;
}

// words of one packed observation record: own[P][NW] vis[P][NW] mountain[NW] city|general[NW] army u16[NA]
#ifdef __CUDACC__
#define GRL_HD __host__ __device__
#else
#define GRL_HD
#endif
GRL_HD static inline int grl_packed_words(const GrlLayout &L) { return ((2 * L.P + 2) * L.NW + L.NA / 2 + 3) & ~3; }

// Kernel parameter block (passed by value, __grid_constant__).
struct GrlKParams {
  uint32_t *state;          // [B][slab_words]
  const uint32_t *statics;  // [B][static_words]
  const uint32_t *geom;     // [3][32]: valid tiles, x != 0, x != W-1 (linear bitmasks)
  const void *actions;      // grl_action [B][A] or nullptr
  float *obs;
  uint32_t *mask_bits;
  float *reward;
  uint8_t *done;
  int8_t *winner;
  uint8_t *step_error;
  int32_t *action_index;
  uint32_t *obs_packed;     // [B][grl_packed_words(L)] packed observation records (include/grlcuda.h)
  unsigned long long policy_seed;
  uint32_t flags;
  int B, W, H, N, P, NW, A;
  int game0, game_end;      // the turn kernel processes envs [game0, game_end) (host-buffer calls pipeline sub-ranges)
  GrlLayout L;
  int fog, pg, pc, pn, grow_interval;
  int env_id_base;
  int prefetch_dist;        // > 0: the warp of game g prefetches the slab of game g + dist into L2
  // Overlapped launches (grl_turn.cuh, "launch overlap"): one word per warp of the grid.  A warp publishes epoch_seq when
  // everything it wrote is visible; with epoch_need != 0 it first waits until the warp that held its games in the previous
  // launch has published at least epoch_need.  nullptr: neither.
  uint32_t *epoch;
  uint32_t epoch_need, epoch_seq;
  float rw[11];
};

// Second parameter block of the turn kernel, read only by its fused gym-step instantiation
// (grl_gym_step: one GeneralsEnv.step() for every env in a single launch).
struct GrlGymK {
  const long long *action;           // [B] Discrete(N*5) indices of player 0, or nullptr: the random agent
  const long long *opponent_action;  // [B] or nullptr: the random opponent
  const float *logtab;               // log(a + 1) / 10 for every uint16 army
  float *obs;                        // [B][P][9][N]
  uint8_t *mask;                     // [B][P][N*5]
  int32_t *stats;                    // [B][P][4]
  int32_t *turns, *calls;            // [B]
  double *reward;                    // [B]
  uint8_t *terminated, *truncated, *valid;
  int32_t *n_finished;
  unsigned long long opponent_seed;
  int max_turns;
  unsigned long long agent_seed;     // action == nullptr: player 0 draws grl_gym_sample(agent_seed, mask, 0) in the launch
  long long *sampled_action;         // [B] the index it drew
};
