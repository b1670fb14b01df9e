/*
 * grlcuda.h — C ABI of libgrlcuda.so, the B200-native batched Generals.io turn engine.
 *
 * Drop-in boundary for the reference's turn-processing hot path.  The reference
 * (mitchelldurbincs/GeneralsReinforcementLearning, Go) has no FFI of its own; the
 * seam this ABI replaces is the public surface of `game.Engine`:
 *
 *   grl_create / grl_reset_*   <- NewEngine            internal/game/engine.go:62-71,
 *                                                      engine_initializer.go:34-87,218-225
 *   grl_step / grl_step_fused  <- Engine.Step          internal/game/engine.go:75
 *                                 (TurnProcessor.ProcessTurn, turn_processor.go:29-77)
 *   grl_mask                   <- Engine.GetLegalActionMask   engine.go:271-280
 *                                 Serializer.GenerateActionMask experience/serializer.go:112-176
 *   grl_visibility             <- Engine.ComputePlayerVisibility visibility.go:153,
 *                                 visibility_optimized.go:166-195
 *   grl_observe                <- Serializer.StateToTensor    experience/serializer.go:37-109
 *                                 CalculateRewardWithConfig   experience/rewards.go:45-85
 *                                 Engine.IsGameOver/GetWinner engine.go:198,248
 *   grl_get_state/grl_set_state<- Engine.GameState()/GameState.Clone  engine.go:197, state.go:37-70
 *
 * One grl_env = B independent games ("envs") resident on ONE CUDA device as
 * struct-of-arrays slabs, stepped in lockstep on one stream.  The handle is not
 * internally locked (the reference serialises each game under one mutex,
 * internal/grpc/gameserver/game_manager.go:25).  Every data pointer is
 * caller-allocated and never retained past the call; pointers may be host or
 * device memory (detected with cudaPointerGetAttributes; host buffers are staged
 * through library scratch and copied on the env's stream, and the call returns
 * after the copy completes).  Calls with device buffers are asynchronous on the
 * env's stream; use grl_sync().  Turn launches (grl_step / grl_step_fused with
 * device buffers) that follow each other on the stream with nothing in between
 * OVERLAP on the device: the later launch starts while the earlier grid drains and
 * its warps wait only for the warp that held their games (programmatic dependent
 * launch + per-warp release/acquire words).  Stream order as the caller sees it is
 * unchanged — anything enqueued after call k, on this stream or behind an event,
 * finds call k complete, and a kernel or copy enqueued between two calls
 * serialises them as always.  GRL_LAUNCH_OVERLAP=0 turns the overlap off.
 *
 * The same ABI, prefixed grlo_, is exported by the CPU oracle (oracle/) — test
 * infrastructure only.
 */
#ifndef GRLCUDA_H
#define GRLCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRL_ABI_VERSION 3 /* 2: grl_step_outputs.obs_packed, grl_obs_packed_words, grl_expand_obs; 3: grl_gym_step_io.agent_seed / sampled_action */

#define GRL_MAX_DIM 32      /* width, height <= 32: one 32-bit word spans a board row */
#define GRL_MAX_PLAYERS 8   /* reference bitfield allows 32 (core/board.go:11); configs need <= 4 */
#define GRL_MAX_ACTIONS 12  /* action slots per env per step: Go's sort.Slice is an insertion
                               sort (stable) only for n <= 12 (processor/action_processor.go:39) */
#define GRL_OBS_CHANNELS 9  /* experience/serializer.go:9-20 */

/* ---- call status ------------------------------------------------------- */
enum {
  GRL_OK = 0,
  GRL_ERR_INVALID_ARG = -1,
  GRL_ERR_CUDA = -2,
  GRL_ERR_NOMEM = -3,
  GRL_ERR_MAPGEN = -4, /* mapgen could not place a general (mapgen/generator.go:252) */
  GRL_ERR_UNSUPPORTED = -5
};

/* ---- per-env step error plane (numbers = proto/common/v1/common.proto:39-55,
 *      sentinels internal/game/core/errors.go:8-17) ------------------------ */
enum {
  GRL_STEP_OK = 0,
  GRL_STEP_INVALID_COORDINATES = 1,
  GRL_STEP_NOT_ADJACENT = 2,
  GRL_STEP_NOT_OWNED = 3,
  GRL_STEP_INSUFFICIENT_ARMY = 4,
  GRL_STEP_GAME_OVER = 5,
  GRL_STEP_MOVE_TO_SELF = 7,
  GRL_STEP_TARGET_IS_MOUNTAIN = 8,
  GRL_STEP_ARMY_OVERFLOW = 100 /* not a reference error: a tile army left the uint16 plane's
                                  range; the value was saturated (never wraps silently) */
};

/* ---- tile types (internal/game/core/board.go:20-26) --------------------- */
enum { GRL_TILE_NORMAL = 0, GRL_TILE_GENERAL = 1, GRL_TILE_CITY = 2, GRL_TILE_MOUNTAIN = 3 };
#define GRL_NEUTRAL (-1)

/* ---- one move (core/action.go:23-35).  8 bytes, [B][max_actions] per step.
 * present == 0 is an empty slot (the reference passes a shorter slice).
 * Slots of one env are applied in the order a stable sort by player_id gives. */
typedef struct grl_action {
  int8_t player_id;
  int8_t from_x, from_y;
  int8_t to_x, to_y;
  uint8_t move_all; /* 1: leave one behind; 0: move half, floor, min 1 (core/movement.go:40-49) */
  uint8_t present;
  uint8_t flags;    /* GRL_ACTION_FLAG_*; read from slot 0 of an env only */
} grl_action;

/* Slot 0 of an env may carry this flag: the env takes NO turn in this call (its game is still
 * waiting for its players' actions — the per-game turn barrier of
 * internal/grpc/gameserver/game_manager.go:559-600; or the gym client rejected the action
 * before submitting it, generals_env.py:226-229).  State and read-outs are left unchanged. */
#define GRL_ACTION_FLAG_SKIP_ENV 1

/* ---- reward weights (experience/rewards.go:9-37) ------------------------ */
typedef struct grl_reward_config {
  float win_game, lose_game;
  float capture_city, lose_city;
  float capture_general, lose_general;
  float territory_gained, territory_lost; /* territory_lost / army_lost are carried but, as in */
  float army_gained, army_lost;           /* the reference, never read by the reward sum       */
  float army_advantage;
} grl_reward_config;

typedef struct grl_config {
  int32_t num_envs;           /* B games on this device */
  int32_t width, height;      /* 1..GRL_MAX_DIM */
  int32_t num_players;        /* 1..GRL_MAX_PLAYERS */
  int32_t device;             /* CUDA ordinal */
  int32_t max_actions;        /* action slots per env per step, 1..GRL_MAX_ACTIONS */
  int32_t fog_of_war;         /* engine_initializer.go:118 hard-wires 1 */
  int32_t env_id_base;        /* global id of env 0 (shard offset); keys the synthetic policy */
  /* internal/config/config.go:198-209 */
  int32_t city_ratio;             /* 20 */
  int32_t city_start_army;        /* 40 */
  int32_t min_general_spacing;    /* 5 */
  int32_t production_general;     /* 1 */
  int32_t production_city;        /* 1 */
  int32_t production_normal;      /* 1 */
  int32_t normal_growth_interval; /* 25 */
  int32_t host_threads;       /* host mapgen / oracle stepping threads; 0 = all cores */
  grl_reward_config reward;
} grl_config;

typedef struct grl_env grl_env;

/* ---- mask read-out variants -------------------------------- */
enum {
  GRL_MASK_ENGINE_URDL = 0,     /* rules/legal_moves.go:19-73: bytes [B][P][N*4], list-based, army>1 */
  GRL_MASK_SERIALIZER_UDLR = 1, /* experience/serializer.go:112-176: bytes [B][P][N*4], ownership scan */
  GRL_MASK_ENGINE_URDL_BITS = 2,/* packed: uint32 [B][P][ceil(4N/32)], bit i of the flat []bool */
  GRL_MASK_ENGINE_HALF_BITS = 3 /* packed tile x 4 dirs x {full,half}: uint32 [B][P][2][ceil(4N/32)];
                                   a half move is legal iff the full move is (core/action.go:56-105) */
};

/* ---- fused step outputs; any NULL member is skipped --------------------- */
typedef struct grl_step_outputs {
  float *obs;            /* [B][P][9][H][W]  StateToTensor of the post-step state        */
  uint32_t *mask_bits;   /* [B][P][ceil(4N/32)] GRL_MASK_ENGINE_URDL_BITS                */
  float *reward;         /* [B][P]  CalculateReward(prev, curr, p)                        */
  uint8_t *done;         /* [B]     Engine.IsGameOver()                                   */
  int8_t *winner;        /* [B]     Engine.GetWinner()                                    */
  uint8_t *step_error;   /* [B]     GRL_STEP_* of this step                               */
  int32_t *action_index; /* [B][P]  Serializer.ActionToIndex of the player's submitted move
                                    (UDLR order, serializer.go:179-198); -1 when the player
                                    submitted none or the turn aborted (no experience)    */
  uint32_t *obs_packed;  /* [B][grl_obs_packed_words()]  everything StateToTensor reads, bit-packed: the
                                    record grl_expand_obs turns into `obs` bit for bit on the host
                                    (for consumers in HOST memory: 1,120 B instead of 28,800 B per
                                    20x20 2-player env-step across PCIe)                  */
} grl_step_outputs;

/* One packed observation record (32-bit words; NW = ceil(N/32), tile t = bit t&31 of word t>>5):
 *   own [P][NW]   tiles owned by player p            vis [P][NW]   Tile.VisibleBitfield bit p (all ones without fog)
 *   mountain [NW]                                    city_or_general [NW]
 *   army uint16 [N rounded up to 8]                  (padded to a multiple of 4 words)
 * Serializer.StateToTensor (experience/serializer.go:37-109) is a function of exactly these planes. */
int32_t grl_obs_packed_words(int32_t width, int32_t height, int32_t num_players);

/* Host-only (no device work): expand `count` packed records into StateToTensor's float32 tensors,
 * obs[count][P][9][H][W], on `threads` host threads (0 = all cores).  Bit-identical to what the same step
 * would have written into grl_step_outputs.obs. */
int grl_expand_obs(int32_t width, int32_t height, int32_t num_players, const uint32_t *packed, int32_t count, float *obs,
                   int32_t threads);

/* ---- full state planes for parity / replay / checkpoint (host memory).
 * Arrays are [count][...]; NULL members are skipped. ----------------------- */
typedef struct grl_state_planes {
  int32_t *owner;        /* [count][N]  -1 neutral                                   */
  int32_t *army;         /* [count][N]                                               */
  int32_t *type;         /* [count][N]  GRL_TILE_*                                   */
  uint32_t *visible;     /* [count][N]  Tile.VisibleBitfield (bit p)                 */
  uint8_t *owned;        /* [count][P][N] membership in Player.OwnedTiles (cached!)  */
  uint8_t *changed;      /* [count][N]  GameState.ChangedTiles                       */
  uint8_t *vis_changed;  /* [count][N]  GameState.VisibilityChangedTiles             */
  int32_t *turn;         /* [count]                                                  */
  int32_t *game_over;    /* [count]     Engine.gameOver                              */
  int32_t *winner;       /* [count]                                                  */
  int32_t *alive;        /* [count][P]                                               */
  int32_t *army_count;   /* [count][P]  Player.ArmyCount (over the cached list)      */
  int32_t *general_idx;  /* [count][P]  Player.GeneralIdx (highest general-type tile in the
                                        list; the reference's tie-break is Go map order) */
  int32_t *step_error;   /* [count]                                                  */
} grl_state_planes;

/* ---- generals_gym read-outs of player p's fog-filtered view (the proto view of
 *      internal/grpc/gameserver/server.go:556-582 fed through
 *      python/generals_gym/generals_env.py:291-387); any NULL member is skipped -------------- */
#define GRL_GYM_CHANNELS 9
typedef struct grl_gym_outputs {
  float *obs;      /* [B][P][9][H][W] _get_observation: visible, ownership 0/.5/1, log(army+1)/10,
                      one-hot normal/mountain/city/general, turn/max_turns, 0 */
  uint8_t *mask;   /* [B][P][N*5] _get_valid_actions_mask: tile*5 + {up,right,down,left,half} */
  int32_t *stats;  /* [B][P][4] PlayerState: army_count, tile_count (cached list length), alive,
                      general_idx (server.go:528-553) */
} grl_gym_outputs;

/* flags for grl_step / grl_step_fused */
enum {
  GRL_STEP_FLAG_NONE = 0,
  GRL_STEP_FLAG_RANDOM_POLICY = 1 /* ignore `actions`; every alive player plays one move drawn
                                     uniformly from its engine mask with the counter-based
                                     generator keyed (policy_seed, global env id, turn, player) */
};

int grl_abi_version(void);
const char *grl_status_string(int status);
const char *grl_last_error(void); /* thread-local detail of the last failing call */

/* Fill *cfg with the reference defaults (config.go:198-209, rewards.go:23-37). */
int grl_default_config(grl_config *cfg);

int grl_create(const grl_config *cfg, grl_env **out);
int grl_destroy(grl_env *env);
int grl_sync(grl_env *env);
/* Run this env's work on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL is
 * CUDA's default stream).  Lets a trainer order the step against its own kernels. */
int grl_set_stream(grl_env *env, void *cuda_stream);
int grl_get_config(const grl_env *env, grl_config *out);

/* Reset `n` envs (env_ids NULL = all B in order) from seeds: map i is what
 * mapgen.NewGenerator(DefaultMapConfig(W,H,P), rand.New(rand.NewSource(seeds[i]))).GenerateMap()
 * yields (mapgen/generator.go:25-253), generated on the host and uploaded once; then the
 * turn-0 full stats + full fog + game-over check of engine_initializer.go:218-225. */
int grl_reset_seeded(grl_env *env, const int32_t *env_ids, int32_t n, const int64_t *seeds);

/* Reset from caller-built boards (the reference's tests build boards by hand):
 * owner/army/type are host int32 [n][N]; visible bits start at 0. */
int grl_reset_boards(grl_env *env, const int32_t *env_ids, int32_t n, const int32_t *owner,
                     const int32_t *army, const int32_t *type);

/* Host-only: generate one map exactly as the reference does (no device work). */
int grl_mapgen(const grl_config *cfg, int64_t seed, int32_t *owner, int32_t *army, int32_t *type);

/* One ProcessTurn for every env.  actions: [B][max_actions] (host or device), may be NULL
 * (every slot empty).  Stepping a finished env mutates nothing and records
 * GRL_STEP_GAME_OVER (turn_processor.go:95-113). */
int grl_step(grl_env *env, const grl_action *actions, uint32_t flags, uint64_t policy_seed);

/* ProcessTurn + all read-outs in one pass over the state (the hot path). */
int grl_step_fused(grl_env *env, const grl_action *actions, uint32_t flags, uint64_t policy_seed,
                   const grl_step_outputs *out);

/* Read-outs of the current state (no step).  reward is the one computed by the last step
 * (0 after reset). */
int grl_observe(grl_env *env, const grl_step_outputs *out);
int grl_mask(grl_env *env, int variant, void *out);
/* visible/fog: bytes [B][P][N] (PlayerVisibility.VisibleTiles / FogTiles). */
int grl_visibility(grl_env *env, uint8_t *visible, uint8_t *fog);

/* generals_gym observation / mask / player stats of the current state for every (env, player).
 * max_turns normalises channel 7 (generals_env.py:337). */
int grl_gym_observe(grl_env *env, int32_t max_turns, const grl_gym_outputs *out);
/* The same read-outs for the n listed envs only (env_ids: host int32[n]).  The planes keep their full
 * [B][P][...] shape; entries of other envs are left untouched.  What a vector env calls after re-seeding the
 * envs whose episode ended (generals_env.py:188-208 reset -> _get_observation / _get_info). */
int grl_gym_observe_envs(grl_env *env, int32_t max_turns, const int32_t *env_ids, int32_t n, const grl_gym_outputs *out);

/* generals_gym action decoding for one player (generals_env.py:389-441): action_idx[b] = tile*5 +
 * {up,right,down,left,half} becomes a grl_action in slot `slot` of env b (a half move goes to the first
 * in-bounds direction in the order up,right,down,left — the client's own simplification).  An index whose
 * entry in `mask` (the [B][P][N*5] plane of grl_gym_observe) is false is REJECTED like the client does
 * before submitting: the slot is left empty, valid[b] = 0 and, when skip_invalid != 0, slot 0 of that env
 * gets GRL_ACTION_FLAG_SKIP_ENV.  libgrlcuda.so requires device pointers here (it is the device-side glue
 * of the vector env); the oracle takes host pointers. */
int grl_gym_encode(grl_env *env, const int64_t *action_idx, int32_t player, int32_t slot, const uint8_t *mask,
                   int32_t skip_invalid, grl_action *actions, uint8_t *valid);

/* A uniformly random VALID action per env for `player`, drawn from the mask plane of grl_gym_observe /
 * grl_gym_step (what the reference's random agent and its random opponent do with `valid_actions_mask`,
 * python/generals_gym/generals_env.py:443-497, python/generals_agent/random_agent.py): action[b] = the k-th set entry of
 * mask[b][player][:] in index order, k = draw(seed, global env id, player) mod the number of set entries; 0 (which the
 * env rejects) when none is set.  libgrlcuda.so requires device pointers; the oracle takes host pointers. */
int grl_gym_sample(grl_env *env, uint64_t seed, const uint8_t *mask, int32_t player, int64_t *action);

/* One GeneralsEnv.step() for every env in a single call (generals_env.py:210-289): decode the agent's
 * (player 0) action and reject it client-side when the mask forbids it (that env then takes no turn,
 * reward -0.1); the opponent (player 1) plays `opponent_action` or, when that is NULL, a uniformly random
 * legal full move; Step; the gym read-outs of the new state; the client-side reward (:499-561, float64);
 * terminated (game left IN_PROGRESS) / truncated (turns or step() calls reached max_turns).
 * libgrlcuda.so requires device pointers for every plane; nothing is copied to the host. */
typedef struct grl_gym_step_io {
  const int64_t *action;          /* [B] in: Discrete(N*5) indices of player 0, or NULL for the random agent */
  const int64_t *opponent_action; /* [B] in: indices of player 1, or NULL for the random opponent           */
  grl_gym_outputs out;            /* in/out: mask and stats of the CURRENT state are read, all three rewritten */
  grl_action *actions;            /* [B][max_actions] scratch                                                */
  int32_t *prev_stats;            /* [B][P][4] scratch                                                       */
  int32_t *turns, *calls;         /* [B] in/out: turns taken / step() calls of the running episode           */
  double *reward;                 /* [B] out                                                                  */
  uint8_t *terminated, *truncated, *valid; /* [B] out                                                        */
  uint8_t *done;                  /* [B] out: Engine.IsGameOver                                              */
  int8_t *winner;                 /* [B] out                                                                  */
  uint8_t *step_error;            /* [B] out                                                                  */
  int32_t *n_finished;            /* [1] out: number of envs with terminated | truncated                      */
  /* The random agent (python/generals_agent/random_agent.py; what the trainers' epsilon-greedy exploration draws with
   * np.random.choice over valid_actions_mask): with action == NULL player 0 plays, in every env, exactly the index
   * grl_gym_sample(agent_seed, out.mask, 0, .) returns for the CURRENT state — drawn inside the step's own launch, so
   * a random-agent step costs no second launch and no read of the N*5 mask bytes.  The index played is written to
   * sampled_action (required in that case; ignored when `action` is given). */
  uint64_t agent_seed;
  int64_t *sampled_action;        /* [B] out                                                                  */
} grl_gym_step_io;
int grl_gym_step(grl_env *env, int32_t max_turns, uint64_t opponent_seed, const grl_gym_step_io *io);

/* The auto-reset of a vector env, without a host round trip: every env whose episode ended in the last grl_gym_step
 * (terminated | truncated) is re-seeded — episode[b] += 1, seed = base_seed + b + episode[b] * num_envs, the reference's map
 * generator and turn-0 set-up (engine_initializer.go:34-87,218-225) — its turn/call counters are zeroed, player 0's last
 * observation is kept in final_obs[b] (what Gymnasium vector envs hand out as "final_observation"), and the gym read-outs of
 * the new games replace the finished ones' rows.  Nothing is read back: a training loop can enqueue step after step.
 * libgrlcuda.so requires device pointers for every plane (one compaction + map generation + set-up + read-out launch
 * sequence on the env's stream, sized by a device-side count); the oracle takes host pointers.  A seed whose map cannot
 * seat every general (mapgen/generator.go:252 — only on boards too small for the general spacing) is reported by
 * grl_reset_seeded as GRL_ERR_MAPGEN; this asynchronous path leaves such an env without generals, i.e. finished, and it is
 * re-seeded again on the next call. */
typedef struct grl_gym_autoreset_io {
  const uint8_t *terminated, *truncated; /* [B] in: the flags of the last grl_gym_step                       */
  int64_t *episode;                      /* [B] in/out: episodes finished so far per env                     */
  int32_t *turns, *calls;                /* [B] in/out: zeroed for the re-seeded envs                        */
  grl_gym_outputs out;                   /* in/out: rows of the re-seeded envs are rewritten                 */
  float *final_obs;                      /* [B][9][H][W] out (may be NULL): player 0's view as the episode ended */
  int32_t *n_reset;                      /* [1] out (may be NULL): number of envs re-seeded                  */
} grl_gym_autoreset_io;
int grl_gym_autoreset(grl_env *env, int32_t max_turns, int64_t base_seed, const grl_gym_autoreset_io *io);

/* The observation rows of one vector step for a replay ring — what python/generals_gym/vector_env.py:170-176 pushes env
 * by env through replay_buffer.py:30-40, for every env at once.  `obs` is the gym observation plane
 * [B][views][obs_floats] AFTER grl_gym_step (and grl_gym_autoreset); view `view` of env b is written to
 *   next_states[(next_row0 + b) % capacity]   — final_obs[b] instead where done[b] != 0: the view the episode ended with,
 *   states[(state_row0 + b) % capacity]       — the state of the env's NEXT transition (a re-seeded env's first view),
 * in ONE pass over the plane (either destination may be NULL).  Rows of a ring are [obs_floats] floats, contiguous.
 * libgrlcuda.so takes device pointers (one launch on the env's stream, no host read); the oracle takes host pointers. */
typedef struct grl_replay_rows_io {
  const float *obs;            /* [B][views][obs_floats]                                   */
  const float *final_obs;      /* [B][obs_floats] or NULL                                   */
  const uint8_t *done;         /* [B] or NULL: no row comes from final_obs                  */
  float *next_states;          /* [capacity][obs_floats] or NULL                            */
  float *states;               /* [capacity][obs_floats] or NULL                            */
  int64_t capacity, next_row0, state_row0;
  int32_t views, view, obs_floats, reserved;
} grl_replay_rows_io;
int grl_replay_push_rows(grl_env *env, const grl_replay_rows_io *io);

/* Draw the synthetic policy's actions for the current state into `actions`
 * ([B][max_actions], slot p = player p's move, empty when it has none). */
int grl_sample_actions(grl_env *env, uint64_t policy_seed, grl_action *actions);

int grl_get_state(grl_env *env, int32_t first_env, int32_t count, const grl_state_planes *out);
int grl_set_state(grl_env *env, int32_t first_env, int32_t count, const grl_state_planes *in);

/* 64-bit digest per env of the full game state (tiles, lists, sets, header); identical
 * function in the oracle, used for parity at sizes too large to copy back. out: [B]. */
int grl_state_hash(grl_env *env, uint64_t *out);
/* Digest of a float/word buffer laid out [rows][row_words] (device or host pointer),
 * one uint64 per row; used to compare observation planes at scale. */
int grl_buffer_hash(grl_env *env, const void *buf, size_t row_words, int32_t rows, uint64_t *out);

/* Episode statistics since create: [0] env-steps executed (not idle), [1] error-turn steps,
 * [2] games finished, [3] steps attempted on finished envs. */
int grl_stats(grl_env *env, uint64_t out[4]);

/* Number of CUDA kernels this env has launched since create (diagnostic; the oracle reports 0). */
int grl_launch_count(grl_env *env, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* GRLCUDA_H */
