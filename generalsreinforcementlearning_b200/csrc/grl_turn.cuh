// grl_turn.cuh — the fused turn kernel (one template, instantiated per board geometry in grl_turn_*.cu).
//
// Reference semantics (file:line into the reference tree):
//   turn order          internal/game/turn_processor.go:29-77,124-135
//   actions             internal/game/processor/action_processor.go:36-99,
//                       internal/game/core/action.go:56-105, core/movement.go:23-118
//   eliminations        internal/game/engine.go:80-152
//   production          internal/game/production_manager.go:26-101
//   cached lists/stats  internal/game/stats.go:8-144
//   fog of war          internal/game/visibility_optimized.go:16-163
//   win check           internal/game/rules/win_conditions.go:21-57
//   legal mask          internal/game/rules/legal_moves.go:19-73
//   observation         internal/experience/serializer.go:37-109
//   reward              internal/experience/rewards.go:45-175
#pragma once
#include <string.h>

#include "grl_device.cuh"
#include "grl_gym.cuh"
#include "grl_obs.cuh"

// ---------------------------------------------------------------------------------------
// The fused turn kernel.  DO_STEP: ProcessTurn.  DO_OUT: observation / mask / reward / done.
// TW/TH > 0 bake the board geometry in (the BASELINE sizes): loop trip counts, channel strides
// and x/y arithmetic become immediates.  TW == 0 reads the geometry from the parameter block.
// ---------------------------------------------------------------------------------------

// ---- rare / optional phases, deliberately NOT inlined: they talk to the kernel through the
// shared-memory slab, so the common path of the turn kernel stays small in the instruction cache.

// Synthetic policy for all players from the pre-turn state; writes decoded moves into s_act.
template <int PT, int LG>
__device__ __noinline__ void policy_phase(const GrlKParams &prm, uint32_t *s, const uint32_t *st, uint32_t *s_act,
                                          uint32_t alive, uint32_t turn_before, int game, Geo g, int W, int H, int N, int NW);

// The random agent of a fused gym step (GrlGymK.action == nullptr): the index grl_gym_sample(seed, mask, 0) returns for
// this game's CURRENT mask — the k-th set entry of player 0's N*5 mask bytes in index order (tile-major; up, right, down,
// left, half per tile), k = policy_draw(seed, env, 0, 0) mod the number of set entries; 0 (which the env rejects) when
// none is set — from the view's direction masks in registers (word `lane` of the group) instead of the mask bytes in
// HBM.  Only the GYM == 2 instantiations of the turn kernel contain it: the step that is handed its actions (GYM == 1) is
// the same code as before.  Group-uniform result.
template <int LG>
__device__ __noinline__ long long gym_random_agent(unsigned long long seed, unsigned long long env_global, uint32_t U, uint32_t R,
                                                   uint32_t D, uint32_t Lm, Geo g) {
  const uint32_t Hf = U | R | D | Lm;  // the half move: valid wherever a full move is (generals_env.py:380-383)
  const int cnt = __popc(U) + __popc(R) + __popc(D) + __popc(Lm) + __popc(Hf);
  const int total = __reduce_add_sync(g.seg, cnt);
  if (total == 0) return 0;
  const uint64_t rr = policy_draw(seed, env_global, 0ull, 0ull);
  const int k = (int)(rr % (uint64_t)total);
  int incl = cnt;  // inclusive prefix sum over the group's lanes
#pragma unroll
  for (int o = 1; o < LG; o <<= 1) {
    const int v = __shfl_up_sync(g.seg, incl, o, LG);
    if (g.lane >= o) incl += v;
  }
  const int excl = incl - cnt;
  const bool mine = k >= excl && k < incl;
  const int kk = k - excl;
  int b = 0;  // the tile bit of this word the kk-th entry belongs to: the largest b with count(bits < b) <= kk
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const int cand = b + step;
    const uint32_t m = (1u << cand) - 1u;  // cand in 1..31
    const int c = __popc(U & m) + __popc(R & m) + __popc(D & m) + __popc(Lm & m) + __popc(Hf & m);
    if (c <= kk) b = cand;
  }
  const uint32_t below = (1u << b) - 1u;
  int rem = kk - (__popc(U & below) + __popc(R & below) + __popc(D & below) + __popc(Lm & below) + __popc(Hf & below));
  const uint32_t five = ((U >> b) & 1u) | (((R >> b) & 1u) << 1) | (((D >> b) & 1u) << 2) | (((Lm >> b) & 1u) << 3) |
                        (((Hf >> b) & 1u) << 4);
  int dir = 0;
#pragma unroll
  for (int d = 0; d < 5; d++) {
    if ((five >> d) & 1u) {
      if (rem == 0) dir = d;
      rem--;
    }
  }
  int packed = mine ? (32 * g.lane + b) * 5 + dir : 0;
  const uint32_t who = __ballot_sync(g.seg, mine) >> g.shift;
  packed = __shfl_sync(g.seg, packed, __ffs(who) - 1, LG);
  return (long long)packed;
}

// Fused gym step, before the turn: decode the agent's (player 0) Discrete(N*5) index against the gym mask of the
// CURRENT state (client-side rejection, generals_env.py:226-229), then the opponent's index or the random
// opponent's draw; decoded moves go to s_act.  AGENT: the agent's index is not given but drawn here
// (gym_random_agent).  Returns whether the agent's action is valid.
template <int PT, int LG, bool AGENT>
__device__ __forceinline__ bool gym_pre_phase(const GrlKParams &prm, const GrlGymK &gk, uint32_t *s, const uint32_t *st,
                                              uint32_t *s_act, uint32_t alive, bool over, uint32_t turn_before, int game, Geo g,
                                              int W, int H, int N, int NW, long long a0, long long a1);
// Elimination orders: tile turnover over the eliminated player's cached list, then the stats
// rebuild of engine.go:101-109.  Reads and writes own/list/changed/vchg words in the slab.
template <int PT, int LG>
__device__ __noinline__ uint32_t elimination_phase(const GrlKParams &prm, uint32_t *s, const uint32_t *st, uint32_t alive,
                                                   int n_orders, uint32_t ord_lo, uint32_t ord_hi, Geo g, int N, int NW);

// One decoded move, staged in shared memory as two words.  The checks that depend only on the
// action itself (core/action.go:58-79) run on one lane per slot, in parallel; the checks that
// depend on the board (ownership, army, mountain) run in the serial phase.
//   word0: fi[0:10) ti[10:20) player[20:23) moveAll[23] staticErr[24:28) present[28]
//   word1: Serializer.ActionToIndex (serializer.go:179-198)
__device__ __forceinline__ uint2 decode_action(uint2 raw, int W, int H, int P) {
  PackedAction a;
  a.lo = raw.x;
  a.hi = raw.y;
  const int pid = a.player();
  if (!a.present() || pid < 0 || pid >= P) return make_uint2(0u, 0xffffffffu);  // action_processor.go:56-60
  const int fx = a.fx(), fy = a.fy(), tx = a.tx(), ty = a.ty();
  const int ddx = tx - fx, ddy = ty - fy;
  int dir = 0;
  if (ddy == -1 && ddx == 0) dir = 0;
  else if (ddy == 1 && ddx == 0) dir = 1;
  else if (ddy == 0 && ddx == -1) dir = 2;
  else if (ddy == 0 && ddx == 1) dir = 3;
  const int aidx = (fy * W + fx) * 4 + dir;
  uint32_t e = 0;
  if ((unsigned)fx >= (unsigned)W || (unsigned)fy >= (unsigned)H) e = GRL_STEP_INVALID_COORDINATES;
  else if ((unsigned)tx >= (unsigned)W || (unsigned)ty >= (unsigned)H) e = GRL_STEP_INVALID_COORDINATES;
  else if (ddx == 0 && ddy == 0) e = GRL_STEP_MOVE_TO_SELF;
  else if (!((ddx == 0 && (ddy == 1 || ddy == -1)) || (ddy == 0 && (ddx == 1 || ddx == -1)))) e = GRL_STEP_NOT_ADJACENT;
  const uint32_t fi = e == GRL_STEP_INVALID_COORDINATES ? 0u : (uint32_t)(fy * W + fx);
  const uint32_t ti = e == GRL_STEP_INVALID_COORDINATES ? 0u : (uint32_t)(ty * W + tx);
  uint32_t w = fi | (ti << 10) | ((uint32_t)pid << 20) | ((a.move_all() ? 1u : 0u) << 23) | (e << 24) | (1u << 28);
  return make_uint2(w, (uint32_t)aidx);
}

template <int PT, int LG>
struct TurnOccupancy {  // CTAs of 256 threads per SM the register budget is tuned for
  // 64 registers per thread: 4 CTAs of 256 threads per SM for two players
  static constexpr int kMinBlocks = PT <= 2 ? 4 : (PT <= 4 ? 3 : 2);
};

template <int PT, int TW, int TH, int LG, bool DO_STEP, bool DO_OUT, int GYM>
__device__ __forceinline__ void grl_turn_body(const GrlKParams &prm, const GrlGymK &gk) {
  static_assert(!GYM || (DO_STEP && DO_OUT), "the fused gym step is a turn plus read-outs");
  constexpr int GPW = 32 / LG;  // games per warp
  static_assert(LG == 32 || LG == 16 || LG == 8 || LG == 4, "a group is 4, 8, 16 or 32 lanes");
  static_assert(LG >= PT || LG == 32, "per-player scalars are written by one lane each");
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t s_bar[GRL_WARPS_PER_CTA * GPW];
  __shared__ __align__(256) float4 s_lut[16];  // nibble -> four 0/1 floats (observation planes)
  if (DO_OUT && threadIdx.x < 16) {
    const uint32_t n = threadIdx.x;
    s_lut[n] = make_float4((n & 1u) ? 1.f : 0.f, (n & 2u) ? 1.f : 0.f, (n & 4u) ? 1.f : 0.f, (n & 8u) ? 1.f : 0.f);
  }
  if (DO_OUT) __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GrlLayout &L = prm.L;
  const int P = prm.P;
  const int W = TW ? TW : prm.W;
  const int H = TW ? TH : prm.H;
  const int N = TW ? TW * TH : prm.N;
  const int NW = TW ? (TW * TH + 31) / 32 : prm.NW;
  const int act_words = 2 * GRL_MAX_ACTIONS;
  // per game: [slab | terrain | pre-turn snapshot of the slab | decoded action slots]
  // One game per warp keeps a pre-turn snapshot of its slab and writes back only the 32-byte sectors the turn changed
  // (about a quarter of them); packed groups write the whole slab back, which keeps four CTAs per SM in shared memory.
  constexpr bool kSnap = LG == 32;
  const int per_game = (kSnap ? 2 : 1) * L.slab_words + L.static_words + act_words;
  // baked geometries with N % 4 != 0 stage channel masks + an army-fraction plane per warp (obs_linear)
  const int obs_scratch = GYM ? grl_gym_smem_words(P, NW, N, grl_gym_emit_mode(TW, TH)) : grl_obs_scratch_words(TW, TH, PT, NW);
  uint32_t *wbase = smem + warp * (GPW * per_game + obs_scratch);
  uint32_t *s_obs = wbase + GPW * per_game;

  const Geo g = make_geo(prm, W, lane, LG);
  const int l = g.lane;            // lane inside the group
  const int sub = g.shift / LG;    // group inside the warp
  uint32_t *s = wbase + sub * per_game;
  uint32_t *st = s + L.slab_words;
  uint32_t *snap = st + L.static_words;
  uint32_t *s_act = snap + (kSnap ? L.slab_words : 0);
  const bool act_lane = l < NW;
  const uint32_t pmask = (1u << P) - 1u;
  const bool use_policy = DO_STEP && (prm.flags & GRL_STEP_FLAG_RANDOM_POLICY) != 0;
  const bool read_actions = DO_STEP && !GYM && !use_policy && prm.actions != nullptr;
  const int warp_game0 = prm.game0 + (blockIdx.x * GRL_WARPS_PER_CTA + warp) * GPW;
  const int game = warp_game0 + sub;
  const int game_end = prm.game_end;
  const bool gv = game < game_end;  // uniform over the group

  uint64_t *bar = &s_bar[warp * GPW + sub];
  if (l == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp(g.seg);

  SlabView S = make_view(s, st, L);
  uint32_t own[PT], lst[PT], vis[PT];
#pragma unroll
  for (int p = 0; p < PT; p++) own[p] = lst[p] = vis[p] = 0u;
  uint32_t M = 0u, alive = 0u, err = 0u;
  bool over = false;

  if (gv) {
    uint32_t *gslab = prm.state + (size_t)game * L.slab_words;
    const uint32_t *gstat = prm.statics + (size_t)game * L.static_words;

    // ---- stage the slab in shared memory ---------------------------------------------
    if (l == 0) {
      const bool want_snap = kSnap && DO_STEP;
      mbar_expect_tx(bar, (uint32_t)(L.slab_words + L.static_words + (want_snap ? L.slab_words : 0)) * 4u);
      tma_load(s, gslab, (uint32_t)L.slab_words * 4u, bar);
      tma_load(st, gstat, (uint32_t)L.static_words * 4u, bar);
      if (want_snap)  // pre-turn snapshot for the dirty-sector write-back (an L2 hit on the same lines)
        tma_load(snap, gslab, (uint32_t)L.slab_words * 4u, bar);
      // warm L2 for a CTA that will be scheduled a couple of waves from now
      if (prm.prefetch_dist > 0 && game + prm.prefetch_dist < prm.B) {
        tma_prefetch_l2(prm.state + (size_t)(game + prm.prefetch_dist) * L.slab_words, (uint32_t)L.slab_words * 4u);
        tma_prefetch_l2(prm.statics + (size_t)(game + prm.prefetch_dist) * L.static_words, (uint32_t)L.static_words * 4u);
      }
    }
    // decode this game's action slots while its slab lands
    long long gym_a0 = 0, gym_a1 = 0;  // the fused gym step's Discrete(N*5) indices: fetched now, tested after the wait
    int gym_tn = 0, gym_cl = 0;        // the env's turn / call counters: read-modify-write at the end, loaded now
    if constexpr (GYM) {
      if constexpr (GYM == 1) gym_a0 = __ldg(gk.action + game);  // GYM == 2: the random agent, drawn in gym_pre_phase
      if (gk.opponent_action) gym_a1 = __ldg(gk.opponent_action + game);
      if (l == 0) {
        gym_tn = gk.turns[game];
        gym_cl = gk.calls[game];
      }
    }
    bool skip = false;  // GRL_ACTION_FLAG_SKIP_ENV on slot 0: this env takes no turn in this call
    if (DO_STEP) {
      uint32_t slot0_hi = 0u;
      for (int sl = l; sl < GRL_MAX_ACTIONS; sl += LG) {
        uint2 d = make_uint2(0u, 0xffffffffu);
        if (read_actions && sl < prm.A) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(prm.actions) + (size_t)game * prm.A + sl);
          if (sl == 0) slot0_hi = raw.y;
          d = decode_action(raw, W, H, P);
        }
        s_act[2 * sl] = d.x;
        s_act[2 * sl + 1] = d.y;
      }
      if (read_actions) skip = ((__shfl_sync(g.seg, slot0_hi, 0, LG) >> 24) & GRL_ACTION_FLAG_SKIP_ENV) != 0u;
    }
    mbar_wait(bar, 0u);
    __syncwarp(g.seg);

    // ---- mask words into registers -----------------------------------------------------
    uint32_t own_prev[PT];
#pragma unroll
    for (int p = 0; p < PT; p++) {
      bool on = act_lane && p < P;
      own[p] = on ? S.own[p * NW + l] : 0u;
      lst[p] = on ? S.list[p * NW + l] : 0u;
      vis[p] = on ? S.vis[p * NW + l] : 0u;
      own_prev[p] = own[p];
    }
    uint32_t chg = act_lane ? S.chg[l] : 0u;
    uint32_t vch = act_lane ? S.vch[l] : 0u;
    M = act_lane ? S.M[l] : 0u;
    const uint32_t C = act_lane ? S.C[l] : 0u;
    const uint32_t G = act_lane ? S.G[l] : 0u;

    uint32_t turn = S.hdr[GRL_HDR_TURN];
    uint32_t flags = S.hdr[GRL_HDR_FLAGS];
    alive = flags & 0xffu;
    over = (flags & GRL_FLAG_OVER) != 0;
    bool stepped = false;
    int prev_true_army[PT];
#pragma unroll
    for (int p = 0; p < PT; p++)
      prev_true_army[p] = p < P ? (int)S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_TRUE_ARMY] : 0;

    // ---- fused gym step: the client's pre-turn PlayerState (reward baseline) and its action decoding ----
    int gym_army0 = 0, gym_tiles0 = 0;
    uint32_t gym_alive0 = 0u;
    if constexpr (GYM) {
      gym_army0 = (int)S.hdr[GRL_HDR_PLAYER0 + GRL_PL_ARMY_COUNT];
      gym_tiles0 = __reduce_add_sync(g.seg, __popc(lst[0]));
      gym_alive0 = alive;
      skip = !gym_pre_phase<PT, LG, GYM == 2>(prm, gk, s, st, s_act, alive, over, turn, game, g, W, H, N, NW, gym_a0, gym_a1);
    }

    if (DO_STEP) {
      if (skip) {
        err = (flags >> GRL_FLAG_ERR_SHIFT) & 0xffu;  // Step is not called: everything stays as it was
      } else if (over) {
        // turn_processor.go:95-113: ErrGameOver, nothing mutated
        err = GRL_STEP_GAME_OVER;
        if (l == 0) {
          S.hdr[GRL_HDR_REJECTED] += 1;
          for (int p = 0; p < P; p++) {
            S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_REWARD] = 0u;
            S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ACTION_INDEX] = 0xffffffffu;
          }
        }
      } else {
        stepped = true;
        const uint32_t turn_before = turn;
        turn += 1;  // turn_processor.go:125

        // ---- the synthetic policy reads the PRE-turn state (all players at once) ----------
        if (use_policy) policy_phase<PT, LG>(prm, s, st, s_act, alive, turn_before, game, g, W, H, N, NW);

        // ---- fog of war, from LAST turn's vchg and the CURRENT lists (Q1) -----------------
        if (prm.fog) {
          int nv = __reduce_add_sync(g.seg, __popc(vch));
          if (nv > N / 10) {  // visibility_optimized.go:22-25 -> full :33-53
#pragma unroll
            for (int p = 0; p < PT; p++)
              if (p < P) vis[p] = ((alive >> p) & 1u) ? dilate3<LG>(lst[p], g) : 0u;
          } else if (nv > 0) {  // incremental :56-97
            uint32_t d3 = dilate3<LG>(vch, g);
            uint32_t d5 = dilate3<LG>(d3, g);
#pragma unroll
            for (int p = 0; p < PT; p++) {
              if (p < P) {
                bool affected = __any_sync(g.seg, (own[p] & d5) != 0u);  // owners read NOW (:100-115)
                vis[p] &= ~d3;                                           // all players' bits cleared (:131-149)
                if (affected && ((alive >> p) & 1u)) vis[p] |= dilate3<LG>(lst[p], g);
              }
            }
          }
        }
        // turn_processor.go:129-134
        chg = 0u;
        vch = 0u;
        if (act_lane) {
          S.chg[l] = 0u;
          S.vch[l] = 0u;
        }
        if (l < P) S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * l + GRL_PL_ACTION_INDEX] = 0xffffffffu;
        __syncwarp(g.seg);

        // ---- actions: serial by definition, one lane, on the shared-memory slab -----------
        uint32_t ord_lo = 0, ord_hi = 0;  // up to 8 orders, one byte each: eliminated | capturer<<4
        int n_orders = 0;
        if (l == 0) {
          uint32_t processed = 0;
          uint32_t overflow = 0;
          const uint32_t alive_start = alive;  // action_processor.go:56-60 reads Alive as of now
          uint32_t *ta = &S.hdr[GRL_HDR_PLAYER0 + GRL_PL_TRUE_ARMY];
          // stable sort by player id == for each id ascending, slots in submission order
          for (int p = 0; p < P; p++) {
            for (int sl = 0; sl < prm.A; sl++) {
              const uint32_t w = s_act[2 * sl];
              if (!(w >> 28) || (int)((w >> 20) & 7u) != p) continue;
              // collectExperiences keys the action map by player: the last submission wins
              S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ACTION_INDEX] = s_act[2 * sl + 1];
              if (!((alive_start >> p) & 1u)) continue;
              uint32_t e = (w >> 24) & 0xfu;
              const int fi = (int)(w & 1023u), ti = (int)((w >> 10) & 1023u);
              uint32_t a_from = 0;
              if (!e) {  // board-dependent half of core/action.go:56-105 Validate
                if (!((S.own[p * NW + (fi >> 5)] >> (fi & 31)) & 1u)) e = GRL_STEP_NOT_OWNED;
                else {
                  a_from = S.army[fi];
                  if (a_from <= 1u) e = GRL_STEP_INSUFFICIENT_ARMY;
                  else if ((S.M[ti >> 5] >> (ti & 31)) & 1u) e = GRL_STEP_TARGET_IS_MOUNTAIN;
                }
              }
              if (e) {
                if (!err) err = e;  // first error remembered, processing continues (:66-77)
                continue;
              }
              // core/movement.go:23-89 ApplyMoveAction
              const bool move_all = (w >> 23) & 1u;
              uint32_t moved = move_all ? a_from - 1u : (a_from / 2u == 0u ? 1u : a_from / 2u);
              S.army[fi] = (uint16_t)(a_from - moved);
              S.chg[fi >> 5] |= 1u << (fi & 31);
              const int tw = ti >> 5;
              const uint32_t tb = 1u << (ti & 31);
              S.chg[tw] |= tb;
              uint32_t a_to = S.army[ti];
              ta[GRL_HDR_PER_PLAYER * p] -= moved;
              if (S.own[p * NW + tw] & tb) {
                uint32_t sum = a_to + moved;
                if (sum > 65535u) {
                  sum = 65535u;
                  overflow = 1;
                }
                S.army[ti] = (uint16_t)sum;
                ta[GRL_HDR_PER_PLAYER * p] += sum - a_to;
              } else {
                int q = -1;
                for (int r = 0; r < P; r++)
                  if (S.own[r * NW + tw] & tb) q = r;
                if (moved > a_to) {  // ties defend (movement.go:69)
                  if (q >= 0) {
                    S.own[q * NW + tw] &= ~tb;
                    ta[GRL_HDR_PER_PLAYER * q] -= a_to;
                  }
                  S.own[p * NW + tw] |= tb;
                  S.army[ti] = (uint16_t)(moved - a_to);
                  ta[GRL_HDR_PER_PLAYER * p] += moved - a_to;
                  S.vch[tw] |= tb;  // action_processor.go:78-87
                  // core.ProcessCaptures movement.go:100-118: first capture of a player's general wins
                  if ((S.G[tw] & tb) && q >= 0 && !((processed >> q) & 1u) && n_orders < 8) {
                    uint32_t o = (uint32_t)q | ((uint32_t)p << 4);
                    if (n_orders < 4) ord_lo |= o << (8 * n_orders);
                    else ord_hi |= o << (8 * (n_orders - 4));
                    n_orders++;
                    processed |= 1u << q;
                  }
                } else {
                  S.army[ti] = (uint16_t)(a_to - moved);
                  if (q >= 0) ta[GRL_HDR_PER_PLAYER * q] -= moved;
                }
              }
            }
          }
          if (overflow) S.hdr[GRL_HDR_OVERFLOW] = 1u;
        }
        __syncwarp(g.seg);
        err = __shfl_sync(g.seg, err, 0, LG);
        n_orders = __shfl_sync(g.seg, n_orders, 0, LG);
#pragma unroll
        for (int p = 0; p < PT; p++)
          if (p < P && act_lane) own[p] = S.own[p * NW + l];
        if (act_lane) {
          chg = S.chg[l];
          vch = S.vch[l];
        }

        // ---- eliminations + tile turnover over the CACHED list (engine.go:118-152) --------
        if (n_orders > 0) {  // rare: kept out of line so the common path stays compact
          ord_lo = __shfl_sync(g.seg, ord_lo, 0, LG);
          ord_hi = __shfl_sync(g.seg, ord_hi, 0, LG);
          alive = elimination_phase<PT, LG>(prm, s, st, alive, n_orders, ord_lo, ord_hi, g, N, NW);
#pragma unroll
          for (int p = 0; p < PT; p++) {
            if (p < P && act_lane) {
              own[p] = S.own[p * NW + l];
              lst[p] = S.list[p * NW + l];
            }
          }
          if (act_lane) {
            chg = S.chg[l];
            vch = S.vch[l];
          }
        }

        if (err == 0) {
          // ---- production over the cached lists (production_manager.go:26-101) ------------
          uint32_t AL = 0;
#pragma unroll
          for (int p = 0; p < PT; p++)
            if (p < P && ((alive >> p) & 1u)) AL |= lst[p];
          const bool grow = prm.grow_interval == 25 ? (turn % 25u) == 0u : (turn % (uint32_t)prm.grow_interval) == 0u;
          uint32_t PG = prm.pg > 0 ? (AL & G) : 0u;
          uint32_t PC = prm.pc > 0 ? (AL & C) : 0u;
          uint32_t PN = (grow && prm.pn > 0) ? (AL & ~(G | C | M)) : 0u;
          uint32_t produced = PG | PC | PN;
          chg |= produced;
          if (__any_sync(g.seg, produced != 0u)) {
            uint32_t overflow = 0;
            if (grow) {  // dense: most owned tiles grow (1 turn in 25)
#pragma unroll 1
              for (int i = 0; i < NW; i++) {
                uint32_t wg = __shfl_sync(g.seg, PG, i, LG), wc = __shfl_sync(g.seg, PC, i, LG), wn = __shfl_sync(g.seg, PN, i, LG);
#pragma unroll
                for (int b = l; b < 32; b += LG) {
                  int t = 32 * i + b;
                  uint32_t add = (((wg >> b) & 1u) ? (uint32_t)prm.pg : 0u) + (((wc >> b) & 1u) ? (uint32_t)prm.pc : 0u) +
                                 (((wn >> b) & 1u) ? (uint32_t)prm.pn : 0u);
                  if (add) {
                    uint32_t a = (uint32_t)S.army[t] + add;
                    if (a > 65535u) {
                      a = 65535u;
                      overflow = 1;
                    }
                    S.army[t] = (uint16_t)a;
                  }
                }
              }
            } else {  // sparse: generals and cities only
              uint32_t w = produced;
              while (w) {
                int b = __ffs(w) - 1;
                w &= w - 1u;
                int t = 32 * l + b;
                uint32_t add = ((PG >> b) & 1u) ? (uint32_t)prm.pg : (uint32_t)prm.pc;
                uint32_t a = (uint32_t)S.army[t] + add;
                if (a > 65535u) {
                  a = 65535u;
                  overflow = 1;
                }
                S.army[t] = (uint16_t)a;
              }
            }
            if (__any_sync(g.seg, overflow != 0u) && l == 0) S.hdr[GRL_HDR_OVERFLOW] = 1u;
#pragma unroll
            for (int p = 0; p < PT; p++) {
              if (p < P) {
                int d = prm.pg * __popc(PG & own[p]) + prm.pc * __popc(PC & own[p]) + prm.pn * __popc(PN & own[p]);
                d = __reduce_add_sync(g.seg, d);
                if (l == 0) S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_TRUE_ARMY] += (uint32_t)d;
              }
            }
            __syncwarp(g.seg);
          }
          // ---- end of turn: stats, game over (turn_processor.go:170-179) --------------------
          stats_update<PT, LG>(prm, S, own, lst, chg, G, alive, g, N, NW);
          int n_alive = __popc(alive & pmask);
          bool now_over = P > 1 ? (n_alive <= 1) : (n_alive == 0);  // win_conditions.go:38-44
          if (now_over && !over && l == 0) S.hdr[GRL_HDR_FINISHED] += 1;
          over = now_over;
        }
        if (l == 0) {
          S.hdr[GRL_HDR_STEPS] += 1;
          if (err) S.hdr[GRL_HDR_ERRORS] += 1;
        }
      }
    }
    const uint32_t turn_err = err;  // the reference's validation error (or game over)
    if (DO_STEP && stepped && err == 0 && S.hdr[GRL_HDR_OVERFLOW]) err = GRL_STEP_ARMY_OVERFLOW;

    // ---- reward: CalculateRewardWithConfig(prev, curr, p) (rewards.go:45-85) -------------
    if (DO_STEP && stepped) {
      int n_alive = __popc(alive & pmask);
      int sole = n_alive == 1 ? (__ffs(alive & pmask) - 1) : -1;
      int total_army = 0;
#pragma unroll
      for (int p = 0; p < PT; p++)
        if (p < P) total_army += (int)S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_TRUE_ARMY];
      uint32_t any_prev = 0;
#pragma unroll
      for (int p = 0; p < PT; p++) any_prev |= own_prev[p];
#pragma unroll
      for (int p = 0; p < PT; p++) {
        if (p < P) {
          int d_tiles = __reduce_add_sync(g.seg, __popc(own[p]) - __popc(own_prev[p]));
          uint32_t gained = own[p] & ~own_prev[p], lost = own_prev[p] & ~own[p];
          int cc = 0, gg = 0;
          if (__any_sync(g.seg, ((gained | lost) & (C | G)) != 0u)) {
            cc = __reduce_add_sync(g.seg, __popc(gained & C) | (__popc(lost & C) << 16));
            gg = __reduce_add_sync(g.seg, __popc(gained & G & any_prev) | (__popc(lost & G) << 16));
          }
          int cur_army = (int)S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_TRUE_ARMY];
          float r = 0.0f;
          bool terminal = false;
          if (n_alive <= 1) {  // state.go:73-100 on the current state
            if (sole == p) {
              r = prm.rw[0];
              terminal = true;
            } else if (sole != -1) {
              r = prm.rw[1];
              terminal = true;
            }
          }
          if (!terminal) {  // one rounding per Go statement, no FMA contraction (Q12)
            r = __fadd_rn(r, __fmul_rn((float)d_tiles, prm.rw[6]));
            r = __fadd_rn(r, __fmul_rn((float)(cur_army - prev_true_army[p]), prm.rw[8]));
            r = __fadd_rn(r, __fmul_rn((float)(cc & 0xffff), prm.rw[2]));
            r = __fadd_rn(r, __fmul_rn((float)(cc >> 16), prm.rw[3]));
            r = __fadd_rn(r, __fmul_rn((float)(gg & 0xffff), prm.rw[4]));
            r = __fadd_rn(r, __fmul_rn((float)(gg >> 16), prm.rw[5]));
            float adv = 0.0f;
            if (total_army != 0) adv = __fdiv_rn((float)(cur_army - (total_army - cur_army)), (float)total_army);
            r = __fadd_rn(r, __fmul_rn(adv, prm.rw[10]));
          }
          if (l == 0) S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_REWARD] = __float_as_uint(r);
        }
      }
      if (turn_err != 0 && l == 0)  // aborted turn: no experience is emitted (engine.go:111-113)
        for (int p = 0; p < P; p++) S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ACTION_INDEX] = 0xffffffffu;
    }

    // ---- write the state back ---------------------------------------------------------------
    if (DO_STEP) {
      if (stepped) {
        if (act_lane) {
#pragma unroll
          for (int p = 0; p < PT; p++) {
            if (p < P) {
              S.own[p * NW + l] = own[p];
              S.list[p * NW + l] = lst[p];
              S.vis[p * NW + l] = vis[p];
            }
          }
          S.chg[l] = chg;
          S.vch[l] = vch;
        }
        if (l == 0) S.hdr[GRL_HDR_TURN] = turn;
      }
      if (l == 0)
        S.hdr[GRL_HDR_FLAGS] = (alive & 0xffu) | (over ? GRL_FLAG_OVER : 0u) | (err << GRL_FLAG_ERR_SHIFT);
      __syncwarp(g.seg);
      if (kSnap) {
        // sector k = words [8k, 8k+8) of the slab (slabs are 32-byte aligned and a whole number of sectors)
        const uint4 *now4 = reinterpret_cast<const uint4 *>(s);
        const uint4 *old4 = reinterpret_cast<const uint4 *>(snap);
        uint4 *dst4 = reinterpret_cast<uint4 *>(gslab);
        for (int k = l; k < L.slab_words / 8; k += LG) {
          const uint4 a0 = now4[2 * k], a1 = now4[2 * k + 1], b0 = old4[2 * k], b1 = old4[2 * k + 1];
          const uint32_t diff = (a0.x ^ b0.x) | (a0.y ^ b0.y) | (a0.z ^ b0.z) | (a0.w ^ b0.w) | (a1.x ^ b1.x) | (a1.y ^ b1.y) |
                                (a1.z ^ b1.z) | (a1.w ^ b1.w);
          if (diff) {
            dst4[2 * k] = a0;
            dst4[2 * k + 1] = a1;
          }
        }
      } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(s);
        uint4 *dst = reinterpret_cast<uint4 *>(gslab);
        for (int k = l; k < L.slab_words / 4; k += LG) dst[k] = src[k];
      }
    } else {
      err = (flags >> GRL_FLAG_ERR_SHIFT) & 0xffu;
    }

    // ---- scalar read-outs (one lane per value) -------------------------------------------------
    if (DO_OUT) {
      if (l == 0) {
        if (prm.done) prm.done[game] = over ? 1 : 0;
        if (prm.winner) {  // engine.go:248-263
          int n_alive = __popc(alive & pmask);
          prm.winner[game] = (int8_t)((over && n_alive == 1) ? (__ffs(alive & pmask) - 1) : -1);
        }
        if (prm.step_error) prm.step_error[game] = (uint8_t)err;
      }
      if (l < P) {
        if (prm.reward)
          prm.reward[(size_t)game * P + l] = __uint_as_float(S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * l + GRL_PL_REWARD]);
        if (prm.action_index)
          prm.action_index[(size_t)game * P + l] =
              (int32_t)S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * l + GRL_PL_ACTION_INDEX];
      }
    }

    // ---- fused gym step: the tail of GeneralsEnv.step (generals_env.py:268-289) and the client's reward
    //      (:499-561, float64) from the PlayerState before and after the turn --------------------------
    if constexpr (GYM) {
      const int tiles1 = __reduce_add_sync(g.seg, act_lane ? __popc(S.list[l]) : 0);
      if (l == 0) {
        const int valid = skip ? 0 : 1;
        const int tn = gym_tn + valid, cl = gym_cl + 1;
        gk.turns[game] = tn;
        gk.calls[game] = cl;
        const bool term = over && valid;
        const bool trunc = (tn >= gk.max_turns && valid) || cl >= gk.max_turns;
        double r = 0.0;
        if (!valid) {
          r = -0.1;
        } else if (term) {
          const int n_alive = __popc(alive & pmask);
          r = (n_alive == 1 && (alive & 1u)) ? 100.0 : -100.0;
        } else {
          // one rounding per Python statement (generals_env.py:523-547): no FMA contraction
          r = __dadd_rn(r, __dmul_rn((double)(tiles1 - gym_tiles0), 1.0));
          r = __dadd_rn(r, __dmul_rn((double)((int)S.hdr[GRL_HDR_PLAYER0 + GRL_PL_ARMY_COUNT] - gym_army0), 0.01));
          for (int q = 1; q < P; q++)
            if (((gym_alive0 >> q) & 1u) && !((alive >> q) & 1u)) r = __dadd_rn(r, 50.0);
        }
        gk.reward[game] = r;
        gk.valid[game] = (uint8_t)valid;
        gk.terminated[game] = term ? 1 : 0;
        gk.truncated[game] = trunc ? 1 : 0;
        if ((term || trunc) && gk.n_finished) atomicAdd(gk.n_finished, 1);
      }
    }
  }
  if (!DO_OUT) return;
  __syncwarp();  // every group's slab in shared memory is final: the plane read-outs below are warp-wide

  if constexpr (GYM) {
    // the client's read-outs of the new state, one game of the warp after the other (obs, N*5 mask, PlayerState)
    const Geo g32 = make_geo(prm, W, lane, 32);
    if constexpr (TW > 0 && ((TW * TH) & 3) != 0 && GPW == 4) {
      // four whole games, every player slot in use: the observation blocks are one 16-byte aligned run with a
      // compile-time schedule (gym_run_game_obs); the mask bytes and PlayerStates follow game by game
      if (gk.obs && P == PT && (warp_game0 & 3) == 0 && warp_game0 + GPW <= game_end) {
        constexpr int NT = TW > 0 ? TW * TH : 5;
        constexpr int DW = ((PT * NT + 31) / 32 + 1 + 3) & ~3, FW = (2 * NT + 8 + 3) & ~3;
        CtLane c;
        c.init(s_obs + 4 * DW + FW, reinterpret_cast<const float *>(s_obs + 4 * DW),
               gk.obs + (size_t)warp_game0 * (PT * GRL_GYM_CHANNELS * NT), lane);
#pragma unroll 1
        for (int gi = 0; gi < GPW; gi++) {  // one copy of the pass, four games
          const uint32_t *sg = wbase + gi * per_game;
          gym_run_game_obs<PT, NT>(prm, gk.max_turns, gk.logtab, sg, sg + L.slab_words, c, s_obs, lane, g32, gi);
        }
#pragma unroll 1
        for (int gi = 0; gi < GPW; gi++) {  // the mask bytes and PlayerStates (F and the stream are their staging area now)
          const uint32_t *sg = wbase + gi * per_game;
          gym_emit_linear<PT, NT>(prm, gk.max_turns, gk.logtab, nullptr, gk.mask, gk.stats, sg, sg + L.slab_words, s_lut, s_obs,
                                  warp_game0 + gi, lane, g32, false, false);
        }
        return;
      }
    }
#pragma unroll 1
    for (int gi = 0; gi < GPW; gi++) {
      const int game_g = warp_game0 + gi;
      if (game_g >= game_end) break;
      const uint32_t *sg = wbase + gi * per_game;
      if constexpr (TW > 0 && ((TW * TH) & 3) == 0)
        gym_emit_quads<PT, (TW > 0 && ((TW * TH) & 3) == 0 ? TW * TH : 4)>(prm, gk.max_turns, gk.logtab, gk.obs, gk.mask, gk.stats, sg,
                                                                          sg + L.slab_words, s_lut, s_obs, game_g, lane, g32);
      else if constexpr (TW > 0)
        gym_emit_linear<PT, (TW > 0 ? TW * TH : 5)>(prm, gk.max_turns, gk.logtab, gk.obs, gk.mask, gk.stats, sg, sg + L.slab_words,
                                                    s_lut, s_obs, game_g, lane, g32, gi > 0, gi + 1 < GPW && game_g + 1 < game_end);
      else
        gym_emit<0>(prm, gk.max_turns, gk.logtab, gk.obs, gk.mask, gk.stats, sg, sg + L.slab_words, s_obs, game_g, lane, g32);
    }
    if (lane == 0) tma_store_wait_read();  // bulk stores of the mask bytes: the stage stays valid until they have read it
    return;
  }

  // packed observation records (include/grlcuda.h): everything StateToTensor reads, as the slab holds it — for consumers
  // in host memory, which expand them with grl_expand_obs (1/26 of the fp32 planes' bytes across PCIe at 20x20)
  if (prm.obs_packed) {
    const int PNW = P * NW, RW = grl_packed_words(L), AW = RW - 2 * PNW - 2 * NW;
#pragma unroll 1
    for (int gi = 0; gi < GPW; gi++) {
      const int game_g = warp_game0 + gi;
      if (game_g >= game_end) break;
      const uint32_t *sg = wbase + gi * per_game, *stg = sg + L.slab_words;
      uint32_t *rec = prm.obs_packed + (size_t)game_g * RW;
      for (int k = lane; k < PNW; k += 32) {
        __stcs(rec + k, sg[L.off_own + k]);
        __stcs(rec + PNW + k, prm.fog ? sg[L.off_vis + k] : prm.geom[k % NW]);
      }
      for (int k = lane; k < NW; k += 32) {
        __stcs(rec + 2 * PNW + k, stg[k]);
        __stcs(rec + 2 * PNW + NW + k, stg[NW + k] | stg[2 * NW + k]);
      }
      for (int k = lane; k < AW; k += 32) __stcs(rec + 2 * PNW + 2 * NW + k, k < L.NA / 2 ? sg[L.off_army + k] : 0u);
    }
  }

  // engine legal-action mask, packed in the reference's flat index order (t*4 + dir, U,R,D,L)
  if (prm.mask_bits) {
    const int words = (4 * N + 31) / 32;
    if (gv) {
      const uint32_t gt1 = army_gt1_mask<LG>(S.army, NW, N, g);
      const DirMasks dm = dir_targets<LG>(M, g);
#pragma unroll
      for (int p = 0; p < PT; p++) {
        if (p < P) {
          const uint32_t src = ((alive >> p) & 1u) ? (lst[p] & own[p] & gt1) : 0u;
          const uint32_t U = src & dm.up, R = src & dm.right, D = src & dm.down, Lm = src & dm.left;
          uint32_t *dst = prm.mask_bits + ((size_t)game * P + p) * words;
          if (LG == 32) {
            for (int k0 = 0; k0 < words; k0 += 32) {
              int k = k0 + l;  // output word k covers tiles 8k..8k+7 = byte k&3 of mask word k>>2
              int srcl = (k >> 2) & 31, sh = (k & 3) * 8;
              uint32_t bu = __shfl_sync(FULL, U, srcl) >> sh, br = __shfl_sync(FULL, R, srcl) >> sh;
              uint32_t bd = __shfl_sync(FULL, D, srcl) >> sh, bl = __shfl_sync(FULL, Lm, srcl) >> sh;
              uint32_t w = spread8(bu) | (spread8(br) << 1) | (spread8(bd) << 2) | (spread8(bl) << 3);
              if (k < words) __stcs(dst + k, w);
            }
          } else {  // packed groups: a lane expands its own word into output words 4l..4l+3
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int k = 4 * l + q, sh = 8 * q;
              const uint32_t w = spread8(U >> sh) | (spread8(R >> sh) << 1) | (spread8(D >> sh) << 2) | (spread8(Lm >> sh) << 3);
              if (act_lane && k < words) __stcs(dst + k, w);
            }
          }
        }
      }
    }
    __syncwarp();
  }

  // observation planes: Serializer.StateToTensor (serializer.go:37-109)
  if (prm.obs) {
    if (TW > 0) {
      // baked boards: the whole warp writes one game's block after the other, from the slabs in shared memory
      if constexpr (TW > 0 && ((TW * TH) & 3) != 0 && GPW == 4) {
        // four whole games, every player slot in use: one 16-byte aligned run with a compile-time schedule
        if (P == PT && (warp_game0 & 3) == 0 && warp_game0 + GPW <= game_end) {
          constexpr int NT = TW > 0 ? TW * TH : 5;
          CtLane c;
          c.init(s_obs + grl_obs_region_words(NT), reinterpret_cast<const float *>(s_obs),
                 prm.obs + (size_t)warp_game0 * (PT * GRL_OBS_CHANNELS * NT), lane);
          uint4 carry = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
          for (int gi = 0; gi < GPW; gi++)  // one copy of the pass, four games
            obs_run_game<PT, NT>(prm, make_view(wbase + gi * per_game, wbase + gi * per_game + L.slab_words, L), c, s_obs, NW, lane,
                                 gi, carry);
          return;
        }
      }
#pragma unroll 1
      for (int gi = 0; gi < GPW; gi++) {
        const int game_g = warp_game0 + gi;
        if (game_g >= game_end) break;
        uint32_t *sg = wbase + gi * per_game;
        const SlabView Sg = make_view(sg, sg + L.slab_words, L);
        if (((TW * TH) & 3) == 0)
          obs_plane_major<PT, (TW > 0 && ((TW * TH) & 3) == 0 ? TW * TH : 4)>(prm, Sg, s_lut, P, NW, game_g, lane);
        else
          obs_linear<PT, (TW > 0 ? TW * TH : 5)>(prm, Sg, s_lut, s_obs, P, NW, game_g, lane,
                                                 gi > 0 ? wbase + (gi - 1) * per_game : nullptr,
                                                 gi + 1 < GPW && game_g + 1 < game_end);
      }
    } else if (gv) {
      // generic geometries (LG == 32: one game per warp), from the mask words in registers
      uint32_t any_own = 0;
#pragma unroll
      for (int p = 0; p < PT; p++) any_own |= own[p];
      const uint32_t C = act_lane ? S.C[l] : 0u;
      const uint32_t G = act_lane ? S.G[l] : 0u;
      const uint32_t CG = C | G;
      float *gbase = prm.obs + (size_t)game * P * GRL_OBS_CHANNELS * N;
      if ((N & 3) == 0) {
        // 128-bit path: a lane writes 4 consecutive tiles of each channel plane; the army
        // conversion and the terrain nibbles are shared by all players' views
        const int cs = N / 4;  // channel stride in float4
        for (int q0 = 0; q0 * 4 < N; q0 += 32) {
          const int q = q0 + lane;
          const int t0 = 4 * q;
          const int srcl = (t0 >> 5) & 31, sh = t0 & 31;
          const uint32_t mM = (__shfl_sync(FULL, M, srcl) >> sh) & 0xfu;
          const uint32_t mCG = (__shfl_sync(FULL, CG, srcl) >> sh) & 0xfu;
          const uint32_t mAny = (__shfl_sync(FULL, any_own, srcl) >> sh) & 0xfu;
          const bool live = t0 < N;
          float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
          if (__any_sync(FULL, live && (mAny & ~mM) != 0u)) {
            if (live) {
              const uint2 aw = *reinterpret_cast<const uint2 *>(S.army + t0);
              f0 = army_frac(aw.x & 0xffffu);
              f1 = army_frac(aw.x >> 16);
              f2 = army_frac(aw.y & 0xffffu);
              f3 = army_frac(aw.y >> 16);
            }
          }
#define NIBF(n) make_float4(((n)&1u) ? 1.f : 0.f, ((n)&2u) ? 1.f : 0.f, ((n)&4u) ? 1.f : 0.f, ((n)&8u) ? 1.f : 0.f)
#define NIBA(n) make_float4(((n)&1u) ? f0 : 0.f, ((n)&2u) ? f1 : 0.f, ((n)&4u) ? f2 : 0.f, ((n)&8u) ? f3 : 0.f)
#pragma unroll
          for (int p = 0; p < PT; p++) {
            if (p < P) {
              const uint32_t nV = prm.fog ? ((__shfl_sync(FULL, vis[p], srcl) >> sh) & 0xfu) : 0xfu;
              const uint32_t nO = (__shfl_sync(FULL, own[p], srcl) >> sh) & 0xfu;
              float4 *o = reinterpret_cast<float4 *>(gbase + (size_t)p * GRL_OBS_CHANNELS * N + t0);
              if (live) {
                const uint32_t nm = nV & ~mM;
                const uint32_t n2 = nm & nO, n3 = nm & mAny & ~nO, n4 = nm & ~mAny, n5 = nm & mCG, n6 = nV & mM;
                const uint32_t n7 = nV, n8 = nV ^ 0xfu;
                __stcs(o + 0 * cs, NIBA(n2));
                __stcs(o + 1 * cs, NIBA(n3));
                __stcs(o + 2 * cs, NIBF(n2));
                __stcs(o + 3 * cs, NIBF(n3));
                __stcs(o + 4 * cs, NIBF(n4));
                __stcs(o + 5 * cs, NIBF(n5));
                __stcs(o + 6 * cs, NIBF(n6));
                __stcs(o + 7 * cs, NIBF(n7));
                __stcs(o + 8 * cs, NIBF(n8));
              }
            }
          }
#undef NIBF
#undef NIBA
        }
      } else {
        // N % 4 != 0: one tile per lane, coalesced 32-bit stores
        for (int i = 0; i < NW; i++) {
          const int t = 32 * i + lane;
          const uint32_t bM = (__shfl_sync(FULL, M, i) >> lane) & 1u, bCG = (__shfl_sync(FULL, CG, i) >> lane) & 1u;
          const uint32_t bAny = (__shfl_sync(FULL, any_own, i) >> lane) & 1u;
          const bool live = t < N;
          const float f = live ? army_frac((uint32_t)S.army[t]) : 0.f;
#pragma unroll
          for (int p = 0; p < PT; p++) {
            if (p < P) {
              const uint32_t bV = prm.fog ? ((__shfl_sync(FULL, vis[p], i) >> lane) & 1u) : 1u;
              const uint32_t bO = (__shfl_sync(FULL, own[p], i) >> lane) & 1u;
              if (live) {
                const uint32_t nm = bV & ~bM;
                const uint32_t b2 = nm & bO, b3 = nm & bAny & ~bO, b4 = nm & ~bAny & 1u, b5 = nm & bCG, b6 = bV & bM;
                float *o = gbase + (size_t)p * GRL_OBS_CHANNELS * N + t;
                __stcs(o + 0 * N, b2 ? f : 0.f);
                __stcs(o + 1 * N, b3 ? f : 0.f);
                __stcs(o + 2 * N, b2 ? 1.f : 0.f);
                __stcs(o + 3 * N, b3 ? 1.f : 0.f);
                __stcs(o + 4 * N, b4 ? 1.f : 0.f);
                __stcs(o + 5 * N, b5 ? 1.f : 0.f);
                __stcs(o + 6 * N, b6 ? 1.f : 0.f);
                __stcs(o + 7 * N, bV ? 1.f : 0.f);
                __stcs(o + 8 * N, bV ? 0.f : 1.f);
              }
            }
          }
        }
      }
    }
  }
}

// ---- launch overlap -------------------------------------------------------------------------------------------------
// A launch costs 12-14 us beyond its per-game time whatever the board (the first wave loads and steps before any store
// is issued, the last one drains alone; t(B) is linear in B with that intercept: profiles/r2_variants.md) — 4 % of the
// headline launch, 12 % of a 10x10 one.  Consecutive turn launches of one env therefore OVERLAP: every CTA lets the next
// launch in the stream start as soon as this grid is fully resident (griddepcontrol.launch_dependents, "programmatic
// dependent launch"), so that its CTAs take the SM slots this grid's last wave leaves idle.  Games are independent, so
// what the next launch must wait for is not this GRID but the WARP that holds the same games: a warp publishes the
// launch's sequence number in its word of prm.epoch (st.release.gpu after a warp barrier: slab, planes and scalars of
// its games are visible) and, when the host marked the launch as overlapping (epoch_need != 0), first waits until its
// word has reached the previous launch's number (ld.acquire.gpu) — no griddepcontrol.wait, which would wait for the
// whole previous grid.  The warp -> games map is the same in every turn launch of an env; any other work on the stream
// (sampling, resets, copies, the caller's own kernels) does not trigger early and is serialised as always.
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int PT, int TW, int TH, int LG, bool DO_STEP, bool DO_OUT, int GYM>
__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32, TurnOccupancy<PT, LG>::kMinBlocks)
    grl_turn_kernel(const __grid_constant__ GrlKParams prm, const __grid_constant__ GrlGymK gk) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t *flag = nullptr;
  if (prm.epoch) {
    constexpr int GPW = 32 / LG;
    flag = prm.epoch + prm.game0 / GPW + blockIdx.x * GRL_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (prm.epoch_need) {
      if ((threadIdx.x & 31) == 0)
        while ((int32_t)(ld_acquire_gpu(flag) - prm.epoch_need) < 0) __nanosleep(40);
      __syncwarp();
      asm volatile("fence.proxy.async.global;" ::: "memory");  // the slab is read by the async proxy (cp.async.bulk)
    }
  }
  grl_turn_body<PT, TW, TH, LG, DO_STEP, DO_OUT, GYM>(prm, gk);
  if (flag) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) st_release_gpu(flag, prm.epoch_seq);
  }
}

template <int PT, int LG>
__device__ __noinline__ void policy_phase(const GrlKParams &prm, uint32_t *s, const uint32_t *st, uint32_t *s_act,
                                          uint32_t alive, uint32_t turn_before, int game, Geo g, int W, int H, int N, int NW) {
  const GrlLayout &L = prm.L;
  const int P = prm.P;
  SlabView S = make_view(s, st, L);
  const bool act_lane = g.lane < NW;
  const uint32_t M = act_lane ? S.M[g.lane] : 0u;
  uint32_t gt1 = army_gt1_mask<LG>(S.army, NW, N, g);
  DirMasks dm = dir_targets<LG>(M, g);
#pragma unroll 1
  for (int p = 0; p < P && p < prm.A; p++) {
    const uint32_t own = act_lane ? S.own[p * NW + g.lane] : 0u;
    const uint32_t lst = act_lane ? S.list[p * NW + g.lane] : 0u;
    uint32_t src = ((alive >> p) & 1u) ? (lst & own & gt1) : 0u;
    PackedAction a = sample_policy_action<LG>(prm, prm.policy_seed, dm, src, p, (uint64_t)(prm.env_id_base + game), turn_before, g);
    if (g.lane == 0 && a.present()) {
      uint2 d = decode_action(make_uint2(a.lo, a.hi), W, H, P);
      s_act[2 * p] = d.x;
      s_act[2 * p + 1] = d.y;
    }
  }
  __syncwarp(g.seg);
}

template <int PT, int LG, bool AGENT>
__device__ __forceinline__ bool gym_pre_phase(const GrlKParams &prm, const GrlGymK &gk, uint32_t *s, const uint32_t *st,
                                              uint32_t *s_act, uint32_t alive, bool over, uint32_t turn_before, int game, Geo g,
                                              int W, int H, int N, int NW, long long a0, long long a1) {
  const GrlLayout &L = prm.L;
  const int P = prm.P;
  SlabView S = make_view(s, st, L);
  const bool act_lane = g.lane < NW;
  const uint32_t M = act_lane ? S.M[g.lane] : 0u;
  const uint32_t gt1 = army_gt1_mask<LG>(S.army, NW, N, g);
  const DirMasks dm = dir_targets<LG>(M, g);
  // _get_valid_actions_mask (generals_env.py:344-387) of player p's fog-filtered view, tested at one index
  auto gym_ok = [&](long long a, int p) -> bool {
    if (a < 0 || a >= (long long)N * 5) return false;  // uniform over the group
    const int t = (int)(a / 5), k = (int)(a % 5);
    const uint32_t own = act_lane ? S.own[p * NW + g.lane] : 0u;
    const uint32_t v = act_lane ? (prm.fog ? S.vis[p * NW + g.lane] : g.valid) : 0u;
    const uint32_t src = v & own & gt1;
    const uint32_t U = src & dm.up, R = src & dm.right, D = src & dm.down, Lm = src & dm.left;
    const uint32_t sel = k == 0 ? U : (k == 1 ? R : (k == 2 ? D : (k == 3 ? Lm : (U | R | D | Lm))));
    return ((__shfl_sync(g.seg, sel, t >> 5, LG) >> (t & 31)) & 1u) != 0u;
  };
  // _action_index_to_game_action (generals_env.py:389-441)
  auto put = [&](long long a, int p, int slot) {
    const int from_idx = (int)(a / 5), info = (int)(a % 5);
    const int fx = from_idx % W, fy = from_idx / W;
    int tx = fx, ty = fy;
    if (info < 4) {
      tx = fx + (info == 1) - (info == 3);
      ty = fy + (info == 2) - (info == 0);
    } else {  // half move: the first in-bounds direction in the order up, right, down, left
      if (fy - 1 >= 0) ty = fy - 1;
      else if (fx + 1 < W) tx = fx + 1;
      else if (fy + 1 < H) ty = fy + 1;
      else tx = fx - 1;
    }
    // Server.SubmitAction -> ValidateCoreAction (internal/grpc/gameserver/server.go:241, action_validator.go:113-137):
    // the server runs MoveAction.Validate on the board at submission and never buffers a refused action; the turn runs
    // without it and the client, which ignores the response, still counts the step.  The client's mask guarantees
    // everything Validate checks except the target of a half move, which it aims at the first in-bounds direction
    // whatever stands there.
    const int ti = ty * W + tx;
    if (info == 4 && ((S.M[ti >> 5] >> (ti & 31)) & 1u)) return;
    const PackedAction pa = pack_action(p, fx, fy, tx, ty, info != 4);
    const uint2 d = decode_action(make_uint2(pa.lo, pa.hi), W, H, P);
    if (g.lane == 0) {
      s_act[2 * slot] = d.x;
      s_act[2 * slot + 1] = d.y;
    }
  };
  if constexpr (AGENT) {  // the random agent: one of the entries gym_ok accepts, or 0 when there is none
    const uint32_t own0 = act_lane ? S.own[g.lane] : 0u;
    const uint32_t v0 = act_lane ? (prm.fog ? S.vis[g.lane] : g.valid) : 0u;
    const uint32_t src0 = v0 & own0 & gt1;
    a0 = gym_random_agent<LG>(gk.agent_seed, (unsigned long long)(prm.env_id_base + game), src0 & dm.up, src0 & dm.right,
                              src0 & dm.down, src0 & dm.left, g);
    if (g.lane == 0) gk.sampled_action[game] = a0;
  }
  const bool ok0 = gym_ok(a0, 0);
  if (ok0) put(a0, 0, 0);
  if (gk.opponent_action) {
    if (gym_ok(a1, 1)) put(a1, 1, 1);
  } else if (!over) {
    // the reference's default opponent (generals_env.py:443-497): a uniformly random legal FULL move; the
    // synthetic policy's draw keyed (opponent_seed, env, turn, player), players beyond 1 keep its half-move bit
#pragma unroll 1
    for (int p = 1; p < P && p < prm.A; p++) {
      const uint32_t own = act_lane ? S.own[p * NW + g.lane] : 0u;
      const uint32_t lst = act_lane ? S.list[p * NW + g.lane] : 0u;
      const uint32_t src = ((alive >> p) & 1u) ? (lst & own & gt1) : 0u;
      PackedAction a = sample_policy_action<LG>(prm, gk.opponent_seed, dm, src, p, (uint64_t)(prm.env_id_base + game),
                                                turn_before, g);
      if (p == 1) a.hi |= 1u << 8;  // move_all
      if (g.lane == 0 && a.present()) {
        const uint2 d = decode_action(make_uint2(a.lo, a.hi), W, H, P);
        s_act[2 * p] = d.x;
        s_act[2 * p + 1] = d.y;
      }
    }
  }
  __syncwarp(g.seg);
  return ok0;
}

template <int PT, int LG>
__device__ __noinline__ uint32_t elimination_phase(const GrlKParams &prm, uint32_t *s, const uint32_t *st, uint32_t alive,
                                                   int n_orders, uint32_t ord_lo, uint32_t ord_hi, Geo g, int N, int NW) {
  const GrlLayout &L = prm.L;
  const int P = prm.P;
  SlabView S = make_view(s, st, L);
  const int lane = g.lane;
  const bool act_lane = lane < NW;
  uint32_t own[PT], lst[PT];
#pragma unroll
  for (int p = 0; p < PT; p++) {
    bool on = act_lane && p < P;
    own[p] = on ? S.own[p * NW + lane] : 0u;
    lst[p] = on ? S.list[p * NW + lane] : 0u;
  }
  uint32_t chg = act_lane ? S.chg[lane] : 0u;
  uint32_t vch = act_lane ? S.vch[lane] : 0u;
  const uint32_t G = act_lane ? S.G[lane] : 0u;
#pragma unroll 1
  for (int o = 0; o < n_orders; o++) {
    uint32_t ob = (o < 4 ? (ord_lo >> (8 * o)) : (ord_hi >> (8 * (o - 4)))) & 0xffu;
    int el = (int)(ob & 0xfu), nw = (int)(ob >> 4);
    uint32_t X = 0;
#pragma unroll
    for (int q = 0; q < PT; q++)
      if (q == el) X = lst[q] & own[q];  // still owned AND in the cached list (engine.go:130-137)
#pragma unroll
    for (int q = 0; q < PT; q++) {
      if (q == el) own[q] &= ~X;
      if (q == nw) own[q] |= X;
    }
    chg |= X;
    vch |= X;
    int moved_army = 0;
    if (__any_sync(g.seg, X != 0u)) moved_army = sum_army_over<LG>(X, S.army, NW, N, g);
    if (lane == 0) {
      S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * el + GRL_PL_TRUE_ARMY] -= (uint32_t)moved_army;
      S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * nw + GRL_PL_TRUE_ARMY] += (uint32_t)moved_army;
      S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * el + GRL_PL_GENERAL_IDX] = 0xffffffffu;
    }
    alive &= ~(1u << el);
  }
  __syncwarp(g.seg);
  stats_update<PT, LG>(prm, S, own, lst, chg, G, alive, g, N, NW);  // engine.go:107
  if (act_lane) {
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P) {
        S.own[p * NW + lane] = own[p];
        S.list[p * NW + lane] = lst[p];
      }
    }
    S.chg[lane] = chg;
    S.vch[lane] = vch;
  }
  __syncwarp(g.seg);
  return alive;
}
// ---------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------
static inline size_t grl_turn_smem_bytes(const GrlLayout &L, int TW, int TH, int PT, int LG, bool gym) {
  const bool snap = LG == 32;
  const int per_game = (snap ? 2 : 1) * L.slab_words + L.static_words + 2 * GRL_MAX_ACTIONS;
  const int scratch = gym ? grl_gym_smem_words(L.P, L.NW, L.N, grl_gym_emit_mode(TW, TH)) : grl_obs_scratch_words(TW, TH, PT, L.NW);
  return (size_t)GRL_WARPS_PER_CTA * (size_t)((32 / LG) * per_game + scratch) * 4u;
}

template <int PT, int TW, int TH, int LG, bool S, bool O, int GYM = 0>
static cudaError_t launch_turn_t(const GrlKParams &prm, cudaStream_t stream, const GrlGymK *gym = nullptr) {
  size_t smem = grl_turn_smem_bytes(prm.L, TW, TH, PT, LG, GYM != 0);
  auto kern = grl_turn_kernel<PT, TW, TH, LG, S, O, GYM>;
  GrlGymK gk;
  memset(&gk, 0, sizeof gk);
  if (gym) gk = *gym;
  static size_t tuned_smem = ~(size_t)0;  // per instantiation
  if (tuned_smem != smem) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    tuned_smem = smem;
  }
  // a CTA steps GRL_WARPS_PER_CTA * (32 / LG) games; CTAs of a wave move through the load -> turn ->
  // store phases out of step with each other, which keeps the observation store stream busy
  const int per_cta = GRL_WARPS_PER_CTA * (32 / LG);
  int grid = (prm.game_end - prm.game0 + per_cta - 1) / per_cta;
  if (prm.epoch && prm.epoch_need) {  // may start while the previous turn launch of the stream drains (launch overlap)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid < 1 ? 1 : grid);
    cfg.blockDim = dim3(GRL_WARPS_PER_CTA * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, prm, gk);
  }
  kern<<<grid < 1 ? 1 : grid, GRL_WARPS_PER_CTA * 32, smem, stream>>>(prm, gk);
  return cudaGetLastError();
}

// One geometry (TW x TH baked in, TW == 0: generic), every player template it is built for.
template <int PT, int TW, int TH, int LG>
static cudaError_t launch_turn_g(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  if (do_step && do_out) return launch_turn_t<PT, TW, TH, LG, true, true>(prm, stream);
  if (do_step) return launch_turn_t<PT, TW, TH, LG, true, false>(prm, stream);
  return launch_turn_t<PT, TW, TH, LG, false, true>(prm, stream);
}

template <int TW, int TH, int LG>
static cudaError_t launch_turn_geo(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  if (prm.P <= 2) return launch_turn_g<2, TW, TH, LG>(prm, do_step, do_out, stream);
  if (prm.P <= 4) return launch_turn_g<4, TW, TH, LG>(prm, do_step, do_out, stream);
  if constexpr (TW == 0) return launch_turn_g<8, TW, TH, LG>(prm, do_step, do_out, stream);
  return cudaErrorInvalidValue;  // baked boards are built for up to four players (the dispatcher sends the rest to generic)
}

template <int TW, int TH, int LG>
static cudaError_t launch_gym_geo(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  // GYM = 1: the agent's indices are given; 2: the random agent is drawn inside the launch (gk.action == nullptr).  The
  // gym contract drives two players, so only the two-player template carries the second instantiation.
  if (gk.action == nullptr) {
    if (prm.P <= 2) return launch_turn_t<2, TW, TH, LG, true, true, 2>(prm, stream, &gk);
    return cudaErrorInvalidValue;
  }
  if (prm.P <= 2) return launch_turn_t<2, TW, TH, LG, true, true, 1>(prm, stream, &gk);
  if (prm.P <= 4) return launch_turn_t<4, TW, TH, LG, true, true, 1>(prm, stream, &gk);
  if constexpr (TW == 0) return launch_turn_t<8, TW, TH, LG, true, true, 1>(prm, stream, &gk);
  return cudaErrorInvalidValue;
}
