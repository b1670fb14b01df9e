// grl_kernels.cu — sm_100a kernels of the batched Generals.io turn engine.
//
// A GROUP of LG lanes owns one game (LG = 32, 16, 8 or 4: the smallest power of two >= the NW
// words a board's bit planes span), so a warp steps 32/LG games at once and small boards do not
// leave most lanes idle.  Every boolean plane of a game (ownership per player, the
// reference's cached OwnedTiles lists, visibility per player, the changed / visibility-
// changed tile sets, terrain) is an N-bit LINEAR bitmask, N = W*H <= 1024, held as one
// 32-bit word per lane.  Stencils (3x3 fog dilation, the 5x5 "affected players" probe,
// the four move directions) are funnel shifts across neighbouring lanes' words; set sizes
// are popc + REDUX.  Only the armies are a per-tile plane (uint16, staged in shared memory).
//
// Reference semantics (file:line into /root/reference):
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
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/grlcuda.h"
#include "grl_launch.h"
#include "grl_layout.h"

#define FULL 0xffffffffu
#ifndef GRL_WARPS_PER_CTA
#define GRL_WARPS_PER_CTA 8
#endif

// Scheduling variants (profiles/ records the comparison):
//   GRL_PERSISTENT  resident CTAs loop over games and prefetch the next slab (TMA double buffer)
//   GRL_CTA_SYNC    with GRL_PERSISTENT: keep a CTA's warps in the same phase (instruction-cache locality)
#ifndef GRL_PERSISTENT
#define GRL_PERSISTENT 0
#endif
#ifndef GRL_CTA_SYNC
#define GRL_CTA_SYNC 0
#endif
// GRL_DIRTY_WB: write back only the 32-byte sectors of the slab that the turn changed (the
// bulk load lands the slab twice; the second copy is the comparison snapshot).  A turn dirties
// ~40 % of a 20x20 slab's sectors; scattered state writes cost the observation store stream
// about three times their byte share (tools/micro/store_readmix.cu, profiles/r1_variants.md).
#ifndef GRL_DIRTY_WB
#define GRL_DIRTY_WB (!GRL_PERSISTENT)
#endif
// GRL_PACKED_SNAPSHOT (default 0): packed groups (LG < 32) skip the snapshot and write the whole slab back,
// which frees shared memory for a fourth CTA per SM (15x15 x 262,144 games: 0.974 -> 0.877 ms).
#ifndef GRL_PACKED_SNAPSHOT
#define GRL_PACKED_SNAPSHOT 0
#endif
// GRL_OBS_LUT: nibble -> float4 through the shared-memory table (1) or eight ALU selects (0)
#ifndef GRL_OBS_LUT
#define GRL_OBS_LUT 1
#endif
#ifndef GRL_OBS_CHUNK_MAJOR
// 1: the per-game scalar results (done, winner, step_error, reward, action_index) are staged per CTA in shared
// memory and written by one warp as contiguous runs; 0 (default): one lane per value straight to global memory.
// Measured on B200 (profiles/r1_variants.md): staging is 0.3 % slower device-resident, 0.3 % slower through
// pinned host planes at 20x20 and 2 % slower at 15x15 — the end-of-CTA barrier costs more than the runs save.
#ifndef GRL_STAGE_SCALARS
#define GRL_STAGE_SCALARS 0
#endif
// observation float4s that straddle two planes (boards with N % 4 != 0): 2 (default) = four per-element evaluations
// inside the sweep; 0 = a short branch-free merge of the two planes' windows.  Measured at 15x15 x 262,144 games:
// 0.880 ms (2) vs 0.899 ms (0); taking them out of the sweep (a later round of 128-bit stores, or scalar stores)
// 1.02-1.03 ms: a warp store with a 16-byte hole, completed later, costs far more than the divergent branch.
// running (plane, tile) counters in obs_linear instead of a division per store: measured SLOWER (0.9125 vs 0.8748 ms
// per 262,144 15x15 games) — the independent index computations schedule better than the loop-carried chain
// store policy of the linear observation writer.  A store-only stream of the 15x15 layout (blocks that start and end
// mid-sector) runs at 6.6 TB/s with the default policy against 6.1 with evict-first (tools/micro/store_holes.cu), but in
// the kernel the default policy lets the observation stream evict the prefetched state: 0.950 vs 0.877 ms per 262,144 games.
// 1 (default): st.global.cs
#ifndef GRL_LINEAR_STCS
#define GRL_LINEAR_STCS 1
#endif
#if GRL_LINEAR_STCS
#define GRL_LIN_ST(p, v) __stcs((p), (v))
#else
#define GRL_LIN_ST(p, v) (*(p) = (v))
#endif
// 1 (default): the games of a warp (lane groups) write their observation blocks as ONE sector-complete run: the 32-byte
// sector shared by two consecutive games' blocks is written whole by the later game's pass (tools/micro/store_holes.cu:
// 7.0 vs 6.1 TB/s for the 15x15 layout)
#ifndef GRL_OBS_JOIN
#define GRL_OBS_JOIN 1
#endif
// 1: the linear observation writer walks plane by plane (see obs_linear): a quarter fewer instructions per game, but
// 25-lane store rounds that split sectors between instructions — measured equal to the flat sweep (0.8298 vs 0.8276 ms per
// 262,144 15x15 games, 0.2324 vs 0.2290 per 65,536), so the flat sweep stays
#ifndef GRL_OBS_PLANEWISE
#define GRL_OBS_PLANEWISE 0
#endif
#ifndef GRL_OBS_INCR
#define GRL_OBS_INCR 0
#endif
#ifndef GRL_STRADDLE_INLINE
#define GRL_STRADDLE_INLINE 3
#endif
#define GRL_OBS_CHUNK_MAJOR 0  // 1: the round-1a tile-chunk-major observation loop (comparison builds)
#endif

// per-warp shared-memory words of the linear observation writer (only baked boards with N % 4 != 0)
__host__ __device__ constexpr int grl_obs_scratch_words(int TW, int TH, int PT, int NW) {
  // channel masks [PT*9][NW+1] + army-fraction plane [N+4], rounded to 16 bytes, + the P*9-1 precomputed float4s that
  // straddle two planes
  return (TW > 0 && ((TW * TH) & 3) != 0)
             ? (((PT * GRL_OBS_CHANNELS * (NW + 1) + TW * TH + 4 + 3) & ~3) + 4 * PT * GRL_OBS_CHANNELS)
             : 0;
}

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

__device__ __forceinline__ uint64_t policy_draw(uint64_t seed, uint64_t env, uint64_t turn, uint64_t player) {
  uint64_t x = mix64(seed ^ (env * 0xD6E8FEB86659FD93ULL));
  return mix64(x ^ (turn * 0xA0761D6478BD642FULL) ^ (player << 56));
}

__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// float32(army) / 1000.0f, correctly rounded (== IEEE division; verified for every army value):
// q = RN(x*r), rem = RN(x - q*1000) exact by FMA, q' = RN(q + rem*r)   with r = RN(1/1000).
__device__ __forceinline__ float army_frac(uint32_t army) {
  const float r = 1.0f / 1000.0f;
  float x = (float)army;
  float q = __fmul_rn(x, r);
  float rem = __fmaf_rn(-q, 1000.0f, x);
  float v = __fmaf_rn(rem, r, q);
  return army >= 1000u ? 1.0f : v;  // serializer.go:84-88 clip
}

// Per-lane geometry words and linear-bitmask stencils.  `lane` is the lane's index INSIDE its
// group; all cross-lane traffic is segmented (width LG) and synchronises on the group's member mask.
struct Geo {
  uint32_t valid, nc0, ncl;  // tiles that exist / x != 0 / x != W-1, word `lane`
  int W;
  int lane;      // 0..LG-1
  int shift;     // first warp lane of the group
  uint32_t seg;  // member mask of the group
};

template <int LG>
__device__ __forceinline__ uint32_t word_prev(uint32_t v, const Geo &g) {
  uint32_t p = __shfl_up_sync(g.seg, v, 1, LG);
  return g.lane == 0 ? 0u : p;
}
template <int LG>
__device__ __forceinline__ uint32_t word_next(uint32_t v, const Geo &g) {
  uint32_t n = __shfl_down_sync(g.seg, v, 1, LG);
  return g.lane == LG - 1 ? 0u : n;
}
// bit t of result = bit (t-k) of v
template <int LG>
__device__ __forceinline__ uint32_t shl_bits(uint32_t v, int k, const Geo &g) {
  return __funnelshift_lc(word_prev<LG>(v, g), v, k);
}
// bit t of result = bit (t+k) of v
template <int LG>
__device__ __forceinline__ uint32_t shr_bits(uint32_t v, int k, const Geo &g) {
  return __funnelshift_rc(v, word_next<LG>(v, g), k);
}
// in-bounds 3x3 neighbourhood union (visibility_optimized.go:9-13,118-128)
template <int LG>
__device__ __forceinline__ uint32_t dilate3(uint32_t v, const Geo &g) {
  uint32_t h = v | (shl_bits<LG>(v, 1, g) & g.nc0) | (shr_bits<LG>(v, 1, g) & g.ncl);
  uint32_t r = h | shl_bits<LG>(h, g.W, g) | shr_bits<LG>(h, g.W, g);
  return r & g.valid;
}

__device__ __forceinline__ Geo make_geo(const GrlKParams &prm, int W, int lane, int LG) {
  Geo g;
  g.lane = lane % LG;
  g.shift = lane - g.lane;
  g.seg = LG == 32 ? FULL : (((1u << (LG & 31)) - 1u) << g.shift);
  g.W = W;
  g.valid = prm.geom[g.lane];
  g.nc0 = prm.geom[32 + g.lane];
  g.ncl = prm.geom[64 + g.lane];
  return g;
}

// bits 0..7 of b spread to bit positions 0,4,8,...,28
__device__ __forceinline__ uint32_t spread8(uint32_t b) {
  uint32_t x = b & 0xffu;
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}

// Views into one game's slab staged in shared memory.
struct SlabView {
  uint32_t *hdr, *own, *list, *vis, *chg, *vch;
  uint16_t *army;
  const uint32_t *M, *C, *G;
};

__device__ __forceinline__ SlabView make_view(uint32_t *s, const uint32_t *st, const GrlLayout &L) {
  SlabView v;
  v.hdr = s;
  v.own = s + L.off_own;
  v.list = s + L.off_list;
  v.vis = s + L.off_vis;
  v.chg = s + L.off_changed;
  v.vch = s + L.off_vchg;
  v.army = reinterpret_cast<uint16_t *>(s + L.off_army);
  v.M = st;
  v.C = st + L.NW;
  v.G = st + 2 * L.NW;
  return v;
}

// sum of army over the tiles of a linear bitmask (word `lane` in x); slow path helper
template <int LG>
__device__ __forceinline__ int sum_army_over(uint32_t x, const uint16_t *army, int NW, int N, const Geo &g) {
  int acc = 0;
#pragma unroll 1
  for (int i = 0; i < NW; i++) {
    uint32_t xw = __shfl_sync(g.seg, x, i, LG);
#pragma unroll
    for (int b = g.lane; b < 32; b += LG) {
      int t = 32 * i + b;
      int a = (t < N) ? (int)army[t] : 0;
      acc += ((xw >> b) & 1u) ? a : 0;
    }
  }
  return __reduce_add_sync(g.seg, acc);
}

// TMA bulk copies (cp.async.bulk, SASS UBLKCP) for the state slabs -------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr(src_smem)),
               "r"(bytes)
               : "memory");
}
// GRL_L2_HINT=1: the state slabs carry an L2::evict_last policy on their bulk loads, prefetches and write-back
// stores (the observation stream is already evict-first through st.global.cs)
#ifndef GRL_L2_HINT
#define GRL_L2_HINT 0
#endif
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_hint_v4(uint4 *a, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2(const void *src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// cached-list statistics (internal/game/stats.go:8-144) on register words.
// armyCount[p] = trueArmy[p] - (armies on tiles p owns that are missing from its list).
// ---------------------------------------------------------------------------------------
template <int PT, int LG>
__device__ __forceinline__ void stats_update(const GrlKParams &prm, SlabView &S, const uint32_t (&own)[PT],
                                             uint32_t (&lst)[PT], uint32_t chg, uint32_t G, uint32_t &alive, const Geo &g,
                                             int N, int NW) {
  int c = __reduce_add_sync(g.seg, __popc(chg));
  if (c == 0) return;               // stats.go:11-15 (turn > 0 inside a step)
  const bool full = c > N / 5;      // stats.go:20-25
#pragma unroll
  for (int p = 0; p < PT; p++) {
    if (p < prm.P) {
      lst[p] = full ? own[p] : (own[p] & (lst[p] | chg));
      uint32_t orphan = own[p] & ~lst[p];
      int true_army = (int)S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_TRUE_ARMY];
      int corr = 0;
      if (__any_sync(g.seg, orphan != 0u)) corr = sum_army_over<LG>(orphan, S.army, NW, N, g);
      uint32_t gen = lst[p] & G;
      int gi = gen ? (32 * g.lane + 31 - __clz(gen)) : -1;
      gi = __reduce_max_sync(g.seg, gi);
      if (g.lane == 0) {
        S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ARMY_COUNT] = (uint32_t)(true_army - corr);
        S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_GENERAL_IDX] = (uint32_t)gi;
      }
      alive = gi >= 0 ? (alive | (1u << p)) : (alive & ~(1u << p));
    }
  }
  __syncwarp(g.seg);
}

// engine legal-move direction masks for one player (rules/legal_moves.go:19-73):
// a tile in the cached list, still owned, army > 1, target in bounds and not a mountain.
struct DirMasks {
  uint32_t up, right, down, left;
};
template <int LG>
__device__ __forceinline__ DirMasks dir_targets(uint32_t M, const Geo &g) {
  uint32_t free_ = g.valid & ~M;
  DirMasks d;
  d.up = shl_bits<LG>(free_, g.W, g);            // tile t-W exists and is not a mountain
  d.down = shr_bits<LG>(free_, g.W, g);          // tile t+W
  d.left = shl_bits<LG>(free_, 1, g) & g.nc0;    // tile t-1, x != 0
  d.right = shr_bits<LG>(free_, 1, g) & g.ncl;   // tile t+1, x != W-1
  return d;
}

// army > 1 per tile as a linear bitmask (word `lane`).  Lane l tests its own 32 tiles from four 128-bit loads of
// the uint16 army plane, two tiles per 32-bit word at a time (halfword != 0 after clearing bit 0, by the carry
// trick), so the cost does not depend on how many games share the warp.  (The first version balloted LG tiles per
// step: 32 serial ballots per 15x15 game, 375 warp instructions per game and 19 % of the stall samples there.)
__device__ __forceinline__ uint32_t gt1_pair(uint32_t x) {  // bit 0: low halfword > 1, bit 1: high halfword > 1
  const uint32_t y = x & 0xfffefffeu;
  const uint32_t z = ((y & 0x7fff7fffu) + 0x7fff7fffu) | y;  // bit 15 / bit 31: halfword != 0
  return ((z >> 15) & 1u) | ((z >> 30) & 2u);
}
#ifndef GRL_GT1_BALLOT
#define GRL_GT1_BALLOT 0
#endif
template <int LG>
__device__ __forceinline__ uint32_t army_gt1_mask(const uint16_t *army, int NW, int N, const Geo &g) {
  // one game per warp keeps the ballot version: 13 ballots for 20x20 cost about the same as 13 active lanes doing the
  // vector version, and measured 1.4 % faster there (0.3423 vs 0.347 ms per 65,536 games)
  if (GRL_GT1_BALLOT || LG == 32) {
  uint32_t mine = 0;
  for (int i = 0; i < NW; i++) {
#pragma unroll
    for (int r = 0; r < 32 / LG; r++) {
      int t = 32 * i + r * LG + g.lane;
      bool gt = (t < N) && army[t] > 1;
      uint32_t w = __ballot_sync(g.seg, gt) >> g.shift;  // LG bits
      if (g.lane == i) mine |= w << (r * LG);
    }
  }
  return mine;
  }
  uint32_t m = 0;
  if (g.lane < NW) {
    const int NA = (N + 7) & ~7;  // the plane holds NA entries (grl_layout.h): a group of 8 tiles is inside it or not at all
    const uint4 *a4 = reinterpret_cast<const uint4 *>(army + 32 * g.lane);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (32 * g.lane + 8 * j < NA) {
        const uint4 q = a4[j];
        m |= (gt1_pair(q.x) | (gt1_pair(q.y) << 2) | (gt1_pair(q.z) << 4) | (gt1_pair(q.w) << 6)) << (8 * j);
      }
    }
  }
  return m & g.valid;
}

struct PackedAction {  // grl_action as one 64-bit word (little endian field order)
  uint32_t lo, hi;
  __device__ __forceinline__ int player() const { return (int)(int8_t)(lo & 0xff); }
  __device__ __forceinline__ int fx() const { return (int)(int8_t)((lo >> 8) & 0xff); }
  __device__ __forceinline__ int fy() const { return (int)(int8_t)((lo >> 16) & 0xff); }
  __device__ __forceinline__ int tx() const { return (int)(int8_t)((lo >> 24) & 0xff); }
  __device__ __forceinline__ int ty() const { return (int)(int8_t)(hi & 0xff); }
  __device__ __forceinline__ bool move_all() const { return ((hi >> 8) & 0xff) != 0; }
  __device__ __forceinline__ bool present() const { return ((hi >> 16) & 0xff) != 0; }
};

__device__ __forceinline__ PackedAction pack_action(int player, int fx, int fy, int tx, int ty, bool move_all) {
  PackedAction a;
  a.lo = (uint32_t)(player & 0xff) | ((uint32_t)(fx & 0xff) << 8) | ((uint32_t)(fy & 0xff) << 16) |
         ((uint32_t)(tx & 0xff) << 24);
  a.hi = (uint32_t)(ty & 0xff) | ((move_all ? 1u : 0u) << 8) | (1u << 16);
  return a;
}

// Synthetic policy: player p draws uniformly from the set bits of its engine mask
// in flat-index order (tile-major, dirs U,R,D,L).  Warp-uniform result.
template <int LG>
__device__ __forceinline__ PackedAction sample_policy_action(const GrlKParams &prm, uint64_t seed, const DirMasks &dm,
                                                             uint32_t src, int p, uint64_t env_global, uint32_t turn,
                                                             const Geo &g) {
  PackedAction none;
  none.lo = none.hi = 0;
  uint32_t U = src & dm.up, R = src & dm.right, D = src & dm.down, Lm = src & dm.left;
  int cnt = __popc(U) + __popc(R) + __popc(D) + __popc(Lm);
  int total = __reduce_add_sync(g.seg, cnt);
  if (total == 0) return none;
  uint64_t r = policy_draw(seed, env_global, (uint64_t)turn, (uint64_t)p);
  int k = (int)((uint32_t)r % (uint32_t)total);
  int incl = cnt;  // inclusive prefix sum over the group's lanes
#pragma unroll
  for (int o = 1; o < LG; o <<= 1) {
    int v = __shfl_up_sync(g.seg, incl, o, LG);
    if (g.lane >= o) incl += v;
  }
  int excl = incl - cnt;
  bool mine = (k >= excl) && (k < incl);
  int kk = k - excl;
  // smallest bit b with count(bits <= b) > kk, by binary search on the prefix count
  int b = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    int cand = b + step;                      // test whether count(bits < cand) <= kk
    uint32_t m = (1u << cand) - 1u;           // cand in 1..31
    int c = __popc(U & m) + __popc(R & m) + __popc(D & m) + __popc(Lm & m);
    if (c <= kk) b = cand;
  }
  uint32_t below = (1u << b) - 1u;
  int rem = kk - (__popc(U & below) + __popc(R & below) + __popc(D & below) + __popc(Lm & below));
  uint32_t nib = ((U >> b) & 1u) | (((R >> b) & 1u) << 1) | (((D >> b) & 1u) << 2) | (((Lm >> b) & 1u) << 3);
  int dir = 0;
#pragma unroll
  for (int d = 0; d < 4; d++) {
    if ((nib >> d) & 1u) {
      if (rem == 0) dir = d;
      rem--;
    }
  }
  int packed = mine ? ((32 * g.lane + b) * 4 + dir) : 0;
  uint32_t who = __ballot_sync(g.seg, mine) >> g.shift;
  packed = __shfl_sync(g.seg, packed, __ffs(who) - 1, LG);
  int tile = packed >> 2;
  dir = packed & 3;
  int fx = tile % prm.W, fy = tile / prm.W;
  int tx = fx + (dir == 1) - (dir == 3);
  int ty = fy + (dir == 2) - (dir == 0);
  return pack_action(p, fx, fy, tx, ty, ((r >> 32) & 1ULL) != 0);
}

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

// Fused gym step, before the turn: decode the agent's (player 0) Discrete(N*5) index against the gym mask of the
// CURRENT state (client-side rejection, generals_env.py:226-229), then the opponent's index or the random
// opponent's draw; decoded moves go to s_act.  Returns whether the agent's action is valid.
template <int PT, int LG>
__device__ __noinline__ bool gym_pre_phase(const GrlKParams &prm, const GrlGymK &gk, uint32_t *s, const uint32_t *st,
                                           uint32_t *s_act, uint32_t alive, bool over, uint32_t turn_before, int game, Geo g,
                                           int W, int H, int N, int NW);
#define GRL_GYM_EMIT_GENERIC 0
#define GRL_GYM_EMIT_QUADS 1
#define GRL_GYM_EMIT_LINEAR 2
// which read-out writer a geometry uses: baked boards with N % 4 == 0 -> quads, other baked boards -> linear
__host__ __device__ constexpr int grl_gym_emit_mode(int TW, int TH) {
  return TW > 0 ? (((TW * TH) & 3) == 0 ? GRL_GYM_EMIT_QUADS : GRL_GYM_EMIT_LINEAR) : GRL_GYM_EMIT_GENERIC;
}
__host__ __device__ inline int grl_gym_smem_words(int P, int NW, int N, int mode);
template <int NT>
__device__ __forceinline__ void gym_emit(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                         float *__restrict__ obs, uint8_t *__restrict__ mask, int32_t *__restrict__ stats,
                                         const uint32_t *s, const uint32_t *stt, uint32_t *sw, int game, int lane,
                                         const Geo &g);
template <int PT, int N>
__device__ __forceinline__ void gym_emit_quads(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                               float *__restrict__ obs, uint8_t *__restrict__ mask,
                                               int32_t *__restrict__ stats, const uint32_t *s, const uint32_t *stt,
                                               const float4 *lut, uint32_t *sw, int game, int lane, const Geo &g);
template <int PT, int N>
__device__ __forceinline__ void gym_emit_linear(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                                float *__restrict__ obs, uint8_t *__restrict__ mask,
                                                int32_t *__restrict__ stats, const uint32_t *s, const uint32_t *stt,
                                                const float4 *lut, uint32_t *sw, int game, int lane, const Geo &g);

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
#ifdef GRL_MIN_BLOCKS
  static constexpr int kMinBlocks = GRL_MIN_BLOCKS;
#else
  // packed groups (LG < 32) carry 32/LG slabs per warp in shared memory: 3 CTAs fit
  static constexpr int kMinBlocks = (LG < 32 && GRL_PACKED_SNAPSHOT) ? 3 : (PT <= 2 ? 4 : (PT <= 4 ? 3 : 2)) * 8 / GRL_WARPS_PER_CTA;
#endif
};

// Observation planes, PLANE-MAJOR store order, for baked geometries with N % 4 == 0
// (Serializer.StateToTensor, serializer.go:37-109).  A warp writes its game's whole
// [P][9][N] fp32 block as ONE linear sweep of 128-bit stores: measured 7.1 TB/s for this
// order against 6.4 TB/s for tile-chunk-major (profiles/r1_variants.md).  Lane l owns the
// tile quads q = l + 32c (c < NCH); every mask's nibbles for those quads are read once from
// the shared-memory slab and packed 4 bits per chunk, so each channel's nibbles for ALL
// chunks come from one LOP3 on the packed fields.  A nibble becomes a float4 of 0/1 through
// a 16-entry shared-memory table (one LDS.128 per store instead of eight ALU selects).
template <int PT, int N>
__device__ __forceinline__ void obs_plane_major(const GrlKParams &prm, const SlabView &S, const float4 *lut, int P, int NW,
                                                int game, int lane) {
  constexpr int NQ = N / 4;
  constexpr int NCH = (NQ + 31) / 32;
  const int bsel = lane >> 1, bsh = 4 * (lane & 1);  // quad q -> byte q>>1, nibble q&1 of a mask's byte array
  const uint8_t *bM = reinterpret_cast<const uint8_t *>(S.M);
  const uint8_t *bC = reinterpret_cast<const uint8_t *>(S.C);
  const uint8_t *bG = reinterpret_cast<const uint8_t *>(S.G);
  uint32_t mM = 0, mCG = 0, mAny = 0, livem = 0;
  uint32_t nV[PT], nO[PT];
#pragma unroll
  for (int p = 0; p < PT; p++) nV[p] = nO[p] = 0;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    const bool live = 32 * c + lane < NQ;
    if (live) {
      const int b = 16 * c + bsel;
      livem |= 0xfu << (4 * c);
      mM |= ((bM[b] >> bsh) & 0xfu) << (4 * c);
      mCG |= (((bC[b] | bG[b]) >> bsh) & 0xfu) << (4 * c);
#pragma unroll
      for (int p = 0; p < PT; p++) {
        if (p < P) {
          const uint8_t *bo = reinterpret_cast<const uint8_t *>(S.own + p * NW);
          const uint8_t *bv = reinterpret_cast<const uint8_t *>(S.vis + p * NW);
          nO[p] |= ((bo[b] >> bsh) & 0xfu) << (4 * c);
          nV[p] |= ((bv[b] >> bsh) & 0xfu) << (4 * c);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PT; p++) mAny |= nO[p];
  // armies are read only where somebody owns a non-mountain tile of the quad
  float f[NCH][4];
  const uint32_t need = mAny & ~mM;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    f[c][0] = f[c][1] = f[c][2] = f[c][3] = 0.f;
    if ((need >> (4 * c)) & 0xfu) {
      const uint2 aw = *reinterpret_cast<const uint2 *>(S.army + 4 * (32 * c + lane));
      f[c][0] = army_frac(aw.x & 0xffffu);
      f[c][1] = army_frac(aw.x >> 16);
      f[c][2] = army_frac(aw.y & 0xffffu);
      f[c][3] = army_frac(aw.y >> 16);
    }
  }
  const char *lutb = reinterpret_cast<const char *>(lut);
  float4 *gq = reinterpret_cast<float4 *>(prm.obs + (size_t)game * P * GRL_OBS_CHANNELS * N) + lane;
#pragma unroll
  for (int p = 0; p < PT; p++) {
    if (p < P) {
      const uint32_t v = prm.fog ? nV[p] : livem;
      const uint32_t nm = v & ~mM;
      uint32_t ch[GRL_OBS_CHANNELS];
      ch[0] = ch[2] = nm & nO[p];            // own (army, ownership)       serializer.go:75-90
      ch[1] = ch[3] = nm & mAny & ~nO[p];    // enemy
      ch[4] = nm & ~mAny;                    // neutral
      ch[5] = nm & mCG;                      // city or general
      ch[6] = v & mM;                        // mountain
      ch[7] = v;                             // visible
      ch[8] = ~v;                            // fog
#pragma unroll
      for (int k = 0; k < GRL_OBS_CHANNELS; k++) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          if (32 * c + lane < NQ) {
#if GRL_OBS_LUT
            const uint32_t idx16 = (c == 0 ? (ch[k] << 4) : (ch[k] >> (4 * c - 4))) & 0xf0u;
            float4 val = *reinterpret_cast<const float4 *>(lutb + idx16);
#else
            const uint32_t nb = ch[k] >> (4 * c);
            float4 val = make_float4((nb & 1u) ? 1.f : 0.f, (nb & 2u) ? 1.f : 0.f, (nb & 4u) ? 1.f : 0.f, (nb & 8u) ? 1.f : 0.f);
#endif
            if (k < 2) {
              val.x *= f[c][0];
              val.y *= f[c][1];
              val.z *= f[c][2];
              val.w *= f[c][3];
            }
            __stcs(gq + (p * GRL_OBS_CHANNELS + k) * NQ + 32 * c, val);
          }
        }
      }
    }
  }
}

// Observation planes for baked geometries with N % 4 != 0 (15x15): the game's [P][9][N] block is
// still ONE linear, 16-byte aligned sweep of 128-bit stores — channel planes start at odd float
// offsets there, so a float4 is addressed by its position e in the BLOCK, not in a plane:
// plane = e / N, tile = e % N.  The nine channel bitmasks of every player are staged in shared
// memory (one pad word each, so a 4-bit window may straddle the last word), armies are converted
// once into a float plane, and each store costs two LDS for the window, one table lookup and — on
// the two army planes — four scalar LDS.  The <= 3 floats before/after the aligned body and the
// float4s that straddle two planes (P*9-1 of them) take a per-element path.
template <int N>
__device__ __forceinline__ float obs_element(const uint32_t *chm, const float *frac, int NWP, int e) {
  const int plane = e / N, t = e - plane * N;
  const uint32_t bit = (chm[plane * NWP + (t >> 5)] >> (t & 31)) & 1u;
  return bit ? ((plane % GRL_OBS_CHANNELS) < 2 ? frac[t] : 1.f) : 0.f;
}

template <int PT, int N>
__device__ __forceinline__ void obs_linear(const GrlKParams &prm, const SlabView &S, const float4 *lut, uint32_t *scratch,
                                           int P, int NW, int game, int lane, const uint32_t *prev_slab = nullptr,
                                           bool next_in_warp = false) {
  const int NWP = NW + 1;
  uint32_t *chm = scratch;                                                   // [P*9][NWP]
  float *frac = reinterpret_cast<float *>(scratch + PT * GRL_OBS_CHANNELS * NWP);  // [N + 4]
  float4 *sf = reinterpret_cast<float4 *>(scratch + ((PT * GRL_OBS_CHANNELS * NWP + N + 4 + 3) & ~3));  // [PT*9] straddlers
  if (lane < NWP) {
    const bool w = lane < NW;
    const uint32_t valid = w ? prm.geom[lane] : 0u;
    const uint32_t M = w ? S.M[lane] : 0u;
    const uint32_t CG = w ? (S.C[lane] | S.G[lane]) : 0u;
    uint32_t any_own = 0;
#pragma unroll
    for (int p = 0; p < PT; p++)
      if (p < P && w) any_own |= S.own[p * NW + lane];
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P) {
        const uint32_t own = w ? S.own[p * NW + lane] : 0u;
        const uint32_t v = w ? (prm.fog ? S.vis[p * NW + lane] : valid) : 0u;
        const uint32_t nm = v & ~M;
        uint32_t *c = chm + p * GRL_OBS_CHANNELS * NWP + lane;
        const uint32_t mine = nm & own, enemy = nm & any_own & ~own;
        c[0 * NWP] = mine;
        c[1 * NWP] = enemy;
        c[2 * NWP] = mine;
        c[3 * NWP] = enemy;
        c[4 * NWP] = nm & ~any_own;
        c[5 * NWP] = nm & CG;
        c[6 * NWP] = v & M;
        c[7 * NWP] = v;
        c[8 * NWP] = ~v & valid;
      }
    }
  }
  for (int t = lane; t < N + 4; t += 32) frac[t] = t < N ? army_frac((uint32_t)S.army[t]) : 0.f;
  __syncwarp();

  const int total = P * GRL_OBS_CHANNELS * N;  // floats in this game's block
  float *base = prm.obs + (size_t)game * total;
  // A 15x15 block is 16,200 bytes: it starts and ends mid-sector, and a 32-byte sector completed by two different store
  // instructions costs the memory system far more than its bytes (tools/micro/store_holes.cu).  Consecutive games of one
  // warp therefore join their blocks into one sector-complete run: the sector two blocks share is written WHOLE by the
  // later game's pass (lanes 0-1, one instruction) — its first `lead` floats are the tail of the previous game's last
  // plane (player P-1, channel 8 = fog), evaluated from that game's slab, which is still in shared memory.
  const int lead = (int)(((size_t)game * total) & 7u);           // floats of this block's first sector that belong to the previous block
  const int trail = (int)(((size_t)(game + 1) * total) & 7u);    // floats of this block in the sector it shares with the next block
  const bool join_prev = GRL_OBS_JOIN && prev_slab != nullptr && lead != 0;
  const bool join_next = GRL_OBS_JOIN && next_in_warp && trail != 0;
  int head, end;
  if (join_prev) {
    if (lane < 2) {
      const uint32_t *pv = prev_slab + prm.L.off_vis + (P - 1) * NW;
      float v4[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int pos = 4 * lane + c;
        if (pos < lead) {
          const int t = N - lead + pos;
          const uint32_t seen = prm.fog ? ((pv[t >> 5] >> (t & 31)) & 1u) : 1u;
          v4[c] = seen ? 0.f : 1.f;
        } else {
          v4[c] = obs_element<N>(chm, frac, NWP, pos - lead);
        }
      }
      GRL_LIN_ST(reinterpret_cast<float4 *>(base - lead) + lane, make_float4(v4[0], v4[1], v4[2], v4[3]));
    }
    head = 8 - lead;
  } else {
    head = (int)((4u - (uint32_t)(((size_t)game * total) & 3u)) & 3u);  // floats before the 16-byte aligned body
    if (lane < head) GRL_LIN_ST(base + lane, obs_element<N>(chm, frac, NWP, lane));
  }
  if (join_next) {
    end = total - trail;  // the shared sector is left to the next game's pass
  } else {
    end = head + 4 * ((total - head) / 4);
    if (lane < total - end) GRL_LIN_ST(base + end + lane, obs_element<N>(chm, frac, NWP, end + lane));
  }
  const int body4 = (end - head) / 4;
#if GRL_STRADDLE_INLINE == 3
  // The P*9-1 float4s that straddle two planes are evaluated here, one per lane, and parked in shared memory: inside
  // the sweep the per-element branch is divergent — one lane straddles in 17 of a 15x15 game's 32 rounds and the whole
  // warp pays four element evaluations each time (592 of 3,183 warp instructions per game).  The store itself stays in
  // the sweep: a 16-byte hole completed later costs far more than any of this (tools/micro/store_holes.cu).
  for (int j = lane; j < P * GRL_OBS_CHANNELS - 1; j += 32) {
    const int b = (j + 1) * N - head;  // plane boundary, in floats from the start of the aligned body
    if ((b & 3) && b > 0 && (b >> 2) < body4) {
      const int e = head + (b & ~3);
      sf[j] = make_float4(obs_element<N>(chm, frac, NWP, e), obs_element<N>(chm, frac, NWP, e + 1),
                          obs_element<N>(chm, frac, NWP, e + 2), obs_element<N>(chm, frac, NWP, e + 3));
    }
  }
  __syncwarp();
#endif
  const char *lutb = reinterpret_cast<const char *>(lut);
  float4 *body = reinterpret_cast<float4 *>(base + head);
#if GRL_OBS_PLANEWISE && GRL_STRADDLE_INLINE == 3
  // Plane by plane instead of one flat index space: the plane, its channel and its mask row are loop constants (no
  // division or modulo per store, a uniform branch for the two army channels), the tile index is a shift of the
  // position.  The float4s wholly inside a plane take two rounds (32 + 23/24 lanes); the one that straddles into the
  // next plane rides on the second round's next lane, from the values parked above — so the stores stay one
  // contiguous run in address order.
  {
    const int planes = P * GRL_OBS_CHANNELS;
    int k = 0;
#pragma unroll 1
    for (int pl = 0; pl < planes; pl++) {
      const int lo = pl * N - head, hi = lo + N;          // this plane's floats, relative to the aligned body
      int i_lo = lo <= 0 ? 0 : (lo + 3) >> 2;             // first float4 that starts inside the plane
      int i_hi = (hi - 4) >> 2;                           // last float4 that ends inside it (hi >= 4 always)
      if (i_hi > body4 - 1) i_hi = body4 - 1;
      const int count = i_hi - i_lo + 1;
      const bool strad = (hi & 3) != 0 && i_hi + 1 < body4 && pl + 1 < planes;
      const uint32_t *row = chm + pl * NWP;
      const bool armyk = k < 2;
      for (int q = lane; q < count + (strad ? 1 : 0); q += 32) {
        const int i = i_lo + q;
        float4 val;
        if (q < count) {
          const int t = 4 * i - lo;
          const uint32_t *wp = row + (t >> 5);
          const uint32_t nib = __funnelshift_r(wp[0], wp[1], t & 31) & 0xfu;
          val = *reinterpret_cast<const float4 *>(lutb + nib * 16u);
          if (armyk && nib) {
            val.x *= frac[t];
            val.y *= frac[t + 1];
            val.z *= frac[t + 2];
            val.w *= frac[t + 3];
          }
        } else {
          val = sf[pl];
        }
        GRL_LIN_ST(body + i, val);
      }
      k = k == GRL_OBS_CHANNELS - 1 ? 0 : k + 1;
    }
    __syncwarp();
    return;
  }
#endif
#if GRL_OBS_INCR
  // (plane, tile, channel) of a lane's float4 advance by 128 floats per round: running counters instead of a
  // division and a modulo per store (N > 128, so a round crosses at most one plane boundary)
  int e = head + 4 * lane;
  int plane = e / N, t = e - plane * N, k = plane % GRL_OBS_CHANNELS;
  const uint32_t *row = chm + plane * NWP;
#endif
#pragma unroll 2
  for (int i = lane; i < body4; i += 32) {
#if !GRL_OBS_INCR
    const int e = head + 4 * i;
    const int plane = e / N, t = e - plane * N;
    const int k = plane % GRL_OBS_CHANNELS;
    const uint32_t *row = chm + plane * NWP;
#endif
    const uint32_t *wp = row + (t >> 5);
    uint32_t nib = __funnelshift_r(wp[0], wp[1], t & 31) & 0xfu;  // rows are zero from bit N on
    float4 val;
    if (t + 3 < N) {
      val = *reinterpret_cast<const float4 *>(lutb + nib * 16u);
      if (k < 2 && nib) {
        val.x *= frac[t];
        val.y *= frac[t + 1];
        val.z *= frac[t + 2];
        val.w *= frac[t + 3];
      }
    } else {
#if GRL_STRADDLE_INLINE == 3  // evaluated before the sweep, one per lane
      val = sf[plane];
#elif GRL_STRADDLE_INLINE == 2  // comparison builds: four per-element evaluations
      val.x = obs_element<N>(chm, frac, NWP, e);
      val.y = obs_element<N>(chm, frac, NWP, e + 1);
      val.z = obs_element<N>(chm, frac, NWP, e + 2);
      val.w = obs_element<N>(chm, frac, NWP, e + 3);
#else
      // the float4 straddles two planes (P*9-1 of them per game, but in 17 of a 15x15 game's 32 rounds one lane
      // of the warp is here): the first r tiles close this plane, the rest open the next one.  Kept short — the
      // whole warp waits for it — and kept a full 128-bit store: completing the sector with a later or scalar
      // store costs more in partial-sector writes than it saves (measured 1.02 ms against 0.88 ms per 262,144 games).
      const int r = N - t;  // 1..3
      nib |= (chm[(plane + 1) * NWP] << r) & 0xfu;
      val = *reinterpret_cast<const float4 *>(lutb + nib * 16u);
      const bool armyA = k < 2, armyB = k == 0 || k == GRL_OBS_CHANNELS - 1;  // the next plane is k+1, or the next view's 0
      if (armyA || armyB) {
        const float f0 = armyA ? frac[t] : 1.f;                                                   // tile t is always ours
        const float f1 = (1 < r) ? (armyA ? frac[t + 1] : 1.f) : (armyB ? frac[1 - r] : 1.f);
        const float f2 = (2 < r) ? (armyA ? frac[t + 2] : 1.f) : (armyB ? frac[2 - r] : 1.f);
        const float f3 = armyB ? frac[3 - r] : 1.f;                                               // tile 3 is always theirs
        val.x *= f0;
        val.y *= f1;
        val.z *= f2;
        val.w *= f3;
      }
#endif
    }
    GRL_LIN_ST(body + i, val);
#if GRL_OBS_INCR
    e += 128;
    t += 128;
    while (t >= N) {  // once at most for the boards that come here (N > 128)
      t -= N;
      plane += 1;
      row += NWP;
      k = k == GRL_OBS_CHANNELS - 1 ? 0 : k + 1;
    }
#endif
  }
  __syncwarp();
}

template <int PT, int TW, int TH, int LG, bool DO_STEP, bool DO_OUT, bool GYM>
__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32, TurnOccupancy<PT, LG>::kMinBlocks)
    grl_turn_kernel(const __grid_constant__ GrlKParams prm, const __grid_constant__ GrlGymK gk) {
  static_assert(!GYM || (DO_STEP && DO_OUT), "the fused gym step is a turn plus read-outs");
  constexpr int GPW = 32 / LG;  // games per warp
  static_assert(LG == 32 || LG == 16 || LG == 8 || LG == 4, "a group is 4, 8, 16 or 32 lanes");
  static_assert(LG >= PT || LG == 32, "per-player scalars are written by one lane each");
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t s_bar[GRL_WARPS_PER_CTA * GPW];
  __shared__ __align__(16) float4 s_lut[16];  // nibble -> four 0/1 floats (observation planes)
#if GRL_STAGE_SCALARS
  constexpr int GPC = GRL_WARPS_PER_CTA * GPW;  // games per CTA
  __shared__ uint8_t s_sc_done[DO_OUT ? GPC : 1], s_sc_winner[DO_OUT ? GPC : 1], s_sc_err[DO_OUT ? GPC : 1];
  __shared__ uint32_t s_sc_reward[DO_OUT ? GPC * PT : 1], s_sc_aidx[DO_OUT ? GPC * PT : 1];
#endif
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
  constexpr bool kSnap = GRL_DIRTY_WB && (LG == 32 || GRL_PACKED_SNAPSHOT);
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
  if (prm.use_tma && l == 0) {
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
    if (prm.use_tma) {
      if (l == 0) {
        const bool want_snap = kSnap && DO_STEP;
        mbar_expect_tx(bar, (uint32_t)(L.slab_words + L.static_words + (want_snap ? L.slab_words : 0)) * 4u);
#if GRL_L2_HINT
        const uint64_t pol = policy_evict_last();
        tma_load_hint(s, gslab, (uint32_t)L.slab_words * 4u, bar, pol);
        tma_load_hint(st, gstat, (uint32_t)L.static_words * 4u, bar, pol);
        if (want_snap) tma_load_hint(snap, gslab, (uint32_t)L.slab_words * 4u, bar, pol);
#else
        tma_load(s, gslab, (uint32_t)L.slab_words * 4u, bar);
        tma_load(st, gstat, (uint32_t)L.static_words * 4u, bar);
        if (want_snap)  // pre-turn snapshot for the dirty-sector write-back (an L2 hit on the same lines)
          tma_load(snap, gslab, (uint32_t)L.slab_words * 4u, bar);
#endif
        // warm L2 for a CTA that will be scheduled a couple of waves from now
        if (prm.prefetch_dist > 0 && game + prm.prefetch_dist < prm.B) {
          tma_prefetch_l2(prm.state + (size_t)(game + prm.prefetch_dist) * L.slab_words, (uint32_t)L.slab_words * 4u);
          tma_prefetch_l2(prm.statics + (size_t)(game + prm.prefetch_dist) * L.static_words, (uint32_t)L.static_words * 4u);
        }
      }
    } else {
      const uint4 *src = reinterpret_cast<const uint4 *>(gslab);
      uint4 *dst = reinterpret_cast<uint4 *>(s);
      for (int k = l; k < L.slab_words / 4; k += LG) dst[k] = src[k];
      const uint4 *src2 = reinterpret_cast<const uint4 *>(gstat);
      uint4 *dst2 = reinterpret_cast<uint4 *>(st);
      for (int k = l; k < L.static_words / 4; k += LG) dst2[k] = __ldg(src2 + k);
    }
    // decode this game's action slots while its slab lands
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
    if (prm.use_tma) mbar_wait(bar, 0u);
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
      skip = !gym_pre_phase<PT, LG>(prm, gk, s, st, s_act, alive, over, turn, game, g, W, H, N, NW);
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
      if (prm.use_tma && kSnap) {
        // sector k = words [8k, 8k+8) of the slab (slabs are 32-byte aligned and a whole number of sectors)
        const uint4 *now4 = reinterpret_cast<const uint4 *>(s);
        const uint4 *old4 = reinterpret_cast<const uint4 *>(snap);
        uint4 *dst4 = reinterpret_cast<uint4 *>(gslab);
        for (int k = l; k < L.slab_words / 8; k += LG) {
          const uint4 a0 = now4[2 * k], a1 = now4[2 * k + 1], b0 = old4[2 * k], b1 = old4[2 * k + 1];
          const uint32_t diff = (a0.x ^ b0.x) | (a0.y ^ b0.y) | (a0.z ^ b0.z) | (a0.w ^ b0.w) | (a1.x ^ b1.x) | (a1.y ^ b1.y) |
                                (a1.z ^ b1.z) | (a1.w ^ b1.w);
          if (diff) {
#if GRL_L2_HINT
            const uint64_t pol = policy_evict_last();
            st_hint_v4(dst4 + 2 * k, a0, pol);
            st_hint_v4(dst4 + 2 * k + 1, a1, pol);
#else
            dst4[2 * k] = a0;
            dst4[2 * k + 1] = a1;
#endif
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
#if GRL_STAGE_SCALARS
      const int gc = warp * GPW + sub;  // this game's index inside the CTA
      if (l == 0) {
        s_sc_done[gc] = over ? 1 : 0;
        const int n_alive = __popc(alive & pmask);  // engine.go:248-263
        s_sc_winner[gc] = (uint8_t)(int8_t)((over && n_alive == 1) ? (__ffs(alive & pmask) - 1) : -1);
        s_sc_err[gc] = (uint8_t)err;
      }
      if (l < P) {
        s_sc_reward[gc * P + l] = S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * l + GRL_PL_REWARD];
        s_sc_aidx[gc * P + l] = S.hdr[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * l + GRL_PL_ACTION_INDEX];
      }
#else
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
#endif
    }

    // ---- fused gym step: the tail of GeneralsEnv.step (generals_env.py:268-289) and the client's reward
    //      (:499-561, float64) from the PlayerState before and after the turn --------------------------
    if constexpr (GYM) {
      const int tiles1 = __reduce_add_sync(g.seg, act_lane ? __popc(S.list[l]) : 0);
      if (l == 0) {
        const int valid = skip ? 0 : 1;
        const int tn = gk.turns[game] + valid, cl = gk.calls[game] + 1;
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

  // The CTA's scalar results leave as contiguous runs written by its first warp once every warp has finished its
  // planes: one transaction per plane and CTA instead of one partial-sector write per game and value (which is
  // also one PCIe write each when the caller's result planes live in pinned host memory).
  auto flush_scalars = [&]() {
#if GRL_STAGE_SCALARS
    __syncthreads();
    if (warp == 0) {
      const int cta_game0 = prm.game0 + blockIdx.x * GPC;
      const int n = min(GPC, game_end - cta_game0);
      for (int i = lane; i < n; i += 32) {
        if (prm.done) prm.done[cta_game0 + i] = s_sc_done[i];
        if (prm.winner) prm.winner[cta_game0 + i] = (int8_t)s_sc_winner[i];
        if (prm.step_error) prm.step_error[cta_game0 + i] = s_sc_err[i];
      }
      for (int i = lane; i < n * P; i += 32) {
        if (prm.reward) prm.reward[(size_t)cta_game0 * P + i] = __uint_as_float(s_sc_reward[i]);
        if (prm.action_index) prm.action_index[(size_t)cta_game0 * P + i] = (int32_t)s_sc_aidx[i];
      }
    }
#endif
  };

  if constexpr (GYM) {
    // the client's read-outs of the new state, one game of the warp after the other (obs, N*5 mask, PlayerState)
    const Geo g32 = make_geo(prm, W, lane, 32);
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
                                                    s_lut, s_obs, game_g, lane, g32);
      else
        gym_emit<0>(prm, gk.max_turns, gk.logtab, gk.obs, gk.mask, gk.stats, sg, sg + L.slab_words, s_obs, game_g, lane, g32);
    }
    flush_scalars();
    return;
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
    if (TW > 0 && !GRL_OBS_CHUNK_MAJOR) {
      // baked boards: the whole warp writes one game's block after the other, from the slabs in shared memory
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
  flush_scalars();
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

template <int PT, int LG>
__device__ __noinline__ bool gym_pre_phase(const GrlKParams &prm, const GrlGymK &gk, uint32_t *s, const uint32_t *st,
                                           uint32_t *s_act, uint32_t alive, bool over, uint32_t turn_before, int game, Geo g,
                                           int W, int H, int N, int NW) {
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
  const long long a0 = gk.action[game];
  const bool ok0 = gym_ok(a0, 0);
  if (ok0) put(a0, 0, 0);
  if (gk.opponent_action) {
    const long long a1 = gk.opponent_action[game];
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
// Reset: freshly uploaded slabs carry ownership, armies and terrain; this kernel performs
// the turn-0 set-up of engine_initializer.go:113-143,218-225 (players alive, full stats,
// full fog, game-over check), preserving the env's lifetime counters.
// ---------------------------------------------------------------------------------------
template <int PT>
__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32)
    grl_reset_kernel(const __grid_constant__ GrlKParams prm, const uint32_t *__restrict__ src_state,
                     const uint32_t *__restrict__ src_static, const int32_t *__restrict__ env_ids, int n,
                     const int *__restrict__ n_dev) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (n_dev) n = min(n, *n_dev);  // device-side count (grl_gym_autoreset)
  const GrlLayout &L = prm.L;
  const int P = prm.P, NW = prm.NW, N = prm.N;
  const Geo g = make_geo(prm, prm.W, lane, 32);
  for (int i = blockIdx.x * GRL_WARPS_PER_CTA + warp; i < n; i += gridDim.x * GRL_WARPS_PER_CTA) {
    const int game = env_ids ? env_ids[i] : i;
    if (game < 0 || game >= prm.B) continue;
    const uint32_t *ss = src_state + (size_t)i * L.slab_words;
    const uint32_t *sst = src_static + (size_t)i * L.static_words;
    uint32_t *ds = prm.state + (size_t)game * L.slab_words;
    uint32_t *dst = const_cast<uint32_t *>(prm.statics) + (size_t)game * L.static_words;
    for (int k = lane; k < L.static_words; k += 32) dst[k] = sst[k];
    for (int k = L.off_army + lane; k < L.slab_words; k += 32) ds[k] = ss[k];
    const bool act = lane < NW;
    const uint32_t G = act ? sst[2 * NW + lane] : 0u;
    const uint16_t *army = reinterpret_cast<const uint16_t *>(ss + L.off_army);
    uint32_t alive = 0;
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P) {
        uint32_t own = act ? ss[L.off_own + p * NW + lane] : 0u;
        int total = sum_army_over<32>(own, army, NW, N, g);
        uint32_t gen = own & G;
        int gi = gen ? (32 * lane + 31 - __clz(gen)) : -1;
        gi = __reduce_max_sync(FULL, gi);
        if (gi >= 0) alive |= 1u << p;
        uint32_t v = (gi >= 0 && prm.fog) ? dilate3<32>(own, g) : 0u;  // players start Alive; stats then sets Alive = has general
        if (act) {
          ds[L.off_own + p * NW + lane] = own;
          ds[L.off_list + p * NW + lane] = own;
          ds[L.off_vis + p * NW + lane] = v;
        }
        if (lane == 0) {
          uint32_t *h = ds + GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p;
          h[GRL_PL_ARMY_COUNT] = (uint32_t)total;
          h[GRL_PL_GENERAL_IDX] = (uint32_t)gi;
          h[GRL_PL_TRUE_ARMY] = (uint32_t)total;
          h[GRL_PL_REWARD] = 0u;
          h[GRL_PL_ACTION_INDEX] = 0xffffffffu;
        }
      }
    }
    if (act) {
      ds[L.off_changed + lane] = 0u;
      ds[L.off_vchg + lane] = 0u;
    }
    if (lane == 0) {
      int n_alive = __popc(alive);
      bool over = P > 1 ? (n_alive <= 1) : (n_alive == 0);
      ds[GRL_HDR_TURN] = 0u;
      ds[GRL_HDR_FLAGS] = alive | (over ? GRL_FLAG_OVER : 0u);
      ds[GRL_HDR_OVERFLOW] = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Read-out kernels that are not on the hot path (one thread per tile / per word).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool bit_of(const uint32_t *w, int t) { return (w[t >> 5] >> (t & 31)) & 1u; }

// variant 0: engine mask bytes (U,R,D,L; list-based; army > 1)   rules/legal_moves.go:19-73
// variant 1: serializer mask bytes (U,D,L,R; ownership scan; army >= 2)  serializer.go:112-176
__global__ void grl_mask_bytes_kernel(const GrlKParams prm, int variant, uint8_t *__restrict__ out) {
  const GrlLayout &L = prm.L;
  const size_t total = (size_t)prm.B * prm.P * prm.N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx % prm.N);
    const int p = (int)((idx / prm.N) % prm.P);
    const int game = (int)(idx / ((size_t)prm.N * prm.P));
    const uint32_t *s = prm.state + (size_t)game * L.slab_words;
    const uint32_t *M = prm.statics + (size_t)game * L.static_words;
    const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
    const int x = t % prm.W, y = t / prm.W;
    bool src = bit_of(s + L.off_own + p * prm.NW, t) && army[t] > 1;
    if (variant == 0) src = src && bit_of(s + L.off_list + p * prm.NW, t) && ((s[GRL_HDR_FLAGS] >> p) & 1u);
    const bool up = src && y > 0 && !bit_of(M, t - prm.W);
    const bool down = src && y < prm.H - 1 && !bit_of(M, t + prm.W);
    const bool left = src && x > 0 && !bit_of(M, t - 1);
    const bool right = src && x < prm.W - 1 && !bit_of(M, t + 1);
    uchar4 v = variant == 0 ? make_uchar4(up, right, down, left) : make_uchar4(up, down, left, right);
    reinterpret_cast<uchar4 *>(out)[idx] = v;
  }
}

// PlayerVisibility (visibility_optimized.go:166-195): visible = bit p; fog = !visible && type != normal
__global__ void grl_visibility_kernel(const GrlKParams prm, uint8_t *__restrict__ visible, uint8_t *__restrict__ fog) {
  const GrlLayout &L = prm.L;
  const size_t total = (size_t)prm.B * prm.P * prm.N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx % prm.N);
    const int p = (int)((idx / prm.N) % prm.P);
    const int game = (int)(idx / ((size_t)prm.N * prm.P));
    const uint32_t *s = prm.state + (size_t)game * L.slab_words;
    const uint32_t *stt = prm.statics + (size_t)game * L.static_words;
    bool v, f = false;
    if (!prm.fog) {
      v = true;
    } else {
      v = bit_of(s + L.off_vis + p * prm.NW, t);
      bool special = bit_of(stt, t) || bit_of(stt + prm.NW, t) || bit_of(stt + 2 * prm.NW, t);
      f = !v && special;
    }
    if (visible) visible[idx] = v;
    if (fog) fog[idx] = f;
  }
}

// generals_gym read-outs (python/generals_gym/generals_env.py:291-387) of every player's
// fog-filtered proto view (internal/grpc/gameserver/server.go:556-582).  One thread per
// (env, player, tile); channel planes are written coalesced over tiles.  Not on the turn path.
__global__ void __launch_bounds__(256)
    grl_gym_kernel(const GrlKParams prm, int max_turns, const float *__restrict__ logtab, float *__restrict__ obs,
                   uint8_t *__restrict__ mask, int32_t *__restrict__ stats) {
  const GrlLayout &L = prm.L;
  const int N = prm.N, P = prm.P, NW = prm.NW, W = prm.W, H = prm.H;
  const size_t total = (size_t)prm.B * P * N;
  const int lane = threadIdx.x & 31;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  // every lane of a warp runs the same number of iterations: the mask bytes are transposed through shuffles
  for (size_t base = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < total; base += step) {
    const size_t idx = base + lane;
    const bool live = idx < total;
    uint32_t flags = 0;  // bit d = mask[idx*5 + d]
    if (live) {
      int t, p, game;
      if (total <= 0xffffffffull) {  // 32-bit index arithmetic whenever it fits (64-bit div/mod is ~100 instructions)
        const uint32_t i32 = (uint32_t)idx, gp = i32 / (uint32_t)N;
        t = (int)(i32 - gp * (uint32_t)N);
        game = (int)(gp / (uint32_t)P);
        p = (int)(gp - (uint32_t)game * (uint32_t)P);
      } else {
        t = (int)(idx % N);
        p = (int)((idx / N) % P);
        game = (int)(idx / ((size_t)N * P));
      }
      const uint32_t *s = prm.state + (size_t)game * L.slab_words;
      const uint32_t *stt = prm.statics + (size_t)game * L.static_words;
      const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
      const bool vis = prm.fog ? bit_of(s + L.off_vis + p * NW, t) : true;
      const bool mine = vis && bit_of(s + L.off_own + p * NW, t);
      bool owned = false;
      for (int q = 0; q < P; q++) owned = owned || bit_of(s + L.off_own + q * NW, t);
      const bool mnt = bit_of(stt, t), city = bit_of(stt + NW, t), gen = bit_of(stt + 2 * NW, t);
      const uint32_t a = vis ? army[t] : 0u;  // hidden and fogged tiles show no owner and no army
      if (obs) {
        float *o = obs + ((size_t)game * P + p) * GRL_GYM_CHANNELS * N + t;
        // min(turn / max_turns, 1.0) is computed in float64 by the client and stored as float32; for integers
        // below 2^24 the correctly rounded float32 quotient is the same number (no double-rounding case exists
        // for denominators below 2^28), and the B200's float64 rate would make this the kernel's hot spot
        const float tf = fminf(__fdiv_rn((float)s[GRL_HDR_TURN], (float)max_turns), 1.0f);
        __stcs(o + 0 * N, vis ? 1.f : 0.f);
        __stcs(o + 1 * N, mine ? 0.5f : ((vis && owned) ? 1.f : 0.f));
        __stcs(o + 2 * N, a > 0u ? logtab[a] : 0.f);
        __stcs(o + 3 * N, (!mnt && !city && !gen) ? 1.f : 0.f);  // a hidden tile is a normal tile by definition
        __stcs(o + 4 * N, mnt ? 1.f : 0.f);
        __stcs(o + 5 * N, city ? 1.f : 0.f);
        __stcs(o + 6 * N, gen ? 1.f : 0.f);
        __stcs(o + 7 * N, tf);
        __stcs(o + 8 * N, 0.f);
      }
      if (mask) {
        const int x = t % W, y = t / W;
        const bool src = mine && a > 1u;
        const bool up = src && y > 0 && !bit_of(stt, t - W);
        const bool right = src && x < W - 1 && !bit_of(stt, t + 1);
        const bool down = src && y < H - 1 && !bit_of(stt, t + W);
        const bool left = src && x > 0 && !bit_of(stt, t - 1);
        flags = (up ? 1u : 0u) | (right ? 2u : 0u) | (down ? 4u : 0u) | (left ? 8u : 0u) | ((up || right || down || left) ? 16u : 0u);
      }
      if (stats && t == 0) {
        int tiles = 0;
        for (int k = 0; k < NW; k++) tiles += __popc(s[L.off_list + p * NW + k]);
        int32_t *so = stats + ((size_t)game * P + p) * 4;
        so[0] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ARMY_COUNT];
        so[1] = tiles;
        so[2] = (int32_t)((s[GRL_HDR_FLAGS] >> p) & 1u);
        so[3] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_GENERAL_IDX];
      }
    }
    if (mask) {
      // the warp's 32 x 5 mask bytes are contiguous in memory (address = idx*5 + d): lane i writes bytes
      // i, i+32, ... so that every store instruction covers one 32-byte run instead of a stride-5 scatter
      uint8_t *mb = mask + base * 5;
#pragma unroll
      for (int r = 0; r < 5; r++) {
        const int j = 32 * r + lane;
        const uint32_t v = __shfl_sync(FULL, flags, j / 5);
        if (base + j / 5 < total) mb[j] = (uint8_t)((v >> (j % 5)) & 1u);
      }
    }
  }
}

// generals_gym read-outs, warp-per-game version: the game's masks are staged once in shared memory
// and every player's [9][N] observation block is written as one linear, 16-byte aligned sweep of
// 128-bit stores (as the turn kernel's observation writer does); the N*5 mask bytes go out as an
// aligned 32-bit sweep.  Shared-memory words per warp: see grl_gym_smem_words().
__host__ __device__ inline int grl_gym_smem_words(int P, int NW, int N, int mode) {
  const int PT = P <= 2 ? 2 : (P <= 4 ? 4 : 8);
  int m;
  if (mode == GRL_GYM_EMIT_QUADS) {  // gym_emit_quads: dir rows + one player's mask words
    m = PT * 4 * NW + 5 * ((N + 3) / 4);
  } else if (mode == GRL_GYM_EMIT_LINEAR) {  // gym_emit_linear: dir rows + (channel masks + log plane | the game's mask bytes)
    const int lin_obs = (((PT * GRL_GYM_CHANNELS + PT) * (NW + 1) + N + 4 + 3) & ~3) + 4 * PT * GRL_GYM_CHANNELS + 4;
    const int lin_mask = (P * N * 5 + 8 + 3) / 4;
    m = PT * 4 * NW + (lin_obs > lin_mask ? lin_obs : lin_mask);
  } else {  // gym_emit
    m = (3 * P + 5 + 5 * P) * (NW + 1) + N + 4;
  }
  return (m + 3) & ~3;
}

struct GymPlanes {  // shared-memory views of one game (each mask has NW + 1 words, the last one zero)
  const uint32_t *vis, *mine, *enemy;  // [P][NWP]
  const uint32_t *normal, *M, *C, *G;  // [NWP]
  const float *logv;                   // [N + 4]
  float tf;
  int NWP, N;
};

__device__ __forceinline__ float gym_value(const GymPlanes &g, int p, int plane, int t) {
  const int w = t >> 5, b = t & 31;
  switch (plane) {
    case 0: return ((g.vis[p * g.NWP + w] >> b) & 1u) ? 1.f : 0.f;
    case 1: return ((g.mine[p * g.NWP + w] >> b) & 1u) ? 0.5f : (((g.enemy[p * g.NWP + w] >> b) & 1u) ? 1.f : 0.f);
    case 2: return ((g.vis[p * g.NWP + w] >> b) & 1u) ? g.logv[t] : 0.f;
    case 3: return ((g.normal[w] >> b) & 1u) ? 1.f : 0.f;
    case 4: return ((g.M[w] >> b) & 1u) ? 1.f : 0.f;
    case 5: return ((g.C[w] >> b) & 1u) ? 1.f : 0.f;
    case 6: return ((g.G[w] >> b) & 1u) ? 1.f : 0.f;
    case 7: return g.tf;
    default: return 0.f;
  }
}

__device__ __forceinline__ uint32_t nib_at(const uint32_t *m, int t) {
  return __funnelshift_r(m[t >> 5], m[(t >> 5) + 1], t & 31) & 0xfu;
}
#define GYM_NIB4(n, a) make_float4(((n)&1u) ? (a) : 0.f, ((n)&2u) ? (a) : 0.f, ((n)&4u) ? (a) : 0.f, ((n)&8u) ? (a) : 0.f)

__device__ __forceinline__ float4 gym_value4(const GymPlanes &g, int p, int plane, int t) {  // t + 3 < N
  switch (plane) {
    case 0: { const uint32_t n = nib_at(g.vis + p * g.NWP, t); return GYM_NIB4(n, 1.f); }
    case 1: {
      const uint32_t a = nib_at(g.mine + p * g.NWP, t), e = nib_at(g.enemy + p * g.NWP, t);
      return make_float4((a & 1u) ? 0.5f : ((e & 1u) ? 1.f : 0.f), (a & 2u) ? 0.5f : ((e & 2u) ? 1.f : 0.f),
                         (a & 4u) ? 0.5f : ((e & 4u) ? 1.f : 0.f), (a & 8u) ? 0.5f : ((e & 8u) ? 1.f : 0.f));
    }
    case 2: {
      const uint32_t n = nib_at(g.vis + p * g.NWP, t);
      return make_float4((n & 1u) ? g.logv[t] : 0.f, (n & 2u) ? g.logv[t + 1] : 0.f, (n & 4u) ? g.logv[t + 2] : 0.f,
                         (n & 8u) ? g.logv[t + 3] : 0.f);
    }
    case 3: { const uint32_t n = nib_at(g.normal, t); return GYM_NIB4(n, 1.f); }
    case 4: { const uint32_t n = nib_at(g.M, t); return GYM_NIB4(n, 1.f); }
    case 5: { const uint32_t n = nib_at(g.C, t); return GYM_NIB4(n, 1.f); }
    case 6: { const uint32_t n = nib_at(g.G, t); return GYM_NIB4(n, 1.f); }
    case 7: return make_float4(g.tf, g.tf, g.tf, g.tf);
    default: return make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// One game's gym read-outs by a whole warp.  `s` / `stt` are the game's slab and terrain words (global
// memory in grl_gym_warp_kernel, the shared-memory copy in the fused gym step); `sw` is the warp's
// scratch of grl_gym_smem_words() words; `g` is the full-warp (LG = 32) geometry.  NT > 0 bakes the tile
// count in (the element -> (plane, tile) divisions become multiplications).
template <int NT>
__device__ __forceinline__ void gym_emit(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                         float *__restrict__ obs, uint8_t *__restrict__ mask, int32_t *__restrict__ stats,
                                         const uint32_t *s, const uint32_t *stt, uint32_t *sw, int game, int lane,
                                         const Geo &g) {
  const GrlLayout &L = prm.L;
  const int N = NT ? NT : prm.N, P = prm.P, NW = NT ? (NT + 31) / 32 : prm.NW, NWP = NW + 1;
  uint32_t *s_vis = sw, *s_mine = s_vis + P * NWP, *s_enemy = s_mine + P * NWP;
  uint32_t *s_normal = s_enemy + P * NWP, *s_M = s_normal + NWP, *s_C = s_M + NWP, *s_G = s_C + NWP, *s_pad = s_G + NWP;
  uint32_t *s_dir = s_pad + NWP;  // [P][5][NWP]: up, right, down, left, any
  float *s_logv = reinterpret_cast<float *>(s_dir + 5 * P * NWP);
  {
    const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
    const bool w = lane < NW;
    const uint32_t M = w ? stt[lane] : 0u, C = w ? stt[NW + lane] : 0u, G = w ? stt[2 * NW + lane] : 0u;
    uint32_t any_own = 0;
    for (int p = 0; p < P; p++) any_own |= w ? s[L.off_own + p * NW + lane] : 0u;
    const uint32_t gt1 = army_gt1_mask<32>(army, NW, N, g);
    const DirMasks dm = dir_targets<32>(M, g);
    if (lane < NWP) {
      s_normal[lane] = g.valid & ~(M | C | G) & (w ? ~0u : 0u);
      s_M[lane] = M;
      s_C[lane] = C;
      s_G[lane] = G;
    }
    for (int p = 0; p < P; p++) {
      const uint32_t own = w ? s[L.off_own + p * NW + lane] : 0u;
      const uint32_t v = w ? (prm.fog ? s[L.off_vis + p * NW + lane] : g.valid) : 0u;
      const uint32_t mine = v & own, src = mine & gt1;
      if (lane < NWP) {
        s_vis[p * NWP + lane] = v;
        s_mine[p * NWP + lane] = mine;
        s_enemy[p * NWP + lane] = v & any_own & ~own;
        uint32_t *d = s_dir + p * 5 * NWP + lane;
        const uint32_t up = src & dm.up, right = src & dm.right, down = src & dm.down, left = src & dm.left;
        d[0 * NWP] = up;
        d[1 * NWP] = right;
        d[2 * NWP] = down;
        d[3 * NWP] = left;
        d[4 * NWP] = up | right | down | left;
      }
    }
    for (int t = lane; t < N + 4; t += 32) s_logv[t] = t < N ? logtab[army[t]] : 0.f;  // logtab[0] == 0
    __syncwarp();

    GymPlanes gp;
    gp.vis = s_vis;
    gp.mine = s_mine;
    gp.enemy = s_enemy;
    gp.normal = s_normal;
    gp.M = s_M;
    gp.C = s_C;
    gp.G = s_G;
    gp.logv = s_logv;
    gp.NWP = NWP;
    gp.N = N;
    gp.tf = fminf(__fdiv_rn((float)s[GRL_HDR_TURN], (float)max_turns), 1.0f);
    if (obs) {
      const int block = GRL_GYM_CHANNELS * N;  // floats per (game, player)
      for (int p = 0; p < P; p++) {
        const size_t off = ((size_t)game * P + p) * block;
        float *base = obs + off;
        const int head = (int)((4u - (uint32_t)(off & 3u)) & 3u);
        const int body4 = (block - head) / 4, tail0 = head + 4 * body4;
        if (lane < head) __stcs(base + lane, gym_value(gp, p, lane / N, lane % N));
        if (lane < block - tail0) __stcs(base + tail0 + lane, gym_value(gp, p, (tail0 + lane) / N, (tail0 + lane) % N));
        float4 *body = reinterpret_cast<float4 *>(base + head);
        for (int i = lane; i < body4; i += 32) {
          const int e = head + 4 * i, plane = e / N, t = e - plane * N;
          float4 val;
          if (t + 3 < N) {
            val = gym_value4(gp, p, plane, t);
          } else {
            val.x = gym_value(gp, p, plane, t);
            val.y = gym_value(gp, p, (e + 1) / N, (e + 1) % N);
            val.z = gym_value(gp, p, (e + 2) / N, (e + 2) % N);
            val.w = gym_value(gp, p, (e + 3) / N, (e + 3) % N);
          }
          __stcs(body + i, val);
        }
      }
    }
    if (mask) {
      const int bytes = N * 5;
      for (int p = 0; p < P; p++) {
        const uint32_t *d = s_dir + p * 5 * NWP;
        const size_t off = ((size_t)game * P + p) * bytes;
        uint8_t *base = mask + off;
        auto flag = [&](int j) -> uint32_t {  // byte j = direction j%5 of tile j/5
          const int t = j / 5, k = j - 5 * t;
          return (d[k * NWP + (t >> 5)] >> (t & 31)) & 1u;
        };
        const int head = (int)((4u - (uint32_t)(off & 3u)) & 3u);
        const int body4 = (bytes - head) / 4, tail0 = head + 4 * body4;
        if (lane < head) base[lane] = (uint8_t)flag(lane);
        if (lane < bytes - tail0) base[tail0 + lane] = (uint8_t)flag(tail0 + lane);
        uint32_t *body = reinterpret_cast<uint32_t *>(base + head);
        for (int i = lane; i < body4; i += 32) {
          const int j = head + 4 * i;
          body[i] = flag(j) | (flag(j + 1) << 8) | (flag(j + 2) << 16) | (flag(j + 3) << 24);
        }
      }
    }
    if (stats && lane < P) {
      int tiles = 0;
      for (int k = 0; k < NW; k++) tiles += __popc(s[L.off_list + lane * NW + k]);
      int32_t *so = stats + ((size_t)game * P + lane) * 4;
      so[0] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * lane + GRL_PL_ARMY_COUNT];
      so[1] = tiles;
      so[2] = (int32_t)((s[GRL_HDR_FLAGS] >> lane) & 1u);
      so[3] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * lane + GRL_PL_GENERAL_IDX];
    }
    __syncwarp();
  }
}

// The same read-outs for baked boards with N % 4 == 0 (10x10, 20x20), from the slab in SHARED memory, in the
// plane-major order of obs_plane_major: lane l owns the tile quads q = l + 32c, reads every mask's nibble for its
// quads once (packed 4 bits per chunk), converts its quads' armies once, and the warp writes the game's
// [P][9][N] block as one linear sweep of 128-bit stores with compile-time addressing.  The N*5 mask bytes of a
// player (tile-major, {up,right,down,left,any} per tile) are 20 bytes per quad: a lane assembles its five words,
// the warp stages them in shared memory and copies them out as a linear sweep.
template <int PT, int N>
__device__ __forceinline__ void gym_emit_quads(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                               float *__restrict__ obs, uint8_t *__restrict__ mask,
                                               int32_t *__restrict__ stats, const uint32_t *s, const uint32_t *stt,
                                               const float4 *lut, uint32_t *sw, int game, int lane, const Geo &g) {
  static_assert(N % 4 == 0 && N <= 512, "quads of four tiles, at most four chunks of 32 quads");
  constexpr int NQ = N / 4, NCH = (NQ + 31) / 32, NW = (N + 31) / 32;
  const GrlLayout &L = prm.L;
  const int P = prm.P;
  const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
  const bool w = lane < NW;
  // ---- the four direction masks of every player in word layout -> shared memory [P][4][NW] ------------
  uint32_t *s_dir = sw, *s_stage = sw + PT * 4 * NW;
  if (mask) {
    const uint32_t M = w ? stt[lane] : 0u;
    const uint32_t gt1 = army_gt1_mask<32>(army, NW, N, g);
    const DirMasks dm = dir_targets<32>(M, g);
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P && w) {
        const uint32_t own = s[L.off_own + p * NW + lane];
        const uint32_t v = prm.fog ? s[L.off_vis + p * NW + lane] : g.valid;
        const uint32_t src = v & own & gt1;
        s_dir[(p * 4 + 0) * NW + lane] = src & dm.up;
        s_dir[(p * 4 + 1) * NW + lane] = src & dm.right;
        s_dir[(p * 4 + 2) * NW + lane] = src & dm.down;
        s_dir[(p * 4 + 3) * NW + lane] = src & dm.left;
      }
    }
    __syncwarp();
  }
  // ---- nibbles of this lane's quads ---------------------------------------------------------------------
  const int bsel = lane >> 1, bsh = 4 * (lane & 1);  // quad q -> byte q>>1, nibble q&1 of a mask's byte array
  const uint8_t *bM = reinterpret_cast<const uint8_t *>(stt);
  const uint8_t *bC = reinterpret_cast<const uint8_t *>(stt + NW);
  const uint8_t *bG = reinterpret_cast<const uint8_t *>(stt + 2 * NW);
  uint32_t mM = 0, mC = 0, mG = 0, mAny = 0, livem = 0;
  uint32_t nV[PT], nO[PT];
#pragma unroll
  for (int p = 0; p < PT; p++) nV[p] = nO[p] = 0;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    if (32 * c + lane < NQ) {
      const int b = 16 * c + bsel;
      livem |= 0xfu << (4 * c);
      mM |= ((bM[b] >> bsh) & 0xfu) << (4 * c);
      mC |= ((bC[b] >> bsh) & 0xfu) << (4 * c);
      mG |= ((bG[b] >> bsh) & 0xfu) << (4 * c);
#pragma unroll
      for (int p = 0; p < PT; p++) {
        if (p < P) {
          const uint8_t *bo = reinterpret_cast<const uint8_t *>(s + L.off_own + p * NW);
          const uint8_t *bv = reinterpret_cast<const uint8_t *>(s + L.off_vis + p * NW);
          nO[p] |= ((bo[b] >> bsh) & 0xfu) << (4 * c);
          nV[p] |= ((bv[b] >> bsh) & 0xfu) << (4 * c);
        }
      }
    }
  }
  uint32_t seen = 0;  // tiles some player sees: the only ones whose army reaches an observation
#pragma unroll
  for (int p = 0; p < PT; p++) {
    mAny |= nO[p];
    if (!prm.fog) nV[p] = livem;
    seen |= nV[p];
  }
  if (obs) {
    float f[NCH][4];  // log(army + 1) / 10 of the lane's quads
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      f[c][0] = f[c][1] = f[c][2] = f[c][3] = 0.f;
      if ((seen >> (4 * c)) & 0xfu) {
        const uint2 aw = *reinterpret_cast<const uint2 *>(army + 4 * (32 * c + lane));
        f[c][0] = __ldg(logtab + (aw.x & 0xffffu));
        f[c][1] = __ldg(logtab + (aw.x >> 16));
        f[c][2] = __ldg(logtab + (aw.y & 0xffffu));
        f[c][3] = __ldg(logtab + (aw.y >> 16));
      }
    }
    const float tf = fminf(__fdiv_rn((float)s[GRL_HDR_TURN], (float)max_turns), 1.0f);
    const float4 tf4 = make_float4(tf, tf, tf, tf), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const char *lutb = reinterpret_cast<const char *>(lut);
    auto nib4 = [&](uint32_t field, int c) -> float4 {  // chunk c's nibble of a packed field -> four 0/1 floats
      const uint32_t idx16 = (c == 0 ? (field << 4) : (field >> (4 * c - 4))) & 0xf0u;
      return *reinterpret_cast<const float4 *>(lutb + idx16);
    };
    float4 *gq = reinterpret_cast<float4 *>(obs + (size_t)game * P * GRL_GYM_CHANNELS * N) + lane;
    const uint32_t mN = livem & ~(mM | mC | mG);
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P) {
        const uint32_t v = nV[p], mine = v & nO[p], enemy = v & mAny & ~nO[p];
        float4 *gp = gq + p * GRL_GYM_CHANNELS * NQ;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          if (32 * c + lane < NQ) {
            const float4 vv = nib4(v, c), a = nib4(mine, c), e = nib4(enemy, c);
            __stcs(gp + 0 * NQ + 32 * c, vv);
            // ownership: 0.5 own, 1.0 enemy (mine and enemy are disjoint)
            __stcs(gp + 1 * NQ + 32 * c, make_float4(__fmaf_rn(a.x, 0.5f, e.x), __fmaf_rn(a.y, 0.5f, e.y),
                                                     __fmaf_rn(a.z, 0.5f, e.z), __fmaf_rn(a.w, 0.5f, e.w)));
            __stcs(gp + 2 * NQ + 32 * c, make_float4(vv.x * f[c][0], vv.y * f[c][1], vv.z * f[c][2], vv.w * f[c][3]));
            __stcs(gp + 3 * NQ + 32 * c, nib4(mN, c));
            __stcs(gp + 4 * NQ + 32 * c, nib4(mM, c));
            __stcs(gp + 5 * NQ + 32 * c, nib4(mC, c));
            __stcs(gp + 6 * NQ + 32 * c, nib4(mG, c));
            __stcs(gp + 7 * NQ + 32 * c, tf4);
            __stcs(gp + 8 * NQ + 32 * c, zero4);
          }
        }
      }
    }
  }
  if (mask) {
    constexpr int MW = 5 * NQ;  // words of one player's mask
    for (int p = 0; p < P; p++) {
      const uint8_t *bd = reinterpret_cast<const uint8_t *>(s_dir + p * 4 * NW);
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        const int q = 32 * c + lane;
        if (q < NQ) {
          const int b = 16 * c + bsel;
          const uint32_t U = (bd[b] >> bsh) & 0xfu, R = (bd[4 * NW + b] >> bsh) & 0xfu;
          const uint32_t D = (bd[8 * NW + b] >> bsh) & 0xfu, Lm = (bd[12 * NW + b] >> bsh) & 0xfu, A = U | R | D | Lm;
          // byte 5i+k of the quad = direction k of its tile i:  U0 R0 D0 L0 | A0 U1 R1 D1 | L1 A1 U2 R2 | D2 L2 A2 U3 | R3 D3 L3 A3.
          // Each nibble is spread to one 0/1 byte per tile (a multiply and a mask), then five byte permutes
          // pairs interleave the direction words into the 5-byte records.
          auto spread = [](uint32_t n) -> uint32_t { return (n * 0x00204081u) & 0x01010101u; };
          const uint32_t Ub = spread(U), Rb = spread(R), Db = spread(D), Lb = spread(Lm), Ab = spread(A);
          const uint32_t UR = __byte_perm(Ub, Rb, 0x5140), URh = __byte_perm(Ub, Rb, 0x7362);  // U0 R0 U1 R1 | U2 R2 U3 R3
          const uint32_t DL = __byte_perm(Db, Lb, 0x5140), DLh = __byte_perm(Db, Lb, 0x7362);  // D0 L0 D1 L1 | D2 L2 D3 L3
          uint32_t *o = s_stage + 5 * q;
          o[0] = __byte_perm(UR, DL, 0x5410);
          o[1] = __byte_perm(__byte_perm(UR, DL, 0x6320), Ab, 0x3214);
          o[2] = __byte_perm(__byte_perm(DL, URh, 0x5403), Ab, 0x3250);
          o[3] = __byte_perm(__byte_perm(DLh, URh, 0x6010), Ab, 0x3610);
          o[4] = __byte_perm(__byte_perm(URh, DLh, 0x0763), Ab, 0x7210);
        }
      }
      __syncwarp();
      uint8_t *base = mask + ((size_t)game * P + p) * (size_t)(N * 5);
      if constexpr (MW % 4 == 0) {
        uint4 *dst = reinterpret_cast<uint4 *>(base);
        const uint4 *src = reinterpret_cast<const uint4 *>(s_stage);
        for (int i = lane; i < MW / 4; i += 32) __stcs(dst + i, src[i]);
      } else {
        uint32_t *dst = reinterpret_cast<uint32_t *>(base);
        for (int i = lane; i < MW; i += 32) __stcs(dst + i, s_stage[i]);
      }
      __syncwarp();
    }
  }
  if (stats && lane < P) {
    int tiles = 0;
    for (int k = 0; k < NW; k++) tiles += __popc(s[L.off_list + lane * NW + k]);
    int32_t *so = stats + ((size_t)game * P + lane) * 4;
    so[0] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * lane + GRL_PL_ARMY_COUNT];
    so[1] = tiles;
    so[2] = (int32_t)((s[GRL_HDR_FLAGS] >> lane) & 1u);
    so[3] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * lane + GRL_PL_GENERAL_IDX];
  }
  __syncwarp();
}

// The same read-outs for baked boards with N % 4 != 0 (15x15), from the slab in SHARED memory: like obs_linear,
// the game's [P][9][N] block is one linear, 16-byte aligned sweep of 128-bit stores addressed by position in the
// block (plane = e / N, tile = e % N, compile-time N), a 4-bit window of the plane's staged channel mask going
// through the nibble table.  The game's [P][N*5] mask bytes are assembled per tile in shared memory, pre-shifted
// by the block's misalignment, and leave as an aligned 32-bit sweep.
template <int PT, int N>
__device__ __forceinline__ void gym_emit_linear(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                                float *__restrict__ obs, uint8_t *__restrict__ mask,
                                                int32_t *__restrict__ stats, const uint32_t *s, const uint32_t *stt,
                                                const float4 *lut, uint32_t *sw, int game, int lane, const Geo &g) {
  constexpr int NW = (N + 31) / 32, NWP = NW + 1, CH = GRL_GYM_CHANNELS;
  const GrlLayout &L = prm.L;
  const int P = prm.P;
  const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
  uint32_t *s_dir = sw;                                          // [PT][4][NW]
  uint32_t *chm = sw + PT * 4 * NW;                              // [PT*9][NWP] channel masks
  uint32_t *minem = chm + PT * CH * NWP;                         // [PT][NWP]   own tiles in sight (the 0.5 of plane 1)
  float *logv = reinterpret_cast<float *>(minem + PT * NWP);     // [N + 4]
  float4 *sf = reinterpret_cast<float4 *>(sw + ((PT * 4 * NW + (PT * CH + PT) * NWP + N + 4 + 3) & ~3));  // [PT*9] straddlers
  uint8_t *stage = reinterpret_cast<uint8_t *>(chm);             // the mask bytes reuse the observation staging
  const bool w = lane < NW;
  const uint32_t M = w ? stt[lane] : 0u, C = w ? stt[NW + lane] : 0u, G = w ? stt[2 * NW + lane] : 0u;
  if (mask) {
    const uint32_t gt1 = army_gt1_mask<32>(army, NW, N, g);
    const DirMasks dm = dir_targets<32>(M, g);
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P && w) {
        const uint32_t own = s[L.off_own + p * NW + lane];
        const uint32_t v = prm.fog ? s[L.off_vis + p * NW + lane] : g.valid;
        const uint32_t src = v & own & gt1;
        s_dir[(p * 4 + 0) * NW + lane] = src & dm.up;
        s_dir[(p * 4 + 1) * NW + lane] = src & dm.right;
        s_dir[(p * 4 + 2) * NW + lane] = src & dm.down;
        s_dir[(p * 4 + 3) * NW + lane] = src & dm.left;
      }
    }
  }
  if (obs) {
    if (lane < NWP) {
      const uint32_t valid = w ? g.valid : 0u;
      uint32_t any_own = 0;
#pragma unroll
      for (int p = 0; p < PT; p++)
        if (p < P && w) any_own |= s[L.off_own + p * NW + lane];
#pragma unroll
      for (int p = 0; p < PT; p++) {
        if (p < P) {
          const uint32_t own = w ? s[L.off_own + p * NW + lane] : 0u;
          const uint32_t v = w ? (prm.fog ? s[L.off_vis + p * NW + lane] : valid) : 0u;
          uint32_t *c = chm + p * CH * NWP + lane;
          c[0 * NWP] = v;
          c[1 * NWP] = v & any_own & ~own;  // enemy -> 1.0; own tiles come from minem -> 0.5
          c[2 * NWP] = v;                   // x log(army + 1) / 10
          c[3 * NWP] = valid & ~(M | C | G);
          c[4 * NWP] = M;
          c[5 * NWP] = C;
          c[6 * NWP] = G;
          c[7 * NWP] = valid;               // x turn / max_turns
          c[8 * NWP] = 0u;
          minem[p * NWP + lane] = v & own;
        }
      }
    }
    for (int t = lane; t < N + 4; t += 32) logv[t] = t < N ? __ldg(logtab + army[t]) : 0.f;  // logtab[0] == 0
    __syncwarp();

    const float tf = fminf(__fdiv_rn((float)s[GRL_HDR_TURN], (float)max_turns), 1.0f);
    auto elem = [&](int e) -> float {
      const int plane = e / N, t = e - plane * N, k = plane % CH;
      const uint32_t bit = (chm[plane * NWP + (t >> 5)] >> (t & 31)) & 1u;
      if (k == 1) return ((minem[(plane / CH) * NWP + (t >> 5)] >> (t & 31)) & 1u) ? 0.5f : (bit ? 1.f : 0.f);
      if (k == 2) return bit ? logv[t] : 0.f;
      if (k == 7) return tf;
      return bit ? 1.f : 0.f;
    };
    const int total = P * CH * N;  // floats in this game's block
    float *base = obs + (size_t)game * total;
    const int head = (int)((4u - (uint32_t)(((size_t)game * total) & 3u)) & 3u);
    const int body4 = (total - head) / 4;
    const int tail0 = head + 4 * body4;
    if (lane < head) __stcs(base + lane, elem(lane));
    if (lane < total - tail0) __stcs(base + tail0 + lane, elem(tail0 + lane));
    const char *lutb = reinterpret_cast<const char *>(lut);
    float4 *body = reinterpret_cast<float4 *>(base + head);
    for (int j = lane; j < P * CH - 1; j += 32) {  // the float4s that straddle two planes
      const int b = (j + 1) * N - head;
      if ((b & 3) && (b >> 2) < body4) {
        const int e = head + (b & ~3);
        sf[j] = make_float4(elem(e), elem(e + 1), elem(e + 2), elem(e + 3));
      }
    }
    __syncwarp();
#pragma unroll 2
    for (int i = lane; i < body4; i += 32) {
      const int e = head + 4 * i;
      const int plane = e / N, t = e - plane * N;
      const int k = plane % CH;
      const uint32_t *wp = chm + plane * NWP + (t >> 5);
      uint32_t nib = __funnelshift_r(wp[0], wp[1], t & 31) & 0xfu;  // rows are zero from bit N on
      float4 val;
      if (t + 3 < N) {
        val = *reinterpret_cast<const float4 *>(lutb + nib * 16u);
        if (k == 1) {
          const uint32_t *mp = minem + (plane / CH) * NWP + (t >> 5);
          const uint32_t nb2 = __funnelshift_r(mp[0], mp[1], t & 31) & 0xfu;
          const float4 m = *reinterpret_cast<const float4 *>(lutb + nb2 * 16u);
          val = make_float4(__fmaf_rn(m.x, 0.5f, val.x), __fmaf_rn(m.y, 0.5f, val.y), __fmaf_rn(m.z, 0.5f, val.z),
                            __fmaf_rn(m.w, 0.5f, val.w));
        } else if (k == 2) {
          if (nib) {
            val.x *= logv[t];
            val.y *= logv[t + 1];
            val.z *= logv[t + 2];
            val.w *= logv[t + 3];
          }
        } else if (k == 7) {
          val = make_float4(tf, tf, tf, tf);
        }
      } else {
        val = sf[plane];  // evaluated before the sweep, one per lane (see obs_linear)
      }
      __stcs(body + i, val);
    }
  }
  __syncwarp();
  if (mask) {
    const int total = P * N * 5;  // bytes of this game's block [P][N*5]
    const size_t goff = (size_t)game * total;
    const int mis = (int)(goff & 3u);
    for (int p = 0; p < P; p++) {
      const uint32_t *d = s_dir + p * 4 * NW;
      for (int t = lane; t < N; t += 32) {
        const int wd = t >> 5, b = t & 31;
        const uint32_t U = (d[wd] >> b) & 1u, R = (d[NW + wd] >> b) & 1u, D = (d[2 * NW + wd] >> b) & 1u,
                       Lm = (d[3 * NW + wd] >> b) & 1u;
        uint8_t *o = stage + mis + (p * N + t) * 5;
        o[0] = (uint8_t)U;
        o[1] = (uint8_t)R;
        o[2] = (uint8_t)D;
        o[3] = (uint8_t)Lm;
        o[4] = (uint8_t)(U | R | D | Lm);
      }
    }
    __syncwarp();
    uint8_t *base = mask + goff;
    const int head = (4 - mis) & 3;
    const int body4 = (total - head) / 4, tail0 = head + 4 * body4;
    if (lane < head) base[lane] = stage[mis + lane];
    if (lane < total - tail0) base[tail0 + lane] = stage[mis + tail0 + lane];
    uint32_t *dst = reinterpret_cast<uint32_t *>(base + head);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(stage + mis + head);
    for (int i = lane; i < body4; i += 32) __stcs(dst + i, src[i]);
  }
  if (stats && lane < P) {
    int tiles = 0;
    for (int k = 0; k < NW; k++) tiles += __popc(s[L.off_list + lane * NW + k]);
    int32_t *so = stats + ((size_t)game * P + lane) * 4;
    so[0] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * lane + GRL_PL_ARMY_COUNT];
    so[1] = tiles;
    so[2] = (int32_t)((s[GRL_HDR_FLAGS] >> lane) & 1u);
    so[3] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * lane + GRL_PL_GENERAL_IDX];
  }
  __syncwarp();
}

__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32)
    grl_gym_warp_kernel(const __grid_constant__ GrlKParams prm, int max_turns, const float *__restrict__ logtab,
                        float *__restrict__ obs, uint8_t *__restrict__ mask, int32_t *__restrict__ stats,
                        const int32_t *__restrict__ ids, int n_ids, const int *__restrict__ n_dev) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GrlLayout &L = prm.L;
  const Geo g = make_geo(prm, prm.W, lane, 32);
  uint32_t *sw = smem + warp * grl_gym_smem_words(prm.P, prm.NW, prm.N, GRL_GYM_EMIT_GENERIC);
  const int count = ids ? (n_dev ? min(n_ids, *n_dev) : n_ids) : prm.B;  // an id list restricts the read-outs to those envs
  for (int i = blockIdx.x * GRL_WARPS_PER_CTA + warp; i < count; i += gridDim.x * GRL_WARPS_PER_CTA) {
    const int game = ids ? ids[i] : i;
    gym_emit<0>(prm, max_turns, logtab, obs, mask, stats, prm.state + (size_t)game * L.slab_words,
                prm.statics + (size_t)game * L.static_words, sw, game, lane, g);
  }
}
#undef GYM_NIB4

// GeneralsEnv._action_index_to_game_action (generals_env.py:389-441), one thread per env
__global__ void grl_gym_encode_kernel(const GrlKParams prm, const long long *__restrict__ action_idx, int player, int slot,
                                      const uint8_t *__restrict__ mask, int skip_invalid, uint2 *__restrict__ actions,
                                      uint8_t *__restrict__ valid) {
  const int N = prm.N, P = prm.P, W = prm.W, H = prm.H, A = prm.A;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < prm.B; b += gridDim.x * blockDim.x) {
    const long long a = action_idx[b];
    const bool ok = a >= 0 && a < (long long)N * 5 && mask[((size_t)b * P + player) * N * 5 + a] != 0;
    uint2 rec = make_uint2(0u, 0u);
    if (ok) {
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
      const PackedAction pa = pack_action(player, fx, fy, tx, ty, info != 4);
      rec = make_uint2(pa.lo, pa.hi);
    }
    if (!ok && skip_invalid && slot == 0) rec.y |= (uint32_t)GRL_ACTION_FLAG_SKIP_ENV << 24;
    actions[(size_t)b * A + slot] = rec;
    if (!ok && skip_invalid && slot != 0)
      reinterpret_cast<uint8_t *>(actions + (size_t)b * A)[7] |= GRL_ACTION_FLAG_SKIP_ENV;
    if (valid) valid[b] = ok ? 1 : 0;
  }
}

// grl_gym_autoreset, step 1: the envs whose episode ended, compacted into an id list with their next seeds (device-side count)
__global__ void grl_gym_compact_kernel(const GrlKParams prm, const uint8_t *__restrict__ terminated,
                                       const uint8_t *__restrict__ truncated, long long base_seed, long long *__restrict__ episode,
                                       int32_t *__restrict__ turns, int32_t *__restrict__ calls, int32_t *__restrict__ ids,
                                       long long *__restrict__ seeds, int *__restrict__ count) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < prm.B; b += gridDim.x * blockDim.x) {
    if (terminated[b] | truncated[b]) {
      const int pos = atomicAdd(count, 1);
      const long long ep = episode[b] + 1;
      episode[b] = ep;
      ids[pos] = b;
      seeds[pos] = base_seed + b + ep * (long long)prm.B;
      turns[b] = 0;
      calls[b] = 0;
    }
  }
}

// step 2: player 0's last observation of every finished env, before its row is overwritten (one warp per env)
__global__ void __launch_bounds__(256) grl_gym_final_obs_kernel(const GrlKParams prm, const float *__restrict__ obs,
                                                                float *__restrict__ final_obs, const int32_t *__restrict__ ids,
                                                                const int *__restrict__ count) {
  const int lane = threadIdx.x & 31, n = *count, block = GRL_GYM_CHANNELS * prm.N;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += (gridDim.x * blockDim.x) >> 5) {
    const int b = ids[i];
    const float *src = obs + (size_t)b * prm.P * block;
    float *dst = final_obs + (size_t)b * block;
    for (int k = lane; k < block; k += 32) dst[k] = src[k];
  }
}

// step 3: the staging rows the map generator fills have to start zeroed
__global__ void grl_zero_rows_kernel(uint32_t *__restrict__ p, int row_words, int capacity, const int *__restrict__ count) {
  const size_t total = (size_t)min(capacity, *count) * row_words;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) p[k] = 0u;
}

// A uniformly random valid gym action per env: the k-th set byte of the env's N*5 mask bytes, one warp per env
// (the draw is policy_draw(seed, global env, 0, player) mod the number of set bytes).  The row is read as aligned
// 128-bit vectors (512 contiguous bytes per warp instruction, whatever the row's own alignment), each lane turning its
// 16 bytes into a 16-bit mask; the k-th set bit is then found with one warp prefix sum per 512-byte round.
// (One byte per lane per ballot took 72-90 us per 65,536 envs: 32 bytes per load instruction.)
__device__ __forceinline__ uint32_t nonzero_bytes4(uint32_t x) {  // bit i: byte i of x is non-zero
  const uint32_t z = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
  return (((z >> 7) * 0x01020408u) >> 24) & 0xfu;
}

__global__ void __launch_bounds__(256) grl_gym_sample_kernel(const GrlKParams prm, unsigned long long seed,
                                                             const uint8_t *__restrict__ mask, int player,
                                                             long long *__restrict__ action) {
  constexpr int kMaxRounds = (GRL_MAX_DIM * GRL_MAX_DIM * 5 + 15 + 511) / 512 + 1;  // 512-byte rounds of one row
  __shared__ uint16_t s_m[8][kMaxRounds * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int M = prm.N * 5;
  uint16_t *sm = s_m[warp];
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < prm.B; b += (gridDim.x * blockDim.x) >> 5) {
    const uint8_t *row = mask + ((size_t)b * prm.P + player) * M;
    const int mis = (int)(reinterpret_cast<uintptr_t>(row) & 15u);
    const uint4 *base = reinterpret_cast<const uint4 *>(row - mis);
    const int nvec = (mis + M + 15) >> 4, rounds = (nvec + 31) >> 5;
    const bool last_row = b == prm.B - 1 && player == prm.P - 1;  // its last vector may reach past the plane,
    const bool first_row = b == 0 && player == 0 && mis != 0;      // the first row's first vector before it
    int mine = 0;
    for (int r = 0; r < rounds; r++) {
      const int v = 32 * r + lane;
      uint32_t m16 = 0;
      if (v < nvec) {
        const int g0 = 16 * v - mis;  // row-relative index of this vector's first byte
        if ((last_row && v == nvec - 1) || (first_row && v == 0)) {
          for (int j = 0; j < 16; j++)
            if (g0 + j >= 0 && g0 + j < M && row[g0 + j] != 0) m16 |= 1u << j;
        } else {
          const uint4 q = __ldg(base + v);
          m16 = nonzero_bytes4(q.x) | (nonzero_bytes4(q.y) << 4) | (nonzero_bytes4(q.z) << 8) | (nonzero_bytes4(q.w) << 12);
          if (g0 < 0) m16 &= 0xffffu << (-g0);                 // bytes before the row
          if (g0 + 16 > M) m16 &= 0xffffu >> (g0 + 16 - M);    // bytes past its end
        }
      }
      sm[v] = (uint16_t)m16;
      mine += __popc(m16);
    }
    const int total = __reduce_add_sync(FULL, mine);
    __syncwarp();
    long long pick = 0;
    if (total > 0) {
      const uint64_t rr = policy_draw(seed, (uint64_t)(prm.env_id_base + b), 0ull, (uint64_t)player);
      int k = (int)(rr % (uint64_t)total);
      for (int r = 0; r < rounds; r++) {
        const uint32_t w = sm[32 * r + lane];
        const int c = __popc(w);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        const int stripe = __shfl_sync(FULL, incl, 31);
        if (k < stripe) {
          const uint32_t who = __ballot_sync(FULL, k >= incl - c && k < incl);
          const int src = __ffs(who) - 1;
          uint32_t ww = __shfl_sync(FULL, w, src);
          const int kk = k - (__shfl_sync(FULL, incl, src) - __popc(ww));
          for (int j = 0; j < kk; j++) ww &= ww - 1u;  // drop the kk lowest set bits
          pick = 512 * r + 16 * src + (__ffs(ww) - 1) - mis;
          break;
        }
        k -= stripe;
      }
    }
    if (lane == 0) action[b] = pick;
    __syncwarp();
  }
}

// packed engine mask with the half-move replica: [B][P][rep][words]
__global__ void grl_mask_replicate_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t rows, int words,
                                          int rep) {
  const size_t total = rows * (size_t)words * rep;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t row = idx / ((size_t)words * rep);
    int k = (int)(idx % words);
    out[idx] = in[row * words + k];
  }
}

// 64-bit digest of the full game state; identical definition in oracle/grl_oracle.c
__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32) grl_state_hash_kernel(const GrlKParams prm, uint64_t *__restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GrlLayout &L = prm.L;
  for (int game = blockIdx.x * GRL_WARPS_PER_CTA + warp; game < prm.B; game += gridDim.x * GRL_WARPS_PER_CTA) {
    const uint32_t *s = prm.state + (size_t)game * L.slab_words;
    const uint32_t *stt = prm.statics + (size_t)game * L.static_words;
    const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
    uint64_t h = 0;
    for (int t = lane; t < prm.N; t += 32) {
      int owner = -1;
      uint64_t lists = 0, vis = 0;
      for (int p = 0; p < prm.P; p++) {
        if (bit_of(s + L.off_own + p * prm.NW, t)) owner = p;
        if (bit_of(s + L.off_list + p * prm.NW, t)) lists |= 1ULL << p;
        if (bit_of(s + L.off_vis + p * prm.NW, t)) vis |= 1ULL << p;
      }
      uint64_t type = bit_of(stt, t) ? 3 : (bit_of(stt + prm.NW, t) ? 2 : (bit_of(stt + 2 * prm.NW, t) ? 1 : 0));
      uint64_t pack = (uint64_t)(owner + 1) | (type << 4) | ((uint64_t)bit_of(s + L.off_changed, t) << 6) |
                      ((uint64_t)bit_of(s + L.off_vchg, t) << 7) | (vis << 8) | (lists << 16) | ((uint64_t)army[t] << 24);
      h += mix64(pack ^ ((uint64_t)(t + 1) * 0xD6E8FEB86659FD93ULL));
    }
    h = warp_sum64(h);
    if (lane == 0) {
      uint32_t flags = s[GRL_HDR_FLAGS];
      h += mix64(0x1000000000ULL + (uint64_t)s[GRL_HDR_TURN]);
      h += mix64(0x2000000000ULL + (uint64_t)((flags >> 8) & 1u) + ((uint64_t)(flags & 0xffu) << 8));
      for (int p = 0; p < prm.P; p++)
        h += mix64(0x3000000000ULL + ((uint64_t)p << 40) +
                   (uint64_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ARMY_COUNT]);
      out[game] = h;
    }
  }
}

__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32)
    grl_buffer_hash_kernel(const uint32_t *__restrict__ buf, size_t row_words, int rows, uint64_t *__restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * GRL_WARPS_PER_CTA + warp; r < rows; r += gridDim.x * GRL_WARPS_PER_CTA) {
    const uint32_t *w = buf + (size_t)r * row_words;
    uint64_t h = 0;
    for (size_t i = lane; i < row_words; i += 32) h += mix64((uint64_t)w[i] ^ ((uint64_t)(i + 1) * 0xD6E8FEB86659FD93ULL));
    h = warp_sum64(h);
    if (lane == 0) out[r] = h;
  }
}

// synthetic policy as a stand-alone kernel: fills grl_action[B][A]
template <int PT>
__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32) grl_sample_kernel(const __grid_constant__ GrlKParams prm, uint2 *__restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GrlLayout &L = prm.L;
  const int P = prm.P, NW = prm.NW, N = prm.N;
  const Geo g = make_geo(prm, prm.W, lane, 32);
  for (int game = blockIdx.x * GRL_WARPS_PER_CTA + warp; game < prm.B; game += gridDim.x * GRL_WARPS_PER_CTA) {
    const uint32_t *s = prm.state + (size_t)game * L.slab_words;
    const uint32_t *stt = prm.statics + (size_t)game * L.static_words;
    const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
    const bool act = lane < NW;
    const uint32_t M = act ? stt[lane] : 0u;
    const uint32_t flags = s[GRL_HDR_FLAGS];
    const uint32_t turn = s[GRL_HDR_TURN];
    uint32_t gt1 = army_gt1_mask<32>(army, NW, N, g);
    DirMasks dm = dir_targets<32>(M, g);
    if (lane < prm.A) out[(size_t)game * prm.A + lane] = make_uint2(0u, 0u);
    __syncwarp();
    if (flags & GRL_FLAG_OVER) continue;
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P && p < prm.A) {
        uint32_t own = act ? s[L.off_own + p * NW + lane] : 0u;
        uint32_t lst = act ? s[L.off_list + p * NW + lane] : 0u;
        uint32_t src = ((flags >> p) & 1u) ? (lst & own & gt1) : 0u;
        PackedAction a = sample_policy_action<32>(prm, prm.policy_seed, dm, src, p, (uint64_t)(prm.env_id_base + game), turn, g);
        if (lane == 0) out[(size_t)game * prm.A + p] = make_uint2(a.lo, a.hi);
      }
    }
  }
}

// lifetime counters: sum header words 2..5 over all envs
__global__ void grl_stats_kernel(const GrlKParams prm, unsigned long long *__restrict__ out) {
  unsigned long long acc[4] = {0, 0, 0, 0};
  for (int game = blockIdx.x * blockDim.x + threadIdx.x; game < prm.B; game += gridDim.x * blockDim.x) {
    const uint32_t *s = prm.state + (size_t)game * prm.L.slab_words;
    for (int k = 0; k < 4; k++) acc[k] += s[GRL_HDR_STEPS + k];
  }
  for (int k = 0; k < 4; k++) {
    unsigned long long v = warp_sum64(acc[k]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + k, v);
  }
}

// ---------------------------------------------------------------------------------------
// launchers (called from grl_abi.cu)
// ---------------------------------------------------------------------------------------
static inline int grid_for(int items_per_cta_warps, int n) {
  int ctas = (n + items_per_cta_warps - 1) / items_per_cta_warps;
  return ctas < 1 ? 1 : ctas;
}

static size_t grl_turn_smem_bytes(const GrlLayout &L, int TW, int TH, int PT, int LG, bool gym) {
  const bool snap = GRL_DIRTY_WB && (LG == 32 || GRL_PACKED_SNAPSHOT);
  const int per_game = (snap ? 2 : 1) * L.slab_words + L.static_words + 2 * GRL_MAX_ACTIONS;
  const int scratch = gym ? grl_gym_smem_words(L.P, L.NW, L.N, grl_gym_emit_mode(TW, TH)) : grl_obs_scratch_words(TW, TH, PT, L.NW);
  return (size_t)GRL_WARPS_PER_CTA * (size_t)((32 / LG) * per_game + scratch) * 4u;
}

template <int PT, int TW, int TH, int LG, bool S, bool O, bool GYM = false>
static cudaError_t launch_turn_t(const GrlKParams &prm, cudaStream_t stream, const GrlGymK *gym = nullptr) {
  size_t smem = grl_turn_smem_bytes(prm.L, TW, TH, PT, LG, GYM);
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
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid < 1 ? 1 : grid);
  cfg.blockDim = dim3(GRL_WARPS_PER_CTA * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (prm.l2_window_bytes) {  // game state persists in L2; everything else streams through it
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = prm.state;
    attr[0].val.accessPolicyWindow.num_bytes = prm.l2_window_bytes;
    attr[0].val.accessPolicyWindow.hitRatio = prm.l2_hit_ratio;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, prm, gk);
}

template <int PT, int TW, int TH, int LG>
static cudaError_t launch_turn_g(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  if (do_step && do_out) return launch_turn_t<PT, TW, TH, LG, true, true>(prm, stream);
  if (do_step) return launch_turn_t<PT, TW, TH, LG, true, false>(prm, stream);
  return launch_turn_t<PT, TW, TH, LG, false, true>(prm, stream);
}

// the BASELINE board sizes get kernels with the geometry baked in and the lane group sized to the
// board (10x10: 4 words -> 8 games per warp; 15x15: 8 words -> 4 games per warp); everything else is
// generic with one game per warp.  GRL_LANES_PER_GAME=32 (environment) forces one game per warp.
template <int PT>
static cudaError_t launch_turn_p(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  const int lpg = prm.lanes_per_game;
  if constexpr (PT <= 4) {
    if (prm.W == 20 && prm.H == 20) {
      if (lpg == 16) return launch_turn_g<PT, 20, 20, 16>(prm, do_step, do_out, stream);
      return launch_turn_g<PT, 20, 20, 32>(prm, do_step, do_out, stream);
    }
    // measured on B200 (profiles/r1_variants.md): with the packed groups writing the whole slab back (no snapshot in
    // shared memory, four CTAs per SM) 8 lanes per game is fastest for 10x10 and 15x15 at every batch size
    if (prm.W == 15 && prm.H == 15) {
      const int pick = lpg ? lpg : 8;
      if (pick == 8) return launch_turn_g<PT, 15, 15, 8>(prm, do_step, do_out, stream);
      if (pick == 16) return launch_turn_g<PT, 15, 15, 16>(prm, do_step, do_out, stream);
      return launch_turn_g<PT, 15, 15, 32>(prm, do_step, do_out, stream);
    }
    if (prm.W == 10 && prm.H == 10) {
      const int pick = lpg ? lpg : 8;
      if (pick == 4) return launch_turn_g<PT, 10, 10, 4>(prm, do_step, do_out, stream);
      if (pick == 8) return launch_turn_g<PT, 10, 10, 8>(prm, do_step, do_out, stream);
      return launch_turn_g<PT, 10, 10, 32>(prm, do_step, do_out, stream);
    }
  }
  return launch_turn_g<PT, 0, 0, 32>(prm, do_step, do_out, stream);
}

static int player_template(int P) { return P <= 2 ? 2 : (P <= 4 ? 4 : 8); }

// The fused gym step (one GeneralsEnv.step() per env in ONE launch): the default lane group of each baked board.
template <int PT>
static cudaError_t launch_gym_step_p(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  if constexpr (PT <= 4) {
    if (prm.W == 20 && prm.H == 20) return launch_turn_t<PT, 20, 20, 32, true, true, true>(prm, stream, &gk);
    if (prm.W == 15 && prm.H == 15) return launch_turn_t<PT, 15, 15, 8, true, true, true>(prm, stream, &gk);
    if (prm.W == 10 && prm.H == 10) return launch_turn_t<PT, 10, 10, 8, true, true, true>(prm, stream, &gk);
  }
  return launch_turn_t<PT, 0, 0, 32, true, true, true>(prm, stream, &gk);
}

cudaError_t grl_launch_gym_step(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  switch (player_template(prm.P)) {
    case 2: return launch_gym_step_p<2>(prm, gk, stream);
    case 4: return launch_gym_step_p<4>(prm, gk, stream);
    default: return launch_gym_step_p<8>(prm, gk, stream);
  }
}

cudaError_t grl_launch_turn(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  switch (player_template(prm.P)) {
    case 2: return launch_turn_p<2>(prm, do_step, do_out, stream);
    case 4: return launch_turn_p<4>(prm, do_step, do_out, stream);
    default: return launch_turn_p<8>(prm, do_step, do_out, stream);
  }
}

cudaError_t grl_launch_reset(const GrlKParams &prm, const uint32_t *src_state, const uint32_t *src_static,
                             const int32_t *env_ids, int n, cudaStream_t stream, const int *n_dev) {
  int grid = grid_for(GRL_WARPS_PER_CTA, n);
  if (n_dev && grid > 148 * 16) grid = 148 * 16;  // sized for the capacity: the kernel strides
  switch (player_template(prm.P)) {
    case 2: grl_reset_kernel<2><<<grid, GRL_WARPS_PER_CTA * 32, 0, stream>>>(prm, src_state, src_static, env_ids, n, n_dev); break;
    case 4: grl_reset_kernel<4><<<grid, GRL_WARPS_PER_CTA * 32, 0, stream>>>(prm, src_state, src_static, env_ids, n, n_dev); break;
    default: grl_reset_kernel<8><<<grid, GRL_WARPS_PER_CTA * 32, 0, stream>>>(prm, src_state, src_static, env_ids, n, n_dev); break;
  }
  return cudaGetLastError();
}

cudaError_t grl_launch_sample(const GrlKParams &prm, void *out, cudaStream_t stream) {
  int grid = grid_for(GRL_WARPS_PER_CTA, prm.B);
  switch (player_template(prm.P)) {
    case 2: grl_sample_kernel<2><<<grid, GRL_WARPS_PER_CTA * 32, 0, stream>>>(prm, (uint2 *)out); break;
    case 4: grl_sample_kernel<4><<<grid, GRL_WARPS_PER_CTA * 32, 0, stream>>>(prm, (uint2 *)out); break;
    default: grl_sample_kernel<8><<<grid, GRL_WARPS_PER_CTA * 32, 0, stream>>>(prm, (uint2 *)out); break;
  }
  return cudaGetLastError();
}

static int flat_grid(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  return g < 1 ? 1 : (int)g;
}

cudaError_t grl_launch_mask_bytes(const GrlKParams &prm, int variant, uint8_t *out, cudaStream_t stream) {
  size_t total = (size_t)prm.B * prm.P * prm.N;
  grl_mask_bytes_kernel<<<flat_grid(total, 256), 256, 0, stream>>>(prm, variant, out);
  return cudaGetLastError();
}

cudaError_t grl_launch_visibility(const GrlKParams &prm, uint8_t *visible, uint8_t *fog, cudaStream_t stream) {
  size_t total = (size_t)prm.B * prm.P * prm.N;
  grl_visibility_kernel<<<flat_grid(total, 256), 256, 0, stream>>>(prm, visible, fog);
  return cudaGetLastError();
}

cudaError_t grl_launch_gym(const GrlKParams &prm, int max_turns, const float *logtab, float *obs, uint8_t *mask, int32_t *stats,
                           cudaStream_t stream, const int32_t *ids, int n_ids, const int *n_dev) {
  // warp-per-game kernel with linear 128-bit sweeps; GRL_GYM_FLAT=1 keeps the thread-per-tile version for comparison
  static const bool flat = [] { const char *e = getenv("GRL_GYM_FLAT"); return e && e[0] == '1'; }();
  if (!flat || ids) {
    const size_t smem = (size_t)GRL_WARPS_PER_CTA * grl_gym_smem_words(prm.P, prm.NW, prm.N, GRL_GYM_EMIT_GENERIC) * 4u;
    static size_t tuned = 0;
    if (smem > 48 * 1024 && smem > tuned) {
      cudaError_t e = cudaFuncSetAttribute(grl_gym_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      tuned = smem;
    }
    int grid = grid_for(GRL_WARPS_PER_CTA, ids ? n_ids : prm.B);
    if (grid > 148 * 16) grid = 148 * 16;
    grl_gym_warp_kernel<<<grid, GRL_WARPS_PER_CTA * 32, smem, stream>>>(prm, max_turns, logtab, obs, mask, stats, ids, n_ids, n_dev);
    return cudaGetLastError();
  }
  if (ids) return cudaErrorNotSupported;  // the flat comparison kernel has no id list
  size_t total = (size_t)prm.B * prm.P * prm.N;
  grl_gym_kernel<<<flat_grid(total, 256), 256, 0, stream>>>(prm, max_turns, logtab, obs, mask, stats);
  return cudaGetLastError();
}

cudaError_t grl_launch_gym_encode(const GrlKParams &prm, const long long *action_idx, int player, int slot, const uint8_t *mask,
                                  int skip_invalid, void *actions, uint8_t *valid, cudaStream_t stream) {
  grl_gym_encode_kernel<<<flat_grid((size_t)prm.B, 256), 256, 0, stream>>>(prm, action_idx, player, slot, mask, skip_invalid,
                                                                           (uint2 *)actions, valid);
  return cudaGetLastError();
}

cudaError_t grl_launch_gym_compact(const GrlKParams &prm, const uint8_t *terminated, const uint8_t *truncated, long long base_seed,
                                   long long *episode, int32_t *turns, int32_t *calls, int32_t *ids, long long *seeds, int *count,
                                   cudaStream_t stream) {
  grl_gym_compact_kernel<<<flat_grid((size_t)prm.B, 256), 256, 0, stream>>>(prm, terminated, truncated, base_seed, episode, turns, calls,
                                                                            ids, seeds, count);
  return cudaGetLastError();
}

cudaError_t grl_launch_gym_final_obs(const GrlKParams &prm, const float *obs, float *final_obs, const int32_t *ids, const int *count,
                                     cudaStream_t stream) {
  grl_gym_final_obs_kernel<<<148 * 4, 256, 0, stream>>>(prm, obs, final_obs, ids, count);
  return cudaGetLastError();
}

cudaError_t grl_launch_zero_rows(uint32_t *p, int row_words, int capacity, const int *count, cudaStream_t stream) {
  grl_zero_rows_kernel<<<148 * 8, 256, 0, stream>>>(p, row_words, capacity, count);
  return cudaGetLastError();
}

cudaError_t grl_launch_gym_sample(const GrlKParams &prm, unsigned long long seed, const uint8_t *mask, int player, long long *action,
                                  cudaStream_t stream) {
  int grid = grid_for(8, prm.B);
  if (grid > 148 * 16) grid = 148 * 16;
  grl_gym_sample_kernel<<<grid, 256, 0, stream>>>(prm, seed, mask, player, action);
  return cudaGetLastError();
}

cudaError_t grl_launch_mask_replicate(const uint32_t *in, uint32_t *out, size_t rows, int words, int rep, cudaStream_t stream) {
  grl_mask_replicate_kernel<<<flat_grid(rows * words * rep, 256), 256, 0, stream>>>(in, out, rows, words, rep);
  return cudaGetLastError();
}

cudaError_t grl_launch_state_hash(const GrlKParams &prm, uint64_t *out, cudaStream_t stream) {
  grl_state_hash_kernel<<<grid_for(GRL_WARPS_PER_CTA, prm.B), GRL_WARPS_PER_CTA * 32, 0, stream>>>(prm, out);
  return cudaGetLastError();
}

cudaError_t grl_launch_buffer_hash(const uint32_t *buf, size_t row_words, int rows, uint64_t *out, cudaStream_t stream) {
  grl_buffer_hash_kernel<<<grid_for(GRL_WARPS_PER_CTA, rows), GRL_WARPS_PER_CTA * 32, 0, stream>>>(buf, row_words, rows, out);
  return cudaGetLastError();
}

cudaError_t grl_launch_stats(const GrlKParams &prm, unsigned long long *out, cudaStream_t stream) {
  grl_stats_kernel<<<flat_grid((size_t)prm.B, 256), 256, 0, stream>>>(prm, out);
  return cudaGetLastError();
}

// envs that were created but never reset reject steps like a finished game
__global__ void grl_mark_over_kernel(const GrlKParams prm) {
  for (int game = blockIdx.x * blockDim.x + threadIdx.x; game < prm.B; game += gridDim.x * blockDim.x) {
    uint32_t *s = prm.state + (size_t)game * prm.L.slab_words;
    s[GRL_HDR_FLAGS] = GRL_FLAG_OVER;
    for (int p = 0; p < prm.P; p++) {
      s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_GENERAL_IDX] = 0xffffffffu;
      s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ACTION_INDEX] = 0xffffffffu;
    }
  }
}

cudaError_t grl_launch_mark_over(const GrlKParams &prm, cudaStream_t stream) {
  grl_mark_over_kernel<<<flat_grid((size_t)prm.B, 256), 256, 0, stream>>>(prm);
  return cudaGetLastError();
}
