// grl_device.cuh — device-side building blocks of the turn engine (included by every kernel translation unit).
//
// A GROUP of LG lanes owns one game (LG = 32 or 8: the smallest power of two >= the NW words a
// board's bit planes span that measured fastest), so a warp steps 32/LG games at once.  Every
// boolean plane of a game (ownership per player, the reference's cached OwnedTiles lists,
// visibility per player, the changed / visibility-changed tile sets, terrain) is an N-bit LINEAR
// bitmask, N = W*H <= 1024, held as one 32-bit word per lane.  Stencils (3x3 fog dilation, the 5x5
// "affected players" probe, the four move directions) are funnel shifts across neighbouring lanes'
// words; set sizes are popc + REDUX.  Only the armies are a per-tile plane (uint16, in shared memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/grlcuda.h"
#include "grl_layout.h"

#define FULL 0xffffffffu
#define GRL_WARPS_PER_CTA 8  // 4 measured equal or slower on every board (profiles/r2_variants.md)

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
// bulk copy shared -> global with the L2 evict-first policy: like st.global.cs, the lines must not displace the state slabs
// that cp.async.bulk.prefetch.L2 parked there (the default policy measured 4 % slower at 20x20)
__device__ __forceinline__ void tma_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(smem_addr(src_smem)), "r"(bytes), "l"(pol)
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
template <int LG>
__device__ __forceinline__ uint32_t army_gt1_mask(const uint16_t *army, int NW, int N, const Geo &g) {
  // one game per warp keeps the ballot version: 13 ballots for 20x20 cost about the same as 13 active lanes doing the
  // vector version, and measured 1.4 % faster there (0.3423 vs 0.347 ms per 65,536 games)
  if (LG == 32) {
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

// ---- bit streams in shared memory (the linear read-out writers of grl_obs.cuh / grl_gym.cuh) ---------------------------
__device__ __forceinline__ uint32_t lds32(uint32_t sa) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t sa) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa));
  return v;
}
__device__ __forceinline__ uint32_t rotl4(uint32_t v) { return __funnelshift_l(v, v, 4); }
// a value the compiler must keep in a register instead of recomputing it at every use (shared-window addresses are
// otherwise rematerialised inside the store loops: S2R + MOV + VIADD + LEA per access)
__device__ __forceinline__ uint32_t pinned_reg(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}

// The linear sweep of the bit-stream writers.  Lane l writes float4 k = k0 + l + 32 r of the run in round r: the 4-bit
// field of its four floats sits at the same offset `sh` of a stream word in every round (words are stored rotated left by
// 4, so rotating right by sh leaves 16 * field at bits 4-7: the byte offset of the field's entry in the 256-byte aligned
// nibble -> float4 table).  Rounds are issued four at a time whenever all four are whole: the loads of a group are
// independent, which hides the shared-memory latency a single round would expose.
struct StreamSweep {
  uint32_t lut_sa, sp;  // shared-window addresses: the table, this lane's stream word of the current round
  int sh, k, k_end;     // field offset; this lane's float4 index in the current round; one past the last float4
  float4 *op;           // this lane's float4 of the current round
  __device__ __forceinline__ void init(const float4 *lut, const uint32_t *strm, float *base_al, int k_first, int k_end_, int lane) {
    lut_sa = pinned_reg(smem_addr(lut));
    sh = 4 * ((k_first + lane) & 7);
    k = k_first + lane;
    k_end = k_end_;
    sp = smem_addr(strm) + 4u * (uint32_t)(k >> 3);
    op = reinterpret_cast<float4 *>(base_al) + k;
  }
  __device__ __forceinline__ float4 entry(uint32_t w) const { return lds128(lut_sa | (__funnelshift_r(w, w, sh) & 0xf0u)); }
  __device__ __forceinline__ int round_base(int lane) const { return k - lane; }  // first float4 of the current round
  __device__ __forceinline__ void advance(int rounds) { k += 32 * rounds, sp += 16u * rounds, op += 32 * rounds; }
  // `rounds` rounds of plain 0/1 floats
  __device__ __forceinline__ void plain(int rounds, int lane) {
    int r = 0;
    for (; r + 4 <= rounds && round_base(lane) + 127 < k_end; r += 4) {
      const uint32_t w0 = lds32(sp), w1 = lds32(sp + 16), w2 = lds32(sp + 32), w3 = lds32(sp + 48);
      const float4 v0 = entry(w0), v1 = entry(w1), v2 = entry(w2), v3 = entry(w3);
      __stcs(op, v0);
      __stcs(op + 32, v1);
      __stcs(op + 64, v2);
      __stcs(op + 96, v3);
      advance(4);
    }
    for (; r < rounds; r++) {
      if (k < k_end) __stcs(op, entry(lds32(sp)));
      advance(1);
    }
  }
  // the rounds that overlap floats [qa, qz] of the run: their float4s take multipliers from F (F4[0] = float4 qa >> 2)
  __device__ __forceinline__ void scaled(int qa, int qz, uint32_t F_sa, int lane) {
    const int ka = qa >> 2, kz = qz >> 2;
    const int rounds = ((kz - round_base(lane)) >> 5) + 1;
    for (int r = 0; r < rounds; r++) {
      if (k < k_end) {
        float4 val = entry(lds32(sp));
        if (k >= ka && k <= kz) {
          const float4 m = lds128(F_sa + 16u * (uint32_t)(k - ka));
          val.x *= m.x;
          val.y *= m.y;
          val.z *= m.z;
          val.w *= m.w;
        }
        __stcs(op, val);
      }
      advance(1);
    }
  }
  __device__ __forceinline__ int rounds_before(int q, int lane) const {  // whole rounds before the one that holds float q
    const int kb0 = round_base(lane), ka = q >> 2;
    return ka > kb0 ? (ka - kb0) >> 5 : 0;
  }
  __device__ __forceinline__ void finish(int lane) {
    const int kb0 = round_base(lane);
    plain(kb0 < k_end ? (k_end - kb0 + 31) >> 5 : 0, lane);
  }
};

// OR one N-bit mask (word `lane` of it in `word`, lanes 0..NWC; lane NWC holds 0) into the stream at bit offset qs.
// Called by lanes 0..NWC together; masks of one game are disjoint in the stream, words they share are OR-ed atomically.
// ROT: store the words rotated left by 4 (the float writers' table lookup wants a 4-bit field at bits 4-7).
template <int NWC, bool ROT>
__device__ __forceinline__ void stream_or_mask(uint32_t *strm, uint32_t word, int qs, int lane) {
  constexpr uint32_t kLanes = (2u << NWC) - 1u;  // lanes 0..NWC
  uint32_t prev = __shfl_up_sync(kLanes, word, 1);
  if (lane == 0) prev = 0u;
  const uint32_t c = __funnelshift_l(prev, word, qs & 31);  // (word << sh) | (prev >> (32 - sh))
  atomicOr(strm + (qs >> 5) + lane, ROT ? rotl4(c) : c);
}

// ---- the compile-time-scheduled sweep ----------------------------------------------------------------------------------
// When a warp holds GPW whole games of a baked board and every player template slot is in use (P == PT), the games'
// [P][CH][N] blocks form ONE run of GPW * TOTAL floats that starts 16-byte aligned (GPW * TOTAL * 4 bytes is a multiple
// of 16 for GPW = 4).  The run is written in ROUNDS of 32 float4s (128 floats, one STG.128 per lane).  Game gi writes the
// rounds that END inside its block; the floats of its last, incomplete round are carried (as stream bits, in registers)
// into the front of the next game's stream, whose first round completes them: every store of the run except the very last
// covers 512 contiguous bytes.  A game's pass is straight-line code: the round index of every load and store is a
// compile-time constant (LDS / STG.128 with immediate offsets from one base each, no loop control, no address arithmetic);
// what depends on the game — the number `pre` of carried floats in front of its block (0..127), hence where in the stream
// a scaled plane begins and ends — enters as a handful of registers, and the rounds a scaled plane CAN touch for any `pre`
// carry a per-lane range test.  One copy of the pass serves the four games of a warp: four specialised copies (every
// range a constant) executed 13-16 % fewer instructions and ran 3-4 % slower — the kernels wait on instruction fetch
// (`no_instruction` 1.9 -> 4.3 stall cycles per issue with the four copies).
template <int TOTAL, int GPW>
struct CtRun {
  static_assert((GPW * TOTAL) % 4 == 0, "the run of a warp's games is a whole number of float4s");
  static constexpr int NR_MIN = TOTAL / 128;    // every game's pass writes at least this many rounds,
  static constexpr int NR_MAX = NR_MIN + 2;     // at most this many (a carried round in front, the run's last round)
  __host__ __device__ static constexpr int q0(int gi) { return TOTAL * gi; }                     // run float of game gi's first float
  __host__ __device__ static constexpr int r_lo(int gi) { return q0(gi) / 128; }                 // first round game gi's pass writes
  __host__ __device__ static constexpr int r_hi(int gi) { return gi + 1 < GPW ? q0(gi + 1) / 128 : (q0(GPW) + 127) / 128; }
  __host__ __device__ static constexpr int rounds(int gi) { return r_hi(gi) - r_lo(gi); }
  __host__ __device__ static constexpr int pre(int gi) { return q0(gi) - 128 * r_lo(gi); }       // carried floats in front of the block
  __host__ __device__ static constexpr int last_active() { return GPW * TOTAL / 4 - 32 * (r_hi(GPW - 1) - 1); }  // lanes of the run's last round
  __host__ __device__ static constexpr int stream_words() { return 4 * NR_MAX; }
  // rounds [win_lo(c), win_hi(c, len)] are the ones a plane range [pre + c, pre + c + len) can touch for some pre in 0..127
  __host__ __device__ static constexpr int win_lo(int c) { return c / 128; }
  __host__ __device__ static constexpr int win_hi(int c, int len) { return (c + 127 + len - 1) / 128; }
};

struct CtLane {     // per-lane constants of a run
  uint32_t sp;      // shared address of the stream word this lane reads in round 0 (round j: + 16 j)
  uint32_t rot;     // rotate-right amount that leaves the lane's 4-bit field at bits 0-3
  uint32_t F;       // shared address of the multiplier array + 16 * lane
  float4 *op;       // the run's first float4 + lane
  __device__ __forceinline__ void init(const uint32_t *strm, const float *Fp, float *run_base, int lane) {
    sp = pinned_reg(smem_addr(strm) + 4u * (uint32_t)(lane >> 3));
    rot = 4u * (uint32_t)(lane & 7);
    F = pinned_reg(smem_addr(Fp) + 16u * (uint32_t)lane);
    op = reinterpret_cast<float4 *>(run_base) + lane;
  }
};

#define CT_PLAIN 0   // 0/1 floats
#define CT_F 1       // float4s ka..kz of the stream take multipliers from F (F4[0] = float4 ka)
#define CT_SCALAR 2  // stream floats qt .. qt + n - 1 are multiplied by tf
#define CT_ALL_ROUNDS (1 << 20)
// Rounds [j0, j1) of a game's pass; `op` is the lane's float4 of the pass's round 0.  j0 and j1 are constants at the call
// site, so after inlining and unrolling each round is LDS, SHF, four selects and STG.128 with immediate offsets; four
// rounds are loaded before the first is stored, so the shared-memory latencies overlap.  ka, kz, Fb (= c.F - 16 ka), qt
// are the game's registers; `nr` = rounds the pass has (CT_ALL_ROUNDS when [j0, j1) certainly exist), `last_lanes` = lanes
// of round nr - 1 that store.
template <int MODE>
__device__ __forceinline__ void ct_rounds(const CtLane &c, float4 *op, int lane, int j0, int j1, int ka, int kz, uint32_t Fb,
                                          int qt, int n, float tf, int nr, int last_lanes) {
#pragma unroll
  for (int jb = j0; jb < j1; jb += 4) {
    uint32_t w[4];
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (jb + i < j1) w[i] = lds32(c.sp + 16u * (uint32_t)(jb + i));
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (jb + i < j1) {
        // four selects instead of the table look-up of StreamSweep: these kernels wait on the shared-memory pipe, not on
        // issue slots (15x15: main -0.5 %, gym step -1.5 %; the quad writers of 10x10 / 20x20 measured 1-2 % slower this way)
        const uint32_t nib = __funnelshift_r(w[i], w[i], c.rot);  // the lane's 4-bit field at bits 0-3
        v[i] = make_float4((nib & 1u) ? 1.f : 0.f, (nib & 2u) ? 1.f : 0.f, (nib & 4u) ? 1.f : 0.f, (nib & 8u) ? 1.f : 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (jb + i < j1) {
        const int j = jb + i;
        if (MODE == CT_F) {
          if ((uint32_t)(32 * j + lane - ka) <= (uint32_t)(kz - ka)) {  // this lane's float4 is one of ka..kz
            const float4 m = lds128(Fb + (uint32_t)(512 * j));
            v[i].x *= m.x;
            v[i].y *= m.y;
            v[i].z *= m.z;
            v[i].w *= m.w;
          }
        } else if (MODE == CT_SCALAR) {
          const uint32_t d = (uint32_t)(128 * j + 4 * lane - qt);  // float d + i of the range, if below n
          if (d + 0u < (uint32_t)n) v[i].x *= tf;
          if (d + 1u < (uint32_t)n) v[i].y *= tf;
          if (d + 2u < (uint32_t)n) v[i].z *= tf;
          if (d + 3u < (uint32_t)n) v[i].w *= tf;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (jb + i < j1) {
        const int j = jb + i;
        if (nr == CT_ALL_ROUNDS) {
          __stcs(op + 32 * j, v[i]);
        } else if (j < nr) {
          if (j < nr - 1 || lane < last_lanes) __stcs(op + 32 * j, v[i]);
        }
      }
    }
  }
}
