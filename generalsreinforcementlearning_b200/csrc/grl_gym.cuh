// grl_gym.cuh — the generals_gym client's read-outs (python/generals_gym/generals_env.py:291-387 over the proto view of
// internal/grpc/gameserver/server.go:556-582), written by a whole warp per game.
#pragma once
#include "grl_device.cuh"

#define GRL_GYM_EMIT_GENERIC 0
#define GRL_GYM_EMIT_QUADS 1
#define GRL_GYM_EMIT_LINEAR 2
// which read-out writer a geometry uses: baked boards with N % 4 == 0 -> quads, other baked boards -> linear
__host__ __device__ constexpr int grl_gym_emit_mode(int TW, int TH) {
  return TW > 0 ? (((TW * TH) & 3) == 0 ? GRL_GYM_EMIT_QUADS : GRL_GYM_EMIT_LINEAR) : GRL_GYM_EMIT_GENERIC;
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
  } else if (mode == GRL_GYM_EMIT_LINEAR) {  // gym_emit_linear: concatenated direction streams + (F + bit stream | staged mask words)
    const int dirs = 4 * (((PT * N + 31) / 32 + 1 + 3) & ~3);
    const int lin_obs = ((2 * N + 8 + 3) & ~3) + ((((PT * GRL_GYM_CHANNELS * N + 127 + 31) / 32 + 1) + 3) & ~3);
    const int lin_mask = 5 * ((PT * N + 3) / 4) + 12;
    m = dirs + (lin_obs > lin_mask ? lin_obs : lin_mask);
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
// count in (the element -> (plane, tile) divisions become multiplications).  CTA: the whole thread block works on ONE game
// (every thread calls; warp 0 stages the masks, all threads share the sweeps; `sw` is then the block's scratch) — the
// version for a handful of games whose read-outs are on somebody's critical path (grl_gym_autoreset).
template <int NT, bool CTA = false>
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
  const int tid = CTA ? (int)threadIdx.x : lane, nthr = CTA ? (int)blockDim.x : 32;
  {
    const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
    if (!CTA || threadIdx.x < 32) {
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
    }
    for (int t = tid; t < N + 4; t += nthr) s_logv[t] = t < N ? logtab[army[t]] : 0.f;  // logtab[0] == 0
    if (CTA) __syncthreads(); else __syncwarp();

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
        if (tid < head) __stcs(base + tid, gym_value(gp, p, tid / N, tid % N));
        if (tid < block - tail0) __stcs(base + tail0 + tid, gym_value(gp, p, (tail0 + tid) / N, (tail0 + tid) % N));
        float4 *body = reinterpret_cast<float4 *>(base + head);
        for (int i = tid; i < body4; i += nthr) {
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
        if (tid < head) base[tid] = (uint8_t)flag(tid);
        if (tid < bytes - tail0) base[tail0 + tid] = (uint8_t)flag(tail0 + tid);
        uint32_t *body = reinterpret_cast<uint32_t *>(base + head);
        for (int i = tid; i < body4; i += nthr) {
          const int j = head + 4 * i;
          body[i] = flag(j) | (flag(j + 1) << 8) | (flag(j + 2) << 16) | (flag(j + 3) << 24);
        }
      }
    }
    if (stats && tid < P) {
      int tiles = 0;
      for (int k = 0; k < NW; k++) tiles += __popc(s[L.off_list + tid * NW + k]);
      int32_t *so = stats + ((size_t)game * P + tid) * 4;
      so[0] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * tid + GRL_PL_ARMY_COUNT];
      so[1] = tiles;
      so[2] = (int32_t)((s[GRL_HDR_FLAGS] >> tid) & 1u);
      so[3] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * tid + GRL_PL_GENERAL_IDX];
    }
    if (CTA) __syncthreads(); else __syncwarp();
  }
}

// The same read-outs for baked boards with N % 4 == 0 (10x10, 20x20), from the slab in SHARED memory, in the
// plane-major order of obs_plane_major: lane l owns the tile quads q = l + 32c, reads every mask's nibble for its
// quads once (packed 4 bits per chunk), converts its quads' armies once, and the warp writes the game's
// [P][9][N] block as one linear sweep of 128-bit stores with compile-time addressing.  The N*5 mask bytes of a
// player (tile-major, {up,right,down,left,any} per tile) are 20 bytes per quad: a lane assembles its five words,
// the warp stages them in shared memory and copies them out as a linear sweep.
#define GRL_GYM_EMIT_FN __forceinline__  // as real calls (__noinline__) the emitters measured 6-13 % slower
template <int PT, int N>
__device__ GRL_GYM_EMIT_FN void gym_emit_quads(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
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
  auto emit_obs_block = [&]() {
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
  };
  auto emit_mask_block = [&]() {
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
      uint8_t *base = mask + ((size_t)game * P + p) * (size_t)(N * 5);
      __syncwarp();
      // (one bulk copy per view, cp.async.bulk.global.shared::cta, instead of these stores: 20x20 2-4 % slower with either
      // L2 policy and with the copy overlapped by the view's planes; 10x10, whole game as one run, equal.  The 15x15 writer
      // below gains 4 % from it and keeps it.)
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
  };
  // the short mask runs first, then the long observation sweep (20x20: 0.4165 -> 0.4095 ms; the other order of the two
  // output streams, and default-policy instead of streaming stores for the mask words, measured slower)
  emit_mask_block();
  emit_obs_block();
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

// The same read-outs for baked boards with N % 4 != 0 (15x15), from the slab in SHARED memory, with the bit-stream writer
// of obs_linear (grl_obs.cuh): the game's [P][9][N] block is one linear, 16-byte aligned sweep of 128-bit stores; every
// plane is a 0/1 mask in the stream, and the planes that are not 0/1 multiply the table value by an aligned float4 of a
// per-view array F: ownership (0.5 on own tiles) and log-army are adjacent planes [1.. | 0.5 or 1 per tile | log per
// tile | 1..], the turn plane is [1.. | turn/max_turns per tile | ..].  Plane 8 is all zeros, so the sector a block
// shares with the next game of the warp is simply left to that game's pass.
// The N*5 mask bytes of all players are one run of 5-byte tile records: the four direction masks of the views are
// concatenated into P*N-bit streams, a lane expands a quad of four tiles into five words (as gym_emit_quads does), and the
// staged words leave as an aligned 32-bit sweep shifted by the block's misalignment.
template <int PT, int N>
__device__ GRL_GYM_EMIT_FN void gym_emit_linear(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                                float *__restrict__ obs, uint8_t *__restrict__ mask,
                                                int32_t *__restrict__ stats, const uint32_t *s, const uint32_t *stt,
                                                const float4 *lut, uint32_t *sw, int game, int lane, const Geo &g,
                                                bool prev_in_warp, bool next_in_warp) {
  constexpr int NWC = (N + 31) / 32, CH = GRL_GYM_CHANNELS;
  constexpr int DW = ((PT * N + 31) / 32 + 1 + 3) & ~3;                                        // words of one direction stream
  constexpr int FW = (2 * N + 8 + 3) & ~3;
  constexpr int SW = (((PT * GRL_GYM_CHANNELS * N + 127 + 31) / 32 + 1) + 3) & ~3;
  const GrlLayout &L = prm.L;
  const int P = prm.P, NW = NWC;
  const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
  uint32_t *s_dir = sw;                                // [4][DW]: up, right, down, left of every view, concatenated
  float *F = reinterpret_cast<float *>(sw + 4 * DW);   // [2N + 8]
  uint32_t *strm = sw + 4 * DW + FW;                   // [SW]
  uint32_t *stage = sw + 4 * DW;                       // the mask words reuse F + stream
  const bool w = lane < NWC;
  const uint32_t valid = w ? g.valid : 0u;
  const uint32_t M = w ? stt[lane] : 0u, C = w ? stt[NW + lane] : 0u, G = w ? stt[2 * NW + lane] : 0u;
  uint32_t any_own = 0;
#pragma unroll
  for (int p = 0; p < PT; p++)
    if (p < P && w) any_own |= s[L.off_own + p * NW + lane];
  // the previous game's bulk copy of its mask bytes must have read the staging area before it is written again: here when
  // the observation pass comes first (its stream and F share the area), else only before the mask words are staged
  if (obs) {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }

  if (obs) {
    const int planes = P * CH, total = planes * N;
    const size_t off = (size_t)game * total;
    const bool join_prev = prev_in_warp && (off & 7u) != 0;   // the previous block's last floats are zeros (plane 8)
    const int pre = join_prev ? (int)(off & 7u) : (int)(off & 3u);
    const int qend = pre + total;
    for (int k = lane; k < SW / 4; k += 32) reinterpret_cast<uint4 *>(strm)[k] = make_uint4(0u, 0u, 0u, 0u);
    float lg[NWC];  // log(army + 1) / 10 of tiles lane, lane + 32, ...
#pragma unroll
    for (int j = 0; j < NWC; j++) {
      const int t = lane + 32 * j;
      lg[j] = t < N ? __ldg(logtab + army[t]) : 0.f;  // logtab[0] == 0
    }
    const float tf = fminf(__fdiv_rn((float)s[GRL_HDR_TURN], (float)max_turns), 1.0f);
    __syncwarp();
    if (lane <= NWC) {
      int qs = pre;
#pragma unroll
      for (int p = 0; p < PT; p++) {
        if (p < P) {
          const uint32_t v = w ? (prm.fog ? s[L.off_vis + p * NW + lane] : valid) : 0u;
          const uint32_t ch[CH - 1] = {v, v & any_own, v, valid & ~(M | C | G), M, C, G, valid};  // plane 8 stays zero
#pragma unroll
          for (int c = 0; c < CH - 1; c++) stream_or_mask<NWC, true>(strm, ch[c], qs + c * N, lane);
          qs += CH * N;
        }
      }
    }
    __syncwarp();

    float *base_al = obs + (off - pre);
    const int k_first = (pre == 0 || join_prev) ? 0 : 1;
    auto stream_bit = [&](int q) -> uint32_t { return (strm[q >> 5] >> ((q + 4) & 31)) & 1u; };
    if (k_first && pre + lane < 4)  // floats before the aligned body: tiles 0.. of view 0's visibility plane
      __stcs(base_al + pre + lane, stream_bit(pre + lane) ? 1.f : 0.f);
    int k_end;
    if (next_in_warp && (qend & 7) != 0) {
      k_end = (qend >> 3) * 2;  // the shared sector is left to the next game's pass
    } else {
      k_end = qend >> 2;        // floats after the aligned body: zeros (plane 8)
      const int q = 4 * k_end + lane;
      if (q < qend) __stcs(base_al + q, 0.f);
    }

    StreamSweep swp;
    swp.init(lut, strm, base_al, k_first, k_end, lane);
    const uint32_t F_sa = smem_addr(F);
    for (int p = 0; p < P; p++) {
      const int qv = pre + p * CH * N;  // this view's first float
      // ---- planes 1 (ownership: 0.5 on own tiles in sight, 1 on enemy tiles) and 2 (log army in sight) ----------------
      const int qa = qv + N, sa = qa & 3;
      swp.plain(swp.rounds_before(qa, lane), lane);
      __syncwarp();  // F is free
      {
        const uint32_t mine = w ? ((prm.fog ? s[L.off_vis + p * NW + lane] : valid) & s[L.off_own + p * NW + lane]) : 0u;
#pragma unroll
        for (int j = 0; j < NWC; j++) {
          const int t = lane + 32 * j;
          const uint32_t mw = __shfl_sync(FULL, mine, j);
          if (t < N) {
            F[sa + t] = ((mw >> lane) & 1u) ? 0.5f : 1.f;
            F[sa + N + t] = lg[j];
          }
        }
        if (lane < sa) F[lane] = 1.f;
        if (lane < 4) F[sa + 2 * N + lane] = 1.f;
      }
      __syncwarp();
      swp.scaled(qa, qa + 2 * N - 1, F_sa, lane);
      // ---- plane 7: min(turn / max_turns, 1) on every tile ------------------------------------------------------------
      const int qt = qv + 7 * N, st = qt & 3;
      swp.plain(swp.rounds_before(qt, lane), lane);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NWC; j++) {
        const int t = lane + 32 * j;
        if (t < N) F[st + t] = tf;
      }
      if (lane < st) F[lane] = 1.f;
      if (lane < 4) F[st + N + lane] = 1.f;  // plane 8: zeros whatever the multiplier
      __syncwarp();
      swp.scaled(qt, qt + N - 1, F_sa, lane);
    }
    swp.finish(lane);
  }
  __syncwarp();
  if (mask) {
    const uint32_t gt1 = army_gt1_mask<32>(army, NW, N, g);
    const DirMasks dm = dir_targets<32>(M, g);
    for (int k = lane; k < DW; k += 32) reinterpret_cast<uint4 *>(s_dir)[k] = make_uint4(0u, 0u, 0u, 0u);  // 4 * DW words
    __syncwarp();
    if (lane <= NWC) {
#pragma unroll
      for (int p = 0; p < PT; p++) {
        if (p < P) {
          const uint32_t own = w ? s[L.off_own + p * NW + lane] : 0u;
          const uint32_t v = w ? (prm.fog ? s[L.off_vis + p * NW + lane] : valid) : 0u;
          const uint32_t src = v & own & gt1;
          stream_or_mask<NWC, false>(s_dir + 0 * DW, src & dm.up, p * N, lane);
          stream_or_mask<NWC, false>(s_dir + 1 * DW, src & dm.right, p * N, lane);
          stream_or_mask<NWC, false>(s_dir + 2 * DW, src & dm.down, p * N, lane);
          stream_or_mask<NWC, false>(s_dir + 3 * DW, src & dm.left, p * N, lane);
        }
      }
    }
    __syncwarp();
    // The staged words are placed so that stage and destination agree modulo 16 bytes: word 0 of the first quad sits
    // `woff` words into the staging area, where the block's first aligned word sits woff words into a 16-byte line.
    const int tiles = P * N, quads = (tiles + 3) >> 2, bytes = tiles * 5;
    const size_t goff = (size_t)game * bytes;
    const int mis = (int)(goff & 3u), woff = (int)(((goff - mis) >> 2) & 3u);
    const uint8_t *bd = reinterpret_cast<const uint8_t *>(s_dir);
    const int b0 = 4 * woff + mis, b1 = b0 + bytes;            // the block in line coordinates (below)
    uint8_t *line0 = mask + goff - b0;                          // 16-byte aligned
    if ((mis & 1) == 0) {
      // The staging area becomes a byte image of the 16-byte lines the block touches: stage byte i <-> line0[i].  The whole
      // lines leave as ONE bulk copy (cp.async.bulk.global.shared::cta, SASS UBLKCP) issued by lane 0, the ragged ends
      // (at most 14 bytes each) as 16-bit stores.  A block that starts 2 bytes into a word is staged in halfwords.
      uint16_t *img = reinterpret_cast<uint16_t *>(stage);
      if (!obs) {  // (the direction streams above were built while the previous copy drained)
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      for (int q = lane; q < quads; q += 32) {
        const int b = q >> 1, bsh = 4 * (q & 1);
        const uint32_t U = (bd[b] >> bsh) & 0xfu, R = (bd[4 * DW + b] >> bsh) & 0xfu;
        const uint32_t D = (bd[8 * DW + b] >> bsh) & 0xfu, Lm = (bd[12 * DW + b] >> bsh) & 0xfu, A = U | R | D | Lm;
        auto spread = [](uint32_t n) -> uint32_t { return (n * 0x00204081u) & 0x01010101u; };
        const uint32_t Ub = spread(U), Rb = spread(R), Db = spread(D), Lb = spread(Lm), Ab = spread(A);
        const uint32_t UR = __byte_perm(Ub, Rb, 0x5140), URh = __byte_perm(Ub, Rb, 0x7362);
        const uint32_t DL = __byte_perm(Db, Lb, 0x5140), DLh = __byte_perm(Db, Lb, 0x7362);
        const uint32_t o0 = __byte_perm(UR, DL, 0x5410);   // (byte layout: see the word-staged path below)
        const uint32_t o1 = __byte_perm(__byte_perm(UR, DL, 0x6320), Ab, 0x3214);
        const uint32_t o2 = __byte_perm(__byte_perm(DL, URh, 0x5403), Ab, 0x3250);
        const uint32_t o3 = __byte_perm(__byte_perm(DLh, URh, 0x6010), Ab, 0x3610);
        const uint32_t o4 = __byte_perm(__byte_perm(URh, DLh, 0x0763), Ab, 0x7210);
        if (mis == 0) {
          uint32_t *o = stage + woff + 5 * q;
          o[0] = o0, o[1] = o1, o[2] = o2, o[3] = o3, o[4] = o4;
        } else {
          uint16_t *o = img + 2 * (woff + 5 * q) + 1;
          o[0] = (uint16_t)o0, o[1] = (uint16_t)(o0 >> 16), o[2] = (uint16_t)o1, o[3] = (uint16_t)(o1 >> 16);
          o[4] = (uint16_t)o2, o[5] = (uint16_t)(o2 >> 16), o[6] = (uint16_t)o3, o[7] = (uint16_t)(o3 >> 16);
          o[8] = (uint16_t)o4, o[9] = (uint16_t)(o4 >> 16);
        }
      }
      fence_proxy_async_smem();  // every lane's staged bytes become visible to the async proxy
      __syncwarp();
      const int l0 = (b0 + 15) >> 4, l1 = b1 >> 4;                // whole lines [l0, l1)
      if (lane == 0 && l1 > l0) {
        tma_store(line0 + 16 * l0, reinterpret_cast<const uint8_t *>(stage) + 16 * l0, (uint32_t)(16 * (l1 - l0)));
        tma_store_commit();
      }
      for (int x = b0 + 2 * lane; x < 16 * l0 && x < b1; x += 64) *reinterpret_cast<uint16_t *>(line0 + x) = img[x >> 1];
      for (int x = (l1 > l0 ? 16 * l1 : 16 * l0) + 2 * lane; x < b1; x += 64) *reinterpret_cast<uint16_t *>(line0 + x) = img[x >> 1];
    } else {  // odd byte offsets (an odd number of views): staged as words, shifted on the way out
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    if (lane < 4) stage[lane] = 0u;  // the words before the first quad's (read by the shifted copy, never stored)
    for (int q = lane; q < quads; q += 32) {
      const int b = q >> 1, bsh = 4 * (q & 1);
      const uint32_t U = (bd[b] >> bsh) & 0xfu, R = (bd[4 * DW + b] >> bsh) & 0xfu;
      const uint32_t D = (bd[8 * DW + b] >> bsh) & 0xfu, Lm = (bd[12 * DW + b] >> bsh) & 0xfu, A = U | R | D | Lm;
      // byte 5i+k of the quad = direction k of its tile i:  U0 R0 D0 L0 | A0 U1 R1 D1 | L1 A1 U2 R2 | D2 L2 A2 U3 | R3 D3 L3 A3
      auto spread = [](uint32_t n) -> uint32_t { return (n * 0x00204081u) & 0x01010101u; };
      const uint32_t Ub = spread(U), Rb = spread(R), Db = spread(D), Lb = spread(Lm), Ab = spread(A);
      const uint32_t UR = __byte_perm(Ub, Rb, 0x5140), URh = __byte_perm(Ub, Rb, 0x7362);
      const uint32_t DL = __byte_perm(Db, Lb, 0x5140), DLh = __byte_perm(Db, Lb, 0x7362);
      uint32_t *o = stage + 4 + woff + 5 * q;
      o[0] = __byte_perm(UR, DL, 0x5410);
      o[1] = __byte_perm(__byte_perm(UR, DL, 0x6320), Ab, 0x3214);
      o[2] = __byte_perm(__byte_perm(DL, URh, 0x5403), Ab, 0x3250);
      o[3] = __byte_perm(__byte_perm(DLh, URh, 0x6010), Ab, 0x3610);
      o[4] = __byte_perm(__byte_perm(URh, DLh, 0x0763), Ab, 0x7210);
    }
    __syncwarp();
    // copy-out.  In "line" coordinates (16-byte lines from the aligned line that holds the block's first byte) staged
    // word 4 + i holds the bytes of line word i shifted up by `mis` bytes: line word i = bytes [4i - mis, 4i - mis + 4)
    // of the staged run = funnel(staged[4 + i - 1], staged[4 + i]) >> 8 (4 - mis).  The block covers line bytes
    // [b0, b1): whole lines go out as 128-bit stores, the ragged ends as bytes.
    const int shb = (32 - 8 * mis) & 31, adj = mis ? 0 : 1;
    auto line_word = [&](int i) -> uint32_t { return __funnelshift_r(stage[3 + i + adj], stage[4 + i + adj], shb); };
    const int l0 = (b0 + 15) >> 4, l1 = b1 >> 4;                // whole lines [l0, l1)
    for (int l = l0 + lane; l < l1; l += 32)
      __stcs(reinterpret_cast<uint4 *>(line0) + l, make_uint4(line_word(4 * l), line_word(4 * l + 1), line_word(4 * l + 2), line_word(4 * l + 3)));
    // ragged ends: bytes [b0, 16 l0) and [16 l1, b1) (at most 15 each); byte x of the line space = byte x & 3 of line word x >> 2
    for (int x = b0 + lane; x < 16 * l0 && x < b1; x += 32) line0[x] = (uint8_t)(line_word(x >> 2) >> (8 * (x & 3)));
    for (int x = (l1 > l0 ? 16 * l1 : 16 * l0) + lane; x < b1; x += 32) line0[x] = (uint8_t)(line_word(x >> 2) >> (8 * (x & 3)));
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

// ---- the observation half of gym_emit_linear with a compile-time schedule (a warp's four whole games, P == PT; CtRun /
// ct_rounds in grl_device.cuh).  Planes 1-2 of a view take their multipliers (0.5 on own tiles, log-army) from F; plane 7
// (turn fraction, the same on every tile) multiplies by a register, with per-float range tests in the rounds it can
// touch; plane 8 is zeros, so the carried floats of a game's incomplete round are zero bits.  `sw` is the warp's scratch
// as gym_emit_linear lays it out; the mask half (gym_emit_linear with obs == nullptr) reuses F and the stream as its
// staging area.
template <int PT, int N>
__device__ GRL_GYM_EMIT_FN void gym_run_game_obs(const GrlKParams &prm, int max_turns, const float *__restrict__ logtab,
                                                 const uint32_t *s, const uint32_t *stt, const CtLane &c, uint32_t *sw,
                                                 int lane, const Geo &g, int gi) {
  constexpr int NWC = (N + 31) / 32, CH = GRL_GYM_CHANNELS, TOTAL = PT * CH * N, NW = NWC, GPW = 4;
  using R = CtRun<TOTAL, GPW>;
  constexpr int DW = ((PT * N + 31) / 32 + 1 + 3) & ~3;
  constexpr int FW = (2 * N + 8 + 3) & ~3;
  constexpr int SW = (((PT * GRL_GYM_CHANNELS * N + 127 + 31) / 32 + 1) + 3) & ~3;
  static_assert(SW >= R::stream_words(), "the stream holds a carried round in front of the block");
  static_assert(R::win_hi((PT - 1) * CH * N + 7 * N, N) < R::NR_MIN, "plane 7 ends before the rounds only some games have");
  static_assert(N >= 127, "the carried round lies inside the previous game's last plane (zeros)");
  const int q0 = TOTAL * gi, rlo = q0 >> 7, pre = q0 & 127;
  const int nr = (gi == GPW - 1 ? (GPW * TOTAL + 127) >> 7 : (q0 + TOTAL) >> 7) - rlo;  // rounds of this pass
  float4 *op = c.op + 32 * rlo;
  const GrlLayout &L = prm.L;
  const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
  float *F = reinterpret_cast<float *>(sw + 4 * DW);
  uint32_t *strm = sw + 4 * DW + FW;
  uint4 *strm4 = reinterpret_cast<uint4 *>(strm);
  const bool w = lane < NWC;
  const uint32_t valid = w ? g.valid : 0u;
  const uint32_t M = w ? stt[lane] : 0u, C = w ? stt[NW + lane] : 0u, G = w ? stt[2 * NW + lane] : 0u;
  uint32_t any_own = 0;
#pragma unroll
  for (int p = 0; p < PT; p++)
    if (w) any_own |= s[L.off_own + p * NW + lane];
  // the carried floats in front of the block are the previous game's plane 8: zero bits, like the rest of a fresh stream
  for (int k = lane; k < SW / 4; k += 32) strm4[k] = make_uint4(0u, 0u, 0u, 0u);
  float lg[NWC];  // log(army + 1) / 10 of tiles lane, lane + 32, ...
#pragma unroll
  for (int j = 0; j < NWC; j++) {
    const int t = lane + 32 * j;
    lg[j] = t < N ? __ldg(logtab + army[t]) : 0.f;  // logtab[0] == 0
  }
  const float tf = fminf(__fdiv_rn((float)s[GRL_HDR_TURN], (float)max_turns), 1.0f);
  __syncwarp();
  if (lane <= NWC) {
#pragma unroll
    for (int p = 0; p < PT; p++) {
      const uint32_t v = w ? (prm.fog ? s[L.off_vis + p * NW + lane] : valid) : 0u;
      const uint32_t ch[CH - 1] = {v, v & any_own, v, valid & ~(M | C | G), M, C, G, valid};  // plane 8 stays zero
#pragma unroll
      for (int k = 0; k < CH - 1; k++) stream_or_mask<NWC, false>(strm, ch[k], pre + (p * CH + k) * N, lane);
    }
  }
  __syncwarp();
  constexpr int kAll = CT_ALL_ROUNDS;
  int jdone = 0;
#pragma unroll
  for (int p = 0; p < PT; p++) {
    // ---- planes 1 (ownership: 0.5 on own tiles in sight, 1 on enemy tiles) and 2 (log army in sight): stream floats
    //      [qa, qa + 2N), inside rounds ja_lo..ja_hi whatever `pre` is --------------------------------------------------
    const int ja_lo = R::win_lo(p * CH * N + N), ja_hi = R::win_hi(p * CH * N + N, 2 * N);
    const int qa = pre + p * CH * N + N, qz = qa + 2 * N - 1, sa = qa & 3, ka = qa >> 2, kz = qz >> 2;
    ct_rounds<CT_PLAIN>(c, op, lane, jdone, ja_lo, 0, 0, 0u, 0, 0, 0.f, kAll, 32);
    __syncwarp();  // F is free
    {
      const uint32_t mine = w ? ((prm.fog ? s[L.off_vis + p * NW + lane] : valid) & s[L.off_own + p * NW + lane]) : 0u;
#pragma unroll
      for (int j = 0; j < NWC; j++) {
        const int t = lane + 32 * j;
        const uint32_t mw = __shfl_sync(FULL, mine, j);
        if (t < N) {
          F[sa + t] = ((mw >> lane) & 1u) ? 0.5f : 1.f;
          F[sa + N + t] = lg[j];
        }
      }
      if (lane < sa) F[lane] = 1.f;
      if (lane < 4) F[sa + 2 * N + lane] = 1.f;
    }
    __syncwarp();
    ct_rounds<CT_F>(c, op, lane, ja_lo, ja_hi + 1, ka, kz, c.F - 16u * (uint32_t)ka, 0, 0, 0.f, kAll, 32);
    // ---- plane 7: min(turn / max_turns, 1) on every tile: stream floats [qt, qt + N) ----------------------------------
    const int jt_lo = R::win_lo(p * CH * N + 7 * N), jt_hi = R::win_hi(p * CH * N + 7 * N, N);
    ct_rounds<CT_PLAIN>(c, op, lane, ja_hi + 1, jt_lo, 0, 0, 0u, 0, 0, 0.f, kAll, 32);
    ct_rounds<CT_SCALAR>(c, op, lane, jt_lo, jt_hi + 1, 0, 0, 0u, pre + p * CH * N + 7 * N, N, tf, kAll, 32);
    jdone = jt_hi + 1;
  }
  ct_rounds<CT_PLAIN>(c, op, lane, jdone, R::NR_MIN, 0, 0, 0u, 0, 0, 0.f, kAll, 32);
  ct_rounds<CT_PLAIN>(c, op, lane, R::NR_MIN, R::NR_MAX, 0, 0, 0u, 0, 0, 0.f, nr, gi == GPW - 1 ? R::last_active() : 32);
  __syncwarp();
}

#undef GYM_NIB4
