// grl_obs.cuh — the observation writers: Serializer.StateToTensor (internal/experience/serializer.go:37-109)
// as 128-bit streaming stores straight into the caller's [B][P][9][H][W] fp32 tensor.
#pragma once
#include "grl_device.cuh"

// per-warp shared-memory words of the bit-stream observation writer (only baked boards with N % 4 != 0): the
// army-multiplier array F [2N + 8] and the block's bit stream (one bit per float, up to 127 carried lead bits — a whole
// round, see CtRun in grl_device.cuh — and one spare word)
__host__ __device__ constexpr int grl_obs_region_words(int N) { return (2 * N + 8 + 3) & ~3; }
__host__ __device__ constexpr int grl_obs_stream_words(int N, int PT) {
  return (((PT * GRL_OBS_CHANNELS * N + 127 + 31) / 32 + 1) + 3) & ~3;
}
__host__ __device__ constexpr int grl_obs_scratch_words(int TW, int TH, int PT, int NW) {
  return (TW > 0 && ((TW * TH) & 3) != 0) ? grl_obs_region_words(TW * TH) + grl_obs_stream_words(TW * TH, PT) : 0;
}

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
            const uint32_t idx16 = (c == 0 ? (ch[k] << 4) : (ch[k] >> (4 * c - 4))) & 0xf0u;
            float4 val = *reinterpret_cast<const float4 *>(lutb + idx16);
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

// ---- the bit-stream writer (baked geometries with N % 4 != 0: 15x15) -----------------------------------------------------
// Channel planes start at odd float offsets there, so the game's [P][9][N] block is written as ONE linear sweep of
// 16-byte aligned 128-bit stores addressed by position in the block, not in a plane.  All but the two army planes of a
// view hold 0/1, so the block is first laid out as a BIT STREAM in shared memory — bit q of the stream is float q of the
// run that starts at the block's aligned base — by OR-ing every channel mask in at its plane's bit offset (one funnel
// shift per mask word).  A store is then: one stream word, a rotate and a mask for the lane's 4-bit field, one 16-entry
// table lookup, one STG.128 — no division, no plane arithmetic, and the float4s that straddle two planes cost nothing
// extra.  Stream words are kept rotated left by 4 so that rotating right by the field's offset leaves the field at bits
// 4-7: the byte offset of its table entry.  The army planes (channels 0/1 of each view, two adjacent planes) multiply
// the table value by an aligned float4 of F = [1.. | army/1000 per tile | the same again | 1..], rebuilt per view at
// the view's own alignment.
//
// A 15x15 block is 16,200 bytes: it starts and ends mid-sector, and a 32-byte sector completed by two different store
// instructions costs the memory system far more than its bytes (tools/micro/store_holes.cu).  Consecutive games of one
// warp therefore join their blocks into one sector-complete run: the sector two blocks share is written WHOLE by the
// later game's pass — its first floats are the tail of the previous game's last plane (player P-1, channel 8 = fog),
// taken from that game's slab, which is still in shared memory, into the first bits of this game's stream.
template <int PT, int N>
__device__ __forceinline__ void obs_linear(const GrlKParams &prm, const SlabView &S, const float4 *lut, uint32_t *scratch,
                                           int P, int NW, int game, int lane, const uint32_t *prev_slab = nullptr,
                                           bool next_in_warp = false) {
  constexpr int NWC = (N + 31) / 32, CH = GRL_OBS_CHANNELS;
  constexpr int SW = grl_obs_stream_words(N, PT);
  float *F = reinterpret_cast<float *>(scratch);  // [2N + 8]
  uint32_t *strm = scratch + grl_obs_region_words(N);
  const int planes = P * CH, total = planes * N;  // floats in this game's block
  const size_t off = (size_t)game * total;
  // run coordinates: float q of the run is float q - pre of this block; the run starts at a sector (joined to the
  // previous game of the warp) or at the 16-byte boundary at or before the block
  const bool join_prev = prev_slab != nullptr && (off & 7u) != 0;
  const int pre = join_prev ? (int)(off & 7u) : (int)(off & 3u);
  const int qend = pre + total;

  for (int w = lane; w < SW / 4; w += 32) reinterpret_cast<uint4 *>(strm)[w] = make_uint4(0u, 0u, 0u, 0u);
  float fr[NWC];  // army / 1000 of tiles lane, lane + 32, ...
#pragma unroll
  for (int j = 0; j < NWC; j++) {
    const int t = lane + 32 * j;
    fr[j] = t < N ? army_frac((uint32_t)S.army[t]) : 0.f;
  }
  __syncwarp();
  if (lane <= NWC) {
    const bool w = lane < NWC;
    const uint32_t valid = w ? prm.geom[lane] : 0u;
    const uint32_t M = w ? S.M[lane] : 0u;
    const uint32_t CG = w ? (S.C[lane] | S.G[lane]) : 0u;
    uint32_t any_own = 0;
#pragma unroll
    for (int p = 0; p < PT; p++)
      if (p < P && w) any_own |= S.own[p * NW + lane];
    int qs = pre;
#pragma unroll
    for (int p = 0; p < PT; p++) {
      if (p < P) {
        const uint32_t own = w ? S.own[p * NW + lane] : 0u;
        const uint32_t v = w ? (prm.fog ? S.vis[p * NW + lane] : valid) : 0u;
        const uint32_t nm = v & ~M;
        const uint32_t mine = nm & own, enemy = nm & any_own & ~own;   // serializer.go:75-90
        const uint32_t ch[CH] = {mine, enemy, mine, enemy, nm & ~any_own, nm & CG, v & M, v, ~v & valid};
#pragma unroll
        for (int c = 0; c < CH; c++) {
          stream_or_mask<NWC, true>(strm, ch[c], qs, lane);
          qs += N;
        }
      }
    }
    if (lane == 0 && join_prev && prm.fog) {  // the previous block's last `pre` floats: its last view's fog plane
      const uint32_t *pv = prev_slab + prm.L.off_vis + (P - 1) * NW;
      const int t = N - pre, wl = t >> 5;
      const uint32_t lo = ~pv[wl] & prm.geom[wl], hi = wl + 1 < NWC ? (~pv[wl + 1] & prm.geom[wl + 1]) : 0u;
      atomicOr(strm, rotl4(__funnelshift_r(lo, hi, t & 31) & ((1u << pre) - 1u)));
    }
  }
  __syncwarp();

  float *base_al = prm.obs + (off - pre);
  const int k_first = (pre == 0 || join_prev) ? 0 : 1;
  auto stream_bit = [&](int q) -> uint32_t { return (strm[q >> 5] >> ((q + 4) & 31)) & 1u; };
  if (k_first && pre + lane < 4)  // floats before the aligned body: tiles 0.. of view 0's own-army plane
    __stcs(base_al + pre + lane, stream_bit(pre + lane) ? fr[0] : 0.f);
  int k_end;
  if (next_in_warp && (qend & 7) != 0) {
    k_end = (qend >> 3) * 2;  // the shared sector is left to the next game's pass
  } else {
    k_end = qend >> 2;        // floats after the aligned body: the last view's fog plane
    const int q = 4 * k_end + lane;
    if (q < qend) __stcs(base_al + q, stream_bit(q) ? 1.f : 0.f);
  }

  StreamSweep sw;
  sw.init(lut, strm, base_al, k_first, k_end, lane);
  const uint32_t F_sa = smem_addr(F);
  for (int p = 0; p < P; p++) {
    const int qs = pre + p * CH * N, s = qs & 3;  // this view's two army planes: floats [qs, qs + 2N)
    sw.plain(sw.rounds_before(qs, lane), lane);
    __syncwarp();  // the previous view's army rounds are done with F
#pragma unroll
    for (int j = 0; j < NWC; j++) {
      const int t = lane + 32 * j;
      if (t < N) {
        F[s + t] = fr[j];
        F[s + N + t] = fr[j];
      }
    }
    if (lane < s) F[lane] = 1.f;
    if (lane < 4) F[s + 2 * N + lane] = 1.f;
    __syncwarp();
    sw.scaled(qs, qs + 2 * N - 1, F_sa, lane);
  }
  sw.finish(lane);
  __syncwarp();
}

// ---- the same writer with a compile-time schedule (a warp's four whole games, P == PT; CtRun / ct_rounds in
// grl_device.cuh).  Game gi of the run: its stream starts with the `carry` words of the previous game's incomplete round;
// round indices are constants, the bit offsets of the planes and the float4 range of the army planes follow from the
// game's `pre` in registers.  `carry` returns this game's own incomplete round for the next game.
template <int PT, int N>
__device__ __forceinline__ void obs_run_game(const GrlKParams &prm, const SlabView &S, const CtLane &c, uint32_t *scratch,
                                             int NW, int lane, int gi, uint4 &carry) {
  constexpr int NWC = (N + 31) / 32, CH = GRL_OBS_CHANNELS, TOTAL = PT * CH * N, GPW = 4;
  using R = CtRun<TOTAL, GPW>;
  constexpr int SW = grl_obs_stream_words(N, PT);
  static_assert(SW >= R::stream_words(), "the stream holds a carried round in front of the block");
  static_assert(R::win_hi((PT - 1) * CH * N, 2 * N) < R::NR_MIN, "the army planes end before the rounds only some games have");
  const int q0 = TOTAL * gi, rlo = q0 >> 7, pre = q0 & 127;
  const int nr = (gi == GPW - 1 ? (GPW * TOTAL + 127) >> 7 : (q0 + TOTAL) >> 7) - rlo;  // rounds of this pass
  float4 *op = c.op + 32 * rlo;
  float *F = reinterpret_cast<float *>(scratch);  // [2N + 8]
  uint32_t *strm = scratch + grl_obs_region_words(N);
  uint4 *strm4 = reinterpret_cast<uint4 *>(strm);
  for (int w = lane; w < SW / 4; w += 32) strm4[w] = w == 0 ? carry : make_uint4(0u, 0u, 0u, 0u);
  float fr[NWC];  // army / 1000 of tiles lane, lane + 32, ...
#pragma unroll
  for (int j = 0; j < NWC; j++) {
    const int t = lane + 32 * j;
    fr[j] = t < N ? army_frac((uint32_t)S.army[t]) : 0.f;
  }
  __syncwarp();
  if (lane <= NWC) {
    const bool w = lane < NWC;
    const uint32_t valid = w ? prm.geom[lane] : 0u;
    const uint32_t M = w ? S.M[lane] : 0u;
    const uint32_t CG = w ? (S.C[lane] | S.G[lane]) : 0u;
    uint32_t any_own = 0;
#pragma unroll
    for (int p = 0; p < PT; p++)
      if (w) any_own |= S.own[p * NW + lane];
#pragma unroll
    for (int p = 0; p < PT; p++) {
      const uint32_t own = w ? S.own[p * NW + lane] : 0u;
      const uint32_t v = w ? (prm.fog ? S.vis[p * NW + lane] : valid) : 0u;
      const uint32_t nm = v & ~M;
      const uint32_t mine = nm & own, enemy = nm & any_own & ~own;   // serializer.go:75-90
      const uint32_t ch[CH] = {mine, enemy, mine, enemy, nm & ~any_own, nm & CG, v & M, v, ~v & valid};
#pragma unroll
      for (int k = 0; k < CH; k++) stream_or_mask<NWC, false>(strm, ch[k], pre + (p * CH + k) * N, lane);
    }
  }
  __syncwarp();
  int jdone = 0;
#pragma unroll
  for (int p = 0; p < PT; p++) {
    // this view's two army planes: stream floats [qa, qa + 2N), inside rounds j_lo..j_hi whatever `pre` is
    constexpr int kAll = CT_ALL_ROUNDS;
    const int j_lo = R::win_lo(p * CH * N), j_hi = R::win_hi(p * CH * N, 2 * N);
    const int qa = pre + p * CH * N, qz = qa + 2 * N - 1, s = qa & 3, ka = qa >> 2, kz = qz >> 2;
    ct_rounds<CT_PLAIN>(c, op, lane, jdone, j_lo, 0, 0, 0u, 0, 0, 0.f, kAll, 32);
    __syncwarp();  // the previous view's army rounds are done with F
#pragma unroll
    for (int j = 0; j < NWC; j++) {
      const int t = lane + 32 * j;
      if (t < N) {
        F[s + t] = fr[j];
        F[s + N + t] = fr[j];
      }
    }
    if (lane < s) F[lane] = 1.f;
    if (lane < 4) F[s + 2 * N + lane] = 1.f;
    __syncwarp();
    ct_rounds<CT_F>(c, op, lane, j_lo, j_hi + 1, ka, kz, c.F - 16u * (uint32_t)ka, 0, 0, 0.f, kAll, 32);
    jdone = j_hi + 1;
  }
  ct_rounds<CT_PLAIN>(c, op, lane, jdone, R::NR_MIN, 0, 0, 0u, 0, 0, 0.f, CT_ALL_ROUNDS, 32);
  ct_rounds<CT_PLAIN>(c, op, lane, R::NR_MIN, R::NR_MAX, 0, 0, 0u, 0, 0, 0.f, nr, gi == GPW - 1 ? R::last_active() : 32);
  if (gi < GPW - 1) carry = strm4[nr];
  __syncwarp();
}
