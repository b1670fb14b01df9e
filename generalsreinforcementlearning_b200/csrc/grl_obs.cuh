// grl_obs.cuh — the observation writers: Serializer.StateToTensor (internal/experience/serializer.go:37-109)
// as 128-bit streaming stores straight into the caller's [B][P][9][H][W] fp32 tensor.
#pragma once
#include "grl_device.cuh"

// per-warp shared-memory words of the linear observation writer (only baked boards with N % 4 != 0)
__host__ __device__ constexpr int grl_obs_scratch_words(int TW, int TH, int PT, int NW) {
  // channel masks [PT*9][NW+1] + army-fraction plane [N+4], rounded to 16 bytes, + the P*9-1 precomputed float4s that
  // straddle two planes
  return (TW > 0 && ((TW * TH) & 3) != 0)
             ? (((PT * GRL_OBS_CHANNELS * (NW + 1) + TW * TH + 4 + 3) & ~3) + 4 * PT * GRL_OBS_CHANNELS)
             : 0;
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
  const bool join_prev = prev_slab != nullptr && lead != 0;
  const bool join_next = next_in_warp && trail != 0;
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
      __stcs(reinterpret_cast<float4 *>(base - lead) + lane, make_float4(v4[0], v4[1], v4[2], v4[3]));
    }
    head = 8 - lead;
  } else {
    head = (int)((4u - (uint32_t)(((size_t)game * total) & 3u)) & 3u);  // floats before the 16-byte aligned body
    if (lane < head) __stcs(base + lane, obs_element<N>(chm, frac, NWP, lane));
  }
  if (join_next) {
    end = total - trail;  // the shared sector is left to the next game's pass
  } else {
    end = head + 4 * ((total - head) / 4);
    if (lane < total - end) __stcs(base + end + lane, obs_element<N>(chm, frac, NWP, end + lane));
  }
  const int body4 = (end - head) / 4;
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
  const char *lutb = reinterpret_cast<const char *>(lut);
  float4 *body = reinterpret_cast<float4 *>(base + head);
#pragma unroll 2
  for (int i = lane; i < body4; i += 32) {
    const int e = head + 4 * i;
    const int plane = e / N, t = e - plane * N;
    const int k = plane % GRL_OBS_CHANNELS;
    const uint32_t *row = chm + plane * NWP;
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
      val = sf[plane];  // evaluated before the sweep, one per lane
    }
    __stcs(body + i, val);
  }
  __syncwarp();
}
