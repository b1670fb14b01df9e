// grl_kernels.cu — the kernels off the turn path (reset, byte-per-flag read-outs, gym glue, digests, the stand-alone
// synthetic policy) and the launchers grl_abi.cu calls.  The fused turn kernel lives in grl_turn.cuh and is
// instantiated per board geometry in grl_turn_*.cu.
#include <stdlib.h>
#include <string.h>

#include "grl_device.cuh"
#include "grl_gym.cuh"
#include "grl_launch.h"

// ---------------------------------------------------------------------------------------
// Reset: freshly uploaded slabs carry ownership, armies and terrain; this kernel performs
// the turn-0 set-up of engine_initializer.go:113-143,218-225 (players alive, full stats,
// full fog, game-over check), preserving the env's lifetime counters.
// ---------------------------------------------------------------------------------------
// One game's turn-0 set-up by a whole warp: staging slab `ss` / static slab `sst` -> the env's slab `ds` / static slab `dst`.
template <int PT>
__device__ __forceinline__ void reset_one(const GrlKParams &prm, const uint32_t *__restrict__ ss, const uint32_t *__restrict__ sst,
                                          uint32_t *ds, uint32_t *dst, int lane, const Geo &g) {
  const GrlLayout &L = prm.L;
  const int P = prm.P, NW = prm.NW, N = prm.N;
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

template <int PT>
__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32)
    grl_reset_kernel(const __grid_constant__ GrlKParams prm, const uint32_t *__restrict__ src_state,
                     const uint32_t *__restrict__ src_static, const int32_t *__restrict__ env_ids, int n,
                     const int *__restrict__ n_dev) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (n_dev) n = min(n, *n_dev);  // device-side count (grl_gym_autoreset)
  const GrlLayout &L = prm.L;
  const Geo g = make_geo(prm, prm.W, lane, 32);
  for (int i = blockIdx.x * GRL_WARPS_PER_CTA + warp; i < n; i += gridDim.x * GRL_WARPS_PER_CTA) {
    const int game = env_ids ? env_ids[i] : i;
    if (game < 0 || game >= prm.B) continue;
    reset_one<PT>(prm, src_state + (size_t)i * L.slab_words, src_static + (size_t)i * L.static_words,
                  prm.state + (size_t)game * L.slab_words, const_cast<uint32_t *>(prm.statics) + (size_t)game * L.static_words, lane, g);
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

// grl_gym_autoreset, last step: the turn-0 set-up of every re-seeded env and its gym read-outs in ONE launch, one thread
// BLOCK per env: warp 0 runs reset_one, then the block reads the slab it has just written back (block barrier) and shares
// the read-out sweeps (gym_emit<0, true>).  A vector step re-seeds a hundred-odd envs and waits for their read-outs, so what
// counts here is the latency of one env, not throughput: a single warp per env took 28 us, two launches 35 us.
template <int PT>
__global__ void __launch_bounds__(GRL_WARPS_PER_CTA * 32)
    grl_gym_reseed_kernel(const __grid_constant__ GrlKParams prm, const uint32_t *__restrict__ src_state,
                          const uint32_t *__restrict__ src_static, const int32_t *__restrict__ ids, int n,
                          const int *__restrict__ n_dev, int max_turns, const float *__restrict__ logtab, float *__restrict__ obs,
                          uint8_t *__restrict__ mask, int32_t *__restrict__ stats) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const GrlLayout &L = prm.L;
  const Geo g = make_geo(prm, prm.W, lane, 32);
  n = min(n, *n_dev);
  for (int i = blockIdx.x; i < n; i += gridDim.x) {  // uniform over the block
    const int game = ids[i];
    if (game < 0 || game >= prm.B) continue;
    uint32_t *ds = prm.state + (size_t)game * L.slab_words;
    uint32_t *dst = const_cast<uint32_t *>(prm.statics) + (size_t)game * L.static_words;
    if (threadIdx.x < 32)
      reset_one<PT>(prm, src_state + (size_t)i * L.slab_words, src_static + (size_t)i * L.static_words, ds, dst, lane, g);
    __threadfence_block();
    __syncthreads();
    gym_emit<0, true>(prm, max_turns, logtab, obs, mask, stats, ds, dst, smem, game, lane, g);
  }
}

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

// grl_replay_push_rows: one thread block per env copies the env's observation row (view `view` of [B][views][F]) into the
// next_states ring row of the transition that just ended (final_obs[b] where done[b]) and into the states ring row of the
// env's next transition — one read of the plane, two coalesced write streams.  Rows are F floats at 4-byte alignment
// (F = 9 * 225 is odd), so accesses are 32-bit; a warp instruction still covers 128 contiguous bytes.
__global__ void __launch_bounds__(256) grl_replay_rows_kernel(const float *__restrict__ obs, const float *__restrict__ final_obs,
                                                              const uint8_t *__restrict__ done, float *__restrict__ next_states,
                                                              float *__restrict__ states, long long capacity, long long next_row0,
                                                              long long state_row0, int views, int view, int F, int B) {
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float *row = obs + ((size_t)b * views + view) * (size_t)F;
    const bool fin = next_states && done && final_obs && done[b] != 0;
    const float *nsrc = fin ? final_obs + (size_t)b * F : row;
    float *nd = next_states ? next_states + (size_t)((next_row0 + b) % capacity) * F : nullptr;
    float *sd = states ? states + (size_t)((state_row0 + b) % capacity) * F : nullptr;
    for (int k0 = threadIdx.x; k0 < F; k0 += 4 * 256) {  // four independent loads in flight per thread
      float v[4], w[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = k0 + 256 * u;
        if (k < F) {
          v[u] = __ldcs(row + k);
          w[u] = fin ? __ldcs(nsrc + k) : v[u];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = k0 + 256 * u;
        if (k < F) {
          if (nd) __stcs(nd + k, w[u]);
          if (sd) __stcs(sd + k, v[u]);
        }
      }
    }
  }
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
    // the loads of four rounds are issued before the first is consumed: a row is then one or two trips to DRAM (15x15:
    // three rounds, 20x20: four), not one per round — this kernel is bounded by that latency, not by the 74-131 MB it reads
    for (int r0 = 0; r0 < rounds; r0 += 4) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v = 32 * (r0 + u) + lane;
        q[u] = make_uint4(0u, 0u, 0u, 0u);
        if (v < nvec && !((last_row && v == nvec - 1) || (first_row && v == 0))) q[u] = __ldg(base + v);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v = 32 * (r0 + u) + lane;
        if (r0 + u >= rounds) break;
        uint32_t m16 = 0;
        if (v < nvec) {
          const int g0 = 16 * v - mis;  // row-relative index of this vector's first byte
          if ((last_row && v == nvec - 1) || (first_row && v == 0)) {
            for (int j = 0; j < 16; j++)
              if (g0 + j >= 0 && g0 + j < M && row[g0 + j] != 0) m16 |= 1u << j;
          } else {
            m16 = nonzero_bytes4(q[u].x) | (nonzero_bytes4(q[u].y) << 4) | (nonzero_bytes4(q[u].z) << 8) |
                  (nonzero_bytes4(q[u].w) << 12);
            if (g0 < 0) m16 &= 0xffffu << (-g0);                 // bytes before the row
            if (g0 + 16 > M) m16 &= 0xffffu >> (g0 + 16 - M);    // bytes past its end
          }
        }
        sm[v] = (uint16_t)m16;
        mine += __popc(m16);
      }
    }
    const int total = __reduce_add_sync(FULL, mine);
    __syncwarp();
    long long pick = 0;
    if (total > 0) {
      const uint64_t rr = policy_draw(seed, (uint64_t)(prm.env_id_base + b), 0ull, (uint64_t)player);
      int k = (int)(rr % (uint64_t)total);
      for (int r = 0; r < rounds; r++) {  // the round the k-th set byte lies in, then one prefix sum inside it
        const uint32_t w = sm[32 * r + lane];
        const int c = __popc(w);
        const int stripe = __reduce_add_sync(FULL, c);
        if (k < stripe) {
          int incl = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
          }
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

// the BASELINE board sizes get kernels with the geometry baked in and the lane group sized to the board
// (grl_turn_20.cu, grl_turn_15.cu, grl_turn_10.cu); everything else, and more than four players, runs the generic
// instantiation with one game per warp (grl_turn_generic.cu)
cudaError_t grl_launch_turn(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  if (prm.P <= 4) {
    if (prm.W == 20 && prm.H == 20) return grl_launch_turn_20x20(prm, do_step, do_out, stream);
    if (prm.W == 15 && prm.H == 15) return grl_launch_turn_15x15(prm, do_step, do_out, stream);
    if (prm.W == 10 && prm.H == 10) return grl_launch_turn_10x10(prm, do_step, do_out, stream);
  }
  return grl_launch_turn_generic(prm, do_step, do_out, stream);
}

// The fused gym step (one GeneralsEnv.step() per env in ONE launch).
cudaError_t grl_launch_gym_step(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  if (prm.P <= 4) {
    if (prm.W == 20 && prm.H == 20) return grl_launch_gym_step_20x20(prm, gk, stream);
    if (prm.W == 15 && prm.H == 15) return grl_launch_gym_step_15x15(prm, gk, stream);
    if (prm.W == 10 && prm.H == 10) return grl_launch_gym_step_10x10(prm, gk, stream);
  }
  return grl_launch_gym_step_generic(prm, gk, stream);
}

static int player_template(int P) { return P <= 2 ? 2 : (P <= 4 ? 4 : 8); }

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

cudaError_t grl_launch_gym_reseed(const GrlKParams &prm, const uint32_t *src_state, const uint32_t *src_static, const int32_t *ids,
                                  int n, const int *n_dev, int max_turns, const float *logtab, float *obs, uint8_t *mask,
                                  int32_t *stats, cudaStream_t stream) {
  const size_t smem = (size_t)grl_gym_smem_words(prm.P, prm.NW, prm.N, GRL_GYM_EMIT_GENERIC) * 4u;
  const int grid = n < 148 * 8 ? (n < 1 ? 1 : n) : 148 * 8;  // one block per env, strided over the device-side count
#define GRL_RESEED(PT)                                                                                                         \
  {                                                                                                                            \
    static size_t tuned = 0;                                                                                                   \
    if (smem > 48 * 1024 && smem > tuned) {                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(grl_gym_reseed_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                                          \
      tuned = smem;                                                                                                            \
    }                                                                                                                          \
    grl_gym_reseed_kernel<PT><<<grid, GRL_WARPS_PER_CTA * 32, smem, stream>>>(prm, src_state, src_static, ids, n, n_dev, max_turns, \
                                                                              logtab, obs, mask, stats);                      \
  }
  switch (player_template(prm.P)) {
    case 2: GRL_RESEED(2) break;
    case 4: GRL_RESEED(4) break;
    default: GRL_RESEED(8) break;
  }
#undef GRL_RESEED
  return cudaGetLastError();
}

cudaError_t grl_launch_replay_rows(const float *obs, const float *final_obs, const uint8_t *done, float *next_states, float *states,
                                   long long capacity, long long next_row0, long long state_row0, int views, int view, int F, int B,
                                   cudaStream_t stream) {
  const int grid = B < 148 * 64 ? B : 148 * 64;
  grl_replay_rows_kernel<<<grid, 256, 0, stream>>>(obs, final_obs, done, next_states, states, capacity, next_row0, state_row0, views,
                                                   view, F, B);
  return cudaGetLastError();
}

cudaError_t grl_launch_gym_sample(const GrlKParams &prm, unsigned long long seed, const uint8_t *mask, int player, long long *action,
                                  cudaStream_t stream) {
  grl_gym_sample_kernel<<<grid_for(8, prm.B), 256, 0, stream>>>(prm, seed, mask, player, action);
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
