// grl_mapgen_gpu.cu — the reference's procedural maps generated ON THE DEVICE, one warp per map.
//
// Host map generation (grl_mapgen.cpp) costs ~3 us of wall time per 20x20 map on 16 cores — most of
// it Go's math/rand seeding, 1,841 Lehmer steps per seed — which for 65,536 games is as long as
// stepping the whole 500-turn episode.  The algorithm is embarrassingly parallel over maps and
// every draw is integer arithmetic, so the same restatement of mapgen/generator.go:25-253 and of
// go1.24 math/rand runs here.  A warp seeds the generator cooperatively (the k-th Lehmer state is
// 48271^k * x_0, so the 607 state words are independent), keeps the state, the occupancy masks and the
// army plane of its map in SHARED memory, lets lane 0 make the (inherently sequential) draws, and writes
// the finished map into the staging slabs the reset kernel consumes with coalesced stores — a seeded
// reset never touches the host.  (The first version ran one THREAD per map with the state in local memory:
// every draw waited on two dependent local-memory loads, ≈ 130 us per map however few maps there were —
// the critical path of a vector env that re-seeds a hundred-odd envs per step.)
// Bit-identical to the host path by construction; tests compare the two and the oracle.
#include <cuda_runtime.h>
#include <stdint.h>

#include "grl_launch.h"
#include "grl_layout.h"

namespace {

constexpr int kLen = 607;
constexpr int kTap = 273;
constexpr int32_t kInt32Max = 2147483647;

// both tables are indexed per lane (word i, power 21+3i): from constant memory those reads would serialise 32 ways, so
// they live in global memory and are fetched through the read-only path (coalesced, L1/L2 resident)
__device__ const int64_t c_cooked[kLen] = {
#include "go_rng_cooked.inc"
};

// 48271^k mod (2^31-1): x_k = 48271^k * x_0, so the 1,841 seeding steps become independent multiplications
__device__ const int32_t c_lpow[1842] = {
#include "go_rng_lehmer_pow.inc"
};

struct GoRandDev {  // rng.go (*rngSource) + rand.go
  uint64_t *vec;  // [kLen], shared memory
  int tap, feed;

  // every lane of the warp calls this; lane l fills words l, l+32, ...
  __device__ void seed(int64_t s, int lane) {
    tap = 0;
    feed = kLen - kTap;
    s %= kInt32Max;
    if (s < 0) s += kInt32Max;
    if (s == 0) s = 89482311;
    // rng.go seedrand runs 20 warm-up steps of x <- 48271 * x mod (2^31-1), then three per state word; the k-th
    // state is 48271^k * x_0, so word i takes the independent products for k = 21+3i .. 23+3i
    const uint64_t x0 = (uint64_t)s;
#pragma unroll  // 19 rounds, every table read independent of the others: one trip to L2, not nineteen
    for (int j = 0; j < (kLen + 31) / 32; j++) {
      const int i = lane + 32 * j;
      if (i >= kLen) break;
      const int k = 21 + 3 * i;
      const uint64_t u = ((uint64_t)mulmod31(__ldg(c_lpow + k), x0) << 40) ^ ((uint64_t)mulmod31(__ldg(c_lpow + k + 1), x0) << 20) ^
                         (uint64_t)mulmod31(__ldg(c_lpow + k + 2), x0);
      vec[i] = u ^ (uint64_t)__ldg(c_cooked + i);
    }
  }
  __device__ static uint32_t mulmod31(int32_t a, uint64_t x) {  // a * x mod (2^31 - 1), both in [1, 2^31 - 2]
    uint64_t r = (uint64_t)(uint32_t)a * x;
    r = (r >> 31) + (r & 0x7fffffffULL);
    r = (r >> 31) + (r & 0x7fffffffULL);
    return (uint32_t)(r >= 0x7fffffffULL ? r - 0x7fffffffULL : r);
  }
  __device__ uint64_t u64() {
    if (--tap < 0) tap += kLen;
    if (--feed < 0) feed += kLen;
    uint64_t x = vec[feed] + vec[tap];
    vec[feed] = x;
    return x;
  }
  __device__ int64_t int63() { return (int64_t)(u64() & 0x7fffffffffffffffULL); }
  __device__ int32_t int31() { return (int32_t)(int63() >> 32); }
  __device__ uint32_t u32() { return (uint32_t)(int63() >> 31); }
  __device__ int intn(int n) {  // rand.go Intn, n <= 2^31-1
    if ((n & (n - 1)) == 0) return int31() & (n - 1);
    const int32_t limit = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
    int32_t v = int31();
    while (v > limit) v = int31();
    return v % n;
  }
  __device__ static int32_t intn_limit(int n) { return (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n); }
  __device__ int intn(int n, int32_t limit) {  // Intn with the rejection limit of a non-power-of-two n precomputed
    int32_t v = int31();
    if ((n & (n - 1)) == 0) return v & (n - 1);
    while (v > limit) v = int31();
    return v % n;
  }
  __device__ int32_t lemire(int32_t n) {  // rand.go int31n, used by Shuffle
    uint32_t v = u32();
    uint64_t prod = (uint64_t)v * (uint64_t)n;
    uint32_t low = (uint32_t)prod;
    if (low < (uint32_t)n) {
      const uint32_t thresh = (uint32_t)(-n) % (uint32_t)n;
      while (low < thresh) {
        v = u32();
        prod = (uint64_t)v * (uint64_t)n;
        low = (uint32_t)prod;
      }
    }
    return (int32_t)(prod >> 32);
  }
};

struct MapDev {  // occupancy as linear bitmasks (tile t = bit t&31 of word t>>5), N <= 1024; all in shared memory
  uint32_t *taken;      // anything that is not a neutral normal tile
  uint32_t *M, *C, *G;  // terrain masks
  uint16_t *army;
  __device__ bool open(int t) const { return !((taken[t >> 5] >> (t & 31)) & 1u); }
  __device__ void take(int t) { taken[t >> 5] |= 1u << (t & 31); }
};

__device__ bool far_enough(int W, int idx, const int *seats, int n, int spacing) {
  const int x = idx % W, y = idx / W;
  for (int k = 0; k < n; ++k) {
    const int ox = seats[k] % W, oy = seats[k] / W;
    if (abs(x - ox) + abs(y - oy) < spacing) return false;
  }
  return true;
}

}  // namespace

// One warp writes map i into staging slab i / static slab i: every word grl_reset_kernel reads (the ownership masks,
// the army plane to the end of the slab, the whole static slab), so the staging rows need no zero fill.
// With `final_obs` the CTA carries a second set of warps: warp kMapWarps + w saves player 0's observation block of env
// ids[i] (grl_gym_autoreset: the episode's last view, before the re-seeded env's read-outs overwrite it) while warp w
// generates that env's next map.
constexpr int kMapWarps = 4;  // maps per CTA: 4 x (607 x 8 B state + 512 B masks + 2 KB armies) = 29.6 KB of shared memory

__global__ void __launch_bounds__(2 * kMapWarps * 32)
    grl_mapgen_kernel(GrlLayout L, int W, int H, GrlMapParams mp, const long long *__restrict__ seeds, int n,
                      uint32_t *__restrict__ slabs, uint32_t *__restrict__ statics, int *__restrict__ failed,
                      const int *__restrict__ n_dev, const int32_t *__restrict__ ids, const float *__restrict__ obs,
                      float *__restrict__ final_obs, int obs_block) {
  __shared__ uint64_t s_vec[kMapWarps][kLen];
  __shared__ uint32_t s_mask[kMapWarps][4][32];
  __shared__ __align__(16) uint16_t s_army[kMapWarps][1024];
  __shared__ int s_seat[kMapWarps][GRL_MAX_PLAYERS_DEV + 1];  // [players] = 1 when every general found a seat
  const int warp = (threadIdx.x >> 5) % kMapWarps, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kMapWarps + warp;
  if (n_dev) n = min(n, *n_dev);  // the launch covers the capacity; the number of maps lives on the device
  if (i >= n) return;             // uniform over the warp
  const int N = W * H, NW = L.NW;
  uint32_t *slab = slabs + (size_t)i * L.slab_words;
  uint32_t *stat = statics + (size_t)i * L.static_words;
  if (threadIdx.x >= kMapWarps * 32) {  // a copy warp (launched only with final_obs): [P][obs_block] per env in `obs`, [obs_block] per env in `final_obs`; both 4-byte aligned only
    const int b = ids[i];
    const float *src = obs + (size_t)b * mp.players * obs_block;
    float *dst = final_obs + (size_t)b * obs_block;
    const int head = min(obs_block, (int)((4u - (uint32_t)(((size_t)b * obs_block) & 3u)) & 3u));  // dst to 16 bytes
    if (lane < head) dst[lane] = src[lane];
    const int n4 = (obs_block - head) >> 2;
    float4 *d4 = reinterpret_cast<float4 *>(dst + head);
    const float *s1 = src + head;
    for (int k0 = 0; k0 < n4; k0 += 8 * 32) {  // eight independent rounds of loads in flight
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = k0 + 32 * u + lane;
        if (k < n4) v[u] = make_float4(s1[4 * k], s1[4 * k + 1], s1[4 * k + 2], s1[4 * k + 3]);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = k0 + 32 * u + lane;
        if (k < n4) d4[k] = v[u];
      }
    }
    const int t0 = head + 4 * n4;
    if (t0 + lane < obs_block) dst[t0 + lane] = src[t0 + lane];
    return;
  }
  for (int k = lane; k < 4 * 32; k += 32) (&s_mask[warp][0][0])[k] = 0u;
  for (int k = lane; k < 512; k += 32) reinterpret_cast<uint32_t *>(s_army[warp])[k] = 0u;
  GoRandDev rng;
  rng.vec = s_vec[warp];
  rng.seed(seeds[i], lane);
  __syncwarp();
  if (lane == 0) {
  MapDev b;
  b.taken = s_mask[warp][0];
  b.M = s_mask[warp][1];
  b.C = s_mask[warp][2];
  b.G = s_mask[warp][3];
  b.army = s_army[warp];
  int *seats = s_seat[warp];
  seats[mp.players] = 0;
  const int32_t limW = GoRandDev::intn_limit(W), limH = GoRandDev::intn_limit(H);  // two of three draws are Intn(W) / Intn(H)

  // generator.go:77-142 mountain veins
  for (int vein = 0; vein < mp.veins; ++vein) {
    int cx = -1, cy = -1;
    for (int attempt = 0; attempt < 100; ++attempt) {
      const int sx = rng.intn(W, limW);
      const int sy = rng.intn(H, limH);
      if (b.open(sy * W + sx)) {
        cx = sx;
        cy = sy;
        break;
      }
    }
    if (cx < 0) continue;
    int t = cy * W + cx;
    b.take(t);
    b.M[t >> 5] |= 1u << (t & 31);
    int length = mp.min_vein;
    if (mp.max_vein > mp.min_vein) length += rng.intn(mp.max_vein - mp.min_vein + 1);
    for (int step = 1; step < length; ++step) {
      // rand.Shuffle(4, swap) of the directions up, right, down, left, packed 2 bits each
      uint32_t order = 0xE4u;  // slots 3,2,1,0 = 3,2,1,0
      for (int a = 3; a > 0; --a) {
        const int j = rng.lemire(a + 1);
        const uint32_t va = (order >> (2 * a)) & 3u, vj = (order >> (2 * j)) & 3u;
        order = (order & ~((3u << (2 * a)) | (3u << (2 * j)))) | (vj << (2 * a)) | (va << (2 * j));
      }
      int options[4], n_options = 0;
      for (int d = 0; d < 4; ++d) {
        const int dir = (order >> (2 * d)) & 3;
        const int nx = cx + (dir == 1) - (dir == 3), ny = cy + (dir == 2) - (dir == 0);
        if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
        if (b.open(ny * W + nx)) options[n_options++] = ny * W + nx;
      }
      if (n_options == 0) break;
      const int pick = rng.intn(n_options);
      int next = options[0];
      for (int d = 1; d < 4; ++d)
        if (d == pick) next = options[d];
      cx = next % W;
      cy = next / W;
      b.take(next);
      b.M[next >> 5] |= 1u << (next & 31);
    }
  }
  // generator.go:144-164 cities
  {
    const int want = N / mp.city_ratio;
    int placed = 0;
    for (int attempts = 0; placed < want && attempts < want * 20; ++attempts) {
      const int x = rng.intn(W, limW);
      const int y = rng.intn(H, limH);
      const int t = y * W + x;
      if (b.open(t)) {
        b.take(t);
        b.C[t >> 5] |= 1u << (t & 31);
        b.army[t] = (uint16_t)mp.city_start_army;
        ++placed;
      }
    }
  }
  // generator.go:166-253 generals
  bool ok = true;
  for (int pid = 0; pid < mp.players && ok; ++pid) {
    int chosen = -1;
    for (int attempt = 0; attempt < N && chosen < 0; ++attempt) {
      const int x = rng.intn(W, limW);
      const int y = rng.intn(H, limH);
      const int t = y * W + x;
      if (b.open(t) && far_enough(W, t, seats, pid, mp.spacing)) chosen = t;
    }
    for (int t = 0; t < N && chosen < 0; ++t)  // row-major fallback
      if (b.open(t) && far_enough(W, t, seats, pid, mp.spacing)) chosen = t;
    if (chosen < 0) {
      atomicExch(failed, i + 1);
      ok = false;
      break;
    }
    b.take(chosen);
    b.G[chosen >> 5] |= 1u << (chosen & 31);
    b.army[chosen] = 2;
    seats[pid] = chosen;
  }
  seats[mp.players] = ok ? 1 : 0;
  }  // lane 0
  __syncwarp();

  // ---- the finished map leaves shared memory with coalesced stores ---------------------------------------------
  const bool placed = s_seat[warp][mp.players] != 0;  // a map without seats for every general leaves as an empty board (`failed` is set)
  for (int k = lane; k < L.static_words; k += 32)
    stat[k] = (placed && k < 3 * NW) ? (&s_mask[warp][1][0])[(k / NW) * 32 + k % NW] : 0u;
  const uint32_t *sarmy = reinterpret_cast<const uint32_t *>(s_army[warp]);
  for (int k = L.off_army + lane; k < L.slab_words; k += 32)
    slab[k] = (placed && k - L.off_army < L.NA / 2) ? sarmy[k - L.off_army] : 0u;
  for (int k = lane; k < L.P * NW; k += 32) {  // ownership: one general per player
    const int p = k / NW, w = k - p * NW;
    const int t = (placed && p < mp.players) ? s_seat[warp][p] : -1;
    slab[L.off_own + k] = (t >= 0 && (t >> 5) == w) ? 1u << (t & 31) : 0u;
  }
}

cudaError_t grl_launch_mapgen(const GrlLayout &L, int W, int H, const GrlMapParams &mp, const long long *seeds, int n,
                              uint32_t *slabs, uint32_t *statics, int *failed, cudaStream_t stream, const int *n_dev,
                              const int32_t *ids, const float *obs, float *final_obs, int obs_block) {
  grl_mapgen_kernel<<<(n + kMapWarps - 1) / kMapWarps, (final_obs ? 2 : 1) * kMapWarps * 32, 0, stream>>>(L, W, H, mp, seeds, n, slabs, statics, failed, n_dev,
                                                                                   ids, obs, final_obs, obs_block);
  return cudaGetLastError();
}
