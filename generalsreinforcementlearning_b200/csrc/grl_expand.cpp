// grl_expand.cpp — host side of the packed observation read-out (include/grlcuda.h): grl_expand_obs turns packed
// records into Serializer.StateToTensor's float32 tensors (internal/experience/serializer.go:37-109), bit for bit what
// the turn kernel writes into grl_step_outputs.obs.  No device work; plain C++ the compiler vectorises (the bit ->
// float loops run over one 32-bit mask word at a time), one thread per slice of records.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/grlcuda.h"
#include "grl_layout.h"

namespace {

// 32 tiles of one channel: out[b] = bit b of w ? value : 0
__attribute__((target_clones("avx2", "default"))) void expand_ones(uint32_t w, float *out, int n) {
  for (int b = 0; b < n; b++) out[b] = ((w >> b) & 1u) ? 1.0f : 0.0f;
}
__attribute__((target_clones("avx2", "default"))) void expand_vals(uint32_t w, const float *val, float *out, int n) {
  for (int b = 0; b < n; b++) out[b] = ((w >> b) & 1u) ? val[b] : 0.0f;
}

void expand_record(const GrlLayout &L, const uint32_t *rec, float *obs, float *frac) {
  const int N = L.N, P = L.P, NW = L.NW, PNW = P * NW;
  const uint32_t *own = rec, *vis = rec + PNW, *M = rec + 2 * PNW, *CG = M + NW;
  const uint16_t *army = reinterpret_cast<const uint16_t *>(CG + NW);
  // float32(army) / 1000, clipped to 1 (serializer.go:80-88)
  for (int t = 0; t < N; t++) frac[t] = army[t] >= 1000 ? 1.0f : (float)army[t] / 1000.0f;
  for (int p = 0; p < P; p++) {
    float *o = obs + (size_t)p * GRL_OBS_CHANNELS * N;
    for (int j = 0; j < NW; j++) {
      uint32_t any_own = 0;
      for (int q = 0; q < P; q++) any_own |= own[q * NW + j];
      const uint32_t valid = (32 * j + 32 <= N) ? 0xffffffffu : ((1u << (N - 32 * j)) - 1u);
      const uint32_t v = vis[p * NW + j] & valid, nm = v & ~M[j];
      const uint32_t mine = nm & own[p * NW + j], enemy = nm & any_own & ~own[p * NW + j];
      const int n = std::min(32, N - 32 * j);
      float *oj = o + 32 * j;
      expand_vals(mine, frac + 32 * j, oj + 0 * N, n);
      expand_vals(enemy, frac + 32 * j, oj + 1 * N, n);
      expand_ones(mine, oj + 2 * N, n);
      expand_ones(enemy, oj + 3 * N, n);
      expand_ones(nm & ~any_own, oj + 4 * N, n);
      expand_ones(nm & CG[j], oj + 5 * N, n);
      expand_ones(v & M[j], oj + 6 * N, n);
      expand_ones(v, oj + 7 * N, n);
      expand_ones(~v & valid, oj + 8 * N, n);
    }
  }
}

}  // namespace

extern "C" {

int32_t grl_obs_packed_words(int32_t width, int32_t height, int32_t num_players) {
  if (width < 1 || width > GRL_MAX_DIM || height < 1 || height > GRL_MAX_DIM || num_players < 1 || num_players > GRL_MAX_PLAYERS)
    return 0;
  return grl_packed_words(grl_make_layout(width, height, num_players));
}

int grl_expand_obs(int32_t width, int32_t height, int32_t num_players, const uint32_t *packed, int32_t count, float *obs,
                   int32_t threads) {
  if (!packed || !obs || count < 0 || grl_obs_packed_words(width, height, num_players) == 0) return GRL_ERR_INVALID_ARG;
  const GrlLayout L = grl_make_layout(width, height, num_players);
  const int RW = grl_packed_words(L);
  const size_t per = (size_t)L.P * GRL_OBS_CHANNELS * L.N;
  int hw = (int)std::thread::hardware_concurrency();
  int nt = threads > 0 ? threads : (hw > 0 ? hw : 1);
  nt = std::max(1, std::min(nt, count / 64 + 1));
  auto work = [&](int t) {
    std::vector<float> frac((size_t)L.N + 32);
    const int c0 = (int)((int64_t)count * t / nt), c1 = (int)((int64_t)count * (t + 1) / nt);
    for (int c = c0; c < c1; c++) expand_record(L, packed + (size_t)c * RW, obs + (size_t)c * per, frac.data());
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
  }
  return GRL_OK;
}

}  // extern "C"
