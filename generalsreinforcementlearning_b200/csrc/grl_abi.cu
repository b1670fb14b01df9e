// grl_abi.cu — the C ABI of libgrlcuda.so (include/grlcuda.h) over the kernels.
//
// Host responsibilities only: device memory, the stream, staging of host buffers,
// host-side map generation, (un)packing of state slabs for get/set_state.  All game
// logic runs in grl_kernels.cu.  There is no CPU fallback: every entry point that
// touches game state launches CUDA work and fails with GRL_ERR_CUDA if it cannot.
#include <cuda_runtime.h>
#include <cmath>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/grlcuda.h"
#include "grl_launch.h"
#include "grl_layout.h"
#include "grl_mapgen.h"

namespace {

thread_local char g_err[512] = "";

int fail(int status, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return status;
}

#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) return fail(GRL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

enum Slot { SL_ACTIONS, SL_OBS, SL_MASK, SL_REWARD, SL_DONE, SL_WINNER, SL_ERR, SL_AIDX, SL_PACKED, SL_MISC, SL_MISC2, SL_COUNT };

struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct grl_env {
  grl_config cfg;
  GrlLayout L;
  int N;
  cudaStream_t stream = nullptr;      // the stream work is issued on
  cudaStream_t own_stream = nullptr;  // created by grl_create
  uint32_t *d_state = nullptr;
  uint32_t *d_static = nullptr;
  uint32_t *d_geom = nullptr;
  float *d_logtab = nullptr;  // float32(log(a + 1) / 10) for a in 0..65535 (generals_env.py:324), built on first use
  Scratch scratch[SL_COUNT];
  int prefetch_dist = 0;       // > 0: warp of game g prefetches the slab of game g+dist into L2 (GRL_PREFETCH_DIST)
  int host_mapgen = 0;         // GRL_HOST_MAPGEN=1: generate seeded maps on the host (cross-check of the device generator)
  uint64_t launches = 0;
  int host_threads = 1;
  // host-buffer calls split the batch into sub-ranges on these streams so that the copies of one
  // sub-range overlap the kernel of another (created on first use)
  static constexpr int kPipe = 8;
  cudaStream_t pipe[kPipe] = {};
  cudaEvent_t ev_start = nullptr, ev_done[kPipe] = {};
  int pipe_chunks = 6;  // GRL_PIPE_CHUNKS=1 disables the pipelining (e2e: 160 M env-steps/s at 1, 169 M at 4, 170 M at 6-8)
  // launch overlap (grl_turn.cuh): per-warp epoch words, the sequence number of the last turn launch, and that number again
  // while nothing else has been enqueued on the stream since (the next whole-batch turn launch may then overlap it)
  uint32_t *d_epoch = nullptr;
  uint32_t epoch_seq = 0, overlap_prev = 0;
  bool overlap_published = false;  // the launch overlap_prev names wrote its epoch words
  int overlap = 1;      // GRL_LAUNCH_OVERLAP=0 serialises the launches as CUDA does by default
};

namespace {

// Launch overlap bookkeeping.  A turn launch may be marked as overlapping its predecessor only when that predecessor is
// the SAME env's turn launch and nothing this library knows of was enqueued on the stream since — by any env: two envs
// stepped alternately on one stream must not overlap, because the later env's grid does not wait for the earlier env's
// warps and could finish first, and what follows it in the stream would then run before the earlier launch is complete.
std::mutex g_overlap_mu;
std::unordered_map<cudaStream_t, std::pair<const grl_env *, uint32_t>> g_last_turn_launch;  // per stream: env, sequence number

void note_other_work(grl_env *env) {  // anything but a whole-batch turn launch was enqueued on the env's stream
  env->overlap_prev = 0;
  std::lock_guard<std::mutex> lk(g_overlap_mu);
  g_last_turn_launch[env->stream] = {nullptr, 0u};
}
void note_turn_launch(grl_env *env, uint32_t seq) {
  env->overlap_prev = seq;
  std::lock_guard<std::mutex> lk(g_overlap_mu);
  g_last_turn_launch[env->stream] = {seq ? env : nullptr, seq};
}
bool last_on_stream_is(const grl_env *env, uint32_t seq) {
  std::lock_guard<std::mutex> lk(g_overlap_mu);
  auto it = g_last_turn_launch.find(env->stream);
  return it != g_last_turn_launch.end() && it->second.first == env && it->second.second == seq && seq != 0;
}

bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Pinned (page-locked, mapped) host memory is addressable from the device under unified addressing:
// returns the device alias of such a buffer, nullptr for pageable host memory or device memory.
void *pinned_device_alias(const void *p) {
  if (!p) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int ensure(grl_env *env, Slot s, size_t bytes, void **out) {
  Scratch &sc = env->scratch[s];
  if (sc.bytes < bytes) {
    if (sc.ptr) cudaFree(sc.ptr);
    sc.ptr = nullptr;
    sc.bytes = 0;
    cudaError_t e = cudaMalloc(&sc.ptr, bytes);
    if (e != cudaSuccess) return fail(GRL_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    sc.bytes = bytes;
  }
  *out = sc.ptr;
  return GRL_OK;
}

GrlKParams base_params(const grl_env *env) {
  GrlKParams p;
  memset(&p, 0, sizeof p);
  const grl_config &c = env->cfg;
  p.state = env->d_state;
  p.statics = env->d_static;
  p.geom = env->d_geom;
  p.B = c.num_envs;
  p.game0 = 0;
  p.game_end = c.num_envs;
  p.W = c.width;
  p.H = c.height;
  p.N = env->N;
  p.P = c.num_players;
  p.NW = env->L.NW;
  p.A = c.max_actions;
  p.L = env->L;
  p.fog = c.fog_of_war != 0;
  p.pg = c.production_general;
  p.pc = c.production_city;
  p.pn = c.production_normal;
  p.grow_interval = c.normal_growth_interval;
  p.env_id_base = c.env_id_base;
  p.prefetch_dist = env->prefetch_dist;
  const grl_reward_config &r = c.reward;
  const float rw[11] = {r.win_game,        r.lose_game,        r.capture_city, r.lose_city, r.capture_general, r.lose_general,
                        r.territory_gained, r.territory_lost, r.army_gained,  r.army_lost, r.army_advantage};
  memcpy(p.rw, rw, sizeof rw);
  return p;
}

int check_config(const grl_config *c) {
  if (!c) return fail(GRL_ERR_INVALID_ARG, "null config");
  if (c->num_envs < 1) return fail(GRL_ERR_INVALID_ARG, "num_envs must be >= 1");
  if (c->width < 1 || c->width > GRL_MAX_DIM || c->height < 1 || c->height > GRL_MAX_DIM)
    return fail(GRL_ERR_INVALID_ARG, "width/height must be in 1..%d", GRL_MAX_DIM);
  if (c->num_players < 1 || c->num_players > GRL_MAX_PLAYERS)
    return fail(GRL_ERR_INVALID_ARG, "num_players must be in 1..%d", GRL_MAX_PLAYERS);
  if (c->max_actions < 1 || c->max_actions > GRL_MAX_ACTIONS)
    return fail(GRL_ERR_INVALID_ARG, "max_actions must be in 1..%d", GRL_MAX_ACTIONS);
  if (c->city_ratio < 1 || c->normal_growth_interval < 1) return fail(GRL_ERR_INVALID_ARG, "bad game constants");
  if (c->production_general < 0 || c->production_city < 0 || c->production_normal < 0)
    return fail(GRL_ERR_INVALID_ARG, "negative production");
  return GRL_OK;
}

// ---- host <-> slab packing (get/set_state, reset staging) --------------------------------

struct HostGame {  // one env, unpacked
  std::vector<int32_t> owner, army, type;
  std::vector<uint32_t> visible;
  std::vector<uint8_t> owned, changed, vis_changed;  // owned: [P][N]
  int32_t turn = 0, game_over = 0, step_error = 0;
  int32_t alive[GRL_MAX_PLAYERS], army_count[GRL_MAX_PLAYERS], general_idx[GRL_MAX_PLAYERS];
  uint32_t counters[5] = {0, 0, 0, 0, 0};
};

inline bool getbit(const uint32_t *w, int t) { return (w[t >> 5] >> (t & 31)) & 1u; }
inline void setbit(uint32_t *w, int t) { w[t >> 5] |= 1u << (t & 31); }

void unpack_slab(const GrlLayout &L, const uint32_t *s, const uint32_t *st, HostGame &g) {
  const int N = L.N, P = L.P, NW = L.NW;
  g.owner.assign(N, -1);
  g.army.assign(N, 0);
  g.type.assign(N, 0);
  g.visible.assign(N, 0);
  g.owned.assign((size_t)P * N, 0);
  g.changed.assign(N, 0);
  g.vis_changed.assign(N, 0);
  const uint16_t *army = reinterpret_cast<const uint16_t *>(s + L.off_army);
  for (int t = 0; t < N; t++) {
    for (int p = 0; p < P; p++) {
      if (getbit(s + L.off_own + p * NW, t)) g.owner[t] = p;
      if (getbit(s + L.off_list + p * NW, t)) g.owned[(size_t)p * N + t] = 1;
      if (getbit(s + L.off_vis + p * NW, t)) g.visible[t] |= 1u << p;
    }
    g.army[t] = army[t];
    g.type[t] = getbit(st, t) ? GRL_TILE_MOUNTAIN : getbit(st + NW, t) ? GRL_TILE_CITY : getbit(st + 2 * NW, t) ? GRL_TILE_GENERAL : GRL_TILE_NORMAL;
    g.changed[t] = getbit(s + L.off_changed, t);
    g.vis_changed[t] = getbit(s + L.off_vchg, t);
  }
  g.turn = (int32_t)s[GRL_HDR_TURN];
  const uint32_t flags = s[GRL_HDR_FLAGS];
  g.game_over = (flags & GRL_FLAG_OVER) ? 1 : 0;
  g.step_error = (int32_t)((flags >> GRL_FLAG_ERR_SHIFT) & 0xffu);
  for (int p = 0; p < P; p++) {
    g.alive[p] = (flags >> p) & 1u;
    g.army_count[p] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_ARMY_COUNT];
    g.general_idx[p] = (int32_t)s[GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p + GRL_PL_GENERAL_IDX];
  }
  for (int k = 0; k < 5; k++) g.counters[k] = s[GRL_HDR_STEPS + k];
}

// returns 0 or a message
const char *pack_slab(const GrlLayout &L, const HostGame &g, uint32_t *s, uint32_t *st) {
  const int N = L.N, P = L.P, NW = L.NW;
  memset(s, 0, sizeof(uint32_t) * (size_t)L.slab_words);
  memset(st, 0, sizeof(uint32_t) * (size_t)L.static_words);
  uint16_t *army = reinterpret_cast<uint16_t *>(s + L.off_army);
  int64_t true_army[GRL_MAX_PLAYERS] = {0};
  for (int t = 0; t < N; t++) {
    const int o = g.owner[t];
    if (o < -1 || o >= P) return "owner out of range";
    if (g.army[t] < 0 || g.army[t] > 65535) return "army outside the uint16 plane's range 0..65535";
    if (g.type[t] < 0 || g.type[t] > 3) return "tile type out of range";
    army[t] = (uint16_t)g.army[t];
    if (o >= 0) {
      setbit(s + L.off_own + o * NW, t);
      true_army[o] += g.army[t];
    }
    int in_lists = 0;
    for (int p = 0; p < P; p++) {
      if (g.owned[(size_t)p * N + t]) {
        setbit(s + L.off_list + p * NW, t);
        in_lists++;
      }
      if ((g.visible[t] >> p) & 1u) setbit(s + L.off_vis + p * NW, t);
    }
    // reachable reference states keep the cached lists disjoint (every stats rebuild filters by
    // the current owner, stats.go:104-127); overlapping lists are not a state the engine can be in
    if (in_lists > 1) return "a tile may belong to at most one cached OwnedTiles list";
    if (g.type[t] == GRL_TILE_MOUNTAIN) setbit(st, t);
    if (g.type[t] == GRL_TILE_CITY) setbit(st + NW, t);
    if (g.type[t] == GRL_TILE_GENERAL) setbit(st + 2 * NW, t);
    if (g.changed[t]) setbit(s + L.off_changed, t);
    if (g.vis_changed[t]) setbit(s + L.off_vchg, t);
  }
  s[GRL_HDR_TURN] = (uint32_t)g.turn;
  uint32_t flags = g.game_over ? GRL_FLAG_OVER : 0u;
  flags |= ((uint32_t)g.step_error & 0xffu) << GRL_FLAG_ERR_SHIFT;
  for (int p = 0; p < P; p++) {
    if (g.alive[p]) flags |= 1u << p;
    uint32_t *h = s + GRL_HDR_PLAYER0 + GRL_HDR_PER_PLAYER * p;
    h[GRL_PL_ARMY_COUNT] = (uint32_t)g.army_count[p];
    h[GRL_PL_GENERAL_IDX] = (uint32_t)g.general_idx[p];
    h[GRL_PL_TRUE_ARMY] = (uint32_t)true_army[p];
    h[GRL_PL_REWARD] = 0u;
    h[GRL_PL_ACTION_INDEX] = 0xffffffffu;
  }
  s[GRL_HDR_FLAGS] = flags;
  for (int k = 0; k < 5; k++) s[GRL_HDR_STEPS + k] = g.counters[k];
  return nullptr;
}

void make_geom(int W, int H, uint32_t *geom) {
  memset(geom, 0, sizeof(uint32_t) * 96);
  for (int t = 0; t < W * H; t++) {
    setbit(geom, t);
    if (t % W != 0) setbit(geom + 32, t);
    if (t % W != W - 1) setbit(geom + 64, t);
  }
}

template <typename F>
void parallel_for(int n, int threads, F fn) {
  threads = std::max(1, std::min(threads, n));
  if (threads == 1) {
    for (int i = 0; i < n; i++) fn(i);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([=]() {
      for (int i = t; i < n; i += threads) fn(i);
    });
  for (auto &th : pool) th.join();
}

// Upload staged slabs for n envs and run the turn-0 set-up kernel.
int upload_and_reset(grl_env *env, const int32_t *env_ids, int n, const std::vector<uint32_t> &slabs,
                     const std::vector<uint32_t> &statics) {
  const GrlLayout &L = env->L;
  void *d_slabs = nullptr, *d_statics = nullptr;
  int st = ensure(env, SL_MISC, slabs.size() * 4 + 16, &d_slabs);
  if (st) return st;
  st = ensure(env, SL_MISC2, statics.size() * 4 + (size_t)n * 4 + 16, &d_statics);
  if (st) return st;
  CUDA_TRY(cudaMemcpyAsync(d_slabs, slabs.data(), slabs.size() * 4, cudaMemcpyHostToDevice, env->stream));
  CUDA_TRY(cudaMemcpyAsync(d_statics, statics.data(), statics.size() * 4, cudaMemcpyHostToDevice, env->stream));
  int32_t *d_ids = nullptr;
  if (env_ids) {
    d_ids = reinterpret_cast<int32_t *>(reinterpret_cast<uint32_t *>(d_statics) + statics.size());
    CUDA_TRY(cudaMemcpyAsync(d_ids, env_ids, (size_t)n * 4, cudaMemcpyHostToDevice, env->stream));
  }
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_reset(prm, (const uint32_t *)d_slabs, (const uint32_t *)d_statics, d_ids, n, env->stream));
  env->launches++, note_other_work(env);
  CUDA_TRY(cudaStreamSynchronize(env->stream));
  (void)L;
  return GRL_OK;
}

int reset_from_planes(grl_env *env, const int32_t *env_ids, int n, const int32_t *owner, const int32_t *army,
                      const int32_t *type) {
  const GrlLayout &L = env->L;
  const int N = env->N, P = env->cfg.num_players;
  if (env_ids)
    for (int i = 0; i < n; i++)
      if (env_ids[i] < 0 || env_ids[i] >= env->cfg.num_envs) return fail(GRL_ERR_INVALID_ARG, "env id %d out of range", env_ids[i]);
  if (!env_ids && n > env->cfg.num_envs) return fail(GRL_ERR_INVALID_ARG, "n exceeds num_envs");
  const int chunk = 16384;
  std::vector<uint32_t> slabs, statics;
  for (int c0 = 0; c0 < n; c0 += chunk) {
    const int cn = std::min(chunk, n - c0);
    slabs.assign((size_t)cn * L.slab_words, 0u);
    statics.assign((size_t)cn * L.static_words, 0u);
    std::vector<const char *> errs(cn, nullptr);
    parallel_for(cn, cn < 8 ? 1 : env->host_threads, [&](int i) {
      HostGame g;
      const size_t off = (size_t)(c0 + i) * N;
      g.owner.assign(owner + off, owner + off + N);
      g.army.assign(army + off, army + off + N);
      g.type.assign(type + off, type + off + N);
      g.visible.assign(N, 0);
      g.owned.assign((size_t)P * N, 0);
      g.changed.assign(N, 0);
      g.vis_changed.assign(N, 0);
      for (int p = 0; p < P; p++) g.alive[p] = 1, g.army_count[p] = 0, g.general_idx[p] = -1;
      errs[i] = pack_slab(L, g, slabs.data() + (size_t)i * L.slab_words, statics.data() + (size_t)i * L.static_words);
    });
    for (int i = 0; i < cn; i++)
      if (errs[i]) return fail(GRL_ERR_INVALID_ARG, "board %d: %s", c0 + i, errs[i]);
    if (!env_ids && c0 != 0) {
      // contiguous ids beyond the first chunk: materialise them
      std::vector<int32_t> ids(cn);
      for (int i = 0; i < cn; i++) ids[i] = c0 + i;
      int st = upload_and_reset(env, ids.data(), cn, slabs, statics);
      if (st) return st;
    } else {
      int st = upload_and_reset(env, env_ids ? env_ids + c0 : nullptr, cn, slabs, statics);
      if (st) return st;
    }
  }
  return GRL_OK;
}

// copy a device result into a caller buffer that may live on the host
struct OutBuf {
  void *user = nullptr;
  void *dev = nullptr;
  size_t bytes = 0;
  bool staged = false;
};

int bind_out(grl_env *env, Slot slot, void *user, size_t bytes, OutBuf &ob) {
  ob.user = user;
  ob.bytes = bytes;
  if (!user) {
    ob.dev = nullptr;
    return GRL_OK;
  }
  if (is_device_ptr(user)) {
    ob.dev = user;
    return GRL_OK;
  }
  ob.staged = true;
  return ensure(env, slot, bytes, &ob.dev);
}

int flush_out(grl_env *env, OutBuf &ob, bool &need_sync) {
  if (ob.staged) {
    CUDA_TRY(cudaMemcpyAsync(ob.user, ob.dev, ob.bytes, cudaMemcpyDeviceToHost, env->stream));
    need_sync = true;
  }
  return GRL_OK;
}

int ensure_pipe(grl_env *env) {
  if (env->ev_start) return GRL_OK;
  CUDA_TRY(cudaEventCreateWithFlags(&env->ev_start, cudaEventDisableTiming));
  for (int k = 0; k < grl_env::kPipe; k++) {
    CUDA_TRY(cudaStreamCreateWithFlags(&env->pipe[k], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&env->ev_done[k], cudaEventDisableTiming));
  }
  return GRL_OK;
}

int run_turn(grl_env *env, const grl_action *actions, uint32_t flags, uint64_t policy_seed, const grl_step_outputs *out,
             bool do_step) {
  const grl_config &c = env->cfg;
  CUDA_TRY(cudaSetDevice(c.device));
  const size_t B = (size_t)c.num_envs, P = (size_t)c.num_players, N = (size_t)env->N;
  GrlKParams prm = base_params(env);
  prm.flags = flags;
  prm.policy_seed = policy_seed;
  const size_t act_stride = (size_t)c.max_actions * sizeof(grl_action);
  bool actions_staged = false, zero_copied = false;
  if (do_step && actions && !(flags & GRL_STEP_FLAG_RANDOM_POLICY)) {
    if (is_device_ptr(actions)) {
      prm.actions = actions;
    } else {
      void *d = nullptr;
      int st = ensure(env, SL_ACTIONS, B * act_stride, &d);
      if (st) return st;
      prm.actions = d;
      actions_staged = true;  // the caller's buffer must be consumed before we return
    }
  }
  const bool do_out = out != nullptr;
  struct Plane {
    OutBuf ob;
    size_t stride;  // bytes per env
  } planes[8];
  int n_planes = 0;
  bool any_staged = actions_staged;
  if (do_out) {
    const size_t words = (4 * N + 31) / 32;
    const Slot slots[8] = {SL_OBS, SL_MASK, SL_REWARD, SL_DONE, SL_WINNER, SL_ERR, SL_AIDX, SL_PACKED};
    void *user[8] = {out->obs, out->mask_bits, out->reward, out->done, out->winner, out->step_error, out->action_index,
                     out->obs_packed};
    const size_t strides[8] = {P * GRL_OBS_CHANNELS * N * 4, P * words * 4, P * 4, 1, 1, 1, P * 4,
                               (size_t)grl_packed_words(env->L) * 4};
    for (int i = 0; i < 8; i++) {
      planes[i].stride = strides[i];
      // the small result planes (reward, done, winner, step_error, action_index) are written in place when the caller's
      // buffers are pinned host memory: no D2H copy (e2e +5 %; writing the observation planes in place cost 10 %)
      void *alias = (i >= 2 && i <= 6) ? pinned_device_alias(user[i]) : nullptr;
      if (alias) {
        planes[i].ob.user = user[i];
        planes[i].ob.dev = alias;
        planes[i].ob.bytes = B * strides[i];
        zero_copied = true;
        continue;
      }
      int st = bind_out(env, slots[i], user[i], B * strides[i], planes[i].ob);
      if (st) return st;
      any_staged = any_staged || planes[i].ob.staged;
    }
    n_planes = 8;
    prm.obs = (float *)planes[0].ob.dev;
    prm.mask_bits = (uint32_t *)planes[1].ob.dev;
    prm.reward = (float *)planes[2].ob.dev;
    prm.done = (uint8_t *)planes[3].ob.dev;
    prm.winner = (int8_t *)planes[4].ob.dev;
    prm.step_error = (uint8_t *)planes[5].ob.dev;
    prm.action_index = (int32_t *)planes[6].ob.dev;
    prm.obs_packed = (uint32_t *)planes[7].ob.dev;
  }
  if (!do_step && !do_out) return GRL_OK;

  if (!any_staged) {  // device (or device-addressable) buffers only: one launch on the env's stream
    // it may overlap the previous turn launch when that is still the last thing this library put on the stream (whatever
    // else the caller enqueued in between is serialised by CUDA as always); never while the stream is being captured:
    // a replayed graph would carry stale sequence numbers
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(env->stream, &cap) != cudaSuccess) cudaGetLastError(), cap = cudaStreamCaptureStatusActive;
    // A launch PUBLISHES its epoch words (a release fence at the end of every warp: 1-4 % of an isolated launch) only
    // inside a chain: the stream still has work and the last thing this library enqueued was a turn launch.  The first
    // launch of a chain runs as always, the second publishes, from the third on they overlap their predecessor.
    bool chain = false;
    if (cap == cudaStreamCaptureStatusNone && env->overlap && !zero_copied && last_on_stream_is(env, env->overlap_prev)) {
      const cudaError_t q = cudaStreamQuery(env->stream);
      if (q == cudaErrorNotReady) cudaGetLastError(), chain = true;
      else if (q != cudaSuccess) return fail(GRL_ERR_CUDA, "cudaStreamQuery: %s", cudaGetErrorString(q));
    }
    const uint32_t seq = ++env->epoch_seq ? env->epoch_seq : ++env->epoch_seq;  // never 0
    prm.epoch = chain ? env->d_epoch : nullptr;
    prm.epoch_seq = seq;
    prm.epoch_need = (chain && env->overlap_published) ? env->overlap_prev : 0u;
    CUDA_TRY(grl_launch_turn(prm, do_step, do_out, env->stream));
    env->launches++;
    note_turn_launch(env, cap == cudaStreamCaptureStatusNone ? seq : 0u);
    env->overlap_published = chain;
    if (zero_copied) CUDA_TRY(cudaStreamSynchronize(env->stream));  // host buffers are valid / consumed on return
    return GRL_OK;
  }

  // Host buffers: pipeline sub-ranges of the batch over side streams — H2D(actions) -> kernel ->
  // D2H(results) per sub-range — so PCIe copies overlap the other sub-ranges' kernels.
  int chunks = (B >= 8192 && env->pipe_chunks > 1) ? env->pipe_chunks : 1;
  if (chunks > grl_env::kPipe) chunks = grl_env::kPipe;
  if (chunks > 1) {
    int st = ensure_pipe(env);
    if (st) return st;
    CUDA_TRY(cudaEventRecord(env->ev_start, env->stream));
  }
  const size_t per = ((B + chunks - 1) / chunks + 7) & ~(size_t)7;
  void *first_alias = actions_staged ? pinned_device_alias(actions) : nullptr;
  for (int k = 0; k < chunks; k++) {
    const size_t g0 = std::min(B, (size_t)k * per), g1 = std::min(B, g0 + per);
    if (g0 >= g1) continue;
    cudaStream_t sq = chunks > 1 ? env->pipe[k] : env->stream;
    if (chunks > 1) CUDA_TRY(cudaStreamWaitEvent(sq, env->ev_start, 0));
    // the first sub-range's kernel reads its actions in place from pinned host memory, so it starts without waiting for
    // a copy (the other sub-ranges' copies overlap it)
    const bool first_in_place = actions_staged && k == 0 && chunks > 1 && first_alias != nullptr;
    if (actions_staged && !first_in_place)
      CUDA_TRY(cudaMemcpyAsync((char *)const_cast<void *>(prm.actions) + g0 * act_stride, (const char *)actions + g0 * act_stride,
                               (g1 - g0) * act_stride, cudaMemcpyHostToDevice, sq));
    GrlKParams pk = prm;
    if (first_in_place) pk.actions = first_alias;
    pk.game0 = (int)g0;
    pk.game_end = (int)g1;
    CUDA_TRY(grl_launch_turn(pk, do_step, do_out, sq));
    env->launches++, note_other_work(env);
    for (int i = 0; i < n_planes; i++) {
      const OutBuf &ob = planes[i].ob;
      if (ob.staged)
        CUDA_TRY(cudaMemcpyAsync((char *)ob.user + g0 * planes[i].stride, (const char *)ob.dev + g0 * planes[i].stride,
                                 (g1 - g0) * planes[i].stride, cudaMemcpyDeviceToHost, sq));
    }
    if (chunks > 1) {
      CUDA_TRY(cudaEventRecord(env->ev_done[k], sq));
      CUDA_TRY(cudaStreamWaitEvent(env->stream, env->ev_done[k], 0));
    }
  }
  CUDA_TRY(cudaStreamSynchronize(env->stream));  // host buffers are valid (and consumed) on return
  return GRL_OK;
}

}  // namespace

extern "C" {

int grl_abi_version(void) { return GRL_ABI_VERSION; }

const char *grl_status_string(int s) {
  switch (s) {
    case GRL_OK: return "ok";
    case GRL_ERR_INVALID_ARG: return "invalid argument";
    case GRL_ERR_CUDA: return "cuda error";
    case GRL_ERR_NOMEM: return "out of memory";
    case GRL_ERR_MAPGEN: return "map generation failed";
    case GRL_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown";
  }
}

const char *grl_last_error(void) { return g_err; }

int grl_default_config(grl_config *c) {
  if (!c) return fail(GRL_ERR_INVALID_ARG, "null config");
  memset(c, 0, sizeof(*c));
  c->num_envs = 1;
  c->width = 20;
  c->height = 20;
  c->num_players = 2;
  c->max_actions = 2;
  c->fog_of_war = 1;  // engine_initializer.go:118
  c->city_ratio = 20; // internal/config/config.go:198-209
  c->city_start_army = 40;
  c->min_general_spacing = 5;
  c->production_general = 1;
  c->production_city = 1;
  c->production_normal = 1;
  c->normal_growth_interval = 25;
  grl_reward_config &r = c->reward;  // experience/rewards.go:23-37
  r.win_game = 1.0f;
  r.lose_game = -1.0f;
  r.capture_city = 0.1f;
  r.lose_city = -0.1f;
  r.capture_general = 0.5f;
  r.lose_general = -0.5f;
  r.territory_gained = 0.01f;
  r.territory_lost = -0.01f;
  r.army_gained = 0.001f;
  r.army_lost = -0.001f;
  r.army_advantage = 0.05f;
  return GRL_OK;
}

int grl_create(const grl_config *cfg, grl_env **out) {
  int st = check_config(cfg);
  if (st) return st;
  if (!out) return fail(GRL_ERR_INVALID_ARG, "null out");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(GRL_ERR_INVALID_ARG, "device %d of %d", cfg->device, ndev);
  CUDA_TRY(cudaSetDevice(cfg->device));
  grl_env *env = new grl_env();
  env->cfg = *cfg;
  env->N = cfg->width * cfg->height;
  env->L = grl_make_layout(cfg->width, cfg->height, cfg->num_players);
  int hw = (int)std::thread::hardware_concurrency();
  env->host_threads = cfg->host_threads > 0 ? cfg->host_threads : (hw > 0 ? hw : 1);
  const char *hm = getenv("GRL_HOST_MAPGEN");
  env->host_mapgen = (hm && hm[0] == '1') ? 1 : 0;
  const char *pc = getenv("GRL_PIPE_CHUNKS");
  if (pc && atoi(pc) >= 1) env->pipe_chunks = atoi(pc);
  const char *pf = getenv("GRL_PREFETCH_DIST");
  env->prefetch_dist = pf ? atoi(pf) : 8192;  // ~1.7 waves of resident warps ahead (profiles/r1_variants.md)
  const char *ov = getenv("GRL_LAUNCH_OVERLAP");
  if (ov) env->overlap = atoi(ov) != 0;
  auto bail = [&](int code) {
    grl_destroy(env);
    return code;
  };
  if (cudaStreamCreateWithFlags(&env->own_stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(GRL_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(cudaGetLastError())));
  env->stream = env->own_stream;
  const size_t sbytes = (size_t)cfg->num_envs * env->L.slab_words * 4;
  const size_t tbytes = (size_t)cfg->num_envs * env->L.static_words * 4;
  // one allocation [state | terrain | geometry]
  const size_t sbytes_al = (sbytes + 255) & ~(size_t)255, tbytes_al = (tbytes + 255) & ~(size_t)255;
  if (cudaMalloc((void **)&env->d_state, sbytes_al + tbytes_al + 96 * 4) != cudaSuccess)
    return bail(fail(GRL_ERR_NOMEM, "cudaMalloc of %zu state bytes: %s", sbytes + tbytes, cudaGetErrorString(cudaGetLastError())));
  if (cudaMalloc((void **)&env->d_epoch, (size_t)cfg->num_envs * 4) != cudaSuccess ||
      cudaMemsetAsync(env->d_epoch, 0, (size_t)cfg->num_envs * 4, env->stream) != cudaSuccess)
    return bail(fail(GRL_ERR_NOMEM, "cudaMalloc of the epoch words: %s", cudaGetErrorString(cudaGetLastError())));
  env->d_static = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(env->d_state) + sbytes_al);
  env->d_geom = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(env->d_static) + tbytes_al);
  uint32_t geom[96];
  make_geom(cfg->width, cfg->height, geom);
  if (cudaMemsetAsync(env->d_state, 0, sbytes, env->stream) != cudaSuccess ||
      cudaMemsetAsync(env->d_static, 0, tbytes, env->stream) != cudaSuccess ||
      cudaMemcpyAsync(env->d_geom, geom, sizeof geom, cudaMemcpyHostToDevice, env->stream) != cudaSuccess)
    return bail(fail(GRL_ERR_CUDA, "state init: %s", cudaGetErrorString(cudaGetLastError())));
  GrlKParams prm = base_params(env);
  if (grl_launch_mark_over(prm, env->stream) != cudaSuccess || cudaStreamSynchronize(env->stream) != cudaSuccess)
    return bail(fail(GRL_ERR_CUDA, "state init kernel: %s", cudaGetErrorString(cudaGetLastError())));
  env->launches++, note_other_work(env);
  *out = env;
  return GRL_OK;
}

int grl_destroy(grl_env *env) {
  if (!env) return GRL_OK;
  cudaSetDevice(env->cfg.device);
  if (env->stream) cudaStreamSynchronize(env->stream);
  for (auto &s : env->scratch)
    if (s.ptr) cudaFree(s.ptr);
  if (env->d_logtab) cudaFree(env->d_logtab);
  if (env->d_state) cudaFree(env->d_state);  // d_static and d_geom live in the same allocation
  if (env->d_epoch) cudaFree(env->d_epoch);
  {
    std::lock_guard<std::mutex> lk(g_overlap_mu);
    for (auto it = g_last_turn_launch.begin(); it != g_last_turn_launch.end();)
      it = (it->second.first == env || it->first == env->own_stream) ? g_last_turn_launch.erase(it) : std::next(it);
  }
  for (int k = 0; k < grl_env::kPipe; k++) {
    if (env->pipe[k]) cudaStreamDestroy(env->pipe[k]);
    if (env->ev_done[k]) cudaEventDestroy(env->ev_done[k]);
  }
  if (env->ev_start) cudaEventDestroy(env->ev_start);
  if (env->own_stream) cudaStreamDestroy(env->own_stream);
  delete env;
  return GRL_OK;
}

int grl_sync(grl_env *env) {
  if (!env) return fail(GRL_ERR_INVALID_ARG, "null env");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_set_stream(grl_env *env, void *cuda_stream) {
  if (!env) return fail(GRL_ERR_INVALID_ARG, "null env");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(env->stream));  // work already issued stays ordered
  env->stream = static_cast<cudaStream_t>(cuda_stream);  // NULL is CUDA's default stream
  note_other_work(env);
  return GRL_OK;
}

int grl_get_config(const grl_env *env, grl_config *out) {
  if (!env || !out) return fail(GRL_ERR_INVALID_ARG, "null argument");
  *out = env->cfg;
  return GRL_OK;
}

int grl_mapgen(const grl_config *cfg, int64_t seed, int32_t *owner, int32_t *army, int32_t *type) {
  int st = check_config(cfg);
  if (st) return st;
  if (!owner || !army || !type) return fail(GRL_ERR_INVALID_ARG, "null plane");
  grl::MapParams mp = grl::DefaultMapParams(cfg->width, cfg->height, cfg->num_players, cfg->city_ratio,
                                            cfg->city_start_army, cfg->min_general_spacing);
  if (!grl::GenerateMap(cfg->width, cfg->height, mp, seed, owner, army, type))
    return fail(GRL_ERR_MAPGEN, "unable to place a general (seed %lld)", (long long)seed);
  return GRL_OK;
}

// Seeded reset with the maps generated on the device (grl_mapgen_gpu.cu): seeds up, one mapgen
// thread per map into zero-filled staging slabs, then the turn-0 set-up kernel.  No host mapgen,
// no slab upload.
static int reset_seeded_device(grl_env *env, const int32_t *env_ids, int32_t n, const int64_t *seeds) {
  const grl_config &c = env->cfg;
  const GrlLayout &L = env->L;
  if (env_ids)
    for (int i = 0; i < n; i++)
      if (env_ids[i] < 0 || env_ids[i] >= c.num_envs) return fail(GRL_ERR_INVALID_ARG, "env id %d out of range", env_ids[i]);
  if (!env_ids && n > c.num_envs) return fail(GRL_ERR_INVALID_ARG, "n exceeds num_envs");
  grl::MapParams hp = grl::DefaultMapParams(c.width, c.height, c.num_players, c.city_ratio, c.city_start_army, c.min_general_spacing);
  GrlMapParams mp = {hp.players, hp.city_ratio, hp.city_start_army, hp.spacing, hp.veins, hp.min_vein, hp.max_vein};
  const int chunk = 262144;
  for (int c0 = 0; c0 < n; c0 += chunk) {
    const int cn = std::min(chunk, n - c0);
    void *d_slabs = nullptr, *d_statics = nullptr, *d_seeds = nullptr;
    int st;
    const size_t slab_bytes = (size_t)cn * L.slab_words * 4, stat_bytes = (size_t)cn * L.static_words * 4;
    if ((st = ensure(env, SL_MISC, slab_bytes + 16, &d_slabs))) return st;
    if ((st = ensure(env, SL_MISC2, stat_bytes + (size_t)cn * 4 + 16, &d_statics))) return st;
    if ((st = ensure(env, SL_ACTIONS, (size_t)cn * 8 + 16, &d_seeds))) return st;
    int *d_failed = reinterpret_cast<int *>(reinterpret_cast<char *>(d_seeds) + (size_t)cn * 8);
    CUDA_TRY(cudaMemsetAsync(d_failed, 0, 4, env->stream));
    CUDA_TRY(cudaMemcpyAsync(d_seeds, seeds + c0, (size_t)cn * 8, cudaMemcpyHostToDevice, env->stream));
    int32_t *d_ids = nullptr;
    std::vector<int32_t> ids;
    if (env_ids || c0 != 0) {
      const int32_t *src = env_ids ? env_ids + c0 : nullptr;
      if (!src) {
        ids.resize(cn);
        for (int i = 0; i < cn; i++) ids[i] = c0 + i;
        src = ids.data();
      }
      d_ids = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(d_statics) + stat_bytes);
      CUDA_TRY(cudaMemcpyAsync(d_ids, src, (size_t)cn * 4, cudaMemcpyHostToDevice, env->stream));
    }
    CUDA_TRY(grl_launch_mapgen(L, c.width, c.height, mp, (const long long *)d_seeds, cn, (uint32_t *)d_slabs, (uint32_t *)d_statics,
                               d_failed, env->stream));
    GrlKParams prm = base_params(env);
    CUDA_TRY(grl_launch_reset(prm, (const uint32_t *)d_slabs, (const uint32_t *)d_statics, d_ids, cn, env->stream));
    env->launches += 2, note_other_work(env);
    int failed = 0;
    CUDA_TRY(cudaMemcpyAsync(&failed, d_failed, 4, cudaMemcpyDeviceToHost, env->stream));
    CUDA_TRY(cudaStreamSynchronize(env->stream));
    if (failed) return fail(GRL_ERR_MAPGEN, "unable to place a general (seed %lld)", (long long)seeds[c0 + failed - 1]);
  }
  return GRL_OK;
}

int grl_reset_seeded(grl_env *env, const int32_t *env_ids, int32_t n, const int64_t *seeds) {
  if (!env || !seeds || n < 0) return fail(GRL_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return GRL_OK;
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  // maps are generated on the device unless the batch is tiny or GRL_HOST_MAPGEN=1 asks for the host path.  (One device
  // thread per map takes ~0.2 ms whatever n is; the host generator ~30 us per map, plus ~50 us per worker thread it
  // spawns — a vector env re-seeds a few dozen envs per step, so that path has to stay out of thread-spawn territory.)
  if (!env->host_mapgen && n >= 8) return reset_seeded_device(env, env_ids, n, seeds);
  const grl_config &c = env->cfg;
  const int N = env->N;
  grl::MapParams mp = grl::DefaultMapParams(c.width, c.height, c.num_players, c.city_ratio, c.city_start_army, c.min_general_spacing);
  std::vector<int32_t> owner((size_t)n * N), army((size_t)n * N), type((size_t)n * N);
  std::vector<uint8_t> ok(n, 1);
  parallel_for(n, n < 8 ? 1 : env->host_threads, [&](int i) {
    ok[i] = grl::GenerateMap(c.width, c.height, mp, seeds[i], owner.data() + (size_t)i * N, army.data() + (size_t)i * N,
                             type.data() + (size_t)i * N)
                ? 1
                : 0;
  });
  for (int i = 0; i < n; i++)
    if (!ok[i]) return fail(GRL_ERR_MAPGEN, "unable to place a general (seed %lld)", (long long)seeds[i]);
  return reset_from_planes(env, env_ids, n, owner.data(), army.data(), type.data());
}

int grl_reset_boards(grl_env *env, const int32_t *env_ids, int32_t n, const int32_t *owner, const int32_t *army,
                     const int32_t *type) {
  if (!env || !owner || !army || !type || n < 0) return fail(GRL_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return GRL_OK;
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  return reset_from_planes(env, env_ids, n, owner, army, type);
}

int grl_step(grl_env *env, const grl_action *actions, uint32_t flags, uint64_t policy_seed) {
  if (!env) return fail(GRL_ERR_INVALID_ARG, "null env");
  return run_turn(env, actions, flags, policy_seed, nullptr, true);
}

int grl_step_fused(grl_env *env, const grl_action *actions, uint32_t flags, uint64_t policy_seed,
                   const grl_step_outputs *out) {
  if (!env) return fail(GRL_ERR_INVALID_ARG, "null env");
  return run_turn(env, actions, flags, policy_seed, out, true);
}

int grl_observe(grl_env *env, const grl_step_outputs *out) {
  if (!env || !out) return fail(GRL_ERR_INVALID_ARG, "null argument");
  return run_turn(env, nullptr, 0, 0, out, false);
}

int grl_mask(grl_env *env, int variant, void *out) {
  if (!env || !out) return fail(GRL_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  const grl_config &c = env->cfg;
  const size_t B = (size_t)c.num_envs, P = (size_t)c.num_players, N = (size_t)env->N;
  const size_t words = (4 * N + 31) / 32;
  bool need_sync = false;
  GrlKParams prm = base_params(env);
  OutBuf ob;
  int st;
  switch (variant) {
    case GRL_MASK_ENGINE_URDL:
    case GRL_MASK_SERIALIZER_UDLR:
      if ((st = bind_out(env, SL_MISC, out, B * P * N * 4, ob))) return st;
      CUDA_TRY(grl_launch_mask_bytes(prm, variant == GRL_MASK_ENGINE_URDL ? 0 : 1, (uint8_t *)ob.dev, env->stream));
      env->launches++, note_other_work(env);
      break;
    case GRL_MASK_ENGINE_URDL_BITS:
      if ((st = bind_out(env, SL_MASK, out, B * P * words * 4, ob))) return st;
      prm.mask_bits = (uint32_t *)ob.dev;
      CUDA_TRY(grl_launch_turn(prm, false, true, env->stream));
      env->launches++, note_other_work(env);
      break;
    case GRL_MASK_ENGINE_HALF_BITS: {
      void *tmp = nullptr;
      if ((st = ensure(env, SL_MASK, B * P * words * 4, &tmp))) return st;
      if ((st = bind_out(env, SL_MISC, out, B * P * words * 4 * 2, ob))) return st;
      prm.mask_bits = (uint32_t *)tmp;
      CUDA_TRY(grl_launch_turn(prm, false, true, env->stream));
      CUDA_TRY(grl_launch_mask_replicate((const uint32_t *)tmp, (uint32_t *)ob.dev, B * P, (int)words, 2, env->stream));
      env->launches += 2, note_other_work(env);
      break;
    }
    default:
      return fail(GRL_ERR_INVALID_ARG, "unknown mask variant %d", variant);
  }
  if ((st = flush_out(env, ob, need_sync))) return st;
  if (need_sync) CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_visibility(grl_env *env, uint8_t *visible, uint8_t *fog) {
  if (!env) return fail(GRL_ERR_INVALID_ARG, "null env");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  const grl_config &c = env->cfg;
  const size_t bytes = (size_t)c.num_envs * c.num_players * env->N;
  OutBuf v, f;
  int st;
  if ((st = bind_out(env, SL_MISC, visible, bytes, v))) return st;
  if ((st = bind_out(env, SL_MISC2, fog, bytes, f))) return st;
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_visibility(prm, (uint8_t *)v.dev, (uint8_t *)f.dev, env->stream));
  env->launches++, note_other_work(env);
  bool need_sync = false;
  if ((st = flush_out(env, v, need_sync))) return st;
  if ((st = flush_out(env, f, need_sync))) return st;
  if (need_sync) CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

static int ensure_logtab(grl_env *env) {
  if (env->d_logtab) return GRL_OK;
  // the client computes np.log(army + 1) / 10.0 in float64 and stores it into a float32 array
  std::vector<float> tab(65536);
  for (int a = 0; a < 65536; a++) tab[a] = (float)(std::log((double)a + 1.0) / 10.0);
  if (cudaMalloc((void **)&env->d_logtab, tab.size() * 4) != cudaSuccess)
    return fail(GRL_ERR_NOMEM, "cudaMalloc of the log table: %s", cudaGetErrorString(cudaGetLastError()));
  CUDA_TRY(cudaMemcpyAsync(env->d_logtab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, env->stream));
  CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_gym_observe(grl_env *env, int32_t max_turns, const grl_gym_outputs *out) {
  if (!env || !out || max_turns < 1) return fail(GRL_ERR_INVALID_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  const grl_config &c = env->cfg;
  const size_t B = (size_t)c.num_envs, P = (size_t)c.num_players, N = (size_t)env->N;
  int st;
  if ((st = ensure_logtab(env))) return st;
  OutBuf obs, mask, stats;
  if ((st = bind_out(env, SL_OBS, out->obs, B * P * GRL_GYM_CHANNELS * N * 4, obs))) return st;
  if ((st = bind_out(env, SL_MISC, out->mask, B * P * N * 5, mask))) return st;
  if ((st = bind_out(env, SL_MISC2, out->stats, B * P * 4 * 4, stats))) return st;
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_gym(prm, max_turns, env->d_logtab, (float *)obs.dev, (uint8_t *)mask.dev, (int32_t *)stats.dev, env->stream));
  env->launches++, note_other_work(env);
  bool need_sync = false;
  if ((st = flush_out(env, obs, need_sync))) return st;
  if ((st = flush_out(env, mask, need_sync))) return st;
  if ((st = flush_out(env, stats, need_sync))) return st;
  if (need_sync) CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_gym_observe_envs(grl_env *env, int32_t max_turns, const int32_t *env_ids, int32_t n, const grl_gym_outputs *out) {
  if (!env || !out || !env_ids || max_turns < 1 || n < 0) return fail(GRL_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return GRL_OK;
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  const grl_config &c = env->cfg;
  for (int i = 0; i < n; i++)
    if (env_ids[i] < 0 || env_ids[i] >= c.num_envs) return fail(GRL_ERR_INVALID_ARG, "env id %d out of range", env_ids[i]);
  // the planes are written in place at the listed envs' rows: they have to be device memory
  if ((out->obs && !is_device_ptr(out->obs)) || (out->mask && !is_device_ptr(out->mask)) || (out->stats && !is_device_ptr(out->stats)))
    return fail(GRL_ERR_UNSUPPORTED, "grl_gym_observe_envs takes device pointers for the planes");
  int st;
  if ((st = ensure_logtab(env))) return st;
  void *d_ids = nullptr;
  if ((st = ensure(env, SL_AIDX, (size_t)n * 4, &d_ids))) return st;
  CUDA_TRY(cudaMemcpyAsync(d_ids, env_ids, (size_t)n * 4, cudaMemcpyHostToDevice, env->stream));
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_gym(prm, max_turns, env->d_logtab, out->obs, out->mask, out->stats, env->stream, (const int32_t *)d_ids, n));
  env->launches++, note_other_work(env);
  CUDA_TRY(cudaStreamSynchronize(env->stream));  // env_ids (pageable host memory) is consumed on return
  return GRL_OK;
}

int grl_gym_encode(grl_env *env, const int64_t *action_idx, int32_t player, int32_t slot, const uint8_t *mask,
                   int32_t skip_invalid, grl_action *actions, uint8_t *valid) {
  if (!env || !action_idx || !mask || !actions) return fail(GRL_ERR_INVALID_ARG, "null argument");
  const grl_config &c = env->cfg;
  if (player < 0 || player >= c.num_players || slot < 0 || slot >= c.max_actions)
    return fail(GRL_ERR_INVALID_ARG, "player %d / slot %d out of range", player, slot);
  CUDA_TRY(cudaSetDevice(c.device));
  if (!is_device_ptr(action_idx) || !is_device_ptr(mask) || !is_device_ptr(actions) || (valid && !is_device_ptr(valid)))
    return fail(GRL_ERR_UNSUPPORTED, "grl_gym_encode takes device pointers (it is the device-side glue of the vector env)");
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_gym_encode(prm, (const long long *)action_idx, player, slot, mask, skip_invalid, actions, valid, env->stream));
  env->launches++, note_other_work(env);
  return GRL_OK;
}

int grl_gym_autoreset(grl_env *env, int32_t max_turns, int64_t base_seed, const grl_gym_autoreset_io *io) {
  if (!env || !io || max_turns < 1) return fail(GRL_ERR_INVALID_ARG, "bad argument");
  const grl_config &c = env->cfg;
  const GrlLayout &L = env->L;
  CUDA_TRY(cudaSetDevice(c.device));
  const void *need[] = {io->terminated, io->truncated, io->episode, io->turns, io->calls};
  for (const void *p : need)
    if (!p || !is_device_ptr(p)) return fail(GRL_ERR_UNSUPPORTED, "grl_gym_autoreset takes device pointers for every plane");
  const void *opt[] = {io->out.obs, io->out.mask, io->out.stats, io->final_obs, io->n_reset};
  for (const void *p : opt)
    if (p && !is_device_ptr(p)) return fail(GRL_ERR_UNSUPPORTED, "grl_gym_autoreset takes device pointers for every plane");
  if (io->final_obs && !io->out.obs) return fail(GRL_ERR_INVALID_ARG, "final_obs needs the observation plane");
  int st;
  if ((st = ensure_logtab(env))) return st;
  const int B = c.num_envs;
  // staging for up to B new games (all envs end together when their episodes started together)
  void *d_slabs = nullptr, *d_statics = nullptr, *d_misc = nullptr;
  if ((st = ensure(env, SL_MISC, (size_t)B * L.slab_words * 4 + 16, &d_slabs))) return st;
  if ((st = ensure(env, SL_MISC2, (size_t)B * L.static_words * 4 + 16, &d_statics))) return st;
  if ((st = ensure(env, SL_ACTIONS, (size_t)B * 12 + 64, &d_misc))) return st;
  long long *d_seeds = reinterpret_cast<long long *>(d_misc);
  int32_t *d_ids = reinterpret_cast<int32_t *>(d_seeds + B);
  int *d_count = reinterpret_cast<int *>(d_ids + B);  // [0] count, [1] mapgen failure
  cudaStream_t sq = env->stream;
  GrlKParams prm = base_params(env);
  CUDA_TRY(cudaMemsetAsync(d_count, 0, 8, sq));
  CUDA_TRY(grl_launch_gym_compact(prm, io->terminated, io->truncated, (long long)base_seed, (long long *)io->episode, io->turns, io->calls,
                                  d_ids, d_seeds, d_count, sq));
  // three launches: the finished envs compacted into an id list; their last observations saved and their next maps
  // generated (one warp per env); turn-0 set-up and gym read-outs of the re-seeded envs (one warp per env)
  grl::MapParams hp = grl::DefaultMapParams(c.width, c.height, c.num_players, c.city_ratio, c.city_start_army, c.min_general_spacing);
  GrlMapParams mp = {hp.players, hp.city_ratio, hp.city_start_army, hp.spacing, hp.veins, hp.min_vein, hp.max_vein};
  CUDA_TRY(grl_launch_mapgen(L, c.width, c.height, mp, d_seeds, B, (uint32_t *)d_slabs, (uint32_t *)d_statics, d_count + 1, sq, d_count,
                             d_ids, io->out.obs, io->final_obs, GRL_GYM_CHANNELS * c.width * c.height));
  if (io->out.obs || io->out.mask || io->out.stats)
    CUDA_TRY(grl_launch_gym_reseed(prm, (const uint32_t *)d_slabs, (const uint32_t *)d_statics, d_ids, B, d_count, max_turns, env->d_logtab,
                                   io->out.obs, io->out.mask, io->out.stats, sq));
  else
    CUDA_TRY(grl_launch_reset(prm, (const uint32_t *)d_slabs, (const uint32_t *)d_statics, d_ids, B, sq, d_count));
  if (io->n_reset) CUDA_TRY(cudaMemcpyAsync(io->n_reset, d_count, 4, cudaMemcpyDeviceToDevice, sq));
  env->launches += 3, note_other_work(env);
  return GRL_OK;
}

int grl_replay_push_rows(grl_env *env, const grl_replay_rows_io *io) {
  if (!env || !io || !io->obs) return fail(GRL_ERR_INVALID_ARG, "null argument");
  if (io->capacity < 1 || io->obs_floats < 1 || io->views < 1 || io->view < 0 || io->view >= io->views || io->next_row0 < 0 ||
      io->state_row0 < 0)
    return fail(GRL_ERR_INVALID_ARG, "bad ring geometry");
  if (io->capacity < env->cfg.num_envs)  // two envs of one step would write the same ring row
    return fail(GRL_ERR_INVALID_ARG, "replay capacity %lld below the %d envs of a vector step", (long long)io->capacity, env->cfg.num_envs);
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  const void *ptrs[] = {io->obs, io->final_obs, io->done, io->next_states, io->states};
  for (const void *p : ptrs)
    if (p && !is_device_ptr(p)) return fail(GRL_ERR_UNSUPPORTED, "grl_replay_push_rows takes device pointers");
  if (!io->next_states && !io->states) return GRL_OK;
  CUDA_TRY(grl_launch_replay_rows(io->obs, io->final_obs, io->done, io->next_states, io->states, io->capacity, io->next_row0,
                                  io->state_row0, io->views, io->view, io->obs_floats, env->cfg.num_envs, env->stream));
  env->launches++, note_other_work(env);
  return GRL_OK;
}

int grl_gym_sample(grl_env *env, uint64_t seed, const uint8_t *mask, int32_t player, int64_t *action) {
  if (!env || !mask || !action) return fail(GRL_ERR_INVALID_ARG, "null argument");
  if (player < 0 || player >= env->cfg.num_players) return fail(GRL_ERR_INVALID_ARG, "player %d out of range", player);
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  if (!is_device_ptr(mask) || !is_device_ptr(action)) return fail(GRL_ERR_UNSUPPORTED, "grl_gym_sample takes device pointers");
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_gym_sample(prm, seed, mask, player, (long long *)action, env->stream));
  env->launches++, note_other_work(env);
  return GRL_OK;
}

int grl_gym_step(grl_env *env, int32_t max_turns, uint64_t opponent_seed, const grl_gym_step_io *io) {
  if (!env || !io || max_turns < 1) return fail(GRL_ERR_INVALID_ARG, "bad argument");
  const grl_config &c = env->cfg;
  if (c.num_players < 2 || c.max_actions < 2) return fail(GRL_ERR_INVALID_ARG, "the gym step drives 2 players / 2 action slots");
  // ONE launch: the turn kernel's gym instantiation (io->actions / io->prev_stats are scratch only the oracle uses).
  if (!io->action && c.num_players != 2)  // the turn kernel's in-launch agent is instantiated for the two-player template
    return fail(GRL_ERR_UNSUPPORTED, "grl_gym_step: action == NULL (the in-launch random agent) needs a two-player env, got %d players; "
                                     "use grl_gym_sample + action", c.num_players);
  // action == NULL: the random agent, drawn in the launch (its index goes to sampled_action)
  const void *need[] = {io->action ? (const void *)io->action : (const void *)io->sampled_action, io->out.mask, io->out.stats,
                        io->turns, io->calls, io->reward, io->terminated, io->truncated, io->valid, io->done, io->winner,
                        io->step_error};
  CUDA_TRY(cudaSetDevice(c.device));
  for (const void *p : need)
    if (!p || !is_device_ptr(p)) return fail(GRL_ERR_UNSUPPORTED, "grl_gym_step takes device pointers for every plane");
  if ((io->opponent_action && !is_device_ptr(io->opponent_action)) || (io->out.obs && !is_device_ptr(io->out.obs)) ||
      (io->n_finished && !is_device_ptr(io->n_finished)))
    return fail(GRL_ERR_UNSUPPORTED, "grl_gym_step takes device pointers for every plane");
  cudaStream_t sq = env->stream;
  int st0 = ensure_logtab(env);
  if (st0) return st0;
  GrlGymK gk;
  memset(&gk, 0, sizeof gk);
  gk.action = (const long long *)io->action;
  gk.opponent_action = (const long long *)io->opponent_action;
  gk.logtab = env->d_logtab;
  gk.obs = io->out.obs;
  gk.mask = io->out.mask;
  gk.stats = io->out.stats;
  gk.turns = io->turns;
  gk.calls = io->calls;
  gk.reward = io->reward;
  gk.terminated = io->terminated;
  gk.truncated = io->truncated;
  gk.valid = io->valid;
  gk.n_finished = io->n_finished;
  gk.opponent_seed = opponent_seed;
  gk.max_turns = max_turns;
  gk.agent_seed = io->agent_seed;
  gk.sampled_action = io->action ? nullptr : (long long *)io->sampled_action;
  GrlKParams pt = base_params(env);
  pt.actions = nullptr;
  pt.done = io->done;
  pt.winner = io->winner;
  pt.step_error = io->step_error;
  if (io->n_finished) CUDA_TRY(cudaMemsetAsync(io->n_finished, 0, 4, sq));
  CUDA_TRY(grl_launch_gym_step(pt, gk, sq));
  env->launches += 1, note_other_work(env);
  return GRL_OK;
}

int grl_sample_actions(grl_env *env, uint64_t policy_seed, grl_action *actions) {
  if (!env || !actions) return fail(GRL_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  const grl_config &c = env->cfg;
  OutBuf ob;
  int st;
  if ((st = bind_out(env, SL_ACTIONS, actions, (size_t)c.num_envs * c.max_actions * sizeof(grl_action), ob))) return st;
  GrlKParams prm = base_params(env);
  prm.policy_seed = policy_seed;
  CUDA_TRY(grl_launch_sample(prm, ob.dev, env->stream));
  env->launches++, note_other_work(env);
  bool need_sync = false;
  if ((st = flush_out(env, ob, need_sync))) return st;
  if (need_sync) CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

static int fetch_games(grl_env *env, int first, int count, std::vector<HostGame> &games) {
  const GrlLayout &L = env->L;
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  std::vector<uint32_t> slabs((size_t)count * L.slab_words), statics((size_t)count * L.static_words);
  CUDA_TRY(cudaMemcpyAsync(slabs.data(), env->d_state + (size_t)first * L.slab_words, slabs.size() * 4, cudaMemcpyDeviceToHost,
                           env->stream));
  CUDA_TRY(cudaMemcpyAsync(statics.data(), env->d_static + (size_t)first * L.static_words, statics.size() * 4,
                           cudaMemcpyDeviceToHost, env->stream));
  CUDA_TRY(cudaStreamSynchronize(env->stream));
  games.resize(count);
  parallel_for(count, env->host_threads, [&](int i) {
    unpack_slab(L, slabs.data() + (size_t)i * L.slab_words, statics.data() + (size_t)i * L.static_words, games[i]);
  });
  return GRL_OK;
}

int grl_get_state(grl_env *env, int32_t first, int32_t count, const grl_state_planes *o) {
  if (!env || !o || first < 0 || count < 0 || first + count > env->cfg.num_envs) return fail(GRL_ERR_INVALID_ARG, "bad range");
  if (count == 0) return GRL_OK;
  const int N = env->N, P = env->cfg.num_players;
  std::vector<HostGame> games;
  int st = fetch_games(env, first, count, games);
  if (st) return st;
  for (int c = 0; c < count; c++) {
    const HostGame &g = games[c];
    const size_t k = (size_t)c * N;
    if (o->owner) memcpy(o->owner + k, g.owner.data(), (size_t)N * 4);
    if (o->army) memcpy(o->army + k, g.army.data(), (size_t)N * 4);
    if (o->type) memcpy(o->type + k, g.type.data(), (size_t)N * 4);
    if (o->visible) memcpy(o->visible + k, g.visible.data(), (size_t)N * 4);
    if (o->owned) memcpy(o->owned + (size_t)c * P * N, g.owned.data(), (size_t)P * N);
    if (o->changed) memcpy(o->changed + k, g.changed.data(), (size_t)N);
    if (o->vis_changed) memcpy(o->vis_changed + k, g.vis_changed.data(), (size_t)N);
    if (o->turn) o->turn[c] = g.turn;
    if (o->game_over) o->game_over[c] = g.game_over;
    if (o->step_error) o->step_error[c] = g.step_error;
    int n_alive = 0, last = -1;
    for (int p = 0; p < P; p++) {
      if (g.alive[p]) n_alive++, last = p;
      if (o->alive) o->alive[(size_t)c * P + p] = g.alive[p];
      if (o->army_count) o->army_count[(size_t)c * P + p] = g.army_count[p];
      if (o->general_idx) o->general_idx[(size_t)c * P + p] = g.general_idx[p];
    }
    if (o->winner) o->winner[c] = (g.game_over && n_alive == 1) ? last : -1;  // engine.go:248-263
  }
  return GRL_OK;
}

int grl_set_state(grl_env *env, int32_t first, int32_t count, const grl_state_planes *in) {
  if (!env || !in || first < 0 || count < 0 || first + count > env->cfg.num_envs) return fail(GRL_ERR_INVALID_ARG, "bad range");
  if (count == 0) return GRL_OK;
  const GrlLayout &L = env->L;
  const int N = env->N, P = env->cfg.num_players;
  std::vector<HostGame> games;
  int st = fetch_games(env, first, count, games);
  if (st) return st;
  std::vector<uint32_t> slabs((size_t)count * L.slab_words), statics((size_t)count * L.static_words);
  for (int c = 0; c < count; c++) {
    HostGame &g = games[c];
    const size_t k = (size_t)c * N;
    if (in->owner) g.owner.assign(in->owner + k, in->owner + k + N);
    if (in->army) g.army.assign(in->army + k, in->army + k + N);
    if (in->type) g.type.assign(in->type + k, in->type + k + N);
    if (in->visible) g.visible.assign(in->visible + k, in->visible + k + N);
    if (in->owned) g.owned.assign(in->owned + (size_t)c * P * N, in->owned + (size_t)(c + 1) * P * N);
    if (in->changed) g.changed.assign(in->changed + k, in->changed + k + N);
    if (in->vis_changed) g.vis_changed.assign(in->vis_changed + k, in->vis_changed + k + N);
    if (in->turn) g.turn = in->turn[c];
    if (in->game_over) g.game_over = in->game_over[c];
    if (in->step_error) g.step_error = in->step_error[c];
    for (int p = 0; p < P; p++) {
      if (in->alive) g.alive[p] = in->alive[(size_t)c * P + p];
      if (in->army_count) g.army_count[p] = in->army_count[(size_t)c * P + p];
      if (in->general_idx) g.general_idx[p] = in->general_idx[(size_t)c * P + p];
    }
    const char *msg = pack_slab(L, g, slabs.data() + (size_t)c * L.slab_words, statics.data() + (size_t)c * L.static_words);
    if (msg) return fail(GRL_ERR_INVALID_ARG, "env %d: %s", first + c, msg);
  }
  CUDA_TRY(cudaMemcpyAsync(env->d_state + (size_t)first * L.slab_words, slabs.data(), slabs.size() * 4, cudaMemcpyHostToDevice,
                           env->stream));
  CUDA_TRY(cudaMemcpyAsync(env->d_static + (size_t)first * L.static_words, statics.data(), statics.size() * 4,
                           cudaMemcpyHostToDevice, env->stream));
  CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_state_hash(grl_env *env, uint64_t *out) {
  if (!env || !out) return fail(GRL_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  OutBuf ob;
  int st;
  if ((st = bind_out(env, SL_MISC, out, (size_t)env->cfg.num_envs * 8, ob))) return st;
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_state_hash(prm, (uint64_t *)ob.dev, env->stream));
  env->launches++, note_other_work(env);
  bool need_sync = false;
  if ((st = flush_out(env, ob, need_sync))) return st;
  if (need_sync) CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_buffer_hash(grl_env *env, const void *buf, size_t row_words, int32_t rows, uint64_t *out) {
  if (!env || !buf || !out || rows < 0) return fail(GRL_ERR_INVALID_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  const void *dbuf = buf;
  int st;
  if (!is_device_ptr(buf)) {
    void *d = nullptr;
    if ((st = ensure(env, SL_OBS, row_words * (size_t)rows * 4, &d))) return st;
    CUDA_TRY(cudaMemcpyAsync(d, buf, row_words * (size_t)rows * 4, cudaMemcpyHostToDevice, env->stream));
    dbuf = d;
  }
  OutBuf ob;
  if ((st = bind_out(env, SL_MISC, out, (size_t)rows * 8, ob))) return st;
  CUDA_TRY(grl_launch_buffer_hash((const uint32_t *)dbuf, row_words, rows, (uint64_t *)ob.dev, env->stream));
  env->launches++, note_other_work(env);
  bool need_sync = dbuf != buf;
  if ((st = flush_out(env, ob, need_sync))) return st;
  if (need_sync) CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_stats(grl_env *env, uint64_t out[4]) {
  if (!env || !out) return fail(GRL_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(env->cfg.device));
  void *d = nullptr;
  int st = ensure(env, SL_MISC, 64, &d);
  if (st) return st;
  CUDA_TRY(cudaMemsetAsync(d, 0, 32, env->stream));
  GrlKParams prm = base_params(env);
  CUDA_TRY(grl_launch_stats(prm, (unsigned long long *)d, env->stream));
  env->launches++, note_other_work(env);
  CUDA_TRY(cudaMemcpyAsync(out, d, 32, cudaMemcpyDeviceToHost, env->stream));
  CUDA_TRY(cudaStreamSynchronize(env->stream));
  return GRL_OK;
}

int grl_launch_count(grl_env *env, uint64_t *out) {
  if (!env || !out) return fail(GRL_ERR_INVALID_ARG, "null argument");
  *out = env->launches;
  return GRL_OK;
}

}  // extern "C"
