// grl_launch.h — kernel launchers defined in grl_kernels.cu
#pragma once
#include <cuda_runtime.h>

#include "grl_layout.h"

cudaError_t grl_launch_turn(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream);
// per-geometry instantiations (grl_turn_*.cu)
cudaError_t grl_launch_turn_20x20(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream);
cudaError_t grl_launch_turn_15x15(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream);
cudaError_t grl_launch_turn_10x10(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream);
cudaError_t grl_launch_turn_generic(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream);
cudaError_t grl_launch_gym_step_20x20(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream);
cudaError_t grl_launch_gym_step_15x15(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream);
cudaError_t grl_launch_gym_step_10x10(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream);
cudaError_t grl_launch_gym_step_generic(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream);
cudaError_t grl_launch_reset(const GrlKParams &prm, const uint32_t *src_state, const uint32_t *src_static,
                             const int32_t *env_ids, int n, cudaStream_t stream, const int *n_dev = nullptr);
cudaError_t grl_launch_gym_compact(const GrlKParams &prm, const uint8_t *terminated, const uint8_t *truncated, long long base_seed,
                                   long long *episode, int32_t *turns, int32_t *calls, int32_t *ids, long long *seeds, int *count,
                                   cudaStream_t stream);
cudaError_t grl_launch_gym_reseed(const GrlKParams &prm, const uint32_t *src_state, const uint32_t *src_static, const int32_t *ids,
                                  int n, const int *n_dev, int max_turns, const float *logtab, float *obs, uint8_t *mask,
                                  int32_t *stats, cudaStream_t stream);
cudaError_t grl_launch_sample(const GrlKParams &prm, void *out, cudaStream_t stream);
cudaError_t grl_launch_mask_bytes(const GrlKParams &prm, int variant, uint8_t *out, cudaStream_t stream);
cudaError_t grl_launch_visibility(const GrlKParams &prm, uint8_t *visible, uint8_t *fog, cudaStream_t stream);
cudaError_t grl_launch_gym(const GrlKParams &prm, int max_turns, const float *logtab, float *obs, uint8_t *mask, int32_t *stats,
                           cudaStream_t stream, const int32_t *ids = nullptr, int n_ids = 0, const int *n_dev = nullptr);
cudaError_t grl_launch_gym_step(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream);
cudaError_t grl_launch_gym_encode(const GrlKParams &prm, const long long *action_idx, int player, int slot, const uint8_t *mask,
                                  int skip_invalid, void *actions, uint8_t *valid, cudaStream_t stream);
cudaError_t grl_launch_replay_rows(const float *obs, const float *final_obs, const uint8_t *done, float *next_states, float *states,
                                   long long capacity, long long next_row0, long long state_row0, int views, int view, int F, int B,
                                   cudaStream_t stream);
cudaError_t grl_launch_gym_sample(const GrlKParams &prm, unsigned long long seed, const uint8_t *mask, int player, long long *action,
                                  cudaStream_t stream);
cudaError_t grl_launch_mask_replicate(const uint32_t *in, uint32_t *out, size_t rows, int words, int rep,
                                      cudaStream_t stream);
cudaError_t grl_launch_state_hash(const GrlKParams &prm, uint64_t *out, cudaStream_t stream);
cudaError_t grl_launch_buffer_hash(const uint32_t *buf, size_t row_words, int rows, uint64_t *out, cudaStream_t stream);
cudaError_t grl_launch_stats(const GrlKParams &prm, unsigned long long *out, cudaStream_t stream);
cudaError_t grl_launch_mark_over(const GrlKParams &prm, cudaStream_t stream);

// device-side map generation (grl_mapgen_gpu.cu); mirrors grl::MapParams
#define GRL_MAX_PLAYERS_DEV 8
struct GrlMapParams {
  int players, city_ratio, city_start_army, spacing, veins, min_vein, max_vein;
};
cudaError_t grl_launch_mapgen(const GrlLayout &L, int W, int H, const GrlMapParams &mp, const long long *seeds, int n,
                              uint32_t *slabs, uint32_t *statics, int *failed, cudaStream_t stream, const int *n_dev = nullptr,
                              const int32_t *ids = nullptr, const float *obs = nullptr, float *final_obs = nullptr, int obs_block = 0);
