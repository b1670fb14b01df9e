// grl_turn_10.cu — the turn kernel's instantiations for the 10x10 board (100 tiles = 4 mask words: FOUR lanes per game, eight
// games per warp.  Round 1 measured eight lanes faster; with the rest of round 2 in place four lanes are: 0.1016 -> 0.0828 ms per
// 65,536 games, 0.410 -> 0.3525 per 262,144, the gym step 0.155 -> 0.144, profiles/r2_variants.md).
// One translation unit per geometry so the library builds in parallel.
#include "grl_launch.h"
#include "grl_turn.cuh"

cudaError_t grl_launch_turn_10x10(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  return launch_turn_geo<10, 10, 4>(prm, do_step, do_out, stream);
}

cudaError_t grl_launch_gym_step_10x10(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  return launch_gym_geo<10, 10, 4>(prm, gk, stream);
}
