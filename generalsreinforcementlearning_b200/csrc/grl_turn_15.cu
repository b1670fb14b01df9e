// grl_turn_15.cu — the turn kernel's instantiations for the 15x15 board (225 tiles = 8 mask words: eight lanes per game, four games per warp).
// One translation unit per geometry so the library builds in parallel.
#include "grl_launch.h"
#include "grl_turn.cuh"

cudaError_t grl_launch_turn_15x15(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  return launch_turn_geo<15, 15, 8>(prm, do_step, do_out, stream);
}

cudaError_t grl_launch_gym_step_15x15(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  return launch_gym_geo<15, 15, 8>(prm, gk, stream);
}
