// grl_turn_20.cu — the turn kernel's instantiations for the 20x20 board (400 tiles = 13 mask words: one game per warp).
// One translation unit per geometry so the library builds in parallel.
#include "grl_launch.h"
#include "grl_turn.cuh"

cudaError_t grl_launch_turn_20x20(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  return launch_turn_geo<20, 20, 32>(prm, do_step, do_out, stream);
}

cudaError_t grl_launch_gym_step_20x20(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  return launch_gym_geo<20, 20, 32>(prm, gk, stream);
}
