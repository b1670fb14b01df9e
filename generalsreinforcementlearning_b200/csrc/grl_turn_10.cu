// grl_turn_10.cu — the turn kernel's instantiations for the 10x10 board (100 tiles = 4 mask words: eight lanes per game measured faster than four (profiles/r1_variants.md)).
// One translation unit per geometry so the library builds in parallel.
#include "grl_launch.h"
#include "grl_turn.cuh"

cudaError_t grl_launch_turn_10x10(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  return launch_turn_geo<10, 10, 8>(prm, do_step, do_out, stream);
}

cudaError_t grl_launch_gym_step_10x10(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  return launch_gym_geo<10, 10, 8>(prm, gk, stream);
}
