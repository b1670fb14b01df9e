// grl_turn_generic.cu — the turn kernel's instantiations for generic boards (any other W,H <= 32: geometry read from the parameter block, one game per warp).
// One translation unit per geometry so the library builds in parallel.
#include "grl_launch.h"
#include "grl_turn.cuh"

cudaError_t grl_launch_turn_generic(const GrlKParams &prm, bool do_step, bool do_out, cudaStream_t stream) {
  return launch_turn_geo<0, 0, 32>(prm, do_step, do_out, stream);
}

cudaError_t grl_launch_gym_step_generic(const GrlKParams &prm, const GrlGymK &gk, cudaStream_t stream) {
  return launch_gym_geo<0, 0, 32>(prm, gk, stream);
}
