// store_batched.cu — does separating the state reads / write-backs from the observation store
// stream IN TIME (GPU-wide phases) recover the DRAM efficiency that interleaving loses?
// Persistent cooperative kernel, 1 CTA (1,024 threads) per SM.  Per batch: every CTA loads the
// slabs of its next NB games into shared memory (read burst) | grid sync | each warp writes its
// games' 28,800-byte observation blocks plane-major | slabs written back in one burst | grid sync.
// Compare with store_readmix mode 3 (interleaved): same bytes.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
namespace cg = cooperative_groups;

constexpr int SLAB_U4 = 91;  // 1,456 B

template <bool SYNC>
__global__ void __launch_bounds__(1024, 1) batched(float4 *out, uint4 *slabs, int games, int nb_games) {
  extern __shared__ uint4 sm[];
  cg::grid_group grid = cg::this_grid();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_cta = (games + gridDim.x - 1) / gridDim.x;
  const int g0 = blockIdx.x * per_cta, g1 = min(games, g0 + per_cta);
  const int P = 2, C = 9, cs = 100;
  const int nbatches = (per_cta + nb_games - 1) / nb_games;
  for (int b = 0; b < nbatches; b++) {
    const int b0 = g0 + b * nb_games, b1 = min(g1, b0 + nb_games);
    const int n = max(0, b1 - b0);
    // read burst: the batch's slabs are contiguous in memory
    const uint4 *src = slabs + (size_t)b0 * SLAB_U4;
    for (int k = threadIdx.x; k < n * SLAB_U4; k += blockDim.x) sm[k] = src[k];
    __syncthreads();
    if (SYNC) grid.sync();
    for (int i = warp; i < n; i += 32) {
      const uint4 *s = sm + i * SLAB_U4;
      uint32_t key = __reduce_or_sync(0xffffffffu, s[lane].x ^ s[32 + lane].y);
      float4 v = make_float4(1.f, 0.f, 1.f, __uint_as_float(key));
      float4 *base = out + (size_t)(b0 + i) * P * C * cs;
      for (int p = 0; p < P; p++)
        for (int ch = 0; ch < C; ch++)
#pragma unroll
          for (int q0 = 0; q0 < 128; q0 += 32) {
            int q = q0 + lane;
            if (q < cs) __stcs(base + (p * C + ch) * cs + q, v);
          }
      if (lane == 0) sm[i * SLAB_U4].x += 1;
    }
    __syncthreads();
    uint4 *dst = slabs + (size_t)b0 * SLAB_U4;
    for (int k = threadIdx.x; k < n * SLAB_U4; k += blockDim.x) dst[k] = sm[k];
    __syncthreads();
    if (SYNC) grid.sync();
  }
}

template <bool SYNC>
float run(float4 *buf, uint4 *slabs, int games, int nb_games, int ctas) {
  size_t smem = (size_t)nb_games * SLAB_U4 * 16;
  cudaFuncSetAttribute(batched<SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  void *args[] = {&buf, &slabs, &games, &nb_games};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) cudaLaunchCooperativeKernel((void *)batched<SYNC>, dim3(ctas), dim3(1024), args, smem, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; i++) cudaLaunchCooperativeKernel((void *)batched<SYNC>, dim3(ctas), dim3(1024), args, smem, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 20;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int games = 65536;
  const size_t bytes = (size_t)games * 28800;
  float4 *buf;
  uint4 *slabs;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&slabs, (size_t)games * 1456);
  cudaMemset(slabs, 0, (size_t)games * 1456);
  printf("%d SMs; interleaved reference: store_readmix mode 3\n", sms);
  for (int nb : {32, 64, 111, 148}) {
    float a = run<false>(buf, slabs, games, nb, sms);
    float b = run<true>(buf, slabs, games, nb, sms);
    printf("batch %3d games/SM  no grid sync %.4f ms %5.0f GB/s   grid sync %.4f ms %5.0f GB/s\n", nb, a, bytes / a / 1e6, b, bytes / b / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
