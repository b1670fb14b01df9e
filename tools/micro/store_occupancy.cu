// store_occupancy.cu — what does the observation store stream need from the SM?
// 65,536 warps each write a 28,800-byte block plane-major (one linear sweep of 128-bit st.cs),
// as the turn kernel's observation phase does.  Knobs:
//   ctas/SM   occupancy limit (dynamic shared memory padding), 8 warps per CTA
//   preload   each warp first loads its 1,456-byte state slab (dependent: the stored value
//             is derived from it), as the turn kernel must before it can store anything
//   lds       one LDS.128 table lookup in front of every store (the nibble -> float4 table)
// Prints ms per launch and GB/s (profiles/r1_store_occupancy.txt).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <bool PRELOAD, bool LDS>
__global__ void __launch_bounds__(256) writer(float4 *out, const uint4 *slabs, int games) {
  extern __shared__ float4 pad[];
  __shared__ float4 lut[16];
  if (threadIdx.x < 16) lut[threadIdx.x] = make_float4(threadIdx.x & 1, (threadIdx.x >> 1) & 1, (threadIdx.x >> 2) & 1, threadIdx.x >> 3);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int game = blockIdx.x * 8 + warp;
  if (game >= games) return;
  const int P = 2, C = 9, cs = 100;
  float4 *base = out + (size_t)game * P * C * cs;
  uint32_t key = lane;
  if (PRELOAD) {
    const uint4 *s = slabs + (size_t)game * 91;  // 1,456 B
    uint4 a = s[lane], b = s[32 + lane], c = lane < 27 ? s[64 + lane] : make_uint4(0, 0, 0, 0);
    key = a.x ^ b.y ^ c.z;
    key = __reduce_or_sync(0xffffffffu, key);
  }
  float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
  for (int p = 0; p < P; p++)
    for (int c = 0; c < C; c++)
#pragma unroll
      for (int q0 = 0; q0 < 128; q0 += 32) {
        int q = q0 + lane;
        if (LDS) v = lut[(key >> (c + q0 / 8)) & 15];
        else if (PRELOAD) v.x = __uint_as_float(key);
        if (q < cs) __stcs(base + (p * C + c) * cs + q, v);
      }
}

template <bool PRELOAD, bool LDS>
float run(float4 *buf, const uint4 *slabs, int games, int ctas_per_sm) {
  auto k = writer<PRELOAD, LDS>;
  size_t smem = ctas_per_sm >= 8 ? 0 : (size_t)(220 * 1024 / ctas_per_sm) - 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) k<<<games / 8, 256, smem>>>(buf, slabs, games);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; i++) k<<<games / 8, 256, smem>>>(buf, slabs, games);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 20;
}

int main() {
  const int games = 65536;
  const size_t bytes = (size_t)games * 28800;
  float4 *buf;
  uint4 *slabs;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&slabs, (size_t)games * 1456);
  cudaMemset(slabs, 0, (size_t)games * 1456);
  printf("ctas/SM  store-only        +preload          +preload+lds\n");
  for (int c : {2, 3, 4, 5, 6, 8}) {
    float a = run<false, false>(buf, slabs, games, c);
    float b = run<true, false>(buf, slabs, games, c);
    float d = run<true, true>(buf, slabs, games, c);
    printf("%d        %.3f ms %5.0f GB/s  %.3f ms %5.0f GB/s  %.3f ms %5.0f GB/s\n", c, a, bytes / a / 1e6, b, bytes / b / 1e6, d,
           bytes / d / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
