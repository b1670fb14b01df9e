// store_pattern.cu — how fast can 65,536 warps each write a 28,800-byte observation block?
// Pattern 0: the turn kernel's order (tile-chunk major: for chunk { for player { for channel } })
// Pattern 1: plane major (for player { for channel { for chunk } }): 1,600 contiguous bytes per plane
// Pattern 2: one linear sweep of the block
// No compute: this is the ceiling the store ORDER allows (profiles/r1_store_pattern.txt).
#include <cuda_runtime.h>
#include <cstdio>

template <int PAT, bool CS>
__global__ void __launch_bounds__(256) writer(float4 *out, int games) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int game = blockIdx.x * 8 + warp;
  if (game >= games) return;
  const int N = 400, P = 2, C = 9, cs = N / 4;
  float4 *base = out + (size_t)game * P * C * cs;
  const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
  if (PAT == 0) {
    for (int q0 = 0; q0 < cs; q0 += 32) {
      int q = q0 + lane;
      for (int p = 0; p < P; p++)
        for (int c = 0; c < C; c++)
          if (q < cs) { if (CS) __stcs(base + (p * C + c) * cs + q, v); else base[(p * C + c) * cs + q] = v; }
    }
  } else if (PAT == 1) {
    for (int p = 0; p < P; p++)
      for (int c = 0; c < C; c++)
        for (int q0 = 0; q0 < cs; q0 += 32) {
          int q = q0 + lane;
          if (q < cs) { if (CS) __stcs(base + (p * C + c) * cs + q, v); else base[(p * C + c) * cs + q] = v; }
        }
  } else {
    for (int k = lane; k < P * C * cs; k += 32) { if (CS) __stcs(base + k, v); else base[k] = v; }
  }
}

template <int PAT, bool CS>
float run(float4 *buf, int games) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) writer<PAT, CS><<<games / 8, 256>>>(buf, games);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; i++) writer<PAT, CS><<<games / 8, 256>>>(buf, games);
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
  cudaMalloc(&buf, bytes);
  float t;
  t = run<0, true>(buf, games);  printf("chunk-major  st.cs  %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  t = run<1, true>(buf, games);  printf("plane-major  st.cs  %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  t = run<2, true>(buf, games);  printf("linear       st.cs  %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  t = run<0, false>(buf, games); printf("chunk-major  st     %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  t = run<1, false>(buf, games); printf("plane-major  st     %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  t = run<2, false>(buf, games); printf("linear       st     %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
