// store_holes.cu — what does a 16-byte HOLE in a warp's contiguous store cost when it is filled a moment later?
// 65,536 x 4 warps each sweep a 16,192-byte block (1,012 float4) linearly, 512 contiguous bytes per instruction.
//   0  the whole block in the sweep
//   1  every 56th float4 is skipped in the sweep and written in ONE extra round of 128-bit stores afterwards
//   2  ... written afterwards as four scalar stores each
//   3  the whole block in the sweep, plus the 16 bytes before and after the block written as scalars by this warp
//      (what a block that starts mid-sector does to its neighbours)
//   4  sector-aligned blocks of 1,012 float4, nothing else (the reference point)
//   5  the 15x15 observation layout: contiguous blocks of 4,050 floats (16,200 B: every block starts and ends
//      mid-sector), scalar head / 128-bit body / scalar tail per block
//   6  the same blocks, but a warp sweeps FOUR consecutive blocks (64,800 B = 2,025 whole sectors) as one run
// Same bytes in every case.  (profiles/r1_variants.md, "holes")
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int MODE, bool CS>
__global__ void __launch_bounds__(256) writer(float4 *out, int blocks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= blocks) return;
  const int Q = 1012;
  float4 *base = out + (size_t)b * (Q + 2) + 1;  // one float4 of slack either side (mode 3 writes it)
  const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
  auto st4 = [&](float4 *p) { if (CS) __stcs(p, v); else *p = v; };
  auto st1 = [&](float *p) { if (CS) __stcs(p, 1.f); else *p = 1.f; };
  if (MODE == 4) {
    float4 *a = out + (size_t)b * Q;
    for (int k = lane; k < Q; k += 32) st4(a + k);
    return;
  }
  if (MODE == 5) {
    float *f = reinterpret_cast<float *>(out) + (size_t)b * 4050;
    const int head = (int)((4u - (uint32_t)(((size_t)b * 4050) & 3u)) & 3u), body4 = (4050 - head) / 4, tail0 = head + 4 * body4;
    if (lane < head) st1(f + lane);
    if (lane < 4050 - tail0) st1(f + tail0 + lane);
    float4 *a = reinterpret_cast<float4 *>(f + head);
    for (int k = lane; k < body4; k += 32) st4(a + k);
    return;
  }
  if (MODE == 6) {
    if (b & 3) return;  // warps 0, 4 of the CTA write four blocks each (same bytes per launch)
    float4 *a = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + (size_t)b * 4050);
    for (int k = lane; k < 4050; k += 32) st4(a + k);
    return;
  }
  for (int k = lane; k < Q; k += 32) {
    const bool hole = (MODE == 1 || MODE == 2) && (k % 56 == 55);
    if (!hole) st4(base + k);
  }
  if (MODE == 1) {
    for (int j = lane; j < Q / 56; j += 32) st4(base + 56 * j + 55);
  } else if (MODE == 2) {
    for (int j = lane; j < 4 * (Q / 56); j += 32) st1(reinterpret_cast<float *>(base + 56 * (j >> 2) + 55) + (j & 3));
  } else if (MODE == 3) {
    if (lane < 4) st1(reinterpret_cast<float *>(base - 1) + lane);
    else if (lane < 8) st1(reinterpret_cast<float *>(base + Q) + lane - 4);
  } else {
    if (lane == 0) { st4(base - 1); st4(base + Q); }  // same bytes as mode 3, as full 128-bit stores
  }
}

template <int MODE, bool CS>
float run(float4 *buf, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) writer<MODE, CS><<<blocks / 8, 256>>>(buf, blocks);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; i++) writer<MODE, CS><<<blocks / 8, 256>>>(buf, blocks);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 20;
}

int main() {
  const int blocks = 262144;
  const size_t bytes = (size_t)blocks * 1014 * 16;
  float4 *buf;
  cudaMalloc(&buf, bytes);
  const char *names[7] = {"no holes, mid-sector block edges", "holes, one extra 128-bit round  ", "holes, scalar stores afterwards ",
                          "no holes, scalar block edges    ", "sector-aligned blocks           ", "15x15 layout, block per warp    ",
                          "15x15 layout, 4 blocks per warp "};
  float t[14];
  t[0] = run<0, true>(buf, blocks); t[1] = run<1, true>(buf, blocks); t[2] = run<2, true>(buf, blocks); t[3] = run<3, true>(buf, blocks);
  t[4] = run<4, true>(buf, blocks); t[5] = run<5, true>(buf, blocks); t[6] = run<6, true>(buf, blocks);
  t[7] = run<0, false>(buf, blocks); t[8] = run<1, false>(buf, blocks); t[9] = run<2, false>(buf, blocks); t[10] = run<3, false>(buf, blocks);
  t[11] = run<4, false>(buf, blocks); t[12] = run<5, false>(buf, blocks); t[13] = run<6, false>(buf, blocks);
  for (int i = 0; i < 14; i++) {
    const int m = i % 7;
    const size_t by = m == 4 ? (size_t)blocks * 1012 * 16 : (m >= 5 ? (size_t)blocks * 16200 : bytes);
    printf("%s %s  %.3f ms  %.0f GB/s\n", names[m], i < 7 ? "st.cs" : "st   ", t[i], by / t[i] / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
