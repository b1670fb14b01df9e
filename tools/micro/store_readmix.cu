// store_readmix.cu — why does reading a 1.4 KB slab per warp cost the 28.8 KB store stream 15 %?
// 65,536 warps, each: [load its slab] -> 72 x 128-bit st.cs plane-major -> [write the slab back].
//   mode 0  no slab traffic
//   mode 1  dependent plain loads (stores wait for the slab)            = the turn kernel today
//   mode 2  independent loads (issued first, consumed after the stores) -> separates latency from DRAM efficiency
//   mode 3  mode 1 + slab written back (plain)
//   mode 4  mode 3 with L2::evict_last policy on the slab loads and write-back (state stays in the 126 MB L2)
//   mode 5  mode 4 + the observation stores carry an L2::evict_first policy instead of .cs
//   mode 6  mode 3 but only the first third of the slab is written back (contiguous 432 B)
//   mode 7  mode 3 but only every third 32-byte sector is written back (scattered dirty sectors)
//   mode 8  mode 3 with the slab EMBEDDED behind the game's observation block (one 30,720-byte record per game)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_hint(const uint4 *a, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint(uint4 *a, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_f4(float4 *a, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256) writer(float4 *out, uint4 *slabs, int games) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int game = blockIdx.x * 8 + warp;
  if (game >= games) return;
  const int P = 2, C = 9, cs = 100;
  float4 *base = out + (size_t)game * (MODE == 8 ? 1920 : P * C * cs);
  uint4 *s = MODE == 8 ? reinterpret_cast<uint4 *>(base + 1800) : slabs + (size_t)game * 91;  // 1,456 B
  uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a;
  uint64_t pl = 0, pf = 0;
  if (MODE >= 4) pl = policy_evict_last();
  if (MODE >= 5) pf = policy_evict_first();
  if (MODE >= 1) {
    if (MODE >= 4) {
      a = ld_hint(s + lane, pl);
      b = ld_hint(s + 32 + lane, pl);
      if (lane < 27) c = ld_hint(s + 64 + lane, pl);
    } else {
      a = s[lane];
      b = s[32 + lane];
      if (lane < 27) c = s[64 + lane];
    }
  }
  uint32_t key = 0;
  if (MODE == 1 || MODE >= 3) key = __reduce_or_sync(0xffffffffu, a.x ^ b.y ^ c.z);
  float4 v = make_float4(1.f, 0.f, 1.f, __uint_as_float(key));
  for (int p = 0; p < P; p++)
    for (int ch = 0; ch < C; ch++)
#pragma unroll
      for (int q0 = 0; q0 < 128; q0 += 32) {
        int q = q0 + lane;
        if (q < cs) {
          if (MODE >= 5) st_hint_f4(base + (p * C + ch) * cs + q, v, pf);
          else __stcs(base + (p * C + ch) * cs + q, v);
        }
      }
  if (MODE == 2) {
    key = __reduce_or_sync(0xffffffffu, a.x ^ b.y ^ c.z);
    if (key == 0x12345u) base[0] = v;
  }
  if (MODE >= 3) {
    a.x += 1;
    if (MODE >= 4) {
      st_hint(s + lane, a, pl);
      st_hint(s + 32 + lane, b, pl);
      if (lane < 27) st_hint(s + 64 + lane, c, pl);
    } else if (MODE == 6) {
      if (lane < 27) s[lane] = a;
    } else if (MODE == 7) {
      if ((lane >> 1) % 3 == 0) s[lane] = a;
      if (((32 + lane) >> 1) % 3 == 0) s[32 + lane] = b;
      if (lane < 27 && ((64 + lane) >> 1) % 3 == 0) s[64 + lane] = c;
    } else {
      s[lane] = a;
      s[32 + lane] = b;
      if (lane < 27) s[64 + lane] = c;
    }
  }
}

template <int MODE>
float run(float4 *buf, uint4 *slabs, int games) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 5; i++) writer<MODE><<<games / 8, 256>>>(buf, slabs, games);
  cudaEventRecord(e0);
  for (int i = 0; i < 30; i++) writer<MODE><<<games / 8, 256>>>(buf, slabs, games);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 30;
}

int main() {
  for (int games : {65536, 32768, 16384}) {
    const size_t bytes = (size_t)games * 28800;
    float4 *buf;
    const size_t alloc = (size_t)games * 30720;
    uint4 *slabs;
    cudaMalloc(&buf, alloc);
    cudaMemset(buf, 0, alloc);
    cudaMalloc(&slabs, (size_t)games * 1456);
    cudaMemset(slabs, 0, (size_t)games * 1456);
    float t[9] = {run<0>(buf, slabs, games), run<1>(buf, slabs, games), run<2>(buf, slabs, games),
                  run<3>(buf, slabs, games), run<4>(buf, slabs, games), run<5>(buf, slabs, games),
                  run<6>(buf, slabs, games), run<7>(buf, slabs, games), run<8>(buf, slabs, games)};
    const char *names[9] = {"no slab", "dependent load", "independent load", "load + write-back", "evict_last slab", "evict_last slab + evict_first obs", "write back first third", "write back every 3rd sector", "slab embedded behind obs block"};
    printf("games %d (state %.0f MB)\n", games, games * 1456 / 1e6);
    for (int m = 0; m < 9; m++) printf("  mode %d %-34s %.4f ms  %5.0f GB/s of obs\n", m, names[m], t[m], bytes / t[m] / 1e6);
    cudaFree(buf);
    cudaFree(slabs);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
