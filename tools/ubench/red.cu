// Microbenchmark: cost of the Tx scatter (red.global.add.v2.f32) against plain 8-byte
// stores, for the address patterns the fused epilogue produces.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red red.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ void red2(float2* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}
// NA x N plane; thread -> column j (consecutive lanes = consecutive columns), loops rows.
// MODE 0: plain store own row | 1: RED own row | 2: RED row k = hash(row, j/32) (warp-uniform
// random row, 256 B contiguous) | 3: RED row = hash(row, j) (per-lane random row)
// | 4: RED row = hash(row, j/8) (8-lane runs, the round-1 kernel's pattern)
// | 5: store own row + RED mode 2 (Wx store + Tx scatter together)
template <int MODE>
__global__ void k(float2* W, float2* T, int NA, int N, int frac256) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  int r0 = blockIdx.y * 4;
  for (int r = r0; r < r0 + 4 && r < NA; ++r) {
    float2 v = make_float2(1.f + r, 2.f + j);
    unsigned h;
    if (MODE == 2 || MODE == 5) h = (unsigned)(r * 2654435761u) ^ (unsigned)((j >> 5) * 40503u);
    else if (MODE == 4) h = (unsigned)(r * 2654435761u) ^ (unsigned)((j >> 3) * 40503u);
    else h = (unsigned)(r * 2654435761u) ^ (unsigned)(j * 40503u);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    int kk = (int)(h % (unsigned)NA);
    bool act = ((h >> 20) & 255) < (unsigned)frac256;
    if (MODE == 0 || MODE == 5) W[(size_t)r * N + j] = v;
    if (MODE == 1) red2(&T[(size_t)r * N + j], v.x, v.y);
    if ((MODE >= 2) && act) red2(&T[(size_t)kk * N + j], v.x, v.y);
  }
}
int main() {
  const int NA = 300, N = 160000;
  float2 *W, *T; cudaMalloc(&W, (size_t)NA * N * 8); cudaMalloc(&T, (size_t)NA * N * 8);
  cudaMemset(T, 0, (size_t)NA * N * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  dim3 g((N + 255) / 256, (NA + 3) / 4);
  const char* nm[6] = {"store own row", "RED own row", "RED warp-uniform random row", "RED per-lane random row",
                       "RED 8-lane-run random row", "store + RED warp-uniform"};
  for (int frac : {256, 131}) for (int m = 0; m < 6; ++m) {
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
      cudaMemsetAsync(T, 0, (size_t)NA * N * 8);
      cudaEventRecord(a);
      switch (m) {
        case 0: k<0><<<g, 256>>>(W, T, NA, N, frac); break;
        case 1: k<1><<<g, 256>>>(W, T, NA, N, frac); break;
        case 2: k<2><<<g, 256>>>(W, T, NA, N, frac); break;
        case 3: k<3><<<g, 256>>>(W, T, NA, N, frac); break;
        case 4: k<4><<<g, 256>>>(W, T, NA, N, frac); break;
        case 5: k<5><<<g, 256>>>(W, T, NA, N, frac); break;
      }
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("active %3d/256  %-32s %.3f ms  (%.0f GB/s of 384 MB)\n", frac, nm[m], best, 0.384 / best * 1e3);
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
