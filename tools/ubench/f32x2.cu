// Microbenchmark: issue cost of sm_100 packed float32x2 ops and of their operand
// modifiers (lane swap, per-lane sign, scalar broadcast).  Build: nvcc -gencode
// arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef float2 C;
__device__ __forceinline__ C mk(float x, float y) { return make_float2(x, y); }
#define PK2(op) C r; asm("{\n .reg .b64 a_, b_, c_;\n mov.b64 a_, {%2, %3};\n mov.b64 b_, {%4, %5};\n " op ".rn.f32x2 c_, a_, b_;\n mov.b64 {%0, %1}, c_;\n}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)); return r;
__device__ __forceinline__ C padd(C a, C b) { PK2("add") }
__device__ __forceinline__ C pmul(C a, C b) { PK2("mul") }
__device__ __forceinline__ C pfma(C a, C b, C c) { C r;
  asm("{\n .reg .b64 a_, b_, c_, d_;\n mov.b64 a_, {%2, %3};\n mov.b64 b_, {%4, %5};\n mov.b64 c_, {%6, %7};\n fma.rn.f32x2 d_, a_, b_, c_;\n mov.b64 {%0, %1}, d_;\n}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y)); return r; }

template <int MODE> __global__ void k(C* y, const C* wp, int iters) {
  C a[8], w = wp[threadIdx.x & 7];
  int n[8];
  for (int q = 0; q < 8; ++q) { a[q] = mk(threadIdx.x + q, threadIdx.x - q); n[q] = threadIdx.x * q; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (MODE == 0) { a[q].x += w.x; a[q].y += w.y; }                       // 2 FADD
      if (MODE == 1) a[q] = padd(a[q], w);                                   // FADD2
      if (MODE == 2) a[q] = padd(a[q], mk(-w.x, w.y));                       // sign
      if (MODE == 3) a[q] = padd(a[q], mk(w.y, w.x));                        // swap
      if (MODE == 4) a[q] = padd(a[q], mk(-w.y, w.x));                       // swap + sign
      if (MODE == 5) a[q] = pmul(a[q], mk(w.x, w.x));                        // broadcast
      if (MODE == 6) a[q] = pfma(a[q], mk(w.x, w.x), mk(w.y, w.y));          // fma, 2 broadcasts
      if (MODE == 7) a[q] = pfma(mk(-a[q].y, a[q].x), mk(w.y, w.y), pmul(a[q], mk(w.x, w.x)));  // complex mul
      if (MODE == 8) { a[q].x = a[q].x * w.x - a[q].y * w.y; a[q].y = a[q].x * w.y + a[q].y * w.x; }  // scalar cmul-ish
      if (MODE == 9) { if (q & 1) a[q] = padd(a[q], w); else { a[q].x += w.x; a[q].y += w.y; } }  // mix
      if (MODE == 10) { a[q] = padd(a[q], w); n[q] = (n[q] ^ it) + q; }      // FADD2 + 2 ALU
      if (MODE == 11) { a[q].x += w.x; a[q].y += w.y; n[q] = (n[q] ^ it) + q; }  // 2 FADD + 2 ALU
      if (MODE == 13) a[q] = padd(a[q], mk(a[(q + 1) & 7].y, a[(q + 1) & 7].x));    // swap, reg operand
      if (MODE == 14) a[q] = padd(a[q], a[(q + 1) & 7]);                              // plain, reg operand
      if (MODE == 15) a[q] = padd(a[q], mk(-a[(q + 1) & 7].x, a[(q + 1) & 7].y));    // sign, reg operand
      if (MODE == 12) a[q] = padd(a[q], mk(-a[(q + 1) & 7].y, a[(q + 1) & 7].x));   // swap+sign, reg operand
    }
  }
  C s = mk(0, 0); int m = 0;
  for (int q = 0; q < 8; ++q) { s.x += a[q].x; s.y += a[q].y; m += n[q]; }
  s.x += (float)m;
  y[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* nm, C* y, const C* w) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  k<MODE><<<148 * 4, 512>>>(y, w, 10);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 512>>>(y, w, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.965e9;                               // SM cycles
  double per = cyc / ((double)iters * 8) / 4.0;                   // 16 warps/SM = 4 per SMSP
  printf("%-28s %.3f ms  %.2f cycles per complex op per warp (per SMSP)\n", nm, ms, per);
}
int main() {
  C* y; cudaMalloc(&y, 148 * 4 * 512 * sizeof(C));
  C hw[8]; for (int i = 0; i < 8; ++i) hw[i] = make_float2(1e-3f, 0.999f);
  C* w; cudaMalloc(&w, sizeof(hw)); cudaMemcpy(w, hw, sizeof(hw), cudaMemcpyHostToDevice);
  run<0>("2 FADD", y, w); run<1>("FADD2", y, w); run<2>("FADD2 sign", y, w); run<3>("FADD2 swap", y, w);
  run<4>("FADD2 swap+sign", y, w); run<5>("FMUL2 bcast", y, w); run<6>("FFMA2 2 bcast", y, w);
  run<7>("cmul packed (2 instr)", y, w); run<8>("cmul scalar (4 instr)", y, w); run<9>("mix FADD2 / 2 FADD", y, w);
  run<10>("FADD2 + 2 ALU", y, w); run<11>("2 FADD + 2 ALU", y, w); run<12>("FADD2 swap+sign reg", y, w);
  run<13>("FADD2 swap reg", y, w); run<14>("FADD2 plain reg", y, w); run<15>("FADD2 sign reg", y, w);
  return 0;
}
