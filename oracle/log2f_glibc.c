/* TEST INFRASTRUCTURE ONLY (see oracle/README): restatement of glibc's float32 log2
 * (sysdeps/ieee754/flt-32/e_log2f.c, "optimized routines" algorithm: 16-entry table,
 * degree-4 polynomial, float64 arithmetic, one final rounding).  numba lowers the
 * `np.log2(float32)` of the reference's two-step reassignment (ssqueezepy/algos.py:175,
 * 186, 200, 216) to the `llvm.log2.f32` intrinsic, i.e. a call to this libm function,
 * whose x86-64 build (ifunc __log2f_fma, compiled with -mfma: a*b+c contracts) is what
 * the golden fixtures were produced with.  `log2f_check()` compares the restatement
 * with the libm of the running process over EVERY positive finite float32.
 *
 * gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC -o _build/liblog2f_glibc.so log2f_glibc.c -lm */
#include <stdint.h>
#include <string.h>
#include <math.h>

static const double TAB[16][2] = {
  { 0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2 }, { 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2 },
  { 0x1.49539f0f010bp+0,  -0x1.7418b0a1fb77bp-2 }, { 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2 },
  { 0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2 }, { 0x1.25e227b0b8eap+0,  -0x1.97c1d1b3b7afp-3 },
  { 0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3 }, { 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4 },
  { 0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5 }, { 0x1p+0, 0x0p+0 },
  { 0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4 },  { 0x1.ca4b31f026aap-1,  0x1.476a9543891bap-3 },
  { 0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2 },
  { 0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2 },  { 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2 },
};
static const double POLY[4] = { -0x1.712b6f70a7e4dp-2, 0x1.ecabf496832ep-2, -0x1.715479ffae3dep-1,
                                0x1.715475f35c8b8p0 };

static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* fma_variant != 0: every a*b+c of the source is one fused operation (the -mfma build) */
float log2f_restated(float x, int fma_variant) {
  uint32_t ix = asuint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return -INFINITY;
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
    ix = asuint(x * 0x1p23f);                 /* subnormal: normalise */
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> (23 - 4)) % 16;
  uint32_t top = tmp & 0xff800000u;
  uint32_t iz = ix - top;
  int k = (int32_t)tmp >> 23;
  double invc = TAB[i][0], logc = TAB[i][1];
  double z = (double)asfloat(iz);
  double r, r2, y, p, y0 = logc + (double)k;
  if (fma_variant) {
    r = fma(z, invc, -1.0);
    r2 = r * r;
    y = fma(POLY[1], r, POLY[2]);
    y = fma(POLY[0], r2, y);
    p = fma(POLY[3], r, y0);
    y = fma(y, r2, p);
  } else {
    r = z * invc - 1.0;
    r2 = r * r;
    y = POLY[1] * r + POLY[2];
    y = POLY[0] * r2 + y;
    p = POLY[3] * r + y0;
    y = y * r2 + p;
  }
  return (float)y;
}

/* number of positive finite float32 inputs where the restatement differs from libm's log2f;
 * first_bad receives one such input (bit pattern) */
long long log2f_check(int fma_variant, uint32_t* first_bad) {
  long long bad = 0;
  uint32_t fb = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (long long u = 1; u < 0x7f800000ll; ++u) {
    float x = asfloat((uint32_t)u);
    if (asuint(log2f(x)) != asuint(log2f_restated(x, fma_variant))) {
      ++bad;
#pragma omp critical
      fb = (uint32_t)u;
    }
  }
  if (first_bad) *first_bad = fb;
  return bad;
}
