/* CPU ORACLE (test infrastructure, not product code): plain-C restatement of the
 * reference's fused phase-transform + reassignment loops, parallel over columns
 * with OpenMP exactly like the numba `prange` versions:
 *
 *   ssq_cwt_log           ssqueezepy/algos.py:912-924   _ssq_cwt_log_par
 *   ssq_cwt_log_piecewise ssqueezepy/algos.py:878-895   _ssq_cwt_log_piecewise_par
 *   ssq_cwt_lin           ssqueezepy/algos.py:941-953   _ssq_cwt_lin_par
 *   ssq_stft              ssqueezepy/algos.py:971-984   _ssq_stft_par
 *
 * Typing follows numba's (verified against golden vectors from the real
 * reference in tests/test_oracle_golden.py): for complex64 data the products,
 * difference and sum forming num/den are float32; everything after the cast is
 * float64; rounding is half-to-even (rint); `abs(Wx)` is a float32 hypot.
 * Compile WITHOUT -ffast-math / -ffp-contract=fast (see oracle/Makefile).
 *
 * kind: 0 log, 1 log-piecewise, 2 linear, 3 stft.  const_wide: `const` is used in
 * float64 with complex128 accumulation even for complex64 data (what numba does
 * when `const` is a float64 array, ssqueezing.py:124-129).
 */
#include <math.h>
#include <stdint.h>

typedef struct {
  int kind, omax, flipud, idx1, const_wide;
  double a0, d0, a1, d1, gamma;
} grid_t;

static inline int bin_of(double w, const grid_t* g) {
  double kk;
  if (g->kind == 0) {
    double v = (log2(w) - g->a0) / g->d0;
    v = v > 0 ? v : 0;                 /* max(v, 0); NaN -> 0 */
    kk = rint(v); if (kk > g->omax) kk = g->omax;
  } else if (g->kind == 1) {
    double wl = log2(w);
    if (wl > g->a1) { kk = rint((wl - g->a1) / g->d1) + g->idx1; if (kk > g->omax) kk = g->omax; }
    else            { kk = rint((wl - g->a0) / g->d0); if (!(kk > 0)) kk = 0; }
  } else {
    double v = (w - g->a0) / g->d0;
    v = v > 0 ? v : 0;
    kk = rint(v); if (kk > g->omax) kk = g->omax;
  }
  int k = (int)kk;
  return g->flipud ? g->omax - k : k;
}

/* Wx, dWx, Tx: [na][N] interleaved complex64; cst: [na] float64; Sfs: [na] float32 */
void reassign_c64(const float* Wx, const float* dWx, float* Tx, const double* cst,
                  const float* Sfs, int na, int64_t N, const grid_t* g) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) {
    for (int i = 0; i < na; ++i) {
      int64_t o = ((int64_t)i * N + j) * 2;
      float C = Wx[o], D = Wx[o + 1];
      if (!((double)hypotf(C, D) > g->gamma)) continue;
      float A = dWx[o], B = dWx[o + 1];
      volatile float bc = B * C, ad = A * D, cc = C * C, dd = D * D;
      float num = bc - ad, den = cc + dd;
      double r = (double)num / ((double)den * 6.283185307179586);
      double w = g->kind == 3 ? fabs((double)Sfs[i] - r) : fabs(r);
      int k = bin_of(w, g);
      int64_t t = ((int64_t)k * N + j) * 2;
      if (g->const_wide) {
        Tx[t]     = (float)((double)Tx[t]     + (double)C * cst[i]);
        Tx[t + 1] = (float)((double)Tx[t + 1] + (double)D * cst[i]);
      } else {
        float c32 = (float)cst[i];
        volatile float p0 = C * c32, p1 = D * c32;
        Tx[t] += p0; Tx[t + 1] += p1;
      }
    }
  }
}

/* complex128 variant; Sfs float64 */
void reassign_c128(const double* Wx, const double* dWx, double* Tx, const double* cst,
                   const double* Sfs, int na, int64_t N, const grid_t* g) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) {
    for (int i = 0; i < na; ++i) {
      int64_t o = ((int64_t)i * N + j) * 2;
      double C = Wx[o], D = Wx[o + 1];
      if (!(hypot(C, D) > g->gamma)) continue;
      double A = dWx[o], B = dWx[o + 1];
      volatile double bc = B * C, ad = A * D, cc = C * C, dd = D * D;
      double num = bc - ad, den = cc + dd;
      double r = num / (den * 6.283185307179586);
      double w = g->kind == 3 ? fabs(Sfs[i] - r) : fabs(r);
      int k = bin_of(w, g);
      int64_t t = ((int64_t)k * N + j) * 2;
      volatile double p0 = C * cst[i], p1 = D * cst[i];
      Tx[t] += p0; Tx[t + 1] += p1;
    }
  }
}
