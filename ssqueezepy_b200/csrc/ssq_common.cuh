// Common device helpers for the sm_100a CWT/STFT synchrosqueezing kernels.
//
// Arithmetic contracts restated from the reference's CPU (numba) kernels:
//   * ssqueezepy/algos.py:912-924 (`_ssq_cwt_log_par`): for complex64 input the
//     products / difference / sum that form `num`, `den` are each rounded to
//     float32; the division, `* 6.283185307179586`, log2, subtraction of `vlmin`,
//     division by `dvl` and the round-half-even are float64.
//   * nvcc contracts a*b+c into FMA by default, which would change those
//     roundings; every op on the exact path therefore uses the *_rn intrinsics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace ssqb {

template <typename T> struct V2;
template <> struct V2<float>  { using type = float2; };
template <> struct V2<double> { using type = double2; };

template <typename T> using cx = typename V2<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mkc(T x, T y) {
  cx<T> v; v.x = x; v.y = y; return v;
}
template <typename T> __device__ __forceinline__ cx<T> cadd(cx<T> a, cx<T> b) {
  return mkc<T>(a.x + b.x, a.y + b.y);
}
template <typename T> __device__ __forceinline__ cx<T> csub(cx<T> a, cx<T> b) {
  return mkc<T>(a.x - b.x, a.y - b.y);
}
template <typename T> __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
  return mkc<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by +i
template <typename T> __device__ __forceinline__ cx<T> cmuli(cx<T> a) {
  return mkc<T>(-a.y, a.x);
}
template <typename T> __device__ __forceinline__ cx<T> cscale(cx<T> a, T s) {
  return mkc<T>(a.x * s, a.y * s);
}
template <typename T> __device__ __forceinline__ cx<T> cconj(cx<T> a) {
  return mkc<T>(a.x, -a.y);
}
// b + h * u  (real h)
template <typename T> __device__ __forceinline__ cx<T> caxpy(cx<T> u, T h, cx<T> b) {
  return mkc<T>(fma(h, u.x, b.x), fma(h, u.y, b.y));
}
// acc + z * w
template <typename T> __device__ __forceinline__ cx<T> cmac(cx<T> acc, cx<T> z, cx<T> w) {
  return mkc<T>(acc.x + (z.x * w.x - z.y * w.y), acc.y + (z.x * w.y + z.y * w.x));
}

// ---- float32: packed two-lane arithmetic ------------------------------------------
// sm_100 executes add/mul/fma.rn.f32x2 on a 64-bit register pair as ONE instruction
// (FADD2 / FMUL2 / FFMA2), with per-lane sign, lane-swap and scalar-broadcast operand
// modifiers, so a complex add is 1 issue slot instead of 2 and a complex multiply 2
// instead of 4; `cmuli` and the (s, s) broadcasts below fold into those modifiers.
// Each lane is an IEEE round-to-nearest op, as in the scalar code.
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  float2 r;
  asm("{\n .reg .b64 a_, b_, c_;\n mov.b64 a_, {%2, %3};\n mov.b64 b_, {%4, %5};\n"
      " add.rn.f32x2 c_, a_, b_;\n mov.b64 {%0, %1}, c_;\n}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 f2_sub(float2 a, float2 b) {
  float2 r;
  asm("{\n .reg .b64 a_, b_, c_;\n mov.b64 a_, {%2, %3};\n mov.b64 b_, {%4, %5};\n"
      " sub.rn.f32x2 c_, a_, b_;\n mov.b64 {%0, %1}, c_;\n}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  float2 r;
  asm("{\n .reg .b64 a_, b_, c_;\n mov.b64 a_, {%2, %3};\n mov.b64 b_, {%4, %5};\n"
      " mul.rn.f32x2 c_, a_, b_;\n mov.b64 {%0, %1}, c_;\n}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n .reg .b64 a_, b_, c_, d_;\n mov.b64 a_, {%2, %3};\n mov.b64 b_, {%4, %5};\n"
      " mov.b64 c_, {%6, %7};\n fma.rn.f32x2 d_, a_, b_, c_;\n mov.b64 {%0, %1}, d_;\n}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
template <> __device__ __forceinline__ float2 cadd<float>(float2 a, float2 b) { return f2_add(a, b); }
template <> __device__ __forceinline__ float2 csub<float>(float2 a, float2 b) { return f2_sub(a, b); }
template <> __device__ __forceinline__ float2 cscale<float>(float2 a, float s) {
  return f2_mul(a, make_float2(s, s));
}
// a * b = a * (b.x, b.x) + (-a.y, a.x) * (b.y, b.y)
template <> __device__ __forceinline__ float2 cmul<float>(float2 a, float2 b) {
  return f2_fma(make_float2(-a.y, a.x), make_float2(b.y, b.y), f2_mul(a, make_float2(b.x, b.x)));
}
template <> __device__ __forceinline__ float2 caxpy<float>(float2 u, float h, float2 b) {
  return f2_fma(u, make_float2(h, h), b);
}
template <> __device__ __forceinline__ float2 cmac<float>(float2 acc, float2 z, float2 w) {
  return f2_fma(make_float2(-z.y, z.x), make_float2(w.y, w.y),
                f2_fma(z, make_float2(w.x, w.x), acc));
}

// float64 complex product with a FIXED rounding sequence: the generic form leaves the contraction
// of a*b - c*d to the compiler, which may fuse a different product in two instantiations of one
// kernel (cwt with and without the derivative must return bit-identical Wx)
template <> __device__ __forceinline__ double2 cmul<double>(double2 a, double2 b) {
  return make_double2(__fma_rn(a.x, b.x, -__dmul_rn(a.y, b.y)), __fma_rn(a.x, b.y, __dmul_rn(a.y, b.x)));
}
// (a real scale feeding a butterfly's add is the same case: with the product left to the compiler,
// x0*p0 + x4*p4 is fused or not depending on how many other uses the product has)
template <> __device__ __forceinline__ double2 cscale<double>(double2 a, double s) {
  return make_double2(__dmul_rn(a.x, s), __dmul_rn(a.y, s));
}
template <> __device__ __forceinline__ double2 cmac<double>(double2 acc, double2 z, double2 w) {
  return make_double2(__fma_rn(-z.y, w.y, __fma_rn(z.x, w.x, acc.x)), __fma_rn(z.y, w.x, __fma_rn(z.x, w.y, acc.y)));
}

// ---- exactly-rounded (never FMA-contracted) scalar ops ---------------------
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ float  sub_rn(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

#define SSQB_TWO_PI 6.283185307179586   // literal used at algos.py:918

// ---- reassignment-grid description (host fills; passed by value) -----------
// kind: 0 log (algos.py:920), 1 log-piecewise (algos.py:886-889),
//       2 linear (algos.py:949), 3 stft-linear (algos.py:978-981)
struct ReassignGrid {
  int    kind;
  int    omax;        // na - 1
  int    flipud;
  int    idx1;        // log-piecewise only
  double a0, d0;      // vlmin / dvl   (log, piecewise lower) or vmin / dv (lin)
  double a1, d1;      // vlmin1 / dvl1 (piecewise upper)
  double gamma;       // threshold on |Wx|
  // fast-path helpers (float32 estimate of the bin coordinate + guard band);
  // the exact float64 formula is always used inside the guard band.
  float  fa0, fid0, fa1, fid1, ftol;
  float  fvhi;        // omax + 0.25
  float  fhalf;       // 0.5 - ftol
  float  fidx1;       // (float)idx1
  int    const_wide;  // 1: `const` is float64 and products are taken in float64
                      //    (log-piecewise on float32 data)
};

// |z| > gamma exactly as numba types it (algos.py:915): complex64 -> float32
// magnitude (correctly rounded hypot), compared in float64.
__device__ __forceinline__ bool is_active_exact(float C, float D, double gamma) {
  double dd = (double)C * (double)C + (double)D * (double)D;   // products exact
  float m = (float)sqrt(dd);
  return (double)m > gamma;
}
__device__ __forceinline__ bool is_active_exact(double C, double D, double gamma) {
  return hypot(C, D) > gamma;
}
// `abs(Wx) < gamma` of phase_cwt (algos.py:724) - gamma already cast to dtype
__device__ __forceinline__ bool is_below_exact(float C, float D, float gamma) {
  double dd = (double)C * (double)C + (double)D * (double)D;
  return (float)sqrt(dd) < gamma;
}
__device__ __forceinline__ bool is_below_exact(double C, double D, double gamma) {
  return hypot(C, D) < gamma;
}

// Im(dWx/Wx)/(2 pi) with the reference's roundings; (A,B)=dWx, (C,D)=Wx.
template <typename T>
__device__ __forceinline__ double phase_ratio_exact(T A, T B, T C, T D) {
  T num = sub_rn(mul_rn(B, C), mul_rn(A, D));
  T den = add_rn(mul_rn(C, C), mul_rn(D, D));
  return (double)num / ((double)den * SSQB_TWO_PI);
}

// bin index from the float64 `w` (or from log2(w) if LOGGED).  Returns the row
// after the optional flip.
__device__ __forceinline__ int bin_from_w_exact(double w, const ReassignGrid& g) {
  double kk;
  if (g.kind == 0) {
    double v = (log2(w) - g.a0) / g.d0;
    v = fmax(v, 0.0);
    kk = fmin(rint(v), (double)g.omax);
  } else if (g.kind == 1) {
    double wl = log2(w);
    if (wl > g.a1) kk = fmin(rint((wl - g.a1) / g.d1) + (double)g.idx1, (double)g.omax);
    else           kk = fmax(rint((wl - g.a0) / g.d0), 0.0);
  } else {
    double v = (w - g.a0) / g.d0;
    v = fmax(v, 0.0);
    kk = fmin(rint(v), (double)g.omax);
  }
  if (!(kk == kk)) kk = 0.0;              // NaN guard (undefined in the reference)
  int k = (int)kk;
  return g.flipud ? (g.omax - k) : k;
}

// Fused-path bin index: float32 estimate, exact float64 only near a rounding
// boundary.  Bit-identical to bin_from_w_exact(fabs(phase_ratio_exact)) by
// construction: the estimate is trusted only when it is farther than `ftol`
// (a host-computed bound on its error) from every half-integer and clamp edge.
__device__ __forceinline__ float w_estimate(float num, float den) {
  return fabsf(num) / (den * 6.2831853f);
}
__device__ __forceinline__ float w_estimate(double num, double den) {
  // float64 data: the ratio itself is the exact w; only log2 is estimated
  return (float)(fabs(num) / (den * SSQB_TWO_PI));
}

template <typename T>
__device__ __forceinline__ int bin_fused(T A, T B, T C, T D, const ReassignGrid& g) {
  T num = sub_rn(mul_rn(B, C), mul_rn(A, D));
  T den = add_rn(mul_rn(C, C), mul_rn(D, D));
  if (g.kind <= 1 && g.ftol < 0.25f) {
    float wf = w_estimate(num, den);
    float lf = __log2f(wf);
    float v;
    bool ok = true;
    int k = 0;
    if (g.kind == 0) {
      v = (lf - g.fa0) * g.fid0;
    } else {
      // which branch? decided by wl > vlmin1; guard near the switch point
      float dsw = lf - g.fa1;
      if (fabsf(dsw) * g.fid1 <= g.ftol) ok = false;
      if (dsw > 0.f) v = dsw * g.fid1 + (float)g.idx1;
      else           v = (lf - g.fa0) * g.fid0;
    }
    if (ok) {
      float vm = (float)g.omax;
      if (!(v == v)) ok = false;                         // NaN -> exact path
      else if (v <= -1.0f) k = 0;
      else if (v >= vm + 1.0f) k = g.omax;
      else {
        float r = rintf(v);
        float fr = fabsf(v - r);                         // distance to integer
        if (fr >= 0.5f - g.ftol) ok = false;             // near a half-integer
        else {
          // near the clamp edges rint(v) is still right: max(v,0)->rint, min(.,omax)
          r = fminf(fmaxf(r, 0.f), vm);
          k = (int)r;
        }
      }
    }
    if (ok) return g.flipud ? (g.omax - k) : k;
  }
  double w = fabs((double)num / ((double)den * SSQB_TWO_PI));
  return bin_from_w_exact(w, g);
}

// reflect / zero / symmetric / replicate / wrap index map of
// ssqueezepy/utils/common.py:131-147 (np.pad modes).  Returns -1 for "zero".
__device__ __forceinline__ int64_t pad_src_index(int64_t t, int64_t n1, int64_t N, int padtype) {
  int64_t s = t - n1;
  if (s >= 0 && s < N) return s;
  switch (padtype) {
    case 0: {                                  // reflect (no edge repeat), period 2(N-1)
      if (N == 1) return 0;
      int64_t P = 2 * (N - 1);
      int64_t m = s % P; if (m < 0) m += P;
      return m < N ? m : P - m;
    }
    case 1: return -1;                         // zero
    case 2: {                                  // symmetric (edge repeated), period 2N
      int64_t P = 2 * N;
      int64_t m = s % P; if (m < 0) m += P;
      return m < N ? m : P - 1 - m;
    }
    case 3: return s < 0 ? 0 : N - 1;          // replicate
    default: {                                 // wrap
      int64_t m = s % N; if (m < 0) m += N;
      return m;
    }
  }
}

}  // namespace ssqb
