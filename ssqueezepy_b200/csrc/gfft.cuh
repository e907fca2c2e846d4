// Batched complex FFT of ANY length (sm_100a): the transform lengths the power-of-two
// kernels do not cover -- `padtype=None` on a signal of N = 160 000 = 2^8 5^4 samples
// (ssqueezepy/utils/common.py:131-156 returns the signal unpadded, _cwt.py:261-271 then
// transforms at length N) and STFT frames of n_fft = 598 = 2 * 13 * 23
// (ssqueezepy/examples/benchmarks.py:82).  The reference hands these to pocketfft.
//
//   n <= GFFT_SMEM_MAX with prime factors <= 31 : mixed-radix Stockham in shared memory, any
//        radix list (2 .. 31), R transforms per CTA side by side;
//   n = n1 * n2, both as above                    : two passes through a global scratch
//        (columns of length n1 with the twiddle w_n^(i2 t1), then rows of length n2);
//   anything else (large prime factors)           : Bluestein's chirp-z through a power-of-two
//        convolution length M >= 2n - 1, itself one of the two cases above.
// Twiddles are evaluated with sincospi in float64 and rounded once.  Accuracy, not speed, is
// the point of this path (the benchmarked sizes are powers of two); it is O(n log n).
#pragma once
#include "ssq_common.cuh"

namespace ssqb {

constexpr int GFFT_MAX_STAGES = 16;

struct GfftStages { int n; int nst; int radix[GFFT_MAX_STAGES]; };

// where transform `id` (0 <= id < count) lives: id = outer * inner_n + inner
struct GfftView {
  long long outer_stride, inner_stride, elem_stride;
};

template <typename T>
struct GfftArgs {
  GfftStages S;
  const cx<T>* in; cx<T>* out;
  GfftView vin, vout;
  long long count, inner_n;
  int sign;                    // +1: sum x e^{+2 pi i ...} (inverse, unnormalised), -1: forward
  long long tw_n;              // > 0: output element e of transform (outer, inner) times
                               //      e^{sign 2 pi i inner e / tw_n}
  T scale;                     // output scale
  int R;                       // transforms per CTA
};

template <typename T>
__device__ __forceinline__ cx<T> unit_root(long long m, long long n, int sign) {
  double s, c;
  sincospi(2.0 * (double)m / (double)n, &s, &c);
  return mkc<T>((T)c, (T)(sign > 0 ? s : -s));
}

template <typename T>
__global__ void __launch_bounds__(256)
gfft_smem_kernel(const GfftArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = A.S.n, R = A.R, NT = blockDim.x, tid = threadIdx.x;
  cx<T>* s0 = reinterpret_cast<cx<T>*>(smem_raw);        // [n][R]
  cx<T>* s1 = s0 + (size_t)n * R;                        // [n][R]
  cx<T>* tw = s1 + (size_t)n * R;                        // [n]  e^{sign 2 pi i m / n}
  const long long id0 = (long long)blockIdx.x * R;
  for (int m = tid; m < n; m += NT) tw[m] = unit_root<T>(m, n, A.sign);
  for (int idx = tid; idx < n * R; idx += NT) {
    const int r = idx % R, e = idx / R;
    const long long id = id0 + r;
    cx<T> v = mkc<T>((T)0, (T)0);
    if (id < A.count) {
      const long long o = id / A.inner_n, i = id - o * A.inner_n;
      v = A.in[o * A.vin.outer_stride + i * A.vin.inner_stride + (long long)e * A.vin.elem_stride];
    }
    s0[idx] = v;
  }
  __syncthreads();
  cx<T>* src = s0; cx<T>* dst = s1;
  int NS = 1;
  for (int st = 0; st < A.S.nst; ++st) {
    const int r = A.S.radix[st], nbf = n / r;
    const int tstep = n / (NS * r);                      // index step into the n-th roots
    // twiddle the inputs in place: x_q *= w^(k q), k = j mod NS
    for (int idx = tid; idx < n * R; idx += NT) {
      const int lane = idx % R, e = idx / R;
      const int q = e / nbf, j = e - q * nbf, k = j % NS;
      if (q && k) src[idx] = cmul<T>(src[idx], tw[((long long)k * q * tstep) % n]);
      (void)lane;
    }
    __syncthreads();
    // y_p = sum_q x_q w_r^(p q); outputs (j - k) r + k + p NS
    for (int idx = tid; idx < n * R; idx += NT) {
      const int lane = idx % R, e = idx / R;
      const int p = e / nbf, j = e - p * nbf, k = j % NS;
      cx<T> acc = mkc<T>((T)0, (T)0);
      const int rstep = n / r;
      int m = 0;                                         // (p q mod r) * (n / r)
      for (int q = 0; q < r; ++q) {
        acc = cmac<T>(acc, src[(j + q * nbf) * R + lane], tw[m]);
        m += p * rstep; if (m >= n) m -= n * (m / n);
      }
      dst[((j - k) * r + k + p * NS) * R + lane] = acc;
    }
    __syncthreads();
    cx<T>* t = src; src = dst; dst = t;
    NS *= r;
  }
  for (int idx = tid; idx < n * R; idx += NT) {
    const int r = idx % R, e = idx / R;
    const long long id = id0 + r;
    if (id >= A.count) continue;
    const long long o = id / A.inner_n, i = id - o * A.inner_n;
    cx<T> v = src[idx];
    if (A.tw_n > 0) v = cmul<T>(v, unit_root<T>((i * (long long)e) % A.tw_n, A.tw_n, A.sign));
    A.out[o * A.vout.outer_stride + i * A.vout.inner_stride + (long long)e * A.vout.elem_stride] =
        cscale<T>(v, A.scale);
  }
}

// ---- element-wise helpers of the Bluestein route and of the generic CWT plan -----------------
// Bluestein: e^{s 2 pi i j k / n} = e^{s i pi j^2/n} e^{s i pi k^2/n} e^{-s i pi (k-j)^2/n}   (s = sign)
// a[b][j] = x[b][j] * e^{s i pi j^2 / n} for j < n, 0 for n <= j < M
template <typename T>
__global__ void __launch_bounds__(256)
gfft_chirp_in_kernel(const cx<T>* __restrict__ x, cx<T>* __restrict__ a, long long n, long long M,
                     long long batch, int sign) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * M) return;
  const long long b = idx / M, j = idx - b * M;
  cx<T> v = mkc<T>((T)0, (T)0);
  if (j < n) {
    const long long q = (j * j) % (2 * n);             // j^2 mod 2n: e^{-+ i pi j^2 / n} has period 2n
    v = cmul<T>(x[b * n + j], unit_root<T>(q, 2 * n, sign));
  }
  a[idx] = v;
}
// kernel of the convolution: bk[d] = e^{-s i pi d^2 / n} for |d| < n placed circularly in M
template <typename T>
__global__ void __launch_bounds__(256)
gfft_chirp_kernel_kernel(cx<T>* __restrict__ bk, long long n, long long M, int sign) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  cx<T> v = mkc<T>((T)0, (T)0);
  long long d = -1;
  if (j < n) d = j; else if (M - j < n) d = M - j;
  if (d >= 0) v = unit_root<T>((d * d) % (2 * n), 2 * n, -sign);
  bk[j] = v;
}
template <typename T>
__global__ void __launch_bounds__(256)
gfft_mul_kernel(cx<T>* __restrict__ a, const cx<T>* __restrict__ bh, long long M, long long batch) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * M) return;
  a[idx] = cmul<T>(a[idx], bh[idx % M]);
}
// X[b][k] = conv[b][k] * e^{s i pi k^2 / n} * scale
template <typename T>
__global__ void __launch_bounds__(256)
gfft_chirp_out_kernel(const cx<T>* __restrict__ conv, cx<T>* __restrict__ X, long long n,
                      long long M, long long batch, int sign, T scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * n) return;
  const long long b = idx / n, k = idx - b * n;
  const cx<T> c = unit_root<T>((k * k) % (2 * n), 2 * n, sign);
  X[idx] = cscale<T>(cmul<T>(conv[b * M + k], c), scale);
}

}  // namespace ssqb
