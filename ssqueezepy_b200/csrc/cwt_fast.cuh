// Fast-path CWT kernels for large padded lengths (n_up >= 2^13, F = 512).
//
// (1) psih_band_kernel   -- once per plan: samples psih(scale_a * xi_i) (and
//     psih * xi / dt for the derivative) on each scale's non-negligible band
//     [band_lo, band_lo + band_len).  This is the reference's `Wavelet.Psih` cache
//     (ssqueezepy/wavelets.py:135-160) restricted to the bins that matter, so the
//     transcendental work leaves the per-call path.
//
// (2) cwt_direct_kernel  -- rows whose band is at most QMAX*F bins: the whole
//     inverse FFT is ONE pass.  With i = i1 + F*i2 and only <= QMAX non-zero i2 per
//     i1, the length-I2 transforms of pass 1 collapse to a QMAX-term sum that each
//     CTA evaluates for its own 16 (8) output phases t2:
//        A[i1][t2] = w_n^(i1*t2') * sum_q Z[i] * w_I2^(q*t2)    (i = lo + m0 + F*q)
//     followed by the length-F transform over i1 and the fused epilogue (unpad,
//     store Wx, phase transform, bin, red.global.add into Tx).  No scratch, no
//     second kernel, one launch for all such rows of the whole batch.
//
// Both arrays (W and dW) go through the FFT together and share twiddles/indices.
#pragma once
#include "cwt_kernels.cuh"

namespace ssqb {

// ---- two-array Stockham stage (shared twiddles / index math) -------------------
template <typename T, int LOG_M, int R, int NT, int STRIDE, int RADIX, int NS, int NARR>
__device__ __forceinline__ void stockham_stage_n(cx<T>* s, const cx<T>* __restrict__ tw) {
  constexpr int M = 1 << LOG_M;
  constexpr int ASTR = M * STRIDE;              // elements between arrays
  constexpr int NBF = (M / RADIX) * R;
  static_assert(NBF % NT == 0, "butterflies must divide evenly over threads");
  constexpr int BPT = NBF / NT;
  const int tid = threadIdx.x;
  cx<T> v[NARR][BPT][RADIX];
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    int lin = tid + b * NT;
    int r = lin % R, j = lin / R;
#pragma unroll
    for (int a = 0; a < NARR; ++a)
#pragma unroll
      for (int q = 0; q < RADIX; ++q)
        v[a][b][q] = s[a * ASTR + (j + q * (M / RADIX)) * STRIDE + r];
    if (NS > 1) {
      int k = j & (NS - 1);
      constexpr int TSTEP = M / (NS * RADIX);
#pragma unroll
      for (int q = 1; q < RADIX; ++q) {
        cx<T> w = tw[(k * q * TSTEP) & (M - 1)];
#pragma unroll
        for (int a = 0; a < NARR; ++a) v[a][b][q] = cmul<T>(v[a][b][q], w);
      }
    }
#pragma unroll
    for (int a = 0; a < NARR; ++a) idft<T, RADIX>(v[a][b]);
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    int lin = tid + b * NT;
    int r = lin % R, j = lin / R;
    int k = j & (NS - 1);
    int j0 = (j - k) * RADIX + k;
#pragma unroll
    for (int a = 0; a < NARR; ++a)
#pragma unroll
      for (int q = 0; q < RADIX; ++q)
        s[a * ASTR + (j0 + q * NS) * STRIDE + r] = v[a][b][q];
  }
  __syncthreads();
}

template <typename T, int LOG_M, int R, int NT, int STRIDE, int NS, int NARR>
__device__ __forceinline__ void stockham_from_n(cx<T>* s, const cx<T>* __restrict__ tw) {
  constexpr int M = 1 << LOG_M;
  if constexpr (NS < M) {
    if constexpr (NS * 8 <= M) {
      stockham_stage_n<T, LOG_M, R, NT, STRIDE, 8, NS, NARR>(s, tw);
      stockham_from_n<T, LOG_M, R, NT, STRIDE, NS * 8, NARR>(s, tw);
    } else if constexpr (NS * 4 == M) {
      stockham_stage_n<T, LOG_M, R, NT, STRIDE, 4, NS, NARR>(s, tw);
    } else {
      stockham_stage_n<T, LOG_M, R, NT, STRIDE, 2, NS, NARR>(s, tw);
    }
  }
}

// ---- (1) wavelet band tables ------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
psih_band_kernel(const CwtArgs<T> A, const long long* __restrict__ tab_off,
                 T* __restrict__ tab_p, T* __restrict__ tab_pd) {
  const int a = blockIdx.y;
  const long long L = A.band_len[a];
  const long long lo = A.band_lo[a];
  const long long off = tab_off[a];
  const T sc = A.scales[a];
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < L;
       m += (long long)gridDim.x * blockDim.x) {
    long long i = (lo + m) & (A.n_up - 1);
    T p = psih_eval<T>(A, a, i, sc);
    tab_p[off + m] = p;
    tab_pd[off + m] = p * (xi_of<T>(i, A.n_up) / A.dt);   // factor of `*= 1j*xi/dt`
  }
}

// ---- (2) direct single-pass rows ----------------------------------------------------
template <typename T>
struct FastArgs {
  CwtArgs<T> A;
  const int* rows;             // [n_rows] scale indices handled by this launch
  int n_rows;                  // rows per signal in `rows`
  const long long* tab_off;    // [na]
  const T* tab_p;              // psih on the band
  const T* tab_pd;             // psih * xi / dt on the band
  int write_dWx;
  int ssq;                     // 1: fused synchrosqueezing epilogue, 0: plain cwt
  int narr;                    // 1 or 2 arrays needed
};

template <typename T, int LOGE, int QMAX>
__global__ void __launch_bounds__((1 << LOGE) / 16)
cwt_direct_kernel(const FastArgs<T> P) {
  constexpr int ELEMS = 1 << LOGE;
  constexpr int NT = ELEMS / 16;
  constexpr int LOG_F = 9, F = 512;
  constexpr int R2 = ELEMS / F;
  constexpr int EPT = ELEMS / NT;                    // 16 elements per thread per array
  const CwtArgs<T>& A = P.A;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* s = reinterpret_cast<cx<T>*>(smem_raw);     // [2][F][R2]
  cx<T>* tw = s + 2 * ELEMS;                         // [F]   F-th roots
  cx<T>* tlo = tw + F;                               // [2^log_lo]
  cx<T>* thi = tlo + (1 << A.log_lo);                // [n / 2^log_lo]

  const int tid = threadIdx.x;
  const int n_lo = 1 << A.log_lo, n_hi = 1 << (A.logn - A.log_lo);
  for (int m = tid; m < F; m += NT) tw[m] = A.tw2[m];
  for (int m = tid; m < n_lo; m += NT) tlo[m] = A.tw_lo[m];
  for (int m = tid; m < n_hi; m += NT) thi[m] = A.tw_hi[m];

  const int y = blockIdx.y;
  const int b = y / P.n_rows;
  const int a = P.rows[y - b * P.n_rows];
  const unsigned nmask = (unsigned)(A.n_up - 1);
  const int lo = (int)(A.band_lo[a] & (long long)nmask);   // band start, mod n
  const int L = (int)A.band_len[a];
  const T* __restrict__ tp = P.tab_p + P.tab_off[a];
  const T* __restrict__ tpd = P.tab_pd + P.tab_off[a];
  const cx<T>* __restrict__ xh = A.xh + (long long)b * A.n_up;

  const int c = tid % R2, g = tid / R2;              // output phase lane, element group
  const int t2 = blockIdx.x * R2 + c;                // < I2
  const int I2m1 = (1 << A.logI2) - 1;

  // u_q = w_I2^(q*t2) = w_n^(q*t2*F): per-thread constants
  cx<T> u[QMAX];
  u[0] = mkc<T>((T)1, (T)0);
#pragma unroll
  for (int q = 1; q < QMAX; ++q) {
    unsigned mm = ((unsigned)(q * (t2 & I2m1)) << LOG_F) & nmask;
    u[q] = cmul<T>(A.tw_lo[mm & (n_lo - 1)], A.tw_hi[mm >> A.log_lo]);
  }
  __syncthreads();

  constexpr int GSTEP = NT / R2;                     // 32 element groups
#pragma unroll 4
  for (int k = 0; k < EPT; ++k) {
    const int e = g + k * GSTEP;                     // i1
    const int m0 = (e - lo) & (F - 1);               // first band offset with i == e (mod F)
    cx<T> accW = mkc<T>((T)0, (T)0), accD = mkc<T>((T)0, (T)0);
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      int m = m0 + q * F;
      if (m < L) {
        unsigned i = (unsigned)(lo + m) & nmask;
        cx<T> xv = __ldg(&xh[i]);
        T p = __ldg(&tp[m]), pd = __ldg(&tpd[m]);
        cx<T> zu = cmul<T>(xv, u[q]);
        accW.x += zu.x * p;  accW.y += zu.y * p;     // Psih * xh           (_cwt.py:169)
        accD.x -= zu.y * pd; accD.y += zu.x * pd;    // * (1j * xi / dt)    (_cwt.py:175)
      }
    }
    // common twiddle w_n^(i_base * t2), i_base = lo + m0 (the q = 0 index)
    unsigned ib = (unsigned)(lo + m0) & nmask;
    unsigned mm = (ib * (unsigned)t2) & nmask;
    cx<T> w = cmul<T>(tlo[mm & (n_lo - 1)], thi[mm >> A.log_lo]);
    s[e * R2 + c] = cmul<T>(accW, w);
    s[ELEMS + e * R2 + c] = cmul<T>(accD, w);
  }
  __syncthreads();

  if (P.narr == 2) stockham_from_n<T, LOG_F, R2, NT, R2, 1, 2>(s, tw);
  else             stockham_from_n<T, LOG_F, R2, NT, R2, 1, 1>(s, tw);

  // ---- epilogue: t = I2*e + t2 ----------------------------------------------------
  const long long row = (long long)b * A.na + a;
  cx<T>* __restrict__ Wrow = A.Wx + row * A.Nout;
  cx<T>* __restrict__ dWrow = A.dWx ? A.dWx + row * A.Nout : nullptr;
  cx<T>* __restrict__ Tb = A.Tx ? A.Tx + (long long)b * A.na * A.Nout : nullptr;
  const T mlt = (A.out_mul != nullptr) ? A.out_mul[a] : (T)1;
  T cre = (T)0; double cwide = 0.0;
  if (P.ssq) { cwide = A.cst[a]; cre = (T)cwide; }
  const int off = (int)A.out_off;
#pragma unroll 4
  for (int k = 0; k < EPT; ++k) {
    const int e = g + k * GSTEP;
    const int j = (e << A.logI2) + t2 - off;
    if (j < 0 || j >= (int)A.Nout) continue;
    cx<T> W = s[e * R2 + c];
    cx<T> dW = s[ELEMS + e * R2 + c];
    if (!P.ssq) {
      Wrow[j] = cscale<T>(W, mlt);
      if (P.write_dWx) dWrow[j] = cscale<T>(dW, mlt);
    } else {
      Wrow[j] = W;
      if (P.write_dWx) dWrow[j] = dW;
      if (is_active_fast(W.x, W.y, A.grid.gamma)) {
        int kk = bin_fused<T>(dW.x, dW.y, W.x, W.y, A.grid);
        T re, im;
        if (A.grid.const_wide) { re = (T)((double)W.x * cwide); im = (T)((double)W.y * cwide); }
        else                   { re = W.x * cre; im = W.y * cre; }
        atomic_add_cx<T>(&Tb[(long long)kk * A.Nout + j], re, im);
      }
    }
  }
}

}  // namespace ssqb
