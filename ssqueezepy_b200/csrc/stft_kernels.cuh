// STFT hot path (sm_100a).
//
// Replaces ssqueezepy/_stft.py:127-146 (`_stft`): `buffer` framing
// (utils/stft_utils.py:69-98) x2, window / diff-window multiply, rfft x2 -- and,
// for ssq_stft, the fused reassignment `_ssq_stft_par` (algos.py:971-984).
//
// One CTA transforms R = TILE/n_fft frames.  Both real sequences of a frame are
// packed into ONE complex transform:  c[l] = f[l]*win[l] + i*kappa*f[l]*dwin[l]
// (kappa = power of two balancing the two norms, so the float32 error of the
// small derivative spectrum is not inflated by the large one), then separated
// with the Hermitian symmetry of real-input DFTs.  The forward DFT is obtained
// from the inverse engine by conjugating input and output.
// Output layout [B][n_fft/2+1][n_hops] (frames contiguous), as the reference.
#pragma once
#include "fft_engine.cuh"
#include "cwt_kernels.cuh"   // Tile<>, atomic_add_cx, is_active_fast

namespace ssqb {

template <typename T>
struct StftArgs {
  long long N, n_hops;
  int n_fft, hop, n1, padtype, modulated;
  int B;
  const T* x;               // [B][N]
  const T* win;             // [n_fft] window, already ifftshifted when modulated
  const T* dwin;            // [n_fft] diff window * fs, same shift
  T kappa, inv_kappa;
  cx<T>* Sx; cx<T>* dSx; cx<T>* Tx;
  const T* Sfs;             // [n_fft/2+1]
  const double* cst;        // [n_fft/2+1]
  const cx<T>* tw;          // n_fft-th roots exp(+2 pi i m / n_fft)
  int write_dSx;
  ReassignGrid grid;
};

// frame sample l of frame i  ->  index into the padded signal
// (utils/stft_utils.py:85-98: modulated frames are stored ifftshifted)
__device__ __forceinline__ long long frame_src(int l, long long i, int hop, int seg_len,
                                               int modulated) {
  long long start = (long long)hop * i;
  if (!modulated) return start + l;
  int s20 = (seg_len + 1) / 2;
  int s21 = (seg_len % 2 == 1) ? s20 - 1 : s20;
  return (l < s20) ? start + s21 + l : start + (l - s20);
}

template <typename T>
__device__ __forceinline__ void stft_emit(const StftArgs<T>& A, int b, int k, long long frame,
                                          cx<T> Ck, cx<T> Cmk, bool ssq) {
  // C = FFT(c);  S = (C[k] + conj(C[M-k]))/2 ; kappa*dS = (C[k] - conj(C[M-k]))/(2i)
  T h = (T)0.5;
  cx<T> S  = mkc<T>((Ck.x + Cmk.x) * h, (Ck.y - Cmk.y) * h);
  cx<T> dS = mkc<T>((Ck.y + Cmk.y) * h * A.inv_kappa, (Cmk.x - Ck.x) * h * A.inv_kappa);
  int nrows = A.n_fft / 2 + 1;
  long long o = ((long long)b * nrows + k) * A.n_hops + frame;
  A.Sx[o] = S;
  if (A.write_dSx) A.dSx[o] = dS;
  if (ssq && is_active_exact(S.x, S.y, A.grid.gamma)) {
    double r = phase_ratio_exact<T>(dS.x, dS.y, S.x, S.y);
    double w = fabs((double)A.Sfs[k] - r);
    int kk = bin_from_w_exact(w, A.grid);
    T cc = (T)A.cst[k];
    atomic_add_cx<T>(&A.Tx[((long long)b * nrows + kk) * A.n_hops + frame], S.x * cc, S.y * cc);
  }
}

// ---- power-of-two n_fft -------------------------------------------------------
template <typename T, int LOG_M, bool SSQ>
__global__ void __launch_bounds__(Tile<T>::NT)
stft_pow2_kernel(const StftArgs<T> A) {
  constexpr int NT = Tile<T>::NT;
  constexpr int M = 1 << LOG_M;
  constexpr int R = Tile<T>::ELEMS / M;
  constexpr int STRIDE = R + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* s = reinterpret_cast<cx<T>*>(smem_raw);          // [M][STRIDE]
  cx<T>* tw = s + (size_t)M * STRIDE;                     // [M]
  const int tid = threadIdx.x;
  const long long total_frames = (long long)A.B * A.n_hops;
  const long long f0 = (long long)blockIdx.x * R;

  for (int m = tid; m < M; m += NT) tw[m] = A.tw[m];
#pragma unroll 1
  for (int lin = tid; lin < M * R; lin += NT) {
    // frames along r; each lane walks the frame samples l (stride hop between lanes)
    int r = lin % R, l = lin / R;
    long long fr = f0 + r;
    cx<T> z = mkc<T>((T)0, (T)0);
    if (fr < total_frames) {
      int b = (int)(fr / A.n_hops);
      long long i = fr - (long long)b * A.n_hops;
      long long t = frame_src(l, i, A.hop, M, A.modulated);
      long long src = pad_src_index(t, A.n1, A.N, A.padtype);
      T v = (src >= 0) ? A.x[(long long)b * A.N + src] : (T)0;
      z = mkc<T>(v * A.win[l], -(v * A.dwin[l]) * A.kappa);   // conj(c)
    }
    s[l * STRIDE + r] = z;
  }
  __syncthreads();
  block_ifft<T, LOG_M, R, NT, STRIDE>(s, tw);
  // FFT(c)[k] = conj(s[k])
#pragma unroll 1
  for (int lin = tid; lin < (M / 2 + 1) * R; lin += NT) {
    int r = lin % R, k = lin / R;
    long long fr = f0 + r;
    if (fr >= total_frames) continue;
    int b = (int)(fr / A.n_hops);
    long long i = fr - (long long)b * A.n_hops;
    cx<T> Ck = cconj<T>(s[k * STRIDE + r]);
    cx<T> Cmk = cconj<T>(s[((M - k) & (M - 1)) * STRIDE + r]);
    stft_emit<T>(A, b, k, i, Ck, Cmk, SSQ);
  }
}

// ---- any other n_fft: frames -> generic-length FFT (gfft.cuh) -> Hermitian split ------------
// c[f][l] = x_f[l] win[l] + i kappa x_f[l] dwin[l]   (frames f0 .. f0 + nf of the flattened batch)
template <typename T>
__global__ void __launch_bounds__(256)
stft_frames_kernel(const StftArgs<T> A, cx<T>* __restrict__ c, long long f0, long long nf) {
  const int M = A.n_fft;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nf * M) return;
  const long long fl = idx / M; const int l = (int)(idx - fl * M);
  const long long fr = f0 + fl;
  const int b = (int)(fr / A.n_hops);
  const long long i = fr - (long long)b * A.n_hops;
  const long long t = frame_src(l, i, A.hop, M, A.modulated);
  const long long src = pad_src_index(t, A.n1, A.N, A.padtype);
  const T v = (src >= 0) ? A.x[(long long)b * A.N + src] : (T)0;
  c[idx] = mkc<T>(v * A.win[l], (v * A.dwin[l]) * A.kappa);
}
template <typename T, bool SSQ>
__global__ void __launch_bounds__(256)
stft_emit_kernel(const StftArgs<T> A, const cx<T>* __restrict__ C, long long f0, long long nf) {
  const int M = A.n_fft, nrows = M / 2 + 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nf * nrows) return;
  const int k = (int)(idx / nf); const long long fl = idx - (long long)k * nf;   // frames fastest
  const long long fr = f0 + fl;
  const int b = (int)(fr / A.n_hops);
  const long long i = fr - (long long)b * A.n_hops;
  const cx<T> Ck = C[fl * M + k], Cmk = C[fl * M + (k ? M - k : 0)];
  stft_emit<T>(A, b, k, i, Ck, Cmk, SSQ);
}

}  // namespace ssqb
