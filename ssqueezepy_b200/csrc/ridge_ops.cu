// Ridge extraction on the device: forward-backward penalised ridge tracking of
// ssqueezepy/ridge_extraction.py:11-232 (`extract_ridges`), the main consumer of Tx / Wx.
// Returning ridges instead of planes turns the 768 MB a 160k-sample ssq_cwt produces into
// N x n_ridges indices.
//
//   energy  = |Tf|^2                                            (ridge_extraction.py:124)
//   e       = -log(energy / max_f energy + eps)                 (:135-136)
//   forward:  pen[f, t] = e[f, t] + min_g (pen[g, t-1] + P[f, g]),  P = penalty (ls_f - ls_g)^2
//             (:178-189), ridge_fw[t] = argmin_f pen[f, t] (first minimum, :160-162)
//   backward: for t = N-2 .. 0: val = pen[r, t+1] - e[r, t+1] (r = ridge[t+1]); every f with
//             |val - (pen[f, t] + P[r, f])| < eps overwrites ridge[t] in ascending order, i.e.
//             the LAST such f wins (:211-219, the serial kernel; the reference's prange
//             variant races between those f)
//   then energy[ridge - bw : ridge + bw, t] = 0 with Python slice semantics (a negative start
//   counts from the end, :146-148) and the next ridge is tracked on what is left.
// All arithmetic in the data's real dtype, each operation rounded separately (*_rn), as NumPy /
// numba do; `ls` (log of the scales) and the scalars come from the host so that they are the
// host's NumPy values.  Planes are held time-major ([N][na]) so that every step of the two
// sequential sweeps reads and writes contiguous memory.
#include "host_common.h"
#include "ssq_common.cuh"
#include <vector>

namespace ssqb {

template <typename T> __device__ __forceinline__ T t_logr(T x);
template <> __device__ __forceinline__ float  t_logr<float>(float x)   { return logf(x); }
template <> __device__ __forceinline__ double t_logr<double>(double x) { return log(x); }
template <typename T> __device__ __forceinline__ T t_absc(T x, T y);
template <> __device__ __forceinline__ float t_absc<float>(float x, float y) {
  return (float)sqrt((double)x * (double)x + (double)y * (double)y);     // correctly rounded hypot
}
template <> __device__ __forceinline__ double t_absc<double>(double x, double y) { return hypot(x, y); }
template <typename T> __device__ __forceinline__ T t_inf_();
template <> __device__ __forceinline__ float  t_inf_<float>()  { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double t_inf_<double>() { return __longlong_as_double(0x7ff0000000000000ll); }

// energy[f][t] = |Tf[f][t]|^2   (plane [na][N], row-major like Tf)
template <typename T>
__global__ void __launch_bounds__(256)
ridge_energy_kernel(const cx<T>* __restrict__ Tf, T* __restrict__ energy, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const cx<T> v = Tf[i];
  const T a = t_absc<T>(v.x, v.y);
  energy[i] = mul_rn(a, a);
}
// eT[t][f] = -log(energy[f][t] / max_f energy[., t] + eps); one thread per column
template <typename T>
__global__ void __launch_bounds__(128)
ridge_neglog_kernel(const T* __restrict__ energy, T* __restrict__ eT, int na, long long N, T eps) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  T mx = energy[t];
  for (int f = 1; f < na; ++f) { const T v = energy[(long long)f * N + t]; mx = v > mx ? v : mx; }
  for (int f = 0; f < na; ++f) {
    const T q = energy[(long long)f * N + t] / mx;                       // IEEE division
    eT[t * na + f] = -t_logr<T>(add_rn(q, eps));
  }
}

// forward sweep, one CTA per plane; penT[t][f]; ridge_fw[t] = first argmin_f penT[t][f]
template <typename T>
__global__ void __launch_bounds__(1024)
ridge_forward_kernel(const T* __restrict__ eT, T* __restrict__ penT, long long* __restrict__ ridge,
                     const T* __restrict__ ls, int na, long long N, T penalty, int parts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* prev = reinterpret_cast<T*>(smem_raw);           // [na]
  T* lss = prev + na;                                 // [na]
  T* part = lss + na;                                 // [parts][na]
  int* amin = reinterpret_cast<int*>(part + (size_t)parts * na);   // [32] warp argmins
  T* vmin = reinterpret_cast<T*>(amin + 32);                       // [32]
  const int tid = threadIdx.x, NT = blockDim.x;
  const long long plane = blockIdx.x;
  eT += plane * N * na; penT += plane * N * na; ridge += plane * N;
  for (int f = tid; f < na; f += NT) { lss[f] = ls[f]; const T v = eT[f]; prev[f] = v; penT[f] = v; }
  __syncthreads();
  const int chunk = (na + parts - 1) / parts;
  for (long long t = 0; t < N; ++t) {
    // first minimum of prev[] = penalised energy of column t (argmin for the forward ridge)
    {
      T bv = t_inf_<T>(); int bi = 0x7fffffff;
      for (int f = tid; f < na; f += NT) { const T v = prev[f]; if (v < bv || (v == bv && f < bi)) { bv = v; bi = f; } }
      for (int o = 16; o; o >>= 1) {
        const T ov = __shfl_down_sync(0xffffffffu, bv, o); const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if ((tid & 31) == 0) { vmin[tid >> 5] = bv; amin[tid >> 5] = bi; }
      __syncthreads();
      if (tid == 0) {
        T b = vmin[0]; int i = amin[0];
        for (int w = 1; w < (NT + 31) / 32; ++w) if (vmin[w] < b || (vmin[w] == b && amin[w] < i)) { b = vmin[w]; i = amin[w]; }
        ridge[t] = (i == 0x7fffffff) ? 0 : i;        // all-NaN column: numpy's argmin gives 0
      }
    }
    if (t + 1 >= N) break;
    // partial minima over g of prev[g] + penalty * (ls_f - ls_g)^2
    for (int w = tid; w < parts * na; w += NT) {
      const int p = w / na, f = w - p * na;
      const int g0 = p * chunk, g1 = (g0 + chunk < na) ? g0 + chunk : na;
      const T lf = lss[f];
      T m = t_inf_<T>();
      for (int g = g0; g < g1; ++g) {
        const T dlt = sub_rn(lf, lss[g]);
        const T v = add_rn(prev[g], mul_rn(penalty, mul_rn(dlt, dlt)));
        m = (v < m || v != v) ? v : m;                // NaN propagates like np.amin
      }
      part[w] = m;
    }
    __syncthreads();
    const T* en = eT + (t + 1) * na;
    T* pn = penT + (t + 1) * na;
    for (int f = tid; f < na; f += NT) {
      T m = part[f];
      for (int p = 1; p < parts; ++p) { const T v = part[p * na + f]; m = (v < m || v != v) ? v : m; }
      const T v = add_rn(en[f], m);
      pn[f] = v;
      prev[f] = v;                                    // safe: every reader of prev[] is past the barrier
    }
    __syncthreads();
  }
}

// backward sweep, one CTA per plane
template <typename T>
__global__ void __launch_bounds__(512)
ridge_backward_kernel(const T* __restrict__ eT, const T* __restrict__ penT, long long* __restrict__ ridge,
                      const T* __restrict__ ls, int na, long long N, T penalty, T eps) {
  __shared__ int best[16];
  __shared__ int cur;
  const int tid = threadIdx.x, NT = blockDim.x;
  const long long plane = blockIdx.x;
  eT += plane * N * na; penT += plane * N * na; ridge += plane * N;
  if (tid == 0) cur = (int)ridge[N - 1];
  __syncthreads();
  for (long long t = N - 2; t >= 0; --t) {
    const int r = cur;
    const T val = sub_rn(penT[(t + 1) * na + r], eT[(t + 1) * na + r]);
    const T lr = ls[r];
    int b = -1;
    for (int f = tid; f < na; f += NT) {
      const T dlt = sub_rn(lr, ls[f]);
      const T np_ = mul_rn(penalty, mul_rn(dlt, dlt));
      const T df = sub_rn(val, add_rn(penT[t * na + f], np_));
      if (fabs(df) < eps) b = f;                      // ascending f within a thread: last wins
    }
    for (int o = 16; o; o >>= 1) { const int ob = __shfl_down_sync(0xffffffffu, b, o); b = ob > b ? ob : b; }
    if ((tid & 31) == 0) best[tid >> 5] = b;
    __syncthreads();
    if (tid == 0) {
      int m = best[0];
      for (int w = 1; w < (NT + 31) / 32; ++w) m = best[w] > m ? best[w] : m;
      if (m >= 0) { ridge[t] = m; cur = m; } else cur = (int)ridge[t];
    }
    __syncthreads();
  }
}

// ridge_f / ridge_e of this ridge, then energy[r - bw : r + bw, t] = 0 (Python slice semantics)
template <typename T>
__global__ void __launch_bounds__(256)
ridge_finish_kernel(T* __restrict__ energy, const long long* __restrict__ ridge,
                    long long* __restrict__ out_idx, T* __restrict__ out_f, T* __restrict__ out_e,
                    const T* __restrict__ scales, int na, long long N, int bw, int n_ridges, int i) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  const long long plane = blockIdx.y;
  energy += plane * na * N; ridge += plane * N;
  const int r = (int)ridge[t];
  const long long o = (plane * N + t) * n_ridges + i;
  out_idx[o] = r;
  if (out_f) out_f[o] = scales[r];
  if (out_e) out_e[o] = energy[(long long)r * N + t];
  long long a = (long long)r - bw, z = (long long)r + bw;
  if (a < 0) { a += na; if (a < 0) a = 0; }
  if (z < 0) { z += na; if (z < 0) z = 0; }
  if (a > na) a = na;
  if (z > na) z = na;
  for (long long f = a; f < z; ++f) energy[f * N + t] = (T)0;
}

template <typename T>
static int extract_ridges_t(const void* Tf, long long B, int na, long long N, const double* ls_host,
                            const double* scales_host, double penalty, double eps, int n_ridges, int bw,
                            long long* idx_out, void* f_out, void* e_out, cudaStream_t st) {
  const size_t plane = (size_t)na * (size_t)N;
  DevBuf<T> energy, eT, penT, ls_d, sc_d;
  DevBuf<long long> ridge;
  SSQB_CUDA(energy.ensure(plane * B)); SSQB_CUDA(eT.ensure(plane * B)); SSQB_CUDA(penT.ensure(plane * B));
  SSQB_CUDA(ridge.ensure((size_t)N * B));
  std::vector<T> ls((size_t)na), sc((size_t)na);
  for (int f = 0; f < na; ++f) { ls[f] = (T)ls_host[f]; sc[f] = (T)scales_host[f]; }
  SSQB_CUDA(cudaStreamSynchronize(st));
  SSQB_CUDA(ls_d.upload(ls)); SSQB_CUDA(sc_d.upload(sc));
  const long long total = (long long)plane * B;
  ridge_energy_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const cx<T>*)Tf, energy.p, total);
  SSQB_LAUNCH_CHECK();
  int parts = 1024 / (na > 0 ? na : 1); if (parts < 1) parts = 1; if (parts > 8) parts = 8;
  const size_t smem = ((size_t)(2 + parts) * na) * sizeof(T) + 32 * sizeof(int) + 32 * sizeof(T) + 16;
  if (smem > (size_t)200 * 1024) return set_error(SSQB_E_UNSUPP, "too many rows (%d) for ridge tracking", na);
  SSQB_CUDA(cudaFuncSetAttribute(ridge_forward_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int i = 0; i < n_ridges; ++i) {
    for (long long b = 0; b < B; ++b) {
      ridge_neglog_kernel<T><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(energy.p + b * plane, eT.p + b * plane, na, N, (T)eps);
      SSQB_LAUNCH_CHECK();
    }
    ridge_forward_kernel<T><<<(unsigned)B, 1024, smem, st>>>(eT.p, penT.p, ridge.p, ls_d.p, na, N, (T)penalty, parts);
    SSQB_LAUNCH_CHECK();
    ridge_backward_kernel<T><<<(unsigned)B, 512, 0, st>>>(eT.p, penT.p, ridge.p, ls_d.p, na, N, (T)penalty, (T)eps);
    SSQB_LAUNCH_CHECK();
    ridge_finish_kernel<T><<<dim3((unsigned)((N + 255) / 256), (unsigned)B), 256, 0, st>>>(
        energy.p, ridge.p, idx_out, (T*)f_out, (T*)e_out, sc_d.p, na, N, bw, n_ridges, i);
    SSQB_LAUNCH_CHECK();
  }
  SSQB_CUDA(cudaStreamSynchronize(st));               // the scratch planes die with this call
  return 0;
}

int run_extract_ridges(int dtype, const void* Tf, long long B, int na, long long N, const double* ls_host,
                       const double* scales_host, double penalty, double eps, int n_ridges, int bw,
                       long long* idx_out, void* f_out, void* e_out, cudaStream_t st) {
  if (!Tf || !ls_host || !scales_host || !idx_out) return set_error(SSQB_E_ARG, "null argument");
  if (B < 1 || na < 1 || N < 1 || n_ridges < 1 || bw < 0) return set_error(SSQB_E_ARG, "bad shape");
  return dtype == SSQB_F32 ? extract_ridges_t<float>(Tf, B, na, N, ls_host, scales_host, penalty, eps, n_ridges, bw, idx_out, f_out, e_out, st)
                           : extract_ridges_t<double>(Tf, B, na, N, ls_host, scales_host, penalty, eps, n_ridges, bw, idx_out, f_out, e_out, st);
}

}  // namespace ssqb
