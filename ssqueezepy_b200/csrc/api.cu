// extern "C" surface of libssq_b200.so -- see include/ssq_b200.h for the contract.
#include "host_common.h"
#include <cstring>

namespace ssqb {
thread_local std::string g_last_error;
std::atomic<long long> g_launch_count{0};
}
using namespace ssqb;

struct ssqb_cwt_plan { CwtPlanBase* impl; int dtype; };

extern "C" {

const char* ssqb_version(void) { return "ssq_b200 0.1.0 (sm_100a)"; }
const char* ssqb_last_error(void) { return g_last_error.c_str(); }
long long ssqb_launch_count(void) { return g_launch_count.load(); }

int ssqb_device_check(char* name, int name_len) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(SSQB_E_NODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return set_error(SSQB_E_NODEVICE, "%s", cudaGetErrorString(e));
  if (name && name_len > 0) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
  if (p.major != 10)
    return set_error(SSQB_E_NODEVICE, "device %s is sm_%d%d; this library is built for sm_100a only",
                     p.name, p.major, p.minor);
  return 0;
}

int ssqb_cwt_plan_create(const ssqb_cwt_desc* d, ssqb_cwt_plan** out) {
  if (!d || !out || !d->scales_host) return set_error(SSQB_E_ARG, "null descriptor");
  if (d->dtype != SSQB_F32 && d->dtype != SSQB_F64) return set_error(SSQB_E_ARG, "bad dtype");
  int err = 0;
  CwtPlanBase* impl = d->dtype == SSQB_F32 ? make_cwt_plan_f32(d, &err) : make_cwt_plan_f64(d, &err);
  if (!impl) return err ? err : set_error(SSQB_E_ARG, "plan creation failed");
  *out = new ssqb_cwt_plan{impl, d->dtype};
  return 0;
}

int ssqb_cwt_plan_destroy(ssqb_cwt_plan* p) {
  if (!p) return 0;
  delete p->impl; delete p;
  return 0;
}

int ssqb_cwt_plan_set_reassign(ssqb_cwt_plan* p, const ssqb_reassign_desc* r) {
  if (!p || !r || !r->cst_host) return set_error(SSQB_E_ARG, "null argument");
  return p->impl->set_reassign(r);
}

int ssqb_cwt_exec(ssqb_cwt_plan* p, const void* x, int64_t B, void* Wx, void* dWx,
                  const double* out_mul_host, int rpadded, void* stream) {
  if (!p) return set_error(SSQB_E_ARG, "null plan");
  return p->impl->exec(x, B, Wx, dWx, nullptr, false, out_mul_host, rpadded != 0, (cudaStream_t)stream);
}

int ssqb_ssq_cwt_exec(ssqb_cwt_plan* p, const void* x, int64_t B, void* Wx, void* Tx, void* dWx,
                      void* stream) {
  if (!p) return set_error(SSQB_E_ARG, "null plan");
  return p->impl->exec(x, B, Wx, dWx, Tx, true, nullptr, false, (cudaStream_t)stream);
}

int ssqb_cwt_exec_host(ssqb_cwt_plan* p, const void* x, int64_t B, void* Wx, void* dWx,
                       const double* out_mul_host, int rpadded, void* stream) {
  if (!p || !x || !Wx) return set_error(SSQB_E_ARG, "null argument");
  return p->impl->exec_host(x, B, Wx, dWx, nullptr, false, out_mul_host, rpadded != 0, (cudaStream_t)stream);
}

int ssqb_ssq_cwt_exec_host(ssqb_cwt_plan* p, const void* x, int64_t B, void* Wx, void* Tx,
                           void* dWx, void* stream) {
  if (!p || !x || !Wx || !Tx) return set_error(SSQB_E_ARG, "null argument");
  return p->impl->exec_host(x, B, Wx, dWx, Tx, true, nullptr, false, (cudaStream_t)stream);
}

int ssqb_cwt_debug_xh(ssqb_cwt_plan* p, const void* x, int64_t B, void* xh, void* stream) {
  if (!p || !x || !xh) return set_error(SSQB_E_ARG, "null argument");
  return p->impl->debug_xh(x, B, xh, (cudaStream_t)stream);
}

int ssqb_cwt_backward(ssqb_cwt_plan* p, const void* gWx, const void* gdWx, int64_t B,
                      const double* out_mul_host, int rpadded, void* gx, void* stream) {
  if (!p || !gx || (!gWx && !gdWx)) return set_error(SSQB_E_ARG, "null argument");
  return p->impl->backward(gWx, gdWx, B, out_mul_host, rpadded != 0, gx, (cudaStream_t)stream);
}

int ssqb_cwt_plan_set_profiling(ssqb_cwt_plan* p, int on) {
  if (!p) return set_error(SSQB_E_ARG, "null plan");
  return p->impl->set_profiling(on);
}

int ssqb_cwt_plan_get_profile(ssqb_cwt_plan* p, double* ms, long long* launches, long long* rows) {
  if (!p || !ms || !launches || !rows) return set_error(SSQB_E_ARG, "null argument");
  return p->impl->get_profile(ms, launches, rows);
}

int ssqb_ssqueeze(int dtype, const void* Wx, const void* dWx, void* Tx, int64_t B, int na,
                  int64_t N, const ssqb_reassign_desc* r, const void* Sfs, void* stream) {
  return run_ssqueeze(dtype, Wx, dWx, Tx, B, na, N, r, Sfs, (cudaStream_t)stream);
}

int ssqb_indexed_sum(int dtype, const void* Wx, const void* w, void* Tx, int64_t B, int na,
                     int64_t N, const ssqb_reassign_desc* r, void* stream) {
  return run_indexed_sum(dtype, Wx, w, Tx, B, na, N, r, (cudaStream_t)stream);
}

int ssqb_phase_cwt(int dtype, const void* Wx, const void* dWx, void* w, int64_t total,
                   double gamma, void* stream) {
  return run_phase(dtype, false, Wx, dWx, nullptr, w, total, 1, 1, gamma, (cudaStream_t)stream);
}

int ssqb_phase_stft(int dtype, const void* Sx, const void* dSx, const void* Sfs, void* w,
                    int64_t B, int nrows, int64_t ncols, double gamma, void* stream) {
  return run_phase(dtype, true, Sx, dSx, Sfs, w, (long long)B * nrows * ncols, ncols, nrows,
                   gamma, (cudaStream_t)stream);
}

int ssqb_stft_exec(const ssqb_stft_desc* d, const void* x, int64_t B, void* Sx, void* dSx,
                   void* stream) {
  return run_stft(d, nullptr, x, B, Sx, nullptr, dSx, false, (cudaStream_t)stream);
}

int ssqb_ssq_stft_exec(const ssqb_stft_desc* d, const ssqb_reassign_desc* r, const void* x,
                       int64_t B, void* Sx, void* Tx, void* dSx, void* stream) {
  return run_stft(d, r, x, B, Sx, Tx, dSx, true, (cudaStream_t)stream);
}

int ssqb_ssq_stft_exec_host(const ssqb_stft_desc* d, const ssqb_reassign_desc* r, const void* x,
                            int64_t B, void* Sx, void* Tx, void* dSx, void* stream) {
  if (!d || !x || !Sx || !Tx) return set_error(SSQB_E_ARG, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  size_t es = d->dtype == SSQB_F32 ? 4 : 8;
  long long n_hops = (d->N - 1) / d->hop + 1;
  size_t nx = (size_t)B * (size_t)d->N * es;
  size_t nout = (size_t)B * (size_t)(d->n_fft / 2 + 1) * (size_t)n_hops * 2 * es;
  void *xd = nullptr, *Sd = nullptr, *Td = nullptr, *dSd = nullptr;
  // every exit frees what was staged (stream-ordered), also on the error paths
  auto release = [&]() {
    if (xd) cudaFreeAsync(xd, st);
    if (Sd) cudaFreeAsync(Sd, st);
    if (Td) cudaFreeAsync(Td, st);
    if (dSd) cudaFreeAsync(dSd, st);
    xd = Sd = Td = dSd = nullptr;
  };
  auto fail = [&](cudaError_t e, const char* what) {
    release();
    cudaStreamSynchronize(st);
    return set_error((int)e, "%s failed: %s", what, cudaGetErrorString(e));
  };
  cudaError_t e;
  if ((e = cudaMallocAsync(&xd, nx, st)) != cudaSuccess) return fail(e, "cudaMallocAsync(x)");
  if ((e = cudaMallocAsync(&Sd, nout, st)) != cudaSuccess) return fail(e, "cudaMallocAsync(Sx)");
  if ((e = cudaMallocAsync(&Td, nout, st)) != cudaSuccess) return fail(e, "cudaMallocAsync(Tx)");
  if (dSx && (e = cudaMallocAsync(&dSd, nout, st)) != cudaSuccess) return fail(e, "cudaMallocAsync(dSx)");
  if ((e = cudaMemcpyAsync(xd, x, nx, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e, "H2D copy");
  int rc = run_stft(d, r, xd, B, Sd, Td, dSd, true, st);
  if (rc == 0) {
    if ((e = cudaMemcpyAsync(Sx, Sd, nout, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(e, "D2H copy");
    if ((e = cudaMemcpyAsync(Tx, Td, nout, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(e, "D2H copy");
    if (dSx && (e = cudaMemcpyAsync(dSx, dSd, nout, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
      return fail(e, "D2H copy");
  }
  release();
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess && rc == 0)
    return set_error((int)e, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
  return rc;
}

int ssqb_colsum_real(int dtype, int wide, const void* M, int64_t B, int na, int64_t N,
                     const double* div_host, double scale, int has_scale, void* out,
                     void* stream) {
  return run_colsum_real(dtype, wide, M, B, na, N, div_host, scale, has_scale, out,
                         (cudaStream_t)stream);
}

int ssqb_invert_components(int dtype, const void* M, int na, int64_t N, const int32_t* cc,
                           const int32_t* cw, int K, double scale, double* out, void* stream) {
  return run_invert_components(dtype, M, na, N, cc, cw, K, scale, out, (cudaStream_t)stream);
}

int ssqb_istft_exec(const ssqb_istft_desc* d, const void* Sx, int64_t B, void* x, void* stream) {
  return run_istft(d, Sx, B, x, (cudaStream_t)stream);
}

int ssqb_extract_ridges(int dtype, const void* Tf, int64_t B, int na, int64_t N, const double* ls_host,
                        const double* scales_host, double penalty, double eps, int n_ridges, int bw,
                        int64_t* idx_dev, void* f_dev, void* e_dev, void* stream) {
  return run_extract_ridges(dtype, Tf, B, na, N, ls_host, scales_host, penalty, eps, n_ridges, bw,
                            (long long*)idx_dev, f_dev, e_dev, (cudaStream_t)stream);
}

}  // extern "C"
