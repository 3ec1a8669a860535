// Device-resident control of the FRO-BCD iteration (dist_nmf.py:996-1047, 2-D twin :528-579).
//
// The reference keeps the Lipschitz bounds, the objective, the extrapolation weights and the accept / restore decision
// in Python scalars, i.e. at least four host round trips per iteration.  Here they live in a small float64 state
// vector on the device; the decision `obj >= obj_old` becomes a select inside the kernels that consume it, so the
// iteration is a fixed launch sequence (capturable into a CUDA graph) with the reference's arithmetic:
//   state[0] L_W = ||H H^T||_F   [1] its previous value    [2] L_H = ||W^T W||_F   [3] its previous value
//   state[4] obj_old   [5] t_old   [6] accept (1.0 / 0.0)   [7] ww   [8] wh   [9] obj   [10] rw   [11] restores so far
// The restore branch of the reference recomputes H_old H_old^T and A H_old^T (one more pass over A, :1034-1035); both
// were computed when H_old was accepted, so the accepted copies are kept and swapped back instead.
#include "common.cuh"
#include "launch_passes.cuh"

namespace dnmf {
namespace {

__global__ void bcd_state_kernel(int phase, double* __restrict__ s, const double* __restrict__ in) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (phase == 0) {                       // initWandH (:951-969): obj_old = ||A||^2 / 2, t_old = 1, rw = 1 (:987)
    s[0] = 1.0; s[1] = 1.0; s[2] = 1.0; s[3] = 1.0;
    s[4] = 0.5 * in[0]; s[5] = 1.0; s[6] = 1.0; s[7] = 0.0; s[8] = 0.0; s[9] = 0.0; s[10] = 1.0; s[11] = 0.0;
  } else if (phase == 1) {                // L_W = ||H H^T||_F (:1000-1001)
    s[1] = s[0];
    s[0] = sqrt(in[0]);
  } else if (phase == 2) {                // L_H = ||W^T W||_F (:1016-1017)
    s[3] = s[2];
    s[2] = sqrt(in[0]);
  } else {                                // objective, acceleration weights, accept / restore (:1024-1047)
    const double obj = 0.5 * in[0];
    const double t_old = s[5];
    const double t = (1.0 + sqrt(1.0 + 4.0 * t_old * t_old)) / 2.0;
    s[9] = obj;
    if (obj >= s[4]) {
      s[6] = 0.0;
      s[11] += 1.0;
    } else {
      const double w = (t_old - 1.0) / t;
      s[6] = 1.0;
      s[7] = fmin(w, s[10] * sqrt(s[1] / s[0]));
      s[8] = fmin(w, s[10] * sqrt(s[3] / s[2]));
      s[5] = t;
      s[4] = obj;
    }
  }
}

// accept: Xm = X + w (X - X_old) evaluated as (1 + w) X - w X_old like the reference's expression order, X_old = X
// restore: Xm = X_old
template <typename T>
__global__ void bcd_advance_kernel(const T* __restrict__ X, T* __restrict__ Xm, T* __restrict__ X_old, int64_t count,
                                   const double* __restrict__ s, int which) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  if (s[6] != 0.0) {
    const double w = s[7 + which];
    const T a = (T)(1.0 + w), b = (T)(-w);
    const T x = X[idx];
    Xm[idx] = a * x + b * X_old[idx];
    X_old[idx] = x;
  } else {
    Xm[idx] = X_old[idx];
  }
}

// accept: kept <- cur ; restore: cur <- kept          (H H^T and A H^T of the last accepted H)
template <typename T>
__global__ void bcd_keep_kernel(T* __restrict__ cur, T* __restrict__ kept, int64_t count, const double* __restrict__ s) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  if (s[6] != 0.0) kept[idx] = cur[idx];
  else cur[idx] = kept[idx];
}

}  // namespace
}  // namespace dnmf

using namespace dnmf;

extern "C" {

int dnmf_bcd_state(int phase, double* state, const double* in, void* stream) {
  DNMF_CHECK_ARG(state && in && phase >= 0 && phase <= 3, "null pointer / bad phase");
  bcd_state_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(phase, state, in);
  DNMF_LAUNCH_CHECK("bcd_state_kernel");
  return 0;
}

int dnmf_bcd_advance(const void* X, void* Xm, void* X_old, int64_t count, const double* state, int which, int dtype,
                     void* stream) {
  DNMF_CHECK_ARG(X && Xm && X_old && state && count >= 0 && (which == 0 || which == 1), "null pointer / bad selector");
  DNMF_CHECK_ARG(dtype == DNMF_F32 || dtype == DNMF_F64, "dtype");
  if (count == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DNMF_DISPATCH_T(dtype, (bcd_advance_kernel<T><<<(unsigned)ceil_div(count, 256), 256, 0, st>>>((const T*)X, (T*)Xm, (T*)X_old, count, state, which)));
  DNMF_LAUNCH_CHECK("bcd_advance_kernel");
  return 0;
}

int dnmf_bcd_keep(void* cur, void* kept, int64_t count, const double* state, int dtype, void* stream) {
  DNMF_CHECK_ARG(cur && kept && state && count >= 0, "null pointer");
  DNMF_CHECK_ARG(dtype == DNMF_F32 || dtype == DNMF_F64, "dtype");
  if (count == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DNMF_DISPATCH_T(dtype, (bcd_keep_kernel<T><<<(unsigned)ceil_div(count, 256), 256, 0, st>>>((T*)cur, (T*)kept, count, state)));
  DNMF_LAUNCH_CHECK("bcd_keep_kernel");
  return 0;
}

}  // extern "C"
