// Whole-fit, on-chip multiplicative updates for shards that fit in one SM's shared memory (the NMFk ensemble on a
// 96 x 21 matrix, cfg5 of BASELINE.json; the reference's tests).  One CTA per fit: A, W, H are loaded into shared memory
// once and iterations [it0, it1) of PyNMF.fit's loop body (pyDNMF.py:151-172: update() + the every-10th clamp) run
// without leaving the SM; `batch` independent fits (the perturbations of the ensemble) run as `batch` CTAs of ONE launch.
// At these sizes the per-kernel path is pure launch latency (8-15 launches of 2-80 us per iteration).
//
// FRO-MU (dist_nmf.py:715-771):  W *= (A H^T) / (W (H H^T) + eps);  H *= (W^T A) / ((H^T (W^T W)) + eps)^T
// KL-MU  (dist_nmf.py:803-869):  W *= ((A / (W H + eps)) H^T) / (rowsum(H) + eps);  H *= (W^T (A / (W H + eps))) / (colsum(W) + eps)
#include "common.cuh"
#include "launch_passes.cuh"

using namespace dnmf;
#define DISPATCH_T DNMF_DISPATCH_T

namespace {

constexpr int kResThreads = 512;
constexpr int kGroup = 8;                       // lanes cooperating on one dot product
constexpr int kGroups = kResThreads / kGroup;

struct ResLayout {                              // offsets in elements of T
  int64_t A, W, H, X, vec, total;
};
__host__ __device__ inline ResLayout res_layout(int64_t m, int64_t n, int64_t k, int kl) {
  ResLayout L;
  const int64_t mk = m * k, kn = k * n, f = mk > kn ? mk : kn;
  L.A = 0;
  L.W = m * n;
  L.H = L.W + mk;
  L.X = L.H + kn;
  const int64_t x = kl ? m * n : 2 * f + k * k;   // KL: U;  FRO: [V | next factor | Gram]
  L.vec = L.X + x;
  L.total = L.vec + k + 2;
  return L;
}

// out(o) for o in [0, n_out): sum_{l < len} term(o, l), kGroup lanes per output, uniform shuffles
template <typename T, typename Term, typename Store>
__device__ __forceinline__ void grouped_dots(int n_out, int len, Term term, Store store) {
  const int g = threadIdx.x / kGroup, l0 = threadIdx.x % kGroup;
  for (int base = 0; base < n_out; base += kGroups) {
    const int o = base + g;
    T s = (T)0;
    if (o < n_out)
      for (int l = l0; l < len; l += kGroup) s += term(o, l);
#pragma unroll
    for (int d = kGroup / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (o < n_out && l0 == 0) store(o, s);
  }
}

// vec[kk] = sum over `len` elements (stride `st`) starting at X[kk*off], accumulated in float64 like the library's sums
template <typename T>
__device__ __forceinline__ void k_sums(const T* X, int k, int len, int off, int st, T* vec) {
  const int g = threadIdx.x / kGroup, l0 = threadIdx.x % kGroup;
  for (int base = 0; base < k; base += kGroups) {
    const int o = base + g;
    double s = 0.0;
    if (o < k)
      for (int l = l0; l < len; l += kGroup) s += (double)X[(int64_t)o * off + (int64_t)l * st];
#pragma unroll
    for (int d = kGroup / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (o < k && l0 == 0) vec[o] = (T)s;
  }
}

template <typename T>
__global__ void __launch_bounds__(kResThreads)
mu_fit_resident_kernel(const T* const* __restrict__ Ap, int64_t lda, T* const* __restrict__ Wp, T* const* __restrict__ Hp,
                       int m, int n, int k, int kl, int w_update, int64_t it0, int64_t it1, T eps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const ResLayout L = res_layout(m, n, k, kl);
  T* A = sm + L.A;
  T* W = sm + L.W;
  T* H = sm + L.H;
  T* X = sm + L.X;
  T* vec = sm + L.vec;
  const T* Ag = Ap[blockIdx.x];
  T* Wg = Wp[blockIdx.x];
  T* Hg = Hp[blockIdx.x];
  const int t = threadIdx.x, NT = kResThreads;
  const int mn = m * n, mk = m * k, kn = k * n;
  for (int e = t; e < mn; e += NT) A[e] = Ag[(int64_t)(e / n) * lda + (e % n)];
  for (int e = t; e < mk; e += NT) W[e] = Wg[e];
  for (int e = t; e < kn; e += NT) H[e] = Hg[e];
  __syncthreads();

  for (int64_t it = it0; it < it1; ++it) {
    if (kl) {
      T* U = X;
      if (w_update) {
        k_sums(H, k, n, n, 1, vec);                                  // x2 = H.sum(axis=1)
        for (int e = t; e < mn; e += NT) {                            // U = A / (W H + eps)
          const int i = e / n, j = e % n;
          T s = (T)0;
          for (int l = 0; l < k; ++l) s += W[i * k + l] * H[l * n + j];
          U[e] = A[e] / (s + eps);
        }
        __syncthreads();
        grouped_dots<T>(mk, n, [&](int o, int l) { return U[(o / k) * n + l] * H[(o % k) * n + l]; },
                        [&](int o, T v) { W[o] = W[o] * (v / (vec[o % k] + eps)); });
        __syncthreads();
      }
      k_sums(W, k, m, 1, k, vec);                                     // x = W.sum(axis=0)
      for (int e = t; e < mn; e += NT) {
        const int i = e / n, j = e % n;
        T s = (T)0;
        for (int l = 0; l < k; ++l) s += W[i * k + l] * H[l * n + j];
        U[e] = A[e] / (s + eps);
      }
      __syncthreads();
      grouped_dots<T>(kn, m, [&](int o, int l) { return W[l * k + (o / n)] * U[l * n + (o % n)]; },
                      [&](int o, T y) { H[o] = H[o] * (y / (vec[o / n] + eps)); });
      __syncthreads();
    } else {
      const int f = mk > kn ? mk : kn;
      T* V = X;
      T* Nx = X + f;
      T* G = X + 2 * f;
      if (w_update) {
        grouped_dots<T>(k * k, n, [&](int o, int l) { return H[(o / k) * n + l] * H[(o % k) * n + l]; },
                        [&](int o, T v) { G[o] = v; });                                       // H H^T
        grouped_dots<T>(mk, n, [&](int o, int l) { return A[(o / k) * n + l] * H[(o % k) * n + l]; },
                        [&](int o, T v) { V[o] = v; });                                       // A H^T
        __syncthreads();
        grouped_dots<T>(mk, k, [&](int o, int l) { return W[(o / k) * k + l] * G[l * k + (o % k)]; },
                        [&](int o, T d) { Nx[o] = W[o] * (V[o] / (d + eps)); });
        __syncthreads();
        for (int e = t; e < mk; e += NT) W[e] = Nx[e];
        __syncthreads();
      }
      grouped_dots<T>(k * k, m, [&](int o, int l) { return W[l * k + (o / k)] * W[l * k + (o % k)]; },
                      [&](int o, T v) { G[o] = v; });                                         // W^T W
      grouped_dots<T>(kn, m, [&](int o, int l) { return W[l * k + (o / n)] * A[l * n + (o % n)]; },
                      [&](int o, T v) { V[o] = v; });                                         // W^T A
      __syncthreads();
      grouped_dots<T>(kn, k, [&](int o, int l) { return H[l * n + (o % n)] * G[l * k + (o / n)]; },
                      [&](int o, T d) { Nx[o] = H[o] * (V[o] / (d + eps)); });
      __syncthreads();
      for (int e = t; e < kn; e += NT) H[e] = Nx[e];
      __syncthreads();
    }
    if (it % 10 == 0) {                                                // pyDNMF.py:155-157
      for (int e = t; e < kn; e += NT) H[e] = H[e] > eps ? H[e] : eps;
      for (int e = t; e < mk; e += NT) W[e] = W[e] > eps ? W[e] : eps;
      __syncthreads();
    }
  }
  for (int e = t; e < mk; e += NT) Wg[e] = W[e];
  for (int e = t; e < kn; e += NT) Hg[e] = H[e];
}

inline int64_t res_bytes(int64_t m, int64_t n, int64_t k, int kl, int dtype) {
  return res_layout(m, n, k, kl).total * (dtype == DNMF_F32 ? 4 : 8);
}

}  // namespace

extern "C" {

int64_t dnmf_mu_fit_resident_smem_bytes(int64_t m, int64_t n, int64_t k, int kl, int dtype) {
  if (m < 1 || n < 1 || k < 1 || k > DNMF_MAX_K || (dtype != DNMF_F32 && dtype != DNMF_F64)) return -1;
  if (m * n > (1 << 20)) return -1;
  const int64_t b = res_bytes(m, n, k, kl, dtype);
  return b <= 227 * 1024 ? b : -1;
}

int dnmf_mu_fit_resident(const void* const* A_ptrs, int64_t lda, void* const* W_ptrs, void* const* H_ptrs, int64_t batch,
                         int64_t m, int64_t n, int64_t k, int kl, int w_update, int64_t it_begin, int64_t it_end,
                         double eps, int dtype, void* stream) {
  if (dtype != DNMF_F32 && dtype != DNMF_F64) return fail(DNMF_E_ARG, "dtype must be DNMF_F32 or DNMF_F64");
  DNMF_CHECK_ARG(batch >= 0 && A_ptrs && W_ptrs && H_ptrs && it_begin <= it_end, "batch / null pointer / iteration range");
  const int64_t bytes = dnmf_mu_fit_resident_smem_bytes(m, n, k, kl, dtype);
  if (bytes < 0) return fail(DNMF_E_UNSUPPORTED, "a %lld x %lld, k=%lld fit does not fit in shared memory", (long long)m, (long long)n, (long long)k);
  if (batch == 0 || it_begin == it_end) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DNMF_F32) {
    auto kern = mu_fit_resident_kernel<float>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "mu_fit_resident smem attribute");
    kern<<<(unsigned)batch, kResThreads, (size_t)bytes, st>>>((const float* const*)A_ptrs, lda, (float* const*)W_ptrs, (float* const*)H_ptrs,
                                                             (int)m, (int)n, (int)k, kl, w_update, it_begin, it_end, (float)eps);
  } else {
    auto kern = mu_fit_resident_kernel<double>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "mu_fit_resident smem attribute");
    kern<<<(unsigned)batch, kResThreads, (size_t)bytes, st>>>((const double* const*)A_ptrs, lda, (double* const*)W_ptrs, (double* const*)H_ptrs,
                                                             (int)m, (int)n, (int)k, kl, w_update, it_begin, it_end, eps);
  }
  DNMF_LAUNCH_CHECK("mu_fit_resident_kernel");
  return 0;
}

}  // extern "C"
