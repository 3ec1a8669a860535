// Kernels and extern "C" entry points for the NMFk-level rows: clustering / silhouettes
// (dist_clustering.py), nnsvd initialisation (dist_svd.py).  All of these are factor- or ensemble-sized (m x k x P,
// d x d): HBM-bound elementwise / reduction work, CUDA cores only.  The m-long contractions they need
// (centroid-to-feature similarities, the (kP)^2 similarity Gram, the d x d Gram) run on the A-streaming kernels
// (dnmf_wta / dnmf_ah).
#include <math.h>

#include "common.cuh"
#include "launch_passes.cuh"

using namespace dnmf;
#define DISPATCH_T DNMF_DISPATCH_T

namespace {

inline int check_dtype(int dtype) {
  if (dtype != DNMF_F32 && dtype != DNMF_F64) return fail(DNMF_E_ARG, "dtype must be DNMF_F32 or DNMF_F64");
  return 0;
}

// ---- X[i0,i1,i2] (op)= f(s[i0*s0 + i1*s1 + i2*s2]) ------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
scale_groups_kernel(T* __restrict__ X, int64_t d1, int64_t d2, int64_t total, const T* __restrict__ s, int64_t s0, int64_t s1,
                    int64_t s2, int mode, T eps) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t i2 = idx % d2, r = idx / d2, i1 = r % d1, i0 = r / d1;
  T f = s[i0 * s0 + i1 * s1 + i2 * s2];
  T x = X[idx];
  switch (mode) {
    case 0: x = x * f; break;
    case 1: x = x / f; break;
    case 2: x = x / sqrt(f + eps); break;
    case 3: x = x * sqrt(f + eps); break;
    case 4: x = x / (f + eps); break;
    default: x = x * (f + eps); break;
  }
  X[idx] = x;
}

// ---- greedy assignment (dist_clustering.py:58-69 + change_order :49-55), one block per perturbation ------------
// D is k x (k*P): similarity of centroid r and feature c of perturbation p at D[r*ldd + c*P + p].
// order[p*k + r] = c.  Ties resolve to the smallest row-major index like np.argmax.
template <typename T>
__global__ void __launch_bounds__(256)
greedy_lsa_kernel(const T* __restrict__ D, int64_t ldd, int k, int P, int* __restrict__ order) {
  extern __shared__ unsigned char smem_raw[];
  T* X = reinterpret_cast<T*>(smem_raw);                  // k*k
  __shared__ T best_v[8];
  __shared__ int best_i[8];
  __shared__ int pick;
  const int p = blockIdx.x;
  const int kk = k * k;
  for (int e = threadIdx.x; e < kk; e += blockDim.x) {
    const int r = e / k, c = e % k;
    X[e] = D[(int64_t)r * ldd + (int64_t)c * P + p];
  }
  __syncthreads();
  const T NEG = -INFINITY;
  for (int round = 0; round < k; ++round) {
    T bv = NEG;
    int bi = 0x7fffffff;
    for (int e = threadIdx.x; e < kk; e += blockDim.x) {
      const T v = X[e];
      if (v > bv || (v == bv && e < bi)) { bv = v; bi = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { best_v[threadIdx.x >> 5] = bv; best_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (best_v[w] > bv || (best_v[w] == bv && best_i[w] < bi)) { bv = best_v[w]; bi = best_i[w]; }
      if (bi == 0x7fffffff) bi = 0;                       // everything -inf (np.argmax returns 0)
      pick = bi;
      order[p * k + bi / k] = bi % k;
    }
    __syncthreads();
    const int pr = pick / k, pc = pick % k;
    for (int e = threadIdx.x; e < k; e += blockDim.x) { X[e * k + pc] = NEG; X[pr * k + e] = NEG; }
    __syncthreads();
  }
}

// ---- out[.., r, ..] = in[.., src(p, r), ..] along axis 0 or 1 of a contiguous [d0, d1, P] tensor ------------------
// sequential = 0: src = order[p][r] (a gather, W_sub[:, j] of dist_clustering.py:81).
// sequential = 1: the result of assigning the rows one after another IN PLACE, `for r: X[r] = X[order[p][r]]`, which is
//   what `H_all[:, :, p] = [H_all[:, :, p][k] for k in j]` (dist_clustering.py:116) does under numpy >= 1.20: the list
//   holds views and rows already overwritten are read back.  src(r) = j[r] if j[r] >= r else src(j[r]).
template <typename T>
__global__ void __launch_bounds__(256)
permute_groups_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t d0, int64_t d1, int64_t P, int axis,
                      const int* __restrict__ order, int sequential) {
  const int64_t total = d0 * d1 * P;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t p = idx % P, r = idx / P, i1 = r % d1, i0 = r / d1;
  const int* j = order + p * (axis == 1 ? d1 : d0);
  int cur = (int)(axis == 1 ? i1 : i0);
  int s = j[cur];
  if (sequential)
    while (s < cur) { cur = s; s = j[cur]; }
  const int64_t src = axis == 1 ? (i0 * d1 + s) * P + p : ((int64_t)s * d1 + i1) * P + p;
  out[idx] = in[src];
}

// ---- median (and median absolute deviation) over the last axis, thread per row, P <= kMaxP ----------------------
constexpr int kMaxP = 128;
template <typename T>
__device__ __forceinline__ T median_sorted_insert(T* a, int P) {
  for (int i = 1; i < P; ++i) {
    const T v = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
  return (P & 1) ? a[P / 2] : (a[P / 2 - 1] + a[P / 2]) / (T)2;   // np.median: mean of the two middle values
}
template <typename T>
__global__ void __launch_bounds__(128)
median_last_kernel(const T* __restrict__ X, int64_t rows, int P, T* __restrict__ med, T* __restrict__ mad) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  T a[kMaxP];
  for (int i = 0; i < P; ++i) a[i] = X[row * P + i];
  const T m = median_sorted_insert(a, P);
  if (med) med[row] = m;
  if (mad) {
    for (int i = 0; i < P; ++i) a[i] = fabs(X[row * P + i] - m);
    mad[row] = median_sorted_insert(a, P);
  }
}

// ---- silhouettes from the (kP x kP) cosine Gram (dist_clustering.py:146-159), block per (cluster, perturbation) ---
template <typename T>
__global__ void __launch_bounds__(256)
silhouettes_kernel(const T* __restrict__ G, int64_t ldg, int k, int P, double* __restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  double* tmp = reinterpret_cast<double*>(smem_raw);      // k
  const int kk = blockIdx.x / P, n = blockIdx.x % P;
  const T* row = G + ((int64_t)kk * P + n) * ldg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int k2 = warp; k2 < k; k2 += nw) {
    double s = 0.0;
    for (int n2 = lane; n2 < P; n2 += 32) {
      T g = row[(int64_t)k2 * P + n2];
      g = g < (T)-1 ? (T)-1 : (g > (T)1 ? (T)1 : g);
      s += (double)acos(g);                                // arccos in the data dtype, like np.arccos
    }
    s = warp_sum(s);
    if (lane == 0) tmp[k2] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (k == 1) { out[blockIdx.x] = 1.0; return; }
    const double a = 1.0 / (P - 1) * tmp[kk];
    double mn = INFINITY;
    for (int k2 = 0; k2 < k; ++k2)
      if (k2 != kk && tmp[k2] < mn) mn = tmp[k2];
    const double b = 1.0 / P * mn;
    out[blockIdx.x] = (b - a) / fmax(a, b);
  }
}

// ---- nnsvd helpers ---------------------------------------------------------------------------------------------
// M[i][j] = (T)((double)M[i][j] - sigma * (u[i] * v[j]))          dist_svd.py:160-162
template <typename T>
__global__ void __launch_bounds__(256)
rank1_sub_kernel(T* __restrict__ M, int64_t ldm, int64_t rows, int64_t cols, const double* __restrict__ u,
                 const double* __restrict__ v, const double* __restrict__ sigma) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t i = idx / cols, j = idx % cols;
  // separate roundings like numpy (no FMA contraction): outer = u*v; t = sigma*outer; M - t
  M[i * ldm + j] = (T)__dsub_rn((double)M[i * ldm + j], __dmul_rn(sigma[0], __dmul_rn(u[i], v[j])));
}

// y[i] = sum_j A[i][j] x[j]  (float64 accumulate, A promoted), warp per row
template <typename T>
__global__ void __launch_bounds__(256)
matvec_rows_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, const double* __restrict__ x,
                   double* __restrict__ y) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  double s = 0.0;
  for (int64_t j = lane; j < cols; j += 32) s += (double)A[row * lda + j] * x[j];
  s = warp_sum(s);
  if (lane == 0) y[row] = s;
}
// partial[chunk][j] = sum_{i in chunk} A[i][j] x[i]; grid (ceil(cols/256), chunks)
template <typename T>
__global__ void __launch_bounds__(256)
matvec_cols_partial_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, int64_t rows_per_chunk,
                           const double* __restrict__ x, double* __restrict__ part) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  double s = 0.0;
  for (int64_t i = r0; i < r1; ++i) s += (double)A[i * lda + j] * x[i];
  part[(int64_t)blockIdx.y * cols + j] = s;
}
__global__ void __launch_bounds__(256)
sum_chunks_kernel(const double* __restrict__ part, int chunks, int64_t count, double* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += part[(int64_t)c * count + j];
  out[j] = s;
}

// one block: v_out = y / ||y||, r[0] = <v_out, v_last>              dist_svd.py:121-125
__global__ void __launch_bounds__(1024)
power_normalize_kernel(const double* __restrict__ y, const double* __restrict__ v_last, double* __restrict__ v_out,
                       double* __restrict__ r, int64_t d) {
  __shared__ double red[32];
  __shared__ double nrm;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < d; i += blockDim.x) s += y[i] * y[i];
  s = block_sum<1024>(s, red);
  if (threadIdx.x == 0) nrm = sqrt(s);
  __syncthreads();
  double t = 0.0;
  for (int64_t i = threadIdx.x; i < d; i += blockDim.x) {
    const double v = y[i] / nrm;
    v_out[i] = v;
    t += v * v_last[i];
  }
  t = block_sum<1024>(t, red);
  if (threadIdx.x == 0) r[0] = t;
}

// The whole power iteration of dist_svd.py:117-134 on ONE CTA for small Gram matrices (d <= 512): per step y = B v
// (warp per row, the summation order of matvec_rows_kernel), v' = y / ||y||, r = <v', v> (the order of
// power_normalize_kernel), until |r| > thr -- the same arithmetic as the launch-per-step loop, without its host round trip
// per step.  cur / nxt: 2 d doubles of scratch; v holds the start vector on entry and the result on exit.
template <typename T>
__global__ void __launch_bounds__(1024)
power_iterate_kernel(const T* __restrict__ B, int64_t ldb, int64_t d, double* __restrict__ v, double* __restrict__ scratch,
                     double thr, int cmp_f32, int max_iter, int* __restrict__ iters_out) {
  __shared__ double red[32];
  __shared__ double nrm, rdot;
  double* y = scratch;
  double* nxt = scratch + d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int it = 0;
  for (; it < max_iter;) {
    for (int64_t row = warp; row < d; row += 32) {
      double s = 0.0;
      for (int64_t j = lane; j < d; j += 32) s += (double)B[row * ldb + j] * v[j];
      s = warp_sum(s);
      if (lane == 0) y[row] = s;
    }
    __syncthreads();
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) s += y[i] * y[i];
    s = block_sum<1024>(s, red);
    if (threadIdx.x == 0) nrm = sqrt(s);
    __syncthreads();
    double t = 0.0;
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) {
      const double x = y[i] / nrm;
      nxt[i] = x;
      t += x * v[i];
    }
    t = block_sum<1024>(t, red);
    if (threadIdx.x == 0) rdot = t;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) v[i] = nxt[i];
    ++it;
    // |r| > thr, compared in float32 when the caller's threshold is a float32 scalar (numpy compares a Python float with
    // np.float32(1 - eps) in float32: the reference's stopping rule for fp32 data, dist_svd.py:126)
    const bool done = cmp_f32 ? ((float)fabs(rdot) > (float)thr) : (fabs(rdot) > thr);
    __syncthreads();
    if (done) break;
  }
  if (threadIdx.x == 0 && iters_out != nullptr) *iters_out = it;
}

// dst[i*stride] = src[i] / sqrt(sq[0])       (u = u_unnorm / sig, stored as a column; dist_svd.py:167-176)
__global__ void __launch_bounds__(256)
div_store_kernel(const double* __restrict__ src, const double* __restrict__ sq, double* __restrict__ dst, int64_t n,
                 int64_t stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i * stride] = src[i] / sqrt(sq[0]);
}

// out[j] = sum_i max(X[i][j],0)^2 ; out[k+j] = sum_i max(-X[i][j],0)^2 ; one block per column
__global__ void __launch_bounds__(256)
posneg_colsumsq_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int k, double* __restrict__ out) {
  __shared__ double red[8];
  const int j = blockIdx.x;
  double sp = 0.0, sn = 0.0;
  for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
    const double x = X[i * ldx + j];
    if (x > 0) sp += x * x;
    else if (x < 0) sn += x * x;
  }
  sp = block_sum<256>(sp, red);
  sn = block_sum<256>(sn, red);
  if (threadIdx.x == 0) { out[j] = sp; out[k + j] = sn; }
}

// out[i][j] = pos[j] ? cp[j] * max(X,0) / dp[j] : cn[j] * max(-X,0) / dn[j]     dist_svd.py:241-242
// coef = [cp | dp | cn | dn] (k each); transpose_out writes out[j][i] (the H factor)
__global__ void __launch_bounds__(256)
nnsvd_pick_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int k, const double* __restrict__ coef,
                  const int* __restrict__ pos, double* __restrict__ out, int64_t ldo, int transpose_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * k) return;
  const int64_t i = idx / k;
  const int j = (int)(idx % k);
  const double x = X[i * ldx + j];
  double v;
  if (pos[j]) v = coef[j] * (x > 0 ? x : 0.0) / coef[k + j];
  else v = coef[2 * k + j] * (x < 0 ? -x : 0.0) / coef[3 * k + j];
  if (transpose_out) out[(int64_t)j * ldo + i] = v;
  else out[i * ldo + j] = v;
}

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)ceil_div(n > 0 ? n : 1, bs); }

}  // namespace

extern "C" {

int dnmf_scale_groups(void* X, int64_t d0, int64_t d1, int64_t d2, const void* s, int64_t s0, int64_t s1, int64_t s2,
                      int mode, double eps, int dtype, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(d0 >= 0 && d1 >= 0 && d2 >= 0 && mode >= 0 && mode <= 5, "shape / mode");
  const int64_t total = d0 * d1 * d2;
  if (total == 0) return 0;
  DNMF_CHECK_ARG(X && s, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (scale_groups_kernel<T><<<blocks_for(total, 256), 256, 0, st>>>((T*)X, d1, d2, total, (const T*)s, s0, s1, s2, mode, (T)eps)));
  DNMF_LAUNCH_CHECK("scale_groups_kernel");
  return 0;
}

int dnmf_greedy_lsa(const void* D, int64_t ldd, int64_t k, int64_t P, int32_t* order, int dtype, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(k >= 1 && P >= 1 && D && order, "k, P >= 1 and non-null pointers");
  if (k > DNMF_MAX_K) return fail(DNMF_E_UNSUPPORTED, "k=%lld exceeds DNMF_MAX_K", (long long)k);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sm = (size_t)(k * k) * (dtype == DNMF_F32 ? 4 : 8);
  DISPATCH_T(dtype, (greedy_lsa_kernel<T><<<(unsigned)P, 256, sm, st>>>((const T*)D, ldd, (int)k, (int)P, order)));
  DNMF_LAUNCH_CHECK("greedy_lsa_kernel");
  return 0;
}

int dnmf_permute_groups(const void* in, void* out, int64_t d0, int64_t d1, int64_t P, int axis, const int32_t* order,
                        int sequential, int dtype, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(d0 >= 0 && d1 >= 0 && P >= 0 && (axis == 0 || axis == 1), "shape / axis");
  const int64_t total = d0 * d1 * P;
  if (total == 0) return 0;
  DNMF_CHECK_ARG(in && out && order && in != out, "null or aliased pointers");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (permute_groups_kernel<T><<<blocks_for(total, 256), 256, 0, st>>>((const T*)in, (T*)out, d0, d1, P, axis, order, sequential)));
  DNMF_LAUNCH_CHECK("permute_groups_kernel");
  return 0;
}

int dnmf_median_last(const void* X, int64_t rows, int64_t P, void* med, void* mad, int dtype, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(rows >= 0 && P >= 1, "shape");
  if (P > kMaxP) return fail(DNMF_E_UNSUPPORTED, "median over %lld > %d values", (long long)P, kMaxP);
  if (rows == 0) return 0;
  DNMF_CHECK_ARG(X && (med || mad), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (median_last_kernel<T><<<blocks_for(rows, 128), 128, 0, st>>>((const T*)X, rows, (int)P, (T*)med, (T*)mad)));
  DNMF_LAUNCH_CHECK("median_last_kernel");
  return 0;
}

int dnmf_silhouettes(const void* G, int64_t ldg, int64_t k, int64_t P, double* out, int dtype, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(k >= 1 && P >= 1 && G && out, "k, P >= 1 and non-null pointers");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (silhouettes_kernel<T><<<(unsigned)(k * P), 256, (size_t)k * sizeof(double), st>>>((const T*)G, ldg, (int)k, (int)P, out)));
  DNMF_LAUNCH_CHECK("silhouettes_kernel");
  return 0;
}

int dnmf_rank1_sub(void* M, int64_t ldm, int64_t rows, int64_t cols, const double* u, const double* v, const double* sigma,
                   int dtype, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(rows >= 0 && cols >= 0, "shape");
  if (rows * cols == 0) return 0;
  DNMF_CHECK_ARG(M && u && v && sigma, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (rank1_sub_kernel<T><<<blocks_for(rows * cols, 256), 256, 0, st>>>((T*)M, ldm, rows, cols, u, v, sigma)));
  DNMF_LAUNCH_CHECK("rank1_sub_kernel");
  return 0;
}

int64_t dnmf_matvec_workspace_bytes(int64_t rows, int64_t cols, int trans) {
  if (!trans) return 0;
  int64_t chunks = ceil_div(rows > 0 ? rows : 1, 512);
  if (chunks > 256) chunks = 256;
  return chunks * (cols > 0 ? cols : 1) * (int64_t)sizeof(double);
}

int dnmf_matvec_f64(const void* A, int64_t lda, int64_t rows, int64_t cols, const double* x, double* y, int trans,
                    int dtype, void* ws, int64_t ws_bytes, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(rows >= 0 && cols >= 0 && A && x && y, "shape / null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (!trans) {
    if (rows == 0) return 0;
    DISPATCH_T(dtype, (matvec_rows_kernel<T><<<blocks_for(rows * 32, 256), 256, 0, st>>>((const T*)A, lda, rows, cols, x, y)));
    DNMF_LAUNCH_CHECK("matvec_rows_kernel");
    return 0;
  }
  if (cols == 0) return 0;
  int64_t chunks = ceil_div(rows > 0 ? rows : 1, 512);
  if (chunks > 256) chunks = 256;
  const int64_t per = ceil_div(rows > 0 ? rows : 1, chunks);
  const int64_t need = chunks * cols * (int64_t)sizeof(double);
  if (ws == nullptr || ws_bytes < need) return fail(DNMF_E_WORKSPACE, "matvec needs %lld workspace bytes, got %lld", (long long)need, (long long)ws_bytes);
  dim3 grid(blocks_for(cols, 256), (unsigned)chunks);
  DISPATCH_T(dtype, (matvec_cols_partial_kernel<T><<<grid, 256, 0, st>>>((const T*)A, lda, rows, cols, per, x, (double*)ws)));
  DNMF_LAUNCH_CHECK("matvec_cols_partial_kernel");
  sum_chunks_kernel<<<blocks_for(cols, 256), 256, 0, st>>>((const double*)ws, (int)chunks, cols, y);
  DNMF_LAUNCH_CHECK("sum_chunks_kernel");
  return 0;
}

int dnmf_power_normalize(const double* y, const double* v_last, double* v_out, double* r, int64_t d, void* stream) {
  DNMF_CHECK_ARG(d >= 1 && y && v_last && v_out && r, "shape / null pointer");
  power_normalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(y, v_last, v_out, r, d);
  DNMF_LAUNCH_CHECK("power_normalize_kernel");
  return 0;
}

int dnmf_power_iterate(const void* B, int64_t ldb, int64_t d, double* v, double thr, int cmp_f32, int max_iter,
                       double* scratch, int* iters_out, int dtype, void* stream) {
  if (int rc = check_dtype(dtype)) return rc;
  DNMF_CHECK_ARG(d >= 1 && d <= 512 && ldb >= d && B && v && scratch && max_iter >= 1, "shape / null pointer (d <= 512)");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (power_iterate_kernel<T><<<1, 1024, 0, st>>>((const T*)B, ldb, d, v, scratch, thr, cmp_f32, max_iter, iters_out)));
  DNMF_LAUNCH_CHECK("power_iterate_kernel");
  return 0;
}

int dnmf_div_store(const double* src, const double* sq, double* dst, int64_t n, int64_t stride, void* stream) {
  DNMF_CHECK_ARG(n >= 0 && stride >= 1, "shape");
  if (n == 0) return 0;
  DNMF_CHECK_ARG(src && sq && dst, "null pointer");
  div_store_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, sq, dst, n, stride);
  DNMF_LAUNCH_CHECK("div_store_kernel");
  return 0;
}

int dnmf_posneg_colsumsq(const double* X, int64_t ldx, int64_t rows, int64_t k, double* out, void* stream) {
  DNMF_CHECK_ARG(rows >= 0 && k >= 1 && X && out, "shape / null pointer");
  posneg_colsumsq_kernel<<<(unsigned)k, 256, 0, (cudaStream_t)stream>>>(X, ldx, rows, (int)k, out);
  DNMF_LAUNCH_CHECK("posneg_colsumsq_kernel");
  return 0;
}

int dnmf_nnsvd_pick(const double* X, int64_t ldx, int64_t rows, int64_t k, const double* coef, const int32_t* pos,
                    double* out, int64_t ldo, int transpose_out, void* stream) {
  DNMF_CHECK_ARG(rows >= 0 && k >= 1, "shape");
  if (rows == 0) return 0;
  DNMF_CHECK_ARG(X && coef && pos && out, "null pointer");
  nnsvd_pick_kernel<<<blocks_for(rows * k, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, rows, (int)k, coef, pos, out, ldo, transpose_out);
  DNMF_LAUNCH_CHECK("nnsvd_pick_kernel");
  return 0;
}

}  // extern "C"
