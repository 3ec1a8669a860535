// Whole-fit, on-chip multiplicative updates for shards that fit in shared memory.  A, W, H are loaded once and iterations
// [it0, it1) of PyNMF.fit's loop body (pyDNMF.py:151-172: update() + the every-10th clamp) run without leaving the chip;
// `batch` independent fits (the perturbations of the NMFk ensemble) are the grid of ONE launch.
//   * one CTA per fit: the NMFk example (96 x 21, cfg5 of BASELINE.json), the reference's test matrices;
//   * one thread-block cluster (2..16 CTAs, distributed shared memory) per fit: shards of a few MB (cfg1: swim 1024 x 256).
// At these sizes the per-kernel path is pure launch latency (8-15 launches of 2-80 us per iteration).
//
// FRO-MU (dist_nmf.py:715-771):  W *= (A H^T) / (W (H H^T) + eps);  H *= (W^T A) / ((H^T (W^T W)) + eps)^T
// KL-MU  (dist_nmf.py:803-869):  W *= ((A / (W H + eps)) H^T) / (rowsum(H) + eps);  H *= (W^T (A / (W H + eps))) / (colsum(W) + eps)
#include <cooperative_groups.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "launch_passes.cuh"

using namespace dnmf;
namespace cg = cooperative_groups;
#define DISPATCH_T DNMF_DISPATCH_T

namespace {

constexpr int kResThreads = 512;
constexpr int kMaxCluster = 16;

// Shared-memory layout of one CTA (offsets in elements of T).  Row strides are forced odd (ns, ks) so that the strided
// accesses of the dot-product loops (8 lanes walking a column, 4 groups per warp walking neighbouring rows) fall into
// distinct banks.  `rows` = rows of A / W held by this CTA (all m for the single-CTA variant).
struct ResLayout {
  int64_t ns, ks, A, W, H, X, part, red, vec, total;
};
__host__ __device__ inline ResLayout res_layout(int64_t rows, int64_t n, int64_t k, int kl, int cluster) {
  ResLayout L;
  L.ns = n | 1;
  L.ks = k | 1;
  const int64_t mk = rows * k, kn = k * n, f = mk > kn ? mk : kn;
  L.A = 0;
  L.W = rows * L.ns;
  L.H = L.W + rows * L.ks;
  L.X = L.H + k * L.ns;
  L.part = L.X + (kl ? rows * L.ns : 2 * f + k * k);     // KL: U;  FRO: [V | next factor | Gram]
  L.red = L.part + kn + k * k;                            // [W^T A partial | W^T W or colsum(W) partial]
  L.vec = cluster ? L.red + kn + k * k : L.red;           // single CTA: the partials ARE the sums (red aliases part)
  if (!cluster) L.red = L.part;
  L.total = L.vec + k + 2;
  return L;
}

// a / d for the m x n sized division of the KL update (d = W H + eps >= eps > 0): reciprocal seed + one Newton step +
// one residual correction, all on the FMA pipe after a single MUFU -- within 1 ulp of the IEEE quotient, exact zeros
// stay exact zeros.  float64 uses the IEEE division.
__device__ __forceinline__ float fast_div(float a, float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = fmaf(fmaf(-d, r, 1.0f), r, r);
  const float q = a * r;
  return fmaf(fmaf(-d, q, a), r, q);
}
__device__ __forceinline__ double fast_div(double a, double d) { return a / d; }

// lanes cooperating on one output: all of a warp for a handful of outputs, one thread per output once there are at
// least as many outputs as threads (then consecutive threads walk consecutive outputs: conflict-free row reads)
__device__ __forceinline__ int lanes_per_output(int n_out) {
  int g = 32;
  while (g > 1 && n_out * g > kResThreads) g >>= 1;
  return g;
}

// out(o) for o in [0, n_out): sum_{l < len} term(o, l); G lanes per output, shuffles executed by whole warps
template <typename T, typename Term, typename Store>
__device__ __forceinline__ void grouped_dots(int n_out, int len, Term term, Store store) {
  const int G = lanes_per_output(n_out);
  const int groups = kResThreads / G;
  const int g = threadIdx.x / G, l0 = threadIdx.x % G;
  for (int base = 0; base < n_out; base += groups) {
    const int o = base + g;
    T s = (T)0;
    if (o < n_out) {
#pragma unroll 4
      for (int l = l0; l < len; l += G) s += term(o, l);
    }
    for (int d = G >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (o < n_out && l0 == 0) store(o, s);
  }
}

// vec[kk] = sum over `len` elements (stride `st`) starting at X[kk*off], accumulated in float64 like the library's sums
template <typename T>
__device__ __forceinline__ void k_sums(const T* X, int k, int len, int off, int st, T* vec) {
  const int G = lanes_per_output(k);
  const int groups = kResThreads / G;
  const int g = threadIdx.x / G, l0 = threadIdx.x % G;
  for (int base = 0; base < k; base += groups) {
    const int o = base + g;
    double s = 0.0;
    if (o < k)
      for (int l = l0; l < len; l += G) s += (double)X[(int64_t)o * off + (int64_t)l * st];
    for (int d = G >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (o < k && l0 == 0) vec[o] = (T)s;
  }
}

// One fit per CTA (CLUSTER = false) or per thread-block cluster (CLUSTER = true).  In a cluster the C CTAs are a
// "C x 1 processor grid" in miniature: CTA c keeps rows [c*rpc, (c+1)*rpc) of A and W and a replica of H.  The W
// half-step is local; the H half-step needs W^T W | colsum(W) and W^T A summed over the row blocks: every CTA publishes
// its partials in shared memory, cluster.sync(), every CTA adds the C partials in rank order through distributed shared
// memory (identical result everywhere, so the H replicas never diverge), cluster.sync() before the buffers are reused.
template <typename T, bool CLUSTER>
__global__ void __launch_bounds__(kResThreads)
mu_fit_onchip_kernel(const T* const* __restrict__ Ap, int64_t lda, T* const* __restrict__ Wp, T* const* __restrict__ Hp,
                     int m, int n, int k, int kl, int w_update, int64_t it0, int64_t it1, T eps, int rpc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int C = 1, c = 0;
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    C = (int)cluster.num_blocks();
    c = (int)cluster.block_rank();
  }
  const int fit = blockIdx.x / C;
  T* sm = reinterpret_cast<T*>(smem_raw);
  const ResLayout L = res_layout(rpc, n, k, kl, CLUSTER ? 1 : 0);
  const int ns = (int)L.ns, ks = (int)L.ks;
  T* A = sm + L.A;
  T* W = sm + L.W;
  T* H = sm + L.H;
  T* X = sm + L.X;
  T* part = sm + L.part;
  T* red = sm + L.red;
  T* vec = sm + L.vec;
  const int r0 = min(m, c * rpc), mr = min(m, r0 + rpc) - r0;       // this CTA's rows (possibly none)
  const T* Ag = Ap[fit] + (int64_t)r0 * lda;
  T* Wg = Wp[fit] + (int64_t)r0 * k;
  T* Hg = Hp[fit];
  const int t = threadIdx.x, NT = kResThreads;
  const int mn = mr * n, mk = mr * k, kn = k * n, kk2 = k * k;
  for (int e = t; e < mn; e += NT) A[(e / n) * ns + (e % n)] = Ag[(int64_t)(e / n) * lda + (e % n)];
  for (int e = t; e < mk; e += NT) W[(e / k) * ks + (e % k)] = Wg[e];
  for (int e = t; e < kn; e += NT) H[(e / n) * ns + (e % n)] = Hg[e];
  __syncthreads();
  const int np = kn + (kl ? k : kk2);                                // entries exchanged per H half-step

  auto compute_U = [&](T* U) {                                        // U = A / (W H + eps): warps over rows, lanes over columns
    const int warp = t >> 5, lane = t & 31;
    for (int i = warp; i < mr; i += NT / 32) {
      const T* Wi = W + i * ks;
      for (int j = lane; j < n; j += 32) {
        T s = (T)0;
        for (int l = 0; l < k; ++l) s += Wi[l] * H[l * ns + j];
        U[i * ns + j] = fast_div(A[i * ns + j], s + eps);
      }
    }
  };
  // red[e] = sum over the cluster of part[e]: reduce-scatter (CTA c sums slice c from all peers, in rank order) followed
  // by an all-gather of the slices, both through distributed shared memory.  No third barrier is needed: `part` is next
  // written after every peer passed the second barrier, own slices of `red` after the next first barrier.
  auto exchange = [&]() {
    if (CLUSTER) {
      cg::cluster_group cluster = cg::this_cluster();
      const int sl = (np + C - 1) / C;
      cluster.sync();
      for (int e = c * sl + t; e < min(np, (c + 1) * sl); e += NT) {
        T v[kMaxCluster];
#pragma unroll
        for (int cc = 0; cc < kMaxCluster; ++cc) v[cc] = cc < C ? cluster.map_shared_rank(part, cc)[e] : (T)0;
        T s = (T)0;
#pragma unroll
        for (int cc = 0; cc < kMaxCluster; ++cc) s += v[cc];          // the padding terms are exact zeros
        red[e] = s;
      }
      cluster.sync();
      for (int e = t; e < np; e += NT) {
        const int owner = e / sl;
        if (owner != c) red[e] = cluster.map_shared_rank(red, owner)[e];
      }
      __syncthreads();
    } else {
      __syncthreads();
    }
  };

  for (int64_t it = it0; it < it1; ++it) {
    if (kl) {
      T* U = X;
      if (w_update) {
        k_sums(H, k, n, ns, 1, vec);                                  // x2 = H.sum(axis=1)
        compute_U(U);
        __syncthreads();
        grouped_dots<T>(mk, n, [&](int o, int l) { return U[(o / k) * ns + l] * H[(o % k) * ns + l]; },
                        [&](int o, T v) { T& w = W[(o / k) * ks + (o % k)]; w = w * (v / (vec[o % k] + eps)); });
        __syncthreads();
      }
      k_sums(W, k, mr, 1, ks, part + kn);                             // (partial) x = W.sum(axis=0)
      compute_U(U);
      __syncthreads();
      grouped_dots<T>(kn, mr, [&](int o, int l) { return W[l * ks + (o / n)] * U[l * ns + (o % n)]; },
                      [&](int o, T y) { part[o] = y; });
      exchange();
      for (int e = t; e < kn; e += NT) {
        T& h = H[(e / n) * ns + (e % n)];
        h = h * (red[e] / (red[kn + e / n] + eps));
      }
      __syncthreads();
    } else {
      const int f = (rpc * k) > kn ? (rpc * k) : kn;
      T* V = X;
      T* Nx = X + f;
      T* G = X + 2 * f;
      if (w_update) {
        grouped_dots<T>(kk2, n, [&](int o, int l) { return H[(o / k) * ns + l] * H[(o % k) * ns + l]; },
                        [&](int o, T v) { G[o] = v; });                                       // H H^T
        grouped_dots<T>(mk, n, [&](int o, int l) { return A[(o / k) * ns + l] * H[(o % k) * ns + l]; },
                        [&](int o, T v) { V[o] = v; });                                       // A H^T
        __syncthreads();
        grouped_dots<T>(mk, k, [&](int o, int l) { return W[(o / k) * ks + l] * G[l * k + (o % k)]; },
                        [&](int o, T d) { Nx[o] = W[(o / k) * ks + (o % k)] * (V[o] / (d + eps)); });
        __syncthreads();
        for (int e = t; e < mk; e += NT) W[(e / k) * ks + (e % k)] = Nx[e];
        __syncthreads();
      }
      grouped_dots<T>(kk2, mr, [&](int o, int l) { return W[l * ks + (o / k)] * W[l * ks + (o % k)]; },
                      [&](int o, T v) { part[kn + o] = v; });                                 // (partial) W^T W
      grouped_dots<T>(kn, mr, [&](int o, int l) { return W[l * ks + (o / n)] * A[l * ns + (o % n)]; },
                      [&](int o, T v) { part[o] = v; });                                      // (partial) W^T A
      exchange();
      const T* G2 = red + kn;
      grouped_dots<T>(kn, k, [&](int o, int l) { return H[l * ns + (o % n)] * G2[l * k + (o / n)]; },
                      [&](int o, T d) { X[o] = H[(o / n) * ns + (o % n)] * (red[o] / (d + eps)); });
      __syncthreads();
      for (int e = t; e < kn; e += NT) H[(e / n) * ns + (e % n)] = X[e];
      __syncthreads();
    }
    if (it % 10 == 0) {                                                // pyDNMF.py:155-157
      for (int e = t; e < kn; e += NT) { T& h = H[(e / n) * ns + (e % n)]; h = h > eps ? h : eps; }
      for (int e = t; e < mk; e += NT) { T& w = W[(e / k) * ks + (e % k)]; w = w > eps ? w : eps; }
      __syncthreads();
    }
  }
  for (int e = t; e < mk; e += NT) Wg[e] = W[(e / k) * ks + (e % k)];
  if (c == 0)
    for (int e = t; e < kn; e += NT) Hg[e] = H[(e / n) * ns + (e % n)];
  if (CLUSTER) cg::this_cluster().sync();   // no CTA exits while a peer may still read its shared memory
}

inline int64_t res_bytes(int64_t m, int64_t n, int64_t k, int kl, int dtype) {
  return res_layout(m, n, k, kl, 0).total * (dtype == DNMF_F32 ? 4 : 8);
}

// can a cluster of C CTAs with `bytes` of dynamic shared memory each be scheduled? (cached per dtype / C / size)
template <typename T>
bool clu_schedulable(int C, int64_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, int64_t>, bool> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(C, bytes);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  auto kern = mu_fit_onchip_kernel<T, true>;
  bool ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
  if (ok && C > 8) ok = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  if (ok) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)C);
    cfg.blockDim = dim3(kResThreads);
    cfg.dynamicSmemBytes = (size_t)bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int active = 0;
    ok = cudaOccupancyMaxActiveClusters(&active, kern, &cfg) == cudaSuccess && active >= 1;
  }
  if (!ok) (void)cudaGetLastError();
  cache[key] = ok;
  return ok;
}

// cluster size for an m-row fit: enough CTAs to spread ~64 rows each (up to 16); other sizes if that one does not fit in
// shared memory or cannot be scheduled on this device
inline int clu_pick(int64_t m, int64_t n, int64_t k, int kl, int dtype, int64_t* bytes_out, int64_t* rpc_out) {
  const int es = dtype == DNMF_F32 ? 4 : 8;
  int want = 2;
  while (want < kMaxCluster && m / (want * 2) >= 64) want *= 2;
  int order[8], no = 0;
  for (int C = want; C <= kMaxCluster; C *= 2) order[no++] = C;
  for (int C = want / 2; C >= 2; C /= 2) order[no++] = C;
  for (int i = 0; i < no; ++i) {
    const int C = order[i];
    const int64_t rpc = ceil_div(m, C);
    const int64_t b = res_layout(rpc, n, k, kl, 1).total * es;
    if (b > 227 * 1024) continue;
    const bool ok = dtype == DNMF_F32 ? clu_schedulable<float>(C, b) : clu_schedulable<double>(C, b);
    if (!ok) continue;
    *bytes_out = b;
    *rpc_out = rpc;
    return C;
  }
  return 0;
}

template <typename T>
int launch_onchip(const void* const* A_ptrs, int64_t lda, void* const* W_ptrs, void* const* H_ptrs, int64_t batch, int64_t m,
                  int64_t n, int64_t k, int kl, int w_update, int64_t it0, int64_t it1, double eps, int C, int64_t bytes,
                  int64_t rpc, cudaStream_t st) {
  if (C == 1) {
    auto kern = mu_fit_onchip_kernel<T, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "mu_fit_onchip smem attribute");
    kern<<<(unsigned)batch, kResThreads, (size_t)bytes, st>>>((const T* const*)A_ptrs, lda, (T* const*)W_ptrs, (T* const*)H_ptrs,
                                                             (int)m, (int)n, (int)k, kl, w_update, it0, it1, (T)eps, (int)m);
    DNMF_LAUNCH_CHECK("mu_fit_onchip_kernel");
    return 0;
  }
  auto kern = mu_fit_onchip_kernel<T, true>;     // (attributes were set by clu_schedulable)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(batch * C));
  cfg.blockDim = dim3(kResThreads);
  cfg.dynamicSmemBytes = (size_t)bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const T* const*)A_ptrs, lda, (T* const*)W_ptrs, (T* const*)H_ptrs, (int)m, (int)n,
                                     (int)k, kl, w_update, it0, it1, (T)eps, (int)rpc);
  if (e != cudaSuccess) return cuda_fail(e, "mu_fit_onchip_kernel (cluster) launch");
  DNMF_LAUNCH_CHECK("mu_fit_onchip_kernel");
  return 0;
}

// CTAs per fit (0 = not supported) and the per-CTA shared memory / rows per CTA that go with it
inline int onchip_plan(int64_t m, int64_t n, int64_t k, int kl, int dtype, int64_t* bytes, int64_t* rpc) {
  if (m < 1 || n < 1 || k < 1 || k > DNMF_MAX_K || (dtype != DNMF_F32 && dtype != DNMF_F64)) return 0;
  if (m * n > (1 << 22)) return 0;
  const int64_t b1 = res_bytes(m, n, k, kl, dtype);
  const bool one = b1 <= 227 * 1024;
  if (!(one && m < 128)) {
    const int C = clu_pick(m, n, k, kl, dtype, bytes, rpc);
    if (C > 0) return C;
  }
  if (!one) return 0;
  *bytes = b1;
  *rpc = m;
  return 1;
}

}  // namespace

extern "C" {

int64_t dnmf_mu_fit_resident_smem_bytes(int64_t m, int64_t n, int64_t k, int kl, int dtype) {
  int64_t bytes = 0, rpc = 0;
  return onchip_plan(m, n, k, kl, dtype, &bytes, &rpc) > 0 ? bytes : -1;
}

int dnmf_mu_fit_resident_cluster_size(int64_t m, int64_t n, int64_t k, int kl, int dtype) {
  int64_t bytes = 0, rpc = 0;
  return onchip_plan(m, n, k, kl, dtype, &bytes, &rpc);
}

int dnmf_mu_fit_resident(const void* const* A_ptrs, int64_t lda, void* const* W_ptrs, void* const* H_ptrs, int64_t batch,
                         int64_t m, int64_t n, int64_t k, int kl, int w_update, int64_t it_begin, int64_t it_end,
                         double eps, int dtype, void* stream) {
  if (dtype != DNMF_F32 && dtype != DNMF_F64) return fail(DNMF_E_ARG, "dtype must be DNMF_F32 or DNMF_F64");
  DNMF_CHECK_ARG(batch >= 0 && A_ptrs && W_ptrs && H_ptrs && it_begin <= it_end, "batch / null pointer / iteration range");
  int64_t bytes = 0, rpc = 0;
  const int C = onchip_plan(m, n, k, kl, dtype, &bytes, &rpc);
  if (C < 1) return fail(DNMF_E_UNSUPPORTED, "a %lld x %lld, k=%lld fit does not fit in shared memory", (long long)m, (long long)n, (long long)k);
  if (batch == 0 || it_begin == it_end) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DNMF_F32)
    return launch_onchip<float>(A_ptrs, lda, W_ptrs, H_ptrs, batch, m, n, k, kl, w_update, it_begin, it_end, eps, C, bytes, rpc, st);
  return launch_onchip<double>(A_ptrs, lda, W_ptrs, H_ptrs, batch, m, n, k, kl, w_update, it_begin, it_end, eps, C, bytes, rpc, st);
}

}  // extern "C"
