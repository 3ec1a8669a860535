// Communicators, peer-mapped ("symmetric") memory and the fused half-step exchange of the P x 1 grid.
//
//  * dnmf_comm_*  / dnmf_allreduce / dnmf_allgather / dnmf_reduce_scatter / dnmf_bcast: the MPI communicators of
//    pyDNMFk/dist_comm.py:16-56 as NCCL communicators owned by this library (one process per GPU).  NCCL is resolved
//    at run time from the libnccl.so.2 already loaded into the process (or a given path); collectives are enqueued on
//    the caller's stream, so they are captured into the step's CUDA graph with the kernels around them.
//  * dnmf_symm_*: device buffers exported with CUDA IPC and mapped by every peer of the node: plain ld / st on the
//    mapped pointers travel over NVLink / NVSwitch.
//  * dnmf_xchg_*: the H half-step of the row grid (dist_nmf.py:705-708 allreduce of W^T A, :679-681 allreduce of W^T W,
//    :750-751 update) as ONE exchange over peer memory instead of two NCCL all-reduces and an update kernel:
//      push   : every rank writes the rows of its partial (W_i^T A_i)^T that belong to column-chunk owner q straight
//               into q's receive slot, and its k x k Gram (or k-vector) into every peer's slot; release flags
//      update : the owner of a chunk sums the P slots in rank order (deterministic, identical on every run), applies
//               the multiplicative / HALS update to its columns of H and writes the new columns into the staging
//               replica of EVERY rank; release flags
//      finish : waits for all owners, copies the staging replica over H
//    i.e. reduce-scatter -> column-sharded update -> all-gather, with 1/P of the update work per rank.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstddef>

#include "common.cuh"
#include "generic_small.cuh"
#include "launch_passes.cuh"

namespace dnmf {
namespace {

// ---------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time
// ---------------------------------------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

int nccl_load(const char* path) {
  if (g_nccl.handle) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy torch already mapped, if any
  if (!h && path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { const char* e = getenv("DNMF_NCCL_LIB"); if (e && *e) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(DNMF_E_UNSUPPORTED, "libnccl.so.2 not found (%s)", dlerror());
#define DNMF_NCCL_SYM(field, name)                                                                     \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                                           \
  if (!g_nccl.field) return fail(DNMF_E_UNSUPPORTED, "libnccl.so.2 lacks %s", name)
  DNMF_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  DNMF_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  DNMF_NCCL_SYM(CommSplit, "ncclCommSplit");
  DNMF_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  DNMF_NCCL_SYM(CommCount, "ncclCommCount");
  DNMF_NCCL_SYM(CommUserRank, "ncclCommUserRank");
  DNMF_NCCL_SYM(AllReduce, "ncclAllReduce");
  DNMF_NCCL_SYM(AllGather, "ncclAllGather");
  DNMF_NCCL_SYM(ReduceScatter, "ncclReduceScatter");
  DNMF_NCCL_SYM(Broadcast, "ncclBroadcast");
  DNMF_NCCL_SYM(GroupStart, "ncclGroupStart");
  DNMF_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  DNMF_NCCL_SYM(GetErrorString, "ncclGetErrorString");
  DNMF_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef DNMF_NCCL_SYM
  g_nccl.handle = h;
  return 0;
}

int nccl_fail(ncclResult_t r, const char* where) {
  return fail(DNMF_E_COMM, "%s: NCCL error %d (%s)", where, (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
}
#define DNMF_NCCL_CHECK(call, where)                              \
  do {                                                            \
    ncclResult_t _r = (call);                                     \
    if (_r != ncclSuccess) return nccl_fail(_r, where);           \
  } while (0)

int nccl_dtype(int dtype, ncclDataType_t* out) {
  switch (dtype) {
    case DNMF_F32: *out = ncclFloat32; return 0;
    case DNMF_F64: *out = ncclFloat64; return 0;
    case DNMF_I64: *out = ncclInt64; return 0;
  }
  return fail(DNMF_E_ARG, "collective dtype must be DNMF_F32, DNMF_F64 or DNMF_I64");
}

// ---------------------------------------------------------------------------------------------------------
// exchange region layout (identical on every rank; all offsets 256-byte aligned)
// ---------------------------------------------------------------------------------------------------------
constexpr int XCHG_MAX_P = 16;
struct XchgCtrl {                       // at offset 0 of every rank's region
  unsigned int flag_push[XCHG_MAX_P];   // [q] = epoch of the last push received from rank q
  unsigned int flag_upd[XCHG_MAX_P];    // [q] = epoch of the last updated chunk received from owner q
  unsigned int done_push, done_upd;     // CTA completion counters of the local kernels
  unsigned int epoch;                   // exchanges started so far on this rank
  unsigned int error;                   // != 0: a wait timed out (peer died); results are invalid
  unsigned int hals_epoch;              // HALS W sweeps run so far on this rank
  unsigned int grid_bar;                // grid barrier counter of the sweep kernel (zeroed by its last block)
};
struct XchgLayout {
  int64_t n_chunk, off_recv, off_aux, off_stage, off_hals_sum, off_hals_flag, bytes;
};
XchgLayout xchg_layout(int P, int64_t n, int64_t k, int dtype) {
  XchgLayout L;
  const int64_t es = dtype == DNMF_F32 ? 4 : 8;
  L.n_chunk = ceil_div(n > 0 ? n : 1, P);
  L.off_recv = 256;
  L.off_aux = L.off_recv + round_up((int64_t)P * L.n_chunk * k * es, 256);
  L.off_stage = L.off_aux + round_up((int64_t)P * k * k * es, 256);
  L.off_hals_sum = L.off_stage + round_up(k * n * es, 256);                  // [DNMF_MAX_K][XCHG_MAX_P] float64
  L.off_hals_flag = L.off_hals_sum + DNMF_MAX_K * XCHG_MAX_P * 8;             // [DNMF_MAX_K][XCHG_MAX_P] uint32
  L.bytes = L.off_hals_flag + DNMF_MAX_K * XCHG_MAX_P * 4;
  return L;
}
struct XchgBases { char* p[XCHG_MAX_P]; };

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// threads 0..P-1 of the CTA wait for flags[q] >= epoch; bounded (a dead peer must not hang the GPU)
__device__ __forceinline__ void wait_flags(XchgCtrl* ctrl, const unsigned int* flags, int P, unsigned int epoch) {
  if ((int)threadIdx.x < P) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
      __nanosleep(64);
      if (clock64() - t0 > 20000000000LL) { ctrl->error = 1u; break; }      // ~10 s
    }
  }
  __syncthreads();
}
// called by every thread after its peer stores: the last CTA of the grid publishes `epoch` to slot [me] of every rank
__device__ __forceinline__ void publish_when_grid_done(const XchgBases& B, XchgCtrl* ctrl, unsigned int* done, int P, int me,
                                                       size_t flag_off, unsigned int epoch) {
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < P)
      st_release_sys(reinterpret_cast<unsigned int*>(B.p[threadIdx.x] + flag_off) + me, epoch);
    if (threadIdx.x == 0) *done = 0u;
  }
}

// push: rows [q n_chunk, (q+1) n_chunk) of the local Yt [n x k] -> recv[me] of owner q; aux (aux_len values) -> aux[me]
// of every rank.  grid-stride over the n x k elements.
// (Yt may be given as `splits` split-K partials `sstride` elements apart: summed here in split order, see sum_splits)
template <typename T>
__global__ void __launch_bounds__(256) xchg_push_kernel(XchgBases B, XchgLayout L, int P, int me, const T* __restrict__ Yt,
                                                        int64_t ldy, int64_t n, int k, const T* __restrict__ aux, int aux_len,
                                                        int splits, int64_t sstride) {
  XchgCtrl* ctrl = reinterpret_cast<XchgCtrl*>(B.p[me]);
  const unsigned int epoch = ctrl->epoch + 1u;          // every CTA reads the value left by the previous exchange
  const int64_t total = n * k;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / k;
    const int j = (int)(idx % k);
    const int q = (int)(c / L.n_chunk);
    T* recv = reinterpret_cast<T*>(B.p[q] + L.off_recv) + ((int64_t)me * L.n_chunk + (c - (int64_t)q * L.n_chunk)) * k;
    recv[j] = sum_splits(Yt + c * ldy + j, splits, sstride);
  }
  if (blockIdx.x < (unsigned)P) {                        // CTA q delivers the small operand to rank q
    T* dst = reinterpret_cast<T*>(B.p[blockIdx.x] + L.off_aux) + (int64_t)me * k * k;
    for (int i = threadIdx.x; i < aux_len; i += blockDim.x) dst[i] = aux[i];
  }
  publish_when_grid_done(B, ctrl, &ctrl->done_push, P, me, offsetof(XchgCtrl, flag_push), epoch);
  // the epoch counter itself is advanced by the update kernel (stream order), so that every CTA here saw the old value
}

// update: MODE 0 FRO-MU (aux = W^T W), 2 FRO-HALS (aux = W^T W), 3 KL-MU (aux = colsum(W))
template <typename T, int KP, int MODE>
__global__ void __launch_bounds__(kColUpdThreads) xchg_update_kernel(XchgBases B, XchgLayout L, int P, int me,
                                                                     const T* __restrict__ H, int64_t ldh, int64_t n, int k,
                                                                     T p0, int clamp) {
  XchgCtrl* ctrl = reinterpret_cast<XchgCtrl*>(B.p[me]);
  const unsigned int epoch = ctrl->epoch + 1u;
  wait_flags(ctrl, ctrl->flag_push, P, epoch);
  __shared__ T Gs[KP * KP];
  const T* aux = reinterpret_cast<const T*>(B.p[me] + L.off_aux);
  const int aux_len = (MODE == 3) ? k : k * k;
  for (int idx = threadIdx.x; idx < KP * KP; idx += kColUpdThreads) Gs[idx] = T(0);
  __syncthreads();
  for (int i = threadIdx.x; i < aux_len; i += kColUpdThreads) {
    T s = T(0);
    for (int q = 0; q < P; ++q) s += aux[(int64_t)q * k * k + i];           // rank order: deterministic
    if (MODE == 3) Gs[i] = s;
    else Gs[(i / k) * KP + (i % k)] = s;
  }
  __syncthreads();
  const int64_t c0 = (int64_t)me * L.n_chunk;
  const int64_t c1 = min(n, c0 + L.n_chunk);
  const int64_t c = c0 + (int64_t)blockIdx.x * kColUpdThreads + threadIdx.x;
  if (c < c1) {
    const T* recv = reinterpret_cast<const T*>(B.p[me] + L.off_recv);
    T h[KP], y[KP];
#pragma unroll
    for (int l = 0; l < KP; ++l) {
      h[l] = (l < k) ? H[(int64_t)l * ldh + c] : T(0);
      y[l] = T(0);
    }
    for (int q = 0; q < P; ++q) {
      const T* row = recv + ((int64_t)q * L.n_chunk + (c - c0)) * k;
#pragma unroll
      for (int l = 0; l < KP; ++l)
        if (l < k) y[l] += row[l];
    }
    T out[KP];
    if (MODE == 2) {
#pragma unroll
      for (int kk = 0; kk < KP; ++kk) {
        if (kk < k) {
          T d = T(0);
#pragma unroll
          for (int l = 0; l < KP; ++l) d = fma(Gs[kk * KP + l], h[l], d);
          const T v = h[kk] + y[kk] - d;
          h[kk] = v > p0 ? v : p0;
        }
      }
#pragma unroll
      for (int kk = 0; kk < KP; ++kk) out[kk] = h[kk];
    } else {
#pragma unroll
      for (int kk = 0; kk < KP; ++kk) {
        T res = T(0);
        if (kk < k) {
          if (MODE == 0) {
            T d = T(0);
#pragma unroll
            for (int l = 0; l < KP; ++l) d = fma(h[l], Gs[l * KP + kk], d);
            res = h[kk] * (y[kk] / (d + p0));
          } else {
            res = h[kk] * (y[kk] / (Gs[kk] + p0));
          }
          if (clamp) res = res > p0 ? res : p0;
        }
        out[kk] = res;
      }
    }
    for (int q = 0; q < P; ++q) {
      T* stage = reinterpret_cast<T*>(B.p[q] + L.off_stage);
#pragma unroll
      for (int kk = 0; kk < KP; ++kk)
        if (kk < k) stage[(int64_t)kk * n + c] = out[kk];
    }
  }
  publish_when_grid_done(B, ctrl, &ctrl->done_upd, P, me, offsetof(XchgCtrl, flag_upd), epoch);
}

// finish: wait for every owner's chunk, advance the epoch, copy the staging replica over H
template <typename T>
__global__ void __launch_bounds__(256) xchg_finish_kernel(XchgBases B, XchgLayout L, int P, int me, T* __restrict__ H,
                                                          int64_t ldh, int64_t n, int k) {
  XchgCtrl* ctrl = reinterpret_cast<XchgCtrl*>(B.p[me]);
  const unsigned int epoch = ctrl->epoch + 1u;
  wait_flags(ctrl, ctrl->flag_upd, P, epoch);
  const T* stage = reinterpret_cast<const T*>(B.p[me] + L.off_stage);
  const int64_t total = n * k;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kk = idx / n, c = idx % n;
    H[kk * ldh + c] = stage[idx];
  }
}
__global__ void xchg_epoch_kernel(XchgCtrl* ctrl) { ctrl->epoch += 1u; }

// ---------------------------------------------------------------------------------------------------------
// HALS W sweep in ONE launch (dist_nmf.py:888-893; 2-D :427-432): for kk = 0..k-1, Gauss-Seidel,
//     t = W[:,kk] G[kk,kk] + V[:,kk] - W G[:,kk];  W[:,kk] = max(t, eps);  W[:,kk] /= ||W[:,kk]||_2 (global)
// The k column norms are k dependent global reductions.  The reference (and round 1 of this library) runs them as k x
// (kernel, reduction kernel, all-reduce, scaling kernel); here one persistent grid walks the columns, reduces the
// block partials behind a grid barrier in a fixed order and, on a row grid, exchanges the per-rank sums through the
// peers' exchange regions (one 8-byte store per peer and column) instead of k NCCL all-reduces.
// The scaling of column kk is applied lazily when the row is next touched.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, int KP>
__global__ void __launch_bounds__(256) hals_w_sweep_kernel(T* __restrict__ W, int64_t ldw, const T* __restrict__ V, int64_t ldv,
                                                           const T* __restrict__ G, int64_t m, int k, T eps,
                                                           double* __restrict__ partials, unsigned int* __restrict__ gbar,
                                                           XchgBases B, XchgLayout L, int P, int me) {
  __shared__ T g[KP];
  __shared__ double red[8];
  __shared__ double s_total;
  XchgCtrl* ctrl = P > 1 ? reinterpret_cast<XchgCtrl*>(B.p[me]) : nullptr;
  const unsigned int epoch = P > 1 ? ctrl->hals_epoch + 1u : 0u;
  const int64_t stride = (int64_t)gridDim.x * 256;
  T ss_prev = T(0);
  for (int kk = 0; kk < k; ++kk) {
    if (threadIdx.x < KP) g[threadIdx.x] = ((int)threadIdx.x < k) ? G[threadIdx.x * k + kk] : T(0);
    __syncthreads();
    double sq = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < m; r += stride) {
      T* w = W + r * ldw;
      if (kk > 0 && ss_prev > T(0)) w[kk - 1] = w[kk - 1] / ss_prev;
      T d = T(0);
#pragma unroll
      for (int l = 0; l < KP; ++l)
        if (l < k) d = fma(w[l], g[l], d);
      T v = w[kk] * g[kk] + V[r * ldv + kk] - d;
      v = v > eps ? v : eps;
      w[kk] = v;
      sq += (double)v * (double)v;
    }
    const double bs = block_sum<256>(sq, red);
    if (threadIdx.x == 0) {
      partials[(int64_t)kk * gridDim.x + blockIdx.x] = bs;
      __threadfence();
      atomicAdd(gbar, 1u);
      const unsigned int target = (unsigned int)(kk + 1) * gridDim.x;
      while (ld_acquire_gpu(gbar) < target) {}
      double tot = 0.0;
      for (unsigned int b = 0; b < gridDim.x; ++b) tot += partials[(int64_t)kk * gridDim.x + b];   // fixed order
      if (P > 1) {
        if (blockIdx.x == 0) {
          for (int q = 0; q < P; ++q) {
            double* dst = reinterpret_cast<double*>(B.p[q] + L.off_hals_sum) + kk * XCHG_MAX_P + me;
            *reinterpret_cast<volatile double*>(dst) = tot;
          }
          __threadfence_system();
          for (int q = 0; q < P; ++q)
            st_release_sys(reinterpret_cast<unsigned int*>(B.p[q] + L.off_hals_flag) + kk * XCHG_MAX_P + me, epoch);
        }
        const unsigned int* flags = reinterpret_cast<const unsigned int*>(B.p[me] + L.off_hals_flag) + kk * XCHG_MAX_P;
        const double* sums = reinterpret_cast<const double*>(B.p[me] + L.off_hals_sum) + kk * XCHG_MAX_P;
        tot = 0.0;
        const long long t0 = clock64();
        for (int q = 0; q < P; ++q) {
          while (ld_acquire_sys(flags + q) < epoch) {
            if (clock64() - t0 > 20000000000LL) { ctrl->error = 1u; break; }
          }
          tot += *reinterpret_cast<const volatile double*>(sums + q);                                 // rank order
        }
      }
      s_total = tot;
    }
    __syncthreads();
    ss_prev = (T)sqrt(s_total);
    __syncthreads();
  }
  if (ss_prev > T(0))
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < m; r += stride) W[r * ldw + (k - 1)] /= ss_prev;
  // last block out resets the barrier counter and advances the sweep epoch for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(gbar, 1u);
    if (done == (unsigned int)(k + 1) * gridDim.x - 1u) {
      *gbar = 0u;
      if (P > 1) ctrl->hals_epoch = epoch;
    }
  }
}

struct CommBox {
  ncclComm_t comm;
  int rank, size;
};

}  // namespace
}  // namespace dnmf

using namespace dnmf;

extern "C" {

int dnmf_comm_load(const char* libnccl_path) { return nccl_load(libnccl_path); }

int dnmf_comm_nccl_version(int* version) {
  if (int rc = nccl_load(nullptr)) return rc;
  DNMF_NCCL_CHECK(g_nccl.GetVersion(version), "ncclGetVersion");
  return 0;
}

int dnmf_comm_unique_id(void* id_out_128) {
  if (int rc = nccl_load(nullptr)) return rc;
  DNMF_CHECK_ARG(id_out_128, "null pointer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  DNMF_NCCL_CHECK(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id_out_128, &id, sizeof(id));
  return 0;
}

int dnmf_comm_init_rank(const void* id_128, int nranks, int rank, void** comm_out) {
  if (int rc = nccl_load(nullptr)) return rc;
  DNMF_CHECK_ARG(id_128 && comm_out && nranks >= 1 && rank >= 0 && rank < nranks, "bad id / rank");
  ncclUniqueId id;
  memcpy(&id, id_128, sizeof(id));
  ncclComm_t c;
  DNMF_NCCL_CHECK(g_nccl.CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
  *comm_out = new CommBox{c, rank, nranks};
  return 0;
}

int dnmf_comm_split(void* comm, int color, int key, void** comm_out) {
  DNMF_CHECK_ARG(comm && comm_out, "null communicator");
  CommBox* b = reinterpret_cast<CommBox*>(comm);
  ncclComm_t c = nullptr;
  DNMF_NCCL_CHECK(g_nccl.CommSplit(b->comm, color < 0 ? NCCL_SPLIT_NOCOLOR : color, key, &c, nullptr), "ncclCommSplit");
  if (c == nullptr) { *comm_out = nullptr; return 0; }
  int r = 0, s = 0;
  DNMF_NCCL_CHECK(g_nccl.CommUserRank(c, &r), "ncclCommUserRank");
  DNMF_NCCL_CHECK(g_nccl.CommCount(c, &s), "ncclCommCount");
  *comm_out = new CommBox{c, r, s};
  return 0;
}

int dnmf_comm_rank(void* comm, int* rank, int* size) {
  DNMF_CHECK_ARG(comm, "null communicator");
  CommBox* b = reinterpret_cast<CommBox*>(comm);
  if (rank) *rank = b->rank;
  if (size) *size = b->size;
  return 0;
}

int dnmf_comm_destroy(void* comm) {
  if (!comm) return 0;
  CommBox* b = reinterpret_cast<CommBox*>(comm);
  ncclResult_t r = g_nccl.CommDestroy(b->comm);
  delete b;
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommDestroy");
  return 0;
}

int dnmf_allreduce(void* comm, void* buf, int64_t count, int dtype, void* stream) {
  DNMF_CHECK_ARG(comm && (buf || count == 0) && count >= 0, "null communicator / buffer");
  ncclDataType_t dt;
  if (int rc = nccl_dtype(dtype, &dt)) return rc;
  CommBox* b = reinterpret_cast<CommBox*>(comm);
  DNMF_NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t)count, dt, ncclSum, b->comm, (cudaStream_t)stream), "ncclAllReduce");
  tls().launches++;
  return 0;
}

int dnmf_allgather(void* comm, const void* send, void* recv, int64_t count, int dtype, void* stream) {
  DNMF_CHECK_ARG(comm && send && recv && count >= 0, "null communicator / buffer");
  ncclDataType_t dt;
  if (int rc = nccl_dtype(dtype, &dt)) return rc;
  CommBox* b = reinterpret_cast<CommBox*>(comm);
  DNMF_NCCL_CHECK(g_nccl.AllGather(send, recv, (size_t)count, dt, b->comm, (cudaStream_t)stream), "ncclAllGather");
  tls().launches++;
  return 0;
}

int dnmf_reduce_scatter(void* comm, const void* send, void* recv, int64_t recv_count, int dtype, void* stream) {
  DNMF_CHECK_ARG(comm && send && recv && recv_count >= 0, "null communicator / buffer");
  ncclDataType_t dt;
  if (int rc = nccl_dtype(dtype, &dt)) return rc;
  CommBox* b = reinterpret_cast<CommBox*>(comm);
  DNMF_NCCL_CHECK(g_nccl.ReduceScatter(send, recv, (size_t)recv_count, dt, ncclSum, b->comm, (cudaStream_t)stream),
                  "ncclReduceScatter");
  tls().launches++;
  return 0;
}

int dnmf_bcast(void* comm, void* buf, int64_t count, int dtype, int root, void* stream) {
  DNMF_CHECK_ARG(comm && (buf || count == 0) && count >= 0, "null communicator / buffer");
  ncclDataType_t dt;
  if (int rc = nccl_dtype(dtype, &dt)) return rc;
  CommBox* b = reinterpret_cast<CommBox*>(comm);
  DNMF_CHECK_ARG(root >= 0 && root < b->size, "root out of range");
  DNMF_NCCL_CHECK(g_nccl.Broadcast(buf, buf, (size_t)count, dt, root, b->comm, (cudaStream_t)stream), "ncclBroadcast");
  tls().launches++;
  return 0;
}

int dnmf_group_start(void) {
  if (int rc = nccl_load(nullptr)) return rc;
  DNMF_NCCL_CHECK(g_nccl.GroupStart(), "ncclGroupStart");
  return 0;
}
int dnmf_group_end(void) {
  if (int rc = nccl_load(nullptr)) return rc;
  DNMF_NCCL_CHECK(g_nccl.GroupEnd(), "ncclGroupEnd");
  return 0;
}

// ---- peer-mapped memory ------------------------------------------------------------------------------------
int dnmf_symm_alloc(int64_t bytes, void** ptr, void* handle_out_64) {
  DNMF_CHECK_ARG(bytes > 0 && ptr && handle_out_64, "bad size / null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return cuda_fail(e, "dnmf_symm_alloc cudaMalloc");
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "dnmf_symm_alloc cudaMemset"); }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(handle_out_64, &h, sizeof(h));
  *ptr = p;
  return 0;
}

int dnmf_symm_open(const void* handle_64, void** ptr) {
  DNMF_CHECK_ARG(handle_64 && ptr, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
  *ptr = p;
  return 0;
}

int dnmf_symm_close(void* ptr) {
  if (!ptr) return 0;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaIpcCloseMemHandle");
}

int dnmf_symm_free(void* ptr) {
  if (!ptr) return 0;
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "dnmf_symm_free");
}

// ---- fused H half-step exchange of the P x 1 grid --------------------------------------------------------------
int64_t dnmf_xchg_bytes(int nranks, int64_t n, int64_t k, int dtype) {
  if (nranks < 1 || nranks > XCHG_MAX_P || n < 0 || k < 1 || k > DNMF_MAX_K || (dtype != DNMF_F32 && dtype != DNMF_F64)) {
    fail(DNMF_E_ARG, "dnmf_xchg_bytes: bad arguments");
    return -1;
  }
  return xchg_layout(nranks, n, k, dtype).bytes;
}

int dnmf_xchg_error(const void* local_region, int* error_out, void* stream) {
  DNMF_CHECK_ARG(local_region && error_out, "null pointer");
  unsigned int v = 0;
  const XchgCtrl* c = reinterpret_cast<const XchgCtrl*>(local_region);
  cudaError_t e = cudaMemcpyAsync(&v, &c->error, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "dnmf_xchg_error");
  *error_out = (int)v;
  return 0;
}

// One H half-step: H (k x n, replicated) <- update(H, sum_q Yt_q, sum_q aux_q).  mode 0 FRO-MU (aux = local W^T W,
// p0 = eps), 2 FRO-HALS (aux = local W^T W, p0 = eps), 3 KL-MU (aux = local colsum(W), p0 = eps).
// bases[q] = rank q's exchange region as mapped here (bases[me] = the local allocation).
static int xchg_update_h_impl(void* const* bases, int nranks, int me, int mode, void* H, int64_t ldh, const void* Yt, int64_t ldy,
                              int splits, int64_t sstride, const void* aux, int64_t n, int64_t k, double p0, int clamp, int dtype,
                              void* stream) {
  DNMF_CHECK_ARG(bases && H && Yt && aux, "null pointer");
  DNMF_CHECK_ARG(nranks >= 1 && nranks <= XCHG_MAX_P && me >= 0 && me < nranks, "bad rank count");
  DNMF_CHECK_ARG(mode == 0 || mode == 2 || mode == 3, "mode must be 0 (FRO-MU), 2 (FRO-HALS) or 3 (KL-MU)");
  DNMF_CHECK_ARG(k >= 1 && k <= DNMF_MAX_K && n >= 1 && ldh >= n && ldy >= k, "bad shape");
  DNMF_CHECK_ARG(dtype == DNMF_F32 || dtype == DNMF_F64, "dtype");
  cudaStream_t st = (cudaStream_t)stream;
  XchgBases B;
  for (int q = 0; q < XCHG_MAX_P; ++q) B.p[q] = q < nranks ? reinterpret_cast<char*>(bases[q]) : nullptr;
  const XchgLayout L = xchg_layout(nranks, n, k, dtype);
  const int aux_len = mode == 3 ? (int)k : (int)(k * k);
  const int kp = padded_k(k);
  const unsigned push_grid = (unsigned)std::min<int64_t>(ceil_div(n * k, 256 * 4), (int64_t)sm_count() * 4);
  const unsigned upd_grid = (unsigned)ceil_div(L.n_chunk, kColUpdThreads);
  const unsigned fin_grid = (unsigned)std::min<int64_t>(ceil_div(n * k, 256 * 4), (int64_t)sm_count() * 4);
#define DNMF_XCHG_T(T)                                                                                                        \
  do {                                                                                                                        \
    xchg_push_kernel<T><<<push_grid < (unsigned)nranks ? (unsigned)nranks : push_grid, 256, 0, st>>>(                         \
        B, L, nranks, me, (const T*)Yt, ldy, n, (int)k, (const T*)aux, aux_len, splits, sstride);                             \
    DNMF_LAUNCH_CHECK("xchg_push_kernel");                                                                                    \
    DNMF_DISPATCH_KP(kp, {                                                                                                    \
      if (mode == 0) xchg_update_kernel<T, KP, 0><<<upd_grid, kColUpdThreads, 0, st>>>(B, L, nranks, me, (const T*)H, ldh, n, (int)k, (T)p0, clamp); \
      else if (mode == 2) xchg_update_kernel<T, KP, 2><<<upd_grid, kColUpdThreads, 0, st>>>(B, L, nranks, me, (const T*)H, ldh, n, (int)k, (T)p0, clamp); \
      else xchg_update_kernel<T, KP, 3><<<upd_grid, kColUpdThreads, 0, st>>>(B, L, nranks, me, (const T*)H, ldh, n, (int)k, (T)p0, clamp); \
    });                                                                                                                       \
    DNMF_LAUNCH_CHECK("xchg_update_kernel");                                                                                  \
    xchg_finish_kernel<T><<<fin_grid, 256, 0, st>>>(B, L, nranks, me, (T*)H, ldh, n, (int)k);                                  \
    DNMF_LAUNCH_CHECK("xchg_finish_kernel");                                                                                  \
  } while (0)
  if (dtype == DNMF_F32) DNMF_XCHG_T(float);
  else DNMF_XCHG_T(double);
#undef DNMF_XCHG_T
  xchg_epoch_kernel<<<1, 1, 0, st>>>(reinterpret_cast<XchgCtrl*>(B.p[me]));
  DNMF_LAUNCH_CHECK("xchg_epoch_kernel");
  return 0;
}

int dnmf_xchg_update_h(void* const* bases, int nranks, int me, int mode, void* H, int64_t ldh, const void* Yt, int64_t ldy,
                       const void* aux, int64_t n, int64_t k, double p0, int clamp, int dtype, void* stream) {
  return xchg_update_h_impl(bases, nranks, me, mode, H, ldh, Yt, ldy, 1, 0, aux, n, k, p0, clamp, dtype, stream);
}

int dnmf_xchg_update_h_p(void* const* bases, int nranks, int me, int mode, void* H, int64_t ldh, const int64_t* view4,
                         const void* aux, int64_t n, int64_t k, double p0, int clamp, int dtype, void* stream) {
  DNMF_CHECK_ARG(view4 && view4[0] && view4[2] >= 1, "bad partial view");
  return xchg_update_h_impl(bases, nranks, me, mode, H, ldh, reinterpret_cast<const void*>((uintptr_t)view4[0]), view4[1],
                            (int)view4[2], view4[3], aux, n, k, p0, clamp, dtype, stream);
}

// HALS W sweep, one launch.  nranks == 1 (or bases == NULL): single rank, `scratch` must hold k * 1024 doubles + 1 uint32
// (zeroed once by the caller; the kernel leaves it zeroed).  nranks > 1: the column norms are summed over the ranks
// through the exchange regions `bases` (dnmf_xchg_bytes for the communicator's n, k), scratch as above.
int dnmf_hals_w_sweep(void* W, int64_t ldw, const void* V, int64_t ldv, const void* G, int64_t m, int64_t k, double eps,
                      void* const* bases, int nranks, int me, int64_t xchg_n, void* scratch, int64_t scratch_bytes, int dtype,
                      void* stream) {
  DNMF_CHECK_ARG(W && V && G && scratch, "null pointer");
  DNMF_CHECK_ARG(dtype == DNMF_F32 || dtype == DNMF_F64, "dtype");
  DNMF_CHECK_ARG(k >= 1 && k <= DNMF_MAX_K && m >= 0 && ldw >= k && ldv >= k, "bad shape");
  DNMF_CHECK_ARG(nranks >= 1 && nranks <= XCHG_MAX_P && me >= 0 && me < nranks && (nranks == 1 || bases), "bad rank count");
  DNMF_CHECK_ARG(scratch_bytes >= (int64_t)(k * 1024 * 8 + 256), "scratch too small (k * 1024 doubles + 256 bytes)");
  if (m == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  XchgBases B;
  for (int q = 0; q < XCHG_MAX_P; ++q) B.p[q] = (nranks > 1 && q < nranks) ? reinterpret_cast<char*>(bases[q]) : nullptr;
  XchgLayout L = xchg_layout(nranks, xchg_n > 0 ? xchg_n : 1, k, dtype);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 256);
  unsigned int* gbar = reinterpret_cast<unsigned int*>(scratch);
  int grid = (int)std::min<int64_t>(ceil_div(m, 256), std::min<int64_t>((int64_t)sm_count() * 2, 1024));
  const int kp = padded_k(k);
  {
    void* args[16];
    int ki = (int)k;
    float epsf = (float)eps;
    double epsd = eps;
    args[0] = &W; args[1] = &ldw; args[2] = &V; args[3] = &ldv; args[4] = &G; args[5] = &m; args[6] = &ki;
    args[7] = dtype == DNMF_F32 ? (void*)&epsf : (void*)&epsd;
    args[8] = &partials; args[9] = &gbar; args[10] = &B; args[11] = &L; args[12] = &nranks; args[13] = &me;
    const void* fn = nullptr;
#define DNMF_HALS_FN(T) DNMF_DISPATCH_KP(kp, { fn = (const void*)hals_w_sweep_kernel<T, KP>; })
    if (dtype == DNMF_F32) { DNMF_HALS_FN(float); }
    else { DNMF_HALS_FN(double); }
#undef DNMF_HALS_FN
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0);
    if (e != cudaSuccess) return cuda_fail(e, "hals_w_sweep occupancy");
    const int cap = per_sm * sm_count();
    if (cap < 1) return fail(DNMF_E_UNSUPPORTED, "hals_w_sweep: kernel does not fit on an SM");
    if (grid > cap) grid = cap;
    e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(256), args, 0, st);     // co-residency guaranteed
    if (e != cudaSuccess) return cuda_fail(e, "hals_w_sweep_kernel");
    tls().launches++;
  }
  return 0;
}

}  // extern "C"
