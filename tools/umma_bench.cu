// Micro-benchmark: cycles per tcgen05.mma dispatch on sm_100a as a function of N, operand source (A from shared
// memory = SS, from tensor memory = TS) and kind (tf32 K=8, f16/bf16 K=16).  One CTA per SM, one thread issues
// `iters` MMAs back to back on fixed operands and waits for tcgen05.commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu && ./umma_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// variant: the whole warp runs the loop (uniform control flow), one elected lane issues
__global__ void __launch_bounds__(128, 1) bench_uniform(int n, int ts, int iters, int nacc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
    const uint64_t adesc = make_desc(base), bdesc = make_desc(base + 16384);
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i & (nacc - 1)) * n);
      if (elect_one()) {
        if (ts)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(tmem + 448u), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    // cost of commit + wait round trips with an empty tensor pipe
    uint32_t ph = 0;
    for (int i = 0; i < 200; ++i) {
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      __syncwarp();
      mbar_wait(smem_u32(&bar2), ph);
      ph ^= 1u;
    }
    const long long t3 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = (t3 - t2) / 200; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// pattern: groups of `nm` MMAs (tf32, TS, M=128 or 64) followed by `nc` commits to distinct mbarriers (never waited on)
__global__ void __launch_bounds__(128, 1) bench_pattern(int n, int m, int nm, int nc, int groups, long long* out) {
  __shared__ uint64_t bars[8];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    const uint64_t bdesc = make_desc(base + 16384);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      for (int i = 0; i < nm; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tmem), "r"(tmem + 448u), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
      for (int c = 0; c < nc; ++c)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[c])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// variant C: whole warp executes; the elect.sync predicate guards the MMA inside one asm block (no C++ branch)
__device__ __forceinline__ void mma_ts_elect(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar) : "memory");
}
__global__ void __launch_bounds__(128, 1) bench_c(int n, int nc, int groups, long long* out) {
  __shared__ uint64_t bars[8];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
    const uint64_t bdesc = make_desc(base + 16384);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        mma_ts_elect(tmem, tmem + 448u + kk * 8, bdesc + 2 * kk, idesc, 1u);
        mma_ts_elect(tmem + 64, tmem + 480u + kk * 8, bdesc + 2 * kk, idesc, 1u);
      }
      if (nc > 0) commit_elect(smem_u32(&bars[0]));
      if (nc > 1) commit_elect(smem_u32(&bars[1]));
      if (nc > 2) commit_elect(smem_u32(&bars[2]));
    }
    commit_elect(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int KIND>  // 0 tf32, 1 f16
__global__ void __launch_bounds__(128, 1) bench(int n, int ts, int iters, int nacc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((KIND == 0 ? 2u : 0u) << 7) | ((KIND == 0 ? 2u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
    const uint64_t adesc = make_desc(base), bdesc = make_desc(base + 16384);
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i & (nacc - 1)) * n);   // nacc (power of 2) independent accumulators
      if (ts) {
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(tmem + 448u), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(tmem + 448u), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
      } else {
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 32);
  const int iters = 2000;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("kind  src  N    nacc grid  issue_cyc/mma  total_cyc/mma\n");
  for (int kind = 0; kind < 2; ++kind)
    for (int ts = 0; ts < 2; ++ts)
      for (int n : {32, 64, 128, 256})
        for (int nacc : {1, 2})
          for (int grid : {1}) {
            if (nacc * n > 256) continue;
            if (kind == 0) bench<0><<<grid, 128, 64 * 1024>>>(n, ts, iters, nacc, d);
            else bench<1><<<grid, 128, 64 * 1024>>>(n, ts, iters, nacc, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2] = {0, 0};
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-5s %-4s %-4d %-4d %-5d %8.1f %14.1f %s\n", kind ? "f16" : "tf32", ts ? "TS" : "SS", n, nacc, grid,
                   (double)h[0] / iters, (double)h[1] / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
  cudaFuncSetAttribute(bench_uniform, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("uniform-issue variant (tf32):  src N nacc  issue_cyc/mma total_cyc/mma commit+wait_roundtrip\n");
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {32, 64, 128, 256})
      for (int nacc : {1, 2}) {
        if (nacc * n > 256) continue;
        bench_uniform<<<1, 128, 64 * 1024>>>(n, ts, iters, nacc, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[3] = {0, 0, 0};
        cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("  %-4s %-4d %-4d %8.1f %14.1f %8lld %s\n", ts ? "TS" : "SS", n, nacc, (double)h[0] / iters, (double)h[1] / iters, h[2],
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  cudaFuncSetAttribute(bench_c, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("variant C (warp-uniform, in-asm elect predicate): 8 MMAs (tf32 TS) + nc commits per group\n");
  for (int n : {32, 64, 128})
    for (int nc : {0, 1, 2, 3}) {
      bench_c<<<1, 128, 64 * 1024>>>(n, nc, 1000, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[1] = {0};
      cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
      printf("  N=%-3d commits=%d : %8.1f cyc/group %s\n", n, nc, (double)h[0] / 1000, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  cudaFuncSetAttribute(bench_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("pattern (tf32 TS): M N nMMA nCommit -> cycles per group, per slot\n");
  for (int m : {128, 64})
    for (int n : {32, 64})
      for (int nm : {8, 6, 0})
        for (int nc : {0, 1, 2, 3}) {
          if (nm == 0 && nc == 0) continue;
          bench_pattern<<<1, 128, 64 * 1024>>>(n, m, nm, nc, 500, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[1] = {0};
          cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
          printf("  M=%-4d N=%-3d mma=%d commit=%d : %8.1f cyc/group  %6.1f cyc/slot %s\n", m, n, nm, nc, (double)h[0] / 500,
                 (double)h[0] / 500 / (nm + nc), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
