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


// TMEM load / store throughput: `nw` warps (each on its own lane quarter = warp % 4) issue `iters` x32 loads or stores
// back to back (op 0: tcgen05.ld.32x32b.x32, 1: tcgen05.st.32x32b.x32, 2: ld then st alternating); optionally warp 15
// keeps the tensor pipe busy with dependent TS MMAs (N = 32) at the same time.
__global__ void __launch_bounds__(512, 1) bench_tmem(int nw, int op, int iters, int with_mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ long long cyc[16];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 512) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
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
  if (warp < nw) {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = j + threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (op == 0 || op == 2) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + (uint32_t)((i & 1) * 32)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      if (op == 1 || op == 2) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
            "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
            ::"r"(taddr + (uint32_t)((i & 1) * 32)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
              "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
              "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
              "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
    }
    const long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) x ^= r[j];
    if ((threadIdx.x & 31) == 0) cyc[warp] = (t1 - t0) + (x == 0x12345u ? 1 : 0);
  } else if (warp == 15 && with_mma) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | (8u << 24);
    const uint64_t bdesc = make_desc(base + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters * 4; ++i) mma_ts_elect(tmem + 256u, tmem + 448u, bdesc, idesc, 1u);
    commit_elect(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    if ((threadIdx.x & 31) == 0) cyc[15] = clock64() - t0;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long mx = 0;
    for (int w = 0; w < nw; ++w) mx = cyc[w] > mx ? cyc[w] : mx;
    out[0] = mx;
    out[1] = with_mma ? cyc[15] : 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// the KL kernel's tensor work per tile, two issuing warps on independent accumulators: warp 1 = GEMM-2 pattern
// (4 x {N=64, N=32} dependent), warp 2 = GEMM-1 pattern, either the same 4 x {64, 32} (cat = 0) or 12 x N=32 into one
// accumulator (cat = 1).  Reports cycles per "tile" (one group from each warp).
__global__ void __launch_bounds__(128, 1) bench_kl_pattern(int cat, int two_warps, int groups, long long* out) {
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  __shared__ long long cyc[4];
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1);
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
  const uint32_t id64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
  const uint32_t id32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | (8u << 24);
  if (warp == 1 || (warp == 2 && two_warps)) {
    const uint64_t bdesc = make_desc(base + 16384 + (warp == 2 ? 8192 : 0));
    const uint32_t d = tmem + (warp == 2 ? 128u : 0u), a = tmem + (warp == 2 ? 448u : 384u);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (warp == 2 && cat) {
#pragma unroll
        for (int kk = 0; kk < 12; ++kk) mma_ts_elect(d, a + (kk & 3) * 8 + (kk >= 8 ? 32 : 0), bdesc + 2 * (kk & 3) + (kk >= 4 && kk < 8 ? 256 : 0), id32, 1u);
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          mma_ts_elect(d, a + kk * 8, bdesc + 2 * kk, id64, 1u);
          mma_ts_elect(d + 32, a + 32 + kk * 8, bdesc + 2 * kk, id32, 1u);
        }
      }
      commit_elect(smem_u32(&bar[warp - 1]));     // per-tile commit like the kernel (never waited on except at the end)
    }
    // drain: wait for the final phase of this warp's barrier
    mbar_wait(smem_u32(&bar[warp - 1]), (groups - 1) & 1);
    if ((threadIdx.x & 31) == 0) cyc[warp] = clock64() - t0;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = cyc[1]; out[1] = two_warps ? cyc[2] : 0; }
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
  cudaFuncSetAttribute(bench_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("TMEM x32 load/store (32 lanes x 32 columns x 4 B = 4 KB per warp instruction): warps op(0 ld,1 st,2 ld+st) mma -> cycles per iteration, B/clk/SM\n");
  for (int with_mma : {0, 1})
    for (int op : {0, 1, 2})
      for (int nw : {1, 4, 8}) {
        bench_tmem<<<1, 512, 64 * 1024>>>(nw, op, 2000, with_mma, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double cyc = (double)h[0] / 2000;
        printf("  warps=%d op=%d mma=%d : %8.1f cyc/iter  %8.1f B/clk  (mma warp: %.1f cyc per N=32 MMA) %s\n", nw, op, with_mma, cyc,
               nw * 4096.0 * (op == 2 ? 2 : 1) / cyc, with_mma ? (double)h[1] / 8000 : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  cudaFuncSetAttribute(bench_kl_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("KL tensor work per tile (tf32 TS): cat two_warps -> cycles per tile (warp1, warp2)\n");
  for (int two : {0, 1})
    for (int cat : {0, 1}) {
      bench_kl_pattern<<<1, 128, 64 * 1024>>>(cat, two, 1000, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2] = {0, 0};
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("  cat=%d two_warps=%d : %8.1f  %8.1f %s\n", cat, two, (double)h[0] / 1000, (double)h[1] / 1000, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
