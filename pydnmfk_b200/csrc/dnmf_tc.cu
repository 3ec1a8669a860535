// tcgen05 / TMA / TMEM path of the A-streaming contractions (fp32 data, k in {16, 32, 64}).
//
//   AH  : V[m x k]   = A[m x n] * H[k x n]^T          (dist_nmf.py:198, :730)
//   WTA : Y^T[n x k] = (W[m x k]^T * A[m x n])^T       (dist_nmf.py:166, :749)
//
// Both stream the resident shard A exactly once from HBM with TMA (cp.async.bulk.tensor, 128B swizzle) into a
// multi-stage shared-memory ring and contract it on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators in TMEM).  fp32 accuracy comes from the 3-term split  A*B ~= Ah*Bh + Ah*Bl + Al*Bh :
//   - the small operand (H or W) is split once per call into Bcat = [B_hi | B_lo]  (2k "N" rows, K-major),
//   - the tensor core reads raw fp32 A from shared memory and uses its top 19 bits (= A_hi),
//   - four "splitter" warps read the tile back from shared memory, compute A_lo = A - A_hi and store it straight
//     into TENSOR MEMORY (tcgen05.st), so the third term's operand costs no shared-memory bandwidth or capacity,
//   - per 32-wide K tile the MMA warp issues 4 x { D[:, 0:2k] += Ah[smem] * Bcat^T ;  D[:, k:2k] += Al[tmem] * Bh^T }.
// The epilogue warps add the two halves of the accumulator and write a per-split partial; the splits are summed
// in a fixed order by reduce_partials_kernel (deterministic).
//
// Warp roles (320 threads, one persistent CTA per SM):  w0 TMA producer | w1 MMA issuer + TMEM owner |
// w2-5 A_lo splitters | w6-9 drain (TMEM chunk -> fp32 registers -> global partial).
#include <cuda.h>

#include "generic_passes.cuh"
#include "tc_api.cuh"

namespace dnmf {
namespace {

constexpr int TC_BM = 128;      // outer tile (rows of A for AH, columns of A for WTA) = UMMA M
constexpr int TC_BK = 32;       // reduced-dimension tile: 32 fp32 = one 128-byte swizzle row = 4 UMMA K steps
constexpr int TC_THREADS = 320;
constexpr int TC_CHUNK = 2;      // K-tiles accumulated in TMEM before the drain warps fold them into registers

// Shared memory per stage: the raw A tile (16 KB) + the Bcat tile.  TMEM (512 columns): NBUF accumulator buffers of
// 2K columns, then one 32-column A_lo slot per stage (the A_lo operand of the third term never touches smem).
template <int K>
struct TcCfg {
  static constexpr int N2 = 2 * K;
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
  static constexpr int B_BYTES = N2 * TC_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (K == 16) ? 10 : (K == 32) ? 9 : 6;
  static constexpr int NBUF = (K == 16) ? 4 : (K == 32) ? 3 : 2;
  static constexpr int ALO_COL0 = NBUF * N2;
  static constexpr int TMEM_COLS = 512;
  static_assert(ALO_COL0 + STAGES * TC_BK <= 512, "TMEM has 512 columns");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB opt-in shared memory limit");
};

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout)
// layout_type: 2 = SWIZZLE_128B (16-byte chunks, 8-row period), 1 = SWIZZLE_128B_BASE32B (32-byte chunks, 4-row
// period) -- the only shared-memory layout the tensor core accepts for MN-major 32-bit operands
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);               // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;      // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;      // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                                 // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;                       // [61,64) layout type
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (2u << 7)                     // a_format = TF32
         | (2u << 10)                    // b_format = TF32
         | ((uint32_t)a_mn_major << 15)  // A major: 0 = K, 1 = MN
         | (0u << 16)                    // B major: K
         | ((uint32_t)(n >> 3) << 17)    // N >> 3
         | ((uint32_t)(TC_BM >> 4) << 24);  // M >> 4
}

__device__ __forceinline__ float tf32_hi(float x, int mode) {
  uint32_t u = __float_as_uint(x);
  if (mode) u += 0x0FFFu + ((u >> 13) & 1u);   // round to nearest even (calibration fallback)
  return __uint_as_float(u & 0xFFFFE000u);      // default: the tensor core drops the low 13 mantissa bits
}

// nearest tf32 with ties away from zero: 2 integer ops (the tie bias is ~3e-8 relative)
__device__ __forceinline__ float tf32_round_up(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// ---------------------------------------------------------------------------------------------------------
// persistent tcgen05 kernel
//   MODE 0 (AH):  X = rows of A (m), reduced = columns (n);  A tile = one TMA box {32 cols, 128 rows}, K-major
//   MODE 1 (WTA): X = columns of A (n), reduced = rows (m);  A tile = four TMA boxes {32 cols, 32 rows}, MN-major,
//                 written with the 128B/32B-atom swizzle (the MN-major layout tf32 operands require)
//   B tile = one TMA box {32, 2K} of Bcat[2K][reduced], K-major.
//   Output: P[split][x][K] (ld = K), x < x_len.
// ---------------------------------------------------------------------------------------------------------
template <int K, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_pass_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               float* __restrict__ P, int64_t split_stride, int64_t x_len, int x_blocks, int kt_total,
               int kt_per_split, int num_units, int hi_mode, int dbg) {
  using Cfg = TcCfg<K>;
  // dbg (DNMF_TC_DBG, timing experiments only, results become wrong): 1 = splitter skips its work, 2 = no chunked
  // drain, 4 = skip the A_lo MMA, 8 = skip the main MMA
  const int chunk = (dbg & 2) ? (1 << 30) : TC_CHUNK;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int N2 = Cfg::N2;
  constexpr int NBUF = Cfg::NBUF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  // per stage: [A 16K][Bcat]
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto split_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto accf_bar = [&](int b) { return bars + 8u * (3 * STAGES + b); };
  auto acce_bar = [&](int b) { return bars + 8u * (3 * STAGES + NBUF + b); };
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 2 * NBUF);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 2 * NBUF));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int xb = unit % x_blocks, sp = unit / x_blocks;
        const int kt0 = sp * kt_per_split;
        const int kt1 = min(kt_total, kt0 + kt_per_split);
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sA = base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          mbar_expect_tx(full_bar(stage), Cfg::A_BYTES + Cfg::B_BYTES);
          if (MODE == 0) {
            tma_load_2d(sA, &tmA, full_bar(stage), kt * TC_BK, xb * TC_BM);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              tma_load_2d(sA + g * 4096, &tmA, full_bar(stage), xb * TC_BM + g * 32, kt * TC_BK);
          }
          tma_load_2d(sB, &tmB, full_bar(stage), kt * TC_BK, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_full = make_idesc(N2, MODE);
      constexpr uint32_t idesc_half = make_idesc(K, 0);     // A_lo comes from TMEM: always [M][K]
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t accphase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int sp = unit / x_blocks;
        const int kt0 = sp * kt_per_split;
        const int kt1 = min(kt_total, kt0 + kt_per_split);
        for (int kt = kt0; kt < kt1; ++kt) {
          const int in_chunk = (kt - kt0) % chunk;
          if (in_chunk == 0) {                      // new accumulation chunk: wait for the drain warps
            mbar_wait(acce_bar(buf), accphase ^ 1u);
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N2);
          mbar_wait(split_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          const uint32_t a_lo_tmem = tmem_base + (uint32_t)(Cfg::ALO_COL0 + stage * TC_BK);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint64_t a_hi;
            if (MODE == 0)     // K-major: 8-row groups 1024 B apart, K step = 32 B inside the swizzled row
              a_hi = make_smem_desc(sA + kk * 32, 16, 1024);
            else               // MN-major (SWIZZLE_128B_BASE32B): 32-column groups 4096 B apart (LBO), 4-row K groups
                               // 512 B apart (SBO); one K=8 step = 8 rows = 1024 B
              a_hi = make_smem_desc(sA + kk * 1024, 4096, 512, 1);
            const uint64_t b = make_smem_desc(sB + kk * 32, 16, 1024);
            // cols [0,K): Ah*Bh (large term, alone in its accumulator);  cols [K,2K): Ah*Bl + Al*Bh (small terms)
            if (!(dbg & 8)) umma_tf32(d_tmem, a_hi, b, idesc_full, (in_chunk > 0 || kk > 0) ? 1u : 0u);
            if (!(dbg & 4)) umma_tf32_ts(d_tmem + K, a_lo_tmem + kk * 8, b, idesc_half, 1u);
          }
          umma_commit(empty_bar(stage));          // frees the smem slot when these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          if (in_chunk == chunk - 1 || kt == kt1 - 1) {
            umma_commit(accf_bar(buf));           // chunk complete -> drain warps
            if (++buf == NBUF) { buf = 0; accphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== A_lo splitters (warps 2-5): smem A tile -> A - hi(A) -> TMEM =====================
    // Thread (q, lane) owns accumulator row q*32+lane (a row of A for AH, a column of A for WTA) and writes its 32
    // K-values of the tile into the stage's A_lo slot in tensor memory.
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int sp = unit / x_blocks;
      const int kt0 = sp * kt_per_split;
      const int kt1 = min(kt_total, kt0 + kt_per_split);
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(full_bar(stage), phase);
        const uint8_t* tile = base_ptr + stage * Cfg::STAGE_BYTES;
        uint32_t lo[32];
        if (dbg & 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) lo[j] = 0;
        } else if (MODE == 0) {
          // row r = 128 contiguous bytes, 16-byte chunk c stored at position c ^ (r & 7)  (SWIZZLE_128B)
          const uint8_t* row = tile + r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(row + ((c ^ (r & 7)) << 4));
            lo[4 * c + 0] = __float_as_uint(tf32_round_up(v.x - tf32_hi(v.x, hi_mode)));
            lo[4 * c + 1] = __float_as_uint(tf32_round_up(v.y - tf32_hi(v.y, hi_mode)));
            lo[4 * c + 2] = __float_as_uint(tf32_round_up(v.z - tf32_hi(v.z, hi_mode)));
            lo[4 * c + 3] = __float_as_uint(tf32_round_up(v.w - tf32_hi(v.w, hi_mode)));
          }
        } else {
          // box q holds columns q*32..q*32+31: K-row j at j*128 bytes, 32-byte chunk (lane>>3) stored at position
          // (lane>>3) ^ (j & 3)  (SWIZZLE_128B_ATOM_32B)
          const uint8_t* box = tile + q * 4096 + (lane & 7) * 4;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = *reinterpret_cast<const float*>(box + j * 128 + ((((lane >> 3) ^ (j & 3))) << 5));
            lo[j] = __float_as_uint(tf32_round_up(v - tf32_hi(v, hi_mode)));
          }
        }
        // a - hi(a) is exact; rounding it to tf32 here (nearest) instead of letting the tensor core truncate it
        // removes the one-sided error of the third term
        if (!(dbg & 1)) {
          tmem_st_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::ALO_COL0 + stage * TC_BK), lo);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== drain warps 6-9: TMEM chunk -> fp32 register accumulators -> global partial ==========
    // The tensor core adds into its fp32 accumulator with truncation; draining every TC_CHUNK K-tiles and summing the
    // chunks here with round-to-nearest keeps that bias at the level of plain fp32 arithmetic.
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    int buf = 0;
    uint32_t accphase = 0;
    constexpr int CH = 16;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int xb = unit % x_blocks, sp = unit / x_blocks;
      const int kt0 = sp * kt_per_split;
      const int kt1 = min(kt_total, kt0 + kt_per_split);
      const int nchunks = (dbg & 2) ? 1 : (kt1 - kt0 + TC_CHUNK - 1) / TC_CHUNK;
      float acc[K];
#pragma unroll
      for (int j = 0; j < K; ++j) acc[j] = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(accf_bar(buf), accphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * N2);
        constexpr int HALF = (K >= 32) ? 32 : K;       // columns folded per batch of loads
#pragma unroll
        for (int h0 = 0; h0 < K; h0 += HALF) {
          uint32_t a[HALF], b[HALF];
#pragma unroll
          for (int j0 = 0; j0 < HALF; j0 += CH) {
            tmem_ld_x16(taddr + h0 + j0, *reinterpret_cast<uint32_t(*)[CH]>(&a[j0]));          // Ah*Bh
            tmem_ld_x16(taddr + K + h0 + j0, *reinterpret_cast<uint32_t(*)[CH]>(&b[j0]));      // Ah*Bl + Al*Bh
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < HALF; ++j) acc[h0 + j] += __uint_as_float(a[j]) + __uint_as_float(b[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(buf));   // TMEM buffer may be overwritten
        if (++buf == NBUF) { buf = 0; accphase ^= 1u; }
      }
      const int64_t x = (int64_t)xb * TC_BM + q * 32 + lane;
      if (x < x_len) {
        float* orow = P + (int64_t)sp * split_stride + x * K;
#pragma unroll
        for (int j = 0; j < K; j += 4)
          *reinterpret_cast<float4*>(orow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// small-operand split: Bcat = [hi(B) ; B - hi(B)] with the reduced dimension contiguous (K-major)
// ---------------------------------------------------------------------------------------------------------
// AH: B = H [k x n] (already K-major)
__global__ void __launch_bounds__(256) tc_split_h_kernel(const float* __restrict__ H, int64_t ldh, float* __restrict__ Bcat,
                                                         int64_t ldb, int k, int64_t n) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)k * n) return;
  const int j = (int)(idx / n);
  const int64_t c = idx % n;
  const float h = H[(int64_t)j * ldh + c];
  const float hi = tf32_hi(h, 1);     // nearest: the low part is then signed and half as large
  Bcat[(int64_t)j * ldb + c] = hi;
  Bcat[(int64_t)(k + j) * ldb + c] = h - hi;
}
// WTA: B = W^T, W is [m x k]: transpose through shared memory (coalesced both ways)
__global__ void __launch_bounds__(256) tc_split_wt_kernel(const float* __restrict__ W, int64_t ldw, float* __restrict__ Bcat,
                                                          int64_t ldb, int k, int64_t m) {
  __shared__ float tile[64][65];
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  for (int idx = threadIdx.x; idx < 64 * k; idx += 256) {
    const int r = idx / k, j = idx % k;
    tile[r][j] = (r0 + r < m) ? W[(r0 + r) * ldw + j] : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * k; idx += 256) {
    const int j = idx / 64, r = idx % 64;
    if (r0 + r < m) {
      const float w = tile[r][j];
      const float hi = tf32_hi(w, 1);
      Bcat[(int64_t)j * ldb + r0 + r] = hi;
      Bcat[(int64_t)(k + j) * ldb + r0 + r] = w - hi;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// 2-D fp32 row-major tensor [rows][cols] with leading dimension ld; box = {box_cols, box_rows}; 128B swizzle;
// out-of-bounds elements read as zero (ragged edges need no special casing in the kernel)
int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(DNMF_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DNMF_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

struct TcPlan {
  int x_blocks, kt_total, kt_per_split, splits, num_units, grid;
  int64_t ldb;            // leading dimension of Bcat (reduced length rounded up to 4)
  int64_t bcat_bytes, partial_bytes;
};

TcPlan tc_plan(int64_t x_len, int64_t r_len, int k) {
  TcPlan p;
  p.x_blocks = (int)ceil_div(x_len, TC_BM);
  p.kt_total = (int)ceil_div(r_len, TC_BK);
  const int sms = sm_count();
  int64_t want = ceil_div((int64_t)12 * sms, p.x_blocks);
  if (want < 1) want = 1;
  if (want > 32) want = 32;
  int per = (int)ceil_div(p.kt_total, want);
  if (per < 16) per = 16;                       // >= 512 reduced elements per unit
  if (per > p.kt_total) per = p.kt_total;
  p.kt_per_split = per;
  p.splits = (int)ceil_div(p.kt_total, per);
  p.num_units = p.x_blocks * p.splits;
  p.grid = p.num_units < sms ? p.num_units : sms;
  p.ldb = round_up(r_len, 4);
  p.bcat_bytes = round_up((int64_t)2 * k * p.ldb * 4, 1024);
  p.partial_bytes = (int64_t)p.splits * x_len * k * 4;
  return p;
}

int g_hi_mode = -1;     // -1 unknown, 0 truncate, 1 round-to-nearest-even, 2 = tensor path unusable
int64_t g_min_elems = -1;

template <int K, int MODE>
int launch_pass(const CUtensorMap& tmA, const CUtensorMap& tmB, float* P, int64_t split_stride, int64_t x_len,
                const TcPlan& pl, int hi_mode, cudaStream_t st) {
  auto kern = tc_pass_kernel<K, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<K>::SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail(e, "tc_pass_kernel smem attribute");
  kern<<<pl.grid, TC_THREADS, TcCfg<K>::SMEM_BYTES, st>>>(tmA, tmB, P, split_stride, x_len, pl.x_blocks, pl.kt_total,
                                                            pl.kt_per_split, pl.num_units, hi_mode,
                                                            getenv("DNMF_TC_DBG") ? atoi(getenv("DNMF_TC_DBG")) : 0);
  DNMF_LAUNCH_CHECK("tc_pass_kernel");
  return 0;
}

template <int MODE>
int launch_pass_k(int k, const CUtensorMap& tmA, const CUtensorMap& tmB, float* P, int64_t split_stride, int64_t x_len,
                  const TcPlan& pl, int hi_mode, cudaStream_t st) {
  switch (k) {
    case 16: return launch_pass<16, MODE>(tmA, tmB, P, split_stride, x_len, pl, hi_mode, st);
    case 32: return launch_pass<32, MODE>(tmA, tmB, P, split_stride, x_len, pl, hi_mode, st);
    case 64: return launch_pass<64, MODE>(tmA, tmB, P, split_stride, x_len, pl, hi_mode, st);
  }
  return fail(DNMF_E_UNSUPPORTED, "tcgen05 path: k must be 16, 32 or 64");
}

// V (AH) or Y / Y^T (WTA) through the tensor path.  ws = [Bcat | partials]
int tc_run(int mode, const float* A, int64_t lda, const float* B, int64_t ldbsrc, float* out, int64_t ldo,
           int64_t m, int64_t n, int k, int transposed_out, int hi_mode, void* ws, int64_t ws_bytes, cudaStream_t st) {
  const int64_t x_len = mode == 0 ? m : n;
  const int64_t r_len = mode == 0 ? n : m;
  const TcPlan pl = tc_plan(x_len, r_len, k);
  const int64_t need = pl.bcat_bytes + pl.partial_bytes;
  if (ws == nullptr || ws_bytes < need)
    return fail(DNMF_E_WORKSPACE, "tcgen05 pass needs %lld workspace bytes, got %lld", (long long)need, (long long)ws_bytes);
  if (((uintptr_t)ws % 256) != 0) return fail(DNMF_E_ARG, "workspace must be 256-byte aligned");
  float* Bcat = reinterpret_cast<float*>(ws);
  float* P = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + pl.bcat_bytes);
  if (mode == 0) {
    tc_split_h_kernel<<<(unsigned)ceil_div((int64_t)k * n, 256), 256, 0, st>>>(B, ldbsrc, Bcat, pl.ldb, k, n);
    DNMF_LAUNCH_CHECK("tc_split_h_kernel");
  } else {
    tc_split_wt_kernel<<<(unsigned)ceil_div(m, 64), 256, 0, st>>>(B, ldbsrc, Bcat, pl.ldb, k, m);
    DNMF_LAUNCH_CHECK("tc_split_wt_kernel");
  }
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if (mode == 0) rc = make_map(&tmA, A, m, n, lda, TC_BK, TC_BM);
  else rc = make_map(&tmA, A, m, n, lda, 32, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = make_map(&tmB, Bcat, 2 * k, r_len, pl.ldb, TC_BK, 2 * k);
  if (rc) return rc;
  const int64_t split_stride = x_len * k;
  rc = mode == 0 ? launch_pass_k<0>(k, tmA, tmB, P, split_stride, x_len, pl, hi_mode, st)
                 : launch_pass_k<1>(k, tmA, tmB, P, split_stride, x_len, pl, hi_mode, st);
  if (rc) return rc;
  // fixed-order sum of the split partials P[s][x][kk]  ->  out
  int64_t so_r, so_c;   // strides of (x, kk) in the output
  if (mode == 0) { so_r = ldo; so_c = 1; }                        // V[x][kk]
  else if (transposed_out) { so_r = ldo; so_c = 1; }              // Y^T[x][kk]
  else { so_r = 1; so_c = ldo; }                                  // Y[kk][x]
  reduce_partials_kernel<float><<<(unsigned)ceil_div(x_len * k, 256), 256, 0, st>>>(P, split_stride, pl.splits, x_len, k,
                                                                                     out, so_r, so_c);
  DNMF_LAUNCH_CHECK("reduce_partials_kernel");
  return 0;
}

// One-time probe: does kind::tf32 truncate or round the fp32 bits it reads?  A = 1 + 1.5 * 2^-11 everywhere, H = 1:
// V = 32 * (1 + 1.5 * 2^-11) exactly iff the splitter's notion of A_hi matches the hardware's.
void calibrate() {
  if (g_hi_mode >= 0) return;
  const char* env = getenv("DNMF_TF32_HI_MODE");
  if (env) { g_hi_mode = atoi(env); return; }
  g_hi_mode = 2;
  const int m = 128, n = 32, k = 16;
  const TcPlan pl = tc_plan(m, n, k);
  const int64_t wsb = pl.bcat_bytes + pl.partial_bytes;
  float *dA = nullptr, *dH = nullptr, *dV = nullptr;
  void* dws = nullptr;
  if (cudaMalloc(&dA, m * n * 4) != cudaSuccess || cudaMalloc(&dH, k * n * 4) != cudaSuccess ||
      cudaMalloc(&dV, m * k * 4) != cudaSuccess || cudaMalloc(&dws, wsb) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  const float aval = 1.0f + 1.5f * 0.00048828125f;   // 1 + 1.5 * 2^-11
  float* hA = new float[m * n];
  float* hH = new float[k * n];
  float* hV = new float[m * k];
  for (int i = 0; i < m * n; ++i) hA[i] = aval;
  for (int i = 0; i < k * n; ++i) hH[i] = 1.0f;
  cudaMemcpy(dA, hA, m * n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dH, hH, k * n * 4, cudaMemcpyHostToDevice);
  const float expect = 32.0f * aval;
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dV, 0, m * k * 4);
    const int64_t saved = tls().launches;
    int rc = tc_run(0, dA, n, dH, n, dV, k, m, n, k, 0, mode, dws, wsb, 0);
    tls().launches = saved;
    if (rc != 0 || cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); break; }
    cudaMemcpy(hV, dV, m * k * 4, cudaMemcpyDeviceToHost);
    bool ok = true;
    for (int i = 0; i < m * k; ++i) ok = ok && (hV[i] == expect);
    if (ok) { g_hi_mode = mode; break; }
  }
  if (getenv("DNMF_VERBOSE")) fprintf(stderr, "[libdnmf] tf32 operand mode: %d (0 trunc, 1 rn, 2 tensor path off)\n", g_hi_mode);
  delete[] hA; delete[] hH; delete[] hV;
  cudaFree(dA); cudaFree(dH); cudaFree(dV); cudaFree(dws);
}

bool device_is_sm100() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess)
      cached = (major == 10) ? 1 : 0;
    else {
      cudaGetLastError();
      cached = 0;
    }
  }
  return cached == 1;
}

}  // namespace

bool tc_eligible(int op, const void* A, int64_t lda, int64_t m, int64_t n, int64_t k, int dtype) {
  if (tls().force_generic) return false;
  if (op != DNMF_OP_AH && op != DNMF_OP_WTA) return false;       // KL path: generic kernels for now
  if (dtype != DNMF_F32) return false;
  if (!(k == 16 || k == 32 || k == 64)) return false;
  if (((uintptr_t)A % 16) != 0 || (lda % 4) != 0) return false;
  if (g_min_elems < 0) {
    const char* env = getenv("DNMF_TC_MIN_ELEMS");
    g_min_elems = env ? atoll(env) : (int64_t)1 << 20;
  }
  if (m * n < g_min_elems) return false;
  if (!device_is_sm100()) return false;
  calibrate();
  return g_hi_mode == 0 || g_hi_mode == 1;
}

void tc_set_min_elems(int64_t elems) { g_min_elems = elems < 0 ? 0 : elems; }

int64_t tc_workspace_bytes(int op, int64_t m, int64_t n, int64_t k, int dtype) {
  if (dtype != DNMF_F32 || !(k == 16 || k == 32 || k == 64)) return 0;
  if (op == DNMF_OP_AH) { const TcPlan p = tc_plan(m, n, (int)k); return p.bcat_bytes + p.partial_bytes; }
  if (op == DNMF_OP_WTA) { const TcPlan p = tc_plan(n, m, (int)k); return p.bcat_bytes + p.partial_bytes; }
  return 0;
}

int tc_ah(const float* A, int64_t lda, const float* H, int64_t ldh, float* V, int64_t ldv, int64_t m, int64_t n, int k,
          int math_mode, void* ws, int64_t ws_bytes, cudaStream_t st) {
  return tc_run(0, A, lda, H, ldh, V, ldv, m, n, k, 0, g_hi_mode, ws, ws_bytes, st);
}

int tc_wta(const float* A, int64_t lda, const float* W, int64_t ldw, float* Y, int64_t ldy, int64_t m, int64_t n, int k,
           int transposed_out, int math_mode, void* ws, int64_t ws_bytes, cudaStream_t st) {
  return tc_run(1, A, lda, W, ldw, Y, ldy, m, n, k, transposed_out, g_hi_mode, ws, ws_bytes, st);
}

int tc_kl_uht(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int,
              float, int, void*, int64_t, cudaStream_t) {
  return fail(DNMF_E_UNSUPPORTED, "tcgen05 KL path not built");
}
int tc_kl_wtu(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int,
              float, int, int, void*, int64_t, cudaStream_t) {
  return fail(DNMF_E_UNSUPPORTED, "tcgen05 KL path not built");
}

}  // namespace dnmf
