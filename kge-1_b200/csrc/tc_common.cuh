// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the TMA tensor-map helper shared by the tensor-tile
// kernels (tc_dot.cu, tc_bwd.cu).  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace kgeb {
namespace tc {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per warp: every lane has finished (and fenced) its part, __syncwarp orders the lanes' accesses before
// lane 0's arrive.  512 per-thread arrivals on one barrier word are 512 serialised shared-memory atomics per tile.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (mbarrier.test_wait): true once the phase with the given parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// The retry loop lives INSIDE the asm statement: to the compiler the wait is straight-line code, so a warp that only
// waits and issues (the MMA issuers) stays in provably uniform control flow and its descriptor arithmetic runs on the
// uniform datapath.  A C++ `while (!try_wait)` loop has a per-thread exit condition, after which the compiler treats
// every value as potentially divergent (R2UR.BROADCAST of each tcgen05.mma operand, ~18 instructions per MMA).
// KGEB_TRYWAIT_HINT_NS > 0: try_wait carries a suspend-time hint, so a waiting warp sleeps in hardware until the phase
// completes instead of returning (and re-issuing the probe) after the short default limit.  Every probe is a SYNCS
// instruction in the MIO queue, the same queue the epilogues' MUFU / shared-memory instructions go through: the ncu source
// view counted 2.0 M probes against 3.7 M MUFU instructions in one dTable launch (profiles/README.md).
#ifndef KGEB_TRYWAIT_HINT_NS
#define KGEB_TRYWAIT_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if KGEB_TRYWAIT_HINT_NS > 0
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"((uint32_t)KGEB_TRYWAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
#endif
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int32_t c_inner, int32_t c_row,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_row)
      : "memory");
}
// smem tile -> global with an element-wise fp32 add performed by the L2 (bulk async-group completion)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int32_t c_inner,
                                                  int32_t c_row) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_row)
               : "memory");
}
// smem tile -> global, plain store (bulk async-group completion); rows / columns outside the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c_inner, int32_t c_row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_row)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"((uint32_t)NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)NCOLS) : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, mma_sm100_desc.hpp):
//   [0,14) start>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format TF32 (2) @7/@10,
// a_major @15, b_major @16 (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24
// fmt: 2 = TF32 (kind::tf32), 1 = BF16 (kind::f16)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, int fmt = 2) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// element traits of the two tensor-core operand types: fp32 read as TF32, or BF16 mirrors
template <bool BF16>
struct Elem {
  static constexpr int kBytes = BF16 ? 2 : 4;
  static constexpr int kSlabK = 128 / kBytes;   // elements per 128-byte swizzle row (32 | 64)
  static constexpr int kUmmaK = 32 / kBytes;    // K per tcgen05.mma (8 | 16): always 32 bytes of a row
  static constexpr int kFmt = BF16 ? 1 : 2;
};
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
  if (BF16) umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}
// One lane of a converged warp (elect.sync; the same lane every time for the full mask).  Code predicated on it is
// single-threaded *as far as the compiler knows*, so operands of the uniform-datapath instructions (tcgen05.mma
// descriptors, commit addresses) are moved to uniform registers directly; under a plain `lane == 0` test every such
// instruction is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Single-instruction MUFU forms (flush-to-zero, no range fix-up code around them): the tile epilogues are bound by
// the MUFU pipe (16 lanes/clk/SM) and the warp issue slots, so every guard instruction the fast-math intrinsics add
// (denormal rescale around ex2, the multiply of __fdividef) costs throughput.  Relative error ~2^-22.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 2^x on the FMA / ALU pipes (Cody-Waite split with the round-to-nearest magic constant, minimax polynomial on
// [-0.5, 0.5], exponent inserted with one integer multiply-add).  The tile epilogues are bound by the 16-lane MUFU
// pipe; evaluating a compile-time fraction of the exponentials here (FlashAttention-4's trick) moves that work to
// pipes with spare issue slots.  Relative error: degree 3 -> 7.5e-5 (results that are rounded to bf16 anyway),
// degree 4 -> 2.7e-6 (loss statistics).  x is clamped to the normal exponent range; CLAMP_HI = false when x <= ~0.
template <int DEG, bool CLAMP_HI>
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  if (CLAMP_HI) x = fminf(x, 126.f);
  const float xf = x + 12582912.f;                  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (xf - 12582912.f);            // x - round(x)
  float p;
  if (DEG == 3) {
    p = fmaf(fmaf(fmaf(0.0551716685f, f, 0.2426111251f), f, 0.6932609677f), f, 0.9999280572f);
  } else {
    p = fmaf(fmaf(fmaf(fmaf(0.009570102207f, f, 0.05591785908f), f, 0.2402474433f), f, 0.6931217909f), f,
             0.9999992847f);
  }
  return __int_as_float(__float_as_int(xf) * (1 << 23) + __float_as_int(p));   // p * 2^round(x)
}
// exponentials of columns (c & 7) < N8 of an unrolled epilogue loop go to the polynomial, the others to the MUFU.
// Measured on B200 (tools/build_variant.sh, bench.py + bench_extra.py wd5m-1vsall): every N8 > 0 was slower than
// MUFU-only (BCE dQ 62.9 -> 67.3 / 72.4 us at N8 = 3 / 6, dTable 63.5 -> 64.4 / 67.2 us; KL step 16.4 -> 17.2 /
// 17.6 ms): these epilogues are bound by dependent-issue latency, not by MUFU throughput alone, so the extra
// FMA-pipe instructions cost more than the MUFU slots they free.  Default 0; kept for re-tuning.
#ifndef KGEB_POLY8_BCE
#define KGEB_POLY8_BCE 0
#endif
#ifndef KGEB_POLY8_KL
#define KGEB_POLY8_KL 0
#endif
#ifndef KGEB_POLY8_STATS
#define KGEB_POLY8_STATS 0
#endif
// MUFU diet of the BCE epilogues (tc_bwd.cu, tc_dot.cu): reciprocals of 2 | 4 columns from one rcp of their product;
// one lg2 per KGEB_LG2_GROUP columns (of the product of the 1 + e^-|z| factors) instead of one per column.
// Measured (bench.py, B=4096, tools/build_variant.sh): RCP_GROUP / LG2_GROUP = 1/1 158.7 us per step, 4/1 152.2,
// 2/8 151.3, 4/16 148.9, 4/8 144.8 (dTable kernel alone 64.8 -> 58.4 us, dQ 62.4 -> 58.4 us).
#ifndef KGEB_RCP_GROUP
#define KGEB_RCP_GROUP 4
#endif
#ifndef KGEB_LG2_GROUP
#define KGEB_LG2_GROUP 8
#endif

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// registers -> tensor memory, 32 lanes x 16 columns (the mirror image of tmem_ld16)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// tcgen05.mma with the A operand in tensor memory (K-major, lane = row) and B from a shared-memory descriptor
// (cute::SM100_MMA_F16BF16_TS): the zero "disable output lane" mask keeps all lanes enabled
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __uint_as_float(r);
}

__device__ __forceinline__ int64_t lb_i64(const int64_t* a, int64_t lo, int64_t hi, int64_t v) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}


// 2-D row-major [rows, d] tensor map (fp32 or bf16): box = 128 bytes of a row x box_rows, 128-byte swizzle,
// out-of-bounds elements read as zeros
int make_map(CUtensorMap* map, const void* base, int64_t rows, int d, int box_rows, bool bf16 = false);

}  // namespace tc
}  // namespace kgeb
