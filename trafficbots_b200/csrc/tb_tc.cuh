// sm_100a tensor-core building blocks: tcgen05.mma with TMEM accumulators, operands in 128-byte-swizzled K-major
// shared-memory tiles, weights staged by bulk-async copies (cp.async.bulk, the TMA engine's linear mode) that
// complete on mbarriers, accumulators read back with tcgen05.ld.
//
// Numerics: every logical fp32 GEMM  D = X W^T  is issued as THREE bf16 MMAs with fp32 accumulation,
//     X_hi W_hi^T + X_lo W_hi^T + X_hi W_lo^T,     v_hi = bf16_rn(v), v_lo = bf16_rn(v - v_hi)
// ("bf16x3"): the dropped lo*lo term and the rounding of the lo parts are <= 2^-17 relative per product, so the
// closed-loop rollout stays within the fp32 parity tolerance while running on the tensor pipe.
//
// Layout of a K-major operand tile (R rows x 64 bf16 = one 128-byte swizzle atom wide, "K-block"):
//     byte(r, c16) = (r / 8) * 1024 + (r % 8) * 128 + ((c16 ^ (r % 8)) * 16)          c16 = 16-byte chunk 0..7
// K-blocks of a K = 128 operand follow each other (R * 128 bytes apart).  The UMMA smem descriptor for it:
// start address >> 4, SBO = 1024 B (distance between 8-row groups), SWIZZLE_128B, version 1; a K = 16 MMA step
// advances the start address by 32 bytes inside the atom.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {
namespace tc {

constexpr int BLOCK_BYTES = 65536;  // one packed 128x128 weight block: [hi kb0 | hi kb1 | lo kb0 | lo kb1], 16 KB each
constexpr int KB_BYTES_128 = 16384; // one K-block (64 bf16 wide) of a 128-row tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- warp-uniform issue ---------------------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit take their descriptors and addresses in UNIFORM registers.  If the issuing code runs under
// `if (threadIdx.x == 0)` the compiler cannot prove the operands warp-uniform and wraps every instruction in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~14 instructions, ~90 cycles per MMA).  Instead the whole warp runs the issue
// code on values marked uniform with `uniform()` and one elected lane executes the instruction: 2 instructions per MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- mbarrier -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifndef TB_WATCHDOG_CYCLES
// Limit of the bounded mbarrier waits in SM clock cycles: ~60 s at 1.965 GHz.  Legitimate waits are microseconds; the bound is
// there so that a protocol bug fails the launch instead of hanging the GPU forever, and it is far above anything a
// time-sliced context (a GPU shared with another process) or a stopped debugger session of reasonable length adds.
// -DTB_WATCHDOG_CYCLES=0 compiles the check out (release builds that prefer a hang to a poisoned context);
// larger values for compute-sanitizer runs, which slow the persistent kernels by orders of magnitude.
#define TB_WATCHDOG_CYCLES 120000000000LL
#endif
__device__ __forceinline__ void watchdog_check(long long t0) {
#if TB_WATCHDOG_CYCLES > 0
  if (clock64() - t0 > TB_WATCHDOG_CYCLES) __trap();  // the host sees TB_ERR_LAUNCH on the next call (sticky CUDA error)
#endif
}
// Bounded wait: a protocol bug must trap (fail the launch) instead of hanging the GPU.  The poll carries a suspend-time hint, so
// a waiting warp sleeps in hardware until the phase completes instead of spinning: measured on the decode kernel against a
// plain try_wait spin, -2.4 % step time (spinning roles steal issue / shared-memory slots from the working warps; a
// __nanosleep back-off is worse than either, +3.5 %, because the wake-up latency sits on the serial GEMM -> epilogue chain).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_hint(bar, parity, 1000000u))
    watchdog_check(t0);
}

// ---- bulk async copy global -> shared (TMA linear mode), completes `bytes` on the mbarrier ---------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- fences -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) ---------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor: start[0,14) LBO[16,30) SBO[32,46)
// version[46,48)=1 layout_type[61,64)=2)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                  // leading byte offset: unused for swizzled K-major (canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16, A/B = bf16 K-major, D = fp32 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4)                         // c_format = F32
         | (1u << 7)                       // a_format = BF16
         | (1u << 10)                      // b_format = BF16
         | ((uint32_t)(N >> 3) << 17)      // n_dim
         | ((uint32_t)(M >> 4) << 24);     // m_dim
}

// ---- MMA issue (one thread) -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column = two consecutive bf16 K elements)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier when all MMAs issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[128 x N] (+)= A[128 x K] B[N x K]^T for K-major SW128 tiles whose K-blocks are `a_kb_bytes` / `b_kb_bytes` apart;
// k_elems = K (multiple of 16).  Issues K/16 MMAs.
__device__ __forceinline__ void mma_tile(uint32_t tmem_d, uint32_t a_addr, uint32_t a_kb_bytes, uint32_t b_addr, uint32_t b_kb_bytes,
                                         int k_elems, uint32_t idesc, bool accumulate_first) {
  for (int k = 0; k < k_elems; k += 16) {
    const uint32_t kb = k >> 6, ko = (k & 63) * 2;
    mma_bf16(tmem_d, make_desc_sw128(a_addr + kb * a_kb_bytes + ko), make_desc_sw128(b_addr + kb * b_kb_bytes + ko), idesc,
             (accumulate_first || k > 0) ? 1u : 0u);
  }
}

// ---- TMEM <-> registers: warp w touches lanes 32*(w%4) .. +31, thread = one lane (row), 32 consecutive columns ---------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- bf16x3 operand split and swizzled tile stores -------------------------------------------------------------------------
// 8 consecutive K elements of one row -> one 16-byte chunk of the hi tile and one of the lo tile
// ---- packed fp32 pairs (FFMA2 / FADD2 on sm_100: one issue slot for two lanes; same IEEE rounding as the scalar forms) ----------
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  unsigned long long a, b, c, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void sub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
// bf16x3 split of two floats: h = packed RN-bf16 (v0 low half, v1 high half), l = packed RN-bf16 of the residuals.  Raw
// cvt / shift / mask (6 instructions per pair): the cuda_bf16.hpp struct accessors cost two extra PRMTs per conversion.
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& h, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
  float r0, r1;
  sub2(r0, r1, v0, v1, __uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// 32 consecutive K elements of one row -> 16 packed bf16x2 words of the hi part and 16 of the lo part (TMEM A operand)
__device__ __forceinline__ void split32_packed(const float (&v)[32], float (&hi)[16], float (&lo)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint32_t h, l;
    split_pair(v[2 * i], v[2 * i + 1], h, l);
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(l);
  }
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// byte offset of 16-byte chunk `c16` (0..7) of row `r` inside one K-block of a swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4)); }

// store 32 consecutive fp32 values (columns k0 .. k0+31, k0 multiple of 32) of row r as bf16 hi / lo into 128-row operand
// tiles (`hi` / `lo` = tile base pointers, K-blocks 16 KB apart)
__device__ __forceinline__ void store_row32_split(unsigned char* hi, unsigned char* lo, int r, int k0, const float (&v)[32]) {
  const int kb = k0 >> 6, c0 = (k0 & 63) >> 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 h, l;
    split8(v + 8 * c, h, l);
    const uint32_t off = kb * KB_BYTES_128 + sw128_off(r, c0 + c);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

}  // namespace tc
}  // namespace tb
