// umma.cuh -- tcgen05 (5th-gen tensor core) primitives for sm_100a, written as inline PTX:
// TMEM allocation, shared-memory matrix descriptors, kind::tf32 MMA issue, commit -> mbarrier,
// TMEM <-> register moves.  Bit layouts follow the PTX ISA "tcgen05" matrix / instruction
// descriptors (cross-checked against CUTLASS cute/arch/mma_sm100_desc.hpp).
//
// Operand layout used throughout this code base: SWIZZLE_NONE ("interleaved") canonical layout.
// A row-major matrix Mat[R][C] of 32-bit elements is stored as 8x4 CORE MATRICES (8 rows x 16 bytes,
// 128 contiguous bytes each); core (r/8, c/4) lives at byte offset (r/8)*RG_STRIDE + (c/4)*CG_STRIDE.
// The SAME bytes serve as
//   * a K-major operand (rows = M or N, cols = K):   SBO = RG_STRIDE, LBO = CG_STRIDE; one K=8 MMA
//     step consumes two column groups, so the start address advances by 2*CG_STRIDE per step;
//   * an MN-major operand (rows = K, cols = M or N): SBO = CG_STRIDE, LBO = RG_STRIDE; one K=8 step
//     is one row group, so the start address advances by RG_STRIDE per step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "async.cuh"

namespace tbnn {
namespace umma {

// ---- shared-memory matrix descriptor (64 bit)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr_bytes >> 4) & 0x3FFFu);        // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;    // [16,30) leading-dimension byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;    // [32,46) stride-dimension byte offset >> 4
  d |= (uint64_t)1 << 46;                               // [46,48) descriptor version 1 (Blackwell)
  return d;                                             // base offset 0, LBO mode 0, SWIZZLE_NONE
}

// ---- instruction descriptor (32 bit) for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                            // [4,6)   D format: F32
         | (2u << 7)                          // [7,10)  A format: TF32
         | (2u << 10)                         // [10,13) B format: TF32
         | ((a_mn_major ? 1u : 0u) << 15)     // [15]    A major: 0 = K, 1 = MN
         | ((b_mn_major ? 1u : 0u) << 16)     // [16]    B major
         | ((uint32_t)(N >> 3) << 17)         // [17,23) N >> 3
         | ((uint32_t)(M >> 4) << 24);        // [24,29) M >> 4
}

// ---- TMEM allocation (one full warp executes these)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// ---- MMA issue (ONE thread): D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// The same with the two descriptors passed as 32-bit words (packed to 64 bits inside the asm block).  With 64-bit
// descriptor arithmetic in C++ the compiler keeps the descriptors in vector registers and wraps every MMA in an
// ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY waterfall -- 91 cycles per MMA for ANY N <= 128 (tools/umma_rate.cu), i.e. the
// issuing thread, not the tensor pipe, sets the rate of small-N MMAs.  32-bit words derived from kernel parameters
// and loop counters stay on the uniform datapath (UIADD3 / UMOV feeding UTCHMMA directly).
//   lo word: (address >> 4) & 0x3FFF | (LBO >> 4) << 16        hi word: (SBO >> 4) & 0x3FFF | 1 << 14 (version)
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr_bytes, uint32_t lbo_bytes) {
  return ((saddr_bytes >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void mma_tf32_ss32(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                              uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(d_tmem),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: lane = row, 32-bit column = k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// A from TMEM, B descriptor as 32-bit words (see mma_tf32_ss32)
__device__ __forceinline__ void mma_tf32_ts32(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                              bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 8 / 16 consecutive 32-bit columns (thread t = lane t)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// registers -> TMEM
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// TMEM address of (lane, column) relative to an allocation base
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) {
  return base + ((uint32_t)lane << 16) + (uint32_t)col;
}

// ---- error-compensated split for 3xTF32: x = hi + lo (+ 2^-23 |x|), hi and lo exactly representable in TF32.
// Round to nearest on both parts: the tensor core itself TRUNCATES fp32 operands to TF32 (tools/umma_test.cu), which
// on a truncated split leaves a one-sided error of up to 2^-21 |x| in every product -- visible as a bias in long sums.
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = rn_tf32(x);
  lo = rn_tf32(x - hi);
}

// Split without conversion instructions (cvt.rna.tf32 issues at a fraction of the ALU rate): both parts rounded to
// nearest TF32 with integer arithmetic -- add half an ulp of the 10-bit mantissa, clear the low 13 bits (ties away from
// zero; a carry into the exponent is the correct next binade).  Five full-rate instructions; |lo| <= 2^-12 |x| with
// either sign, so the dropped lo*lo product (<= 2^-24 |a b|) does not pile up one-sidedly.
__device__ __forceinline__ void split_tf32_fast(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
}

// Three-instruction split for the row workers of the training sweep: hi rounded to nearest as above, lo = x - hi left
// as the exact fp32 remainder -- the tensor core truncates it to TF32.  Because hi is rounded to NEAREST, lo has either
// sign, so its truncation (toward zero, <= 2^-21 |x|) does not pile up one-sidedly the way a truncated hi does.
__device__ __forceinline__ void split_tf32_rn_exact(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

// cheaper split for forward-only use (posterior-predictive sweep): hi = truncation (what the tensor core does to
// its operands anyway), lo = exact remainder, truncated again by the hardware -> one-sided error ~2^-21 |x| per
// product, irrelevant next to Monte Carlo error; two ALU instructions instead of two conversions and a subtract
__device__ __forceinline__ void split_tf32_trunc(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

// byte offset of element (r, c) of a row-major matrix stored as 8x4 core matrices
__device__ __forceinline__ uint32_t core_off(int r, int c, uint32_t rg_stride, uint32_t cg_stride) {
  return (uint32_t)(r >> 3) * rg_stride + (uint32_t)(c >> 2) * cg_stride + (uint32_t)(r & 7) * 16u +
         (uint32_t)(c & 3) * 4u;
}

}  // namespace umma
}  // namespace tbnn
