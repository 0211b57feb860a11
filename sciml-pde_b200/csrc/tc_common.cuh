// tcgen05 / TMEM wrappers shared by the tensor-core kernels (head_tc.cu, transform2d_tc.cu).
// Layout facts verified on the hardware by tools/ubench/umma_probe.cu: no-swizzle K-major operands
// (8 x 16-byte core matrices; LBO = byte stride between the 16-byte K chunks -- any multiple of 16,
// 144 keeps shared-memory stores conflict-free --, SBO = byte stride between 8-row groups) work for
// kind::tf32 with M = 128 and N = 32 / 128; MN-major operands returned zeros and are not used.
#pragma once
#include "common.cuh"

namespace fno {
namespace {

// ---- tcgen05 wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst, unsigned ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void tc_mma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                            unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Whole-warp forms: the warp stays converged and one elected lane issues.  With `if (lane == 0) mma(...)`
// the compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY loop and rebuilds the descriptors in the
// uniform datapath (~16 dependent instructions, ~130 cycles per MMA: the issuing thread becomes the
// bottleneck for small-N MMAs -- profiles/r1_h); with the election inside the asm ptxas emits
// back-to-back UTCHMMAs.
__device__ __forceinline__ void tc_mma_tf32_elect(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                                  unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(unsigned long long* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane of its warp's quadrant)
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// A operand read from TENSOR MEMORY (row m in lane m, K along 32-bit columns; verified by
// tools/ubench/umma_tmemA_probe.cu for M = 128), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts_elect(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc,
                                                     unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 registers -> 32 lanes x 16 consecutive fp32 columns (no wait: pair with tmem_st_wait())
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}
// 8 registers -> 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_st8(unsigned taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 / 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 in bits [0,14),
// leading-dimension byte offset >> 4 in [16,30), stride-dimension byte offset >> 4 in [32,46),
// version = 1 in [46,48), layout_type = SWIZZLE_NONE (0) in [61,64)
__device__ __forceinline__ unsigned long long umma_desc(const void* smem, unsigned lbo_bytes, unsigned sbo_bytes) {
  const unsigned long long addr = (smem_u32(smem) & 0x3FFFFu) >> 4;
  return addr | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor for kind::tf32 (cute::UMMA::InstrDescriptor): c_format = F32 (1) at [4,6),
// a_format = b_format = TF32 (2) at [7,10) / [10,13), a_major at 15, b_major at 16 (0 = K-major,
// 1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)a_mn_major << 15) | ((unsigned)b_mn_major << 16) |
         ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// hi = rna_tf32(x) (round to nearest, ties away), lo = x - hi (exact).  cvt.rna.tf32.f32 compiles to four
// instructions (it preserves inf / NaN); the integer form is two and identical for finite values.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

// The same split with the rounding position as run-time values (kernel-uniform, so they cost registers, not
// instructions): (0x1000, 0xffffe000) = tf32 (fp32 and tf32 modes), (0x8000, 0xffff0000) = bfloat16 (bf16 mode: every MMA
// operand is a bf16-representable value, one kind::tf32 pass, fp32 accumulation = the arithmetic of a bf16 tensor-core MMA).
__device__ __forceinline__ void split_rm(float x, float& hi, float& lo, unsigned rnd, unsigned msk) {
  hi = __uint_as_float((__float_as_uint(x) + rnd) & msk);
  lo = x - hi;
}
#define FNO_SPLIT_CONSTS(mode) \
  const unsigned sp_rnd = (mode) == 2 ? 0x8000u : 0x1000u, sp_msk = (mode) == 2 ? 0xffff0000u : 0xffffe000u

}  // namespace
}  // namespace fno
