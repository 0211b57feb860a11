// K1 on the tensor cores: the contiguous-axis half of the pruned forward transform as a truncated-DFT
// GEMM (tcgen05.mma kind::tf32, 3xTF32 split for fp32-mode accuracy, accumulators in TMEM).
//
//   T1[row, q] = sum_w x[row, w] F[q, w],   F[q < m2] = cos(2 pi q w / W),  F[m2 + q] = sin(2 pi q w / W)
//
// over ALL rows of all planes at once: a tensor [planes, H, W] is a [planes*H, W] matrix whose rows are
// contiguous, the W-axis transform is independent per row, so 128-row tiles ignore plane boundaries
// and a tile is one contiguous 128*W*4-byte block of HBM.  Per element the CUDA cores only load,
// (optionally multiply by GELU'), split into hi / lo and store -- ~4 instructions instead of the ~39
// of the FP32 kernel (profiles/r1_e) -- so the kernel streams at the HBM rate.  The strided-axis
// half (130 rows x 24 reals per plane, 18 % of the bytes) is finished by hpass2d_kernel
// (transform2d.cu) on the FP32 pipes.
//
// B = F[q][w] hi / lo lives in the plan, already in the K-major no-swizzle UMMA layout.  (A first version staged the
// A operand hi / lo in shared memory -- 8 B of staging stores + 12 B of operand fetches per 4 B from HBM -- and lost to
// the FP32 kernel; profiles/r1_e.  It is not part of the library any more.)
#include "tc_common.cuh"

namespace fno {
namespace {

constexpr int TW_M = 128;                  // rows per tile = TMEM lanes
constexpr int TW_NQ = 32;                  // accumulator columns (2 * m2 <= 32)

// ------------------------------------------------------------------------------------------
// Plain K1 (no GELU' premultiply): the A operand goes through TENSOR memory.
//
// A 128-row tile of the [planes * H, W] matrix is one contiguous block of HBM: a single cp.async.bulk
// brings it into a 2-deep shared-memory ring, converter warps (thread = row = TMEM lane) read their row
// with conflict-free LDS.64, split it hi / lo in registers and write both halves to TMEM with tcgen05.st
// (row = lane, K along columns is the layout of an A operand in TMEM), and the truncated-DFT GEMM
// T1[row, q] = sum_w x[row, w] F[q, w] runs with A from TMEM and F from shared memory.  Per 4 bytes from HBM
// shared memory carries 4 B in and 4 B out (v1: 8 B of staging stores + 12 B of MMA operand fetches), the
// loads are fully asynchronous (two 66 KB tiles in flight per SM) and the CUDA cores touch each element with
// ~4 instructions.  A is written in two K halves so that the MMAs of one half run under the conversion of
// the other; the accumulator is double-buffered for the epilogue (T1 rows to global, 16-byte stores).
// ------------------------------------------------------------------------------------------
constexpr int TA_CONV_WARPS = 16;          // 4 lane quadrants x 4 K quarters (two quarters per K half)
constexpr int TA_MMA_WARP = 16, TA_LOAD_WARP = 17, TA_EPI_WARP0 = 18;
constexpr int TA_THREADS = 32 * 22;
constexpr int TA_MAXK = 136;               // K padding limit: 17 k-steps; A hi | lo = 272 TMEM columns
constexpr unsigned TA_TM_LO = TA_MAXK;     // column of the lo half of A
constexpr unsigned TA_TM_D = 288;          // 2 x 32 accumulator columns
constexpr unsigned TA_TM_COLS = 512;

__global__ void __launch_bounds__(TA_THREADS, 1)
fwd2d_tca_kernel(const float* __restrict__ x, float* __restrict__ T1, const float* __restrict__ Fhi,
                 const float* __restrict__ Flo, long R, int W, int nch, int TQ, int total_tiles, int single) {
  FNO_SPLIT_CONSTS(single);                  // single: 0 = 3xTF32 (fp32 mode), 1 = tf32, 2 = bf16 operands
  extern __shared__ __align__(128) unsigned char wsm[];
  const int raw_bytes = TW_M * W * 4;                 // one tile of rows (multiple of 512)
  const int f_bytes = 4 * nch * 128;
  float* raw = reinterpret_cast<float*>(wsm);         // [2][128 * W]
  unsigned char* f_hi = wsm + 2 * raw_bytes;
  unsigned char* f_lo = f_hi + f_bytes;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(f_lo + f_bytes);
  unsigned long long* raw_full = bars;      // [2] bulk copy landed
  unsigned long long* raw_free = bars + 2;  // [2] converters done reading
  unsigned long long* a_ready = bars + 4;   // [2 K halves] hi / lo of this half are in TMEM
  unsigned long long* a_free = bars + 6;    // [2 K halves] the MMAs that read this half are done
  unsigned long long* d_full = bars + 8;    // [2] accumulator complete
  unsigned long long* d_free = bars + 10;   // [2] accumulator read back
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 12);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(raw_full + s, 1);
      mbar_init(raw_free + s, TA_CONV_WARPS);
      mbar_init(a_ready + s, TA_CONV_WARPS / 2);
      mbar_init(a_free + s, 1);
      mbar_init(d_full + s, 1);
      mbar_init(d_free + s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TA_MMA_WARP) tmem_alloc(tmem_slot, TA_TM_COLS);
  for (int i = tid; i < f_bytes / 16; i += TA_THREADS) {
    reinterpret_cast<float4*>(f_hi)[i] = __ldg(reinterpret_cast<const float4*>(Fhi) + i);
    reinterpret_cast<float4*>(f_lo)[i] = __ldg(reinterpret_cast<const float4*>(Flo) + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  const int ksteps = nch / 2;                          // K = 8 per MMA
  const int kh0 = (ksteps + 1) / 2;                    // k-steps of the first K half
  const int ntl = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == TA_LOAD_WARP) {
    // ---- producer: one bulk copy per tile ---------------------------------------------------------
    for (int it = 0; it < ntl; ++it) {
      const int s = it & 1;
      mbar_wait(raw_free + s, (((unsigned)it >> 1) & 1u) ^ 1u);
      const long row0 = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M;
      const long rows = (R - row0 < TW_M) ? R - row0 : TW_M;
      const unsigned bytes = (unsigned)(rows * W * 4);   // multiple of 16 (checked on the host)
      if (lane == 0) {
        mbar_arrive_expect_tx(raw_full + s, bytes);
        bulk_g2s(raw + (size_t)s * TW_M * W, x + (size_t)row0 * W, bytes, raw_full + s);
      }
      __syncwarp();
    }
  } else if (warp < TA_CONV_WARPS) {
    // ---- converters: thread = row (TMEM lane), one K half ------------------------------------------
    const int quad = warp & 3, kq = warp >> 2, half = kq >> 1;
    const int row = quad * 32 + lane;
    const int hs0 = half ? kh0 : 0, hs1 = half ? ksteps : kh0, hm = (hs0 + hs1 + 1) / 2;   // this half, split in two
    const int ks0 = (kq & 1) ? hm : hs0, ks1 = (kq & 1) ? hs1 : hm;
    const unsigned ta = tmem_base + ((unsigned)(quad * 32) << 16);
    for (int it = 0; it < ntl; ++it) {
      const int s = it & 1;
      const long row0 = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M;
      const bool live = row0 + row < R;
      mbar_wait(raw_full + s, ((unsigned)it >> 1) & 1u);
      mbar_wait(a_free + half, ((unsigned)it & 1u) ^ 1u);     // the previous tile's MMAs on this half are done
      tc_fence_after();
      const float* __restrict__ rp = raw + (size_t)s * TW_M * W + (size_t)row * W;
#pragma unroll 2
      for (int ks = ks0; ks < ks1; ++ks) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const int w = 8 * ks + e;                            // W is even: pairs are all-in or all-out
          float2 v = make_float2(0.f, 0.f);
          if (live && w < W) v = *reinterpret_cast<const float2*>(rp + w);
          split_rm(v.x, hi[e], lo[e], sp_rnd, sp_msk);
          split_rm(v.y, hi[e + 1], lo[e + 1], sp_rnd, sp_msk);
        }
        tmem_st8(ta + (unsigned)(8 * ks), hi);
        if (!single) tmem_st8(ta + TA_TM_LO + (unsigned)(8 * ks), lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(raw_free + s);
        mbar_arrive(a_ready + half);
      }
    }
  } else if (warp == TA_MMA_WARP) {
    // ---- MMA issuer (whole warp converged, one elected lane issues) ----------------------------------
    constexpr unsigned idesc = umma_idesc_tf32(TW_M, TW_NQ, 0, 0);
    const unsigned long long d_fh = umma_desc(f_hi, 128, nch * 128), d_fl = umma_desc(f_lo, 128, nch * 128);
    for (int it = 0; it < ntl; ++it) {
      const int d = it & 1;
      mbar_wait(d_free + d, (((unsigned)it >> 1) & 1u) ^ 1u);
      const unsigned td = tmem_base + TA_TM_D + (unsigned)(d * TW_NQ);
      for (int half = 0; half < 2; ++half) {
        mbar_wait(a_ready + half, (unsigned)it & 1u);
        tc_fence_after();
        __syncwarp();
        const int ks0 = half ? kh0 : 0, ks1 = half ? ksteps : kh0;
#pragma unroll 1
        for (int ks = ks0; ks < ks1; ++ks) {
          const unsigned long long fo = (unsigned long long)(ks * (256 >> 4));
          const unsigned ah = tmem_base + (unsigned)(8 * ks), al = ah + TA_TM_LO;
          if (!single) {                                               // tf32 mode: the hi * hi pass alone
            tc_mma_tf32_ts_elect(td, al, d_fh + fo, idesc, ks != 0);   // lo * hi
            tc_mma_tf32_ts_elect(td, ah, d_fl + fo, idesc, 1u);        // hi * lo
          }
          tc_mma_tf32_ts_elect(td, ah, d_fh + fo, idesc, single ? (unsigned)(ks != 0) : 1u);   // hi * hi
        }
        tc_commit_elect(a_free + half);
      }
      tc_commit_elect(d_full + d);
    }
  } else {
    // ---- epilogue: T1 rows to global memory ----------------------------------------------------------
    const int quad = warp & 3;
    for (int it = 0; it < ntl; ++it) {
      const int d = it & 1;
      mbar_wait(d_full + d, ((unsigned)it >> 1) & 1u);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((unsigned)(quad * 32) << 16) + TA_TM_D + (unsigned)(d * TW_NQ), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_free + d);
      const long row = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M + quad * 32 + lane;
      if (row < R) {
        float4* __restrict__ o = reinterpret_cast<float4*>(T1 + row * TQ);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (4 * q < TQ) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TA_MMA_WARP) tmem_dealloc(tmem_base, TA_TM_COLS);
}

// ------------------------------------------------------------------------------------------
// The same with the GELU' premultiply (backward of F.gelu fused in front of the transform): x = g, preact = s,
// dS = g * gelu'(s) is both the transform's input and an output tensor.  Same structure as the plain
// kernel; the shared-memory ring holds HALF tiles (64 rows of g and of s per slot, lane quadrants 0-1 work on
// slot 0, quadrants 2-3 on slot 1), the converters overwrite g with dS in place and the producer warp sends
// the slot back to HBM with one bulk store before it refills it.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TA_THREADS, 1)
fwd2d_tcap_kernel(const float* __restrict__ x, const float* __restrict__ preact, float* __restrict__ ds_out,
                  float* __restrict__ T1, const float* __restrict__ Fhi,
                 const float* __restrict__ Flo, long R, int W, int nch, int TQ, int total_tiles, int single) {
  FNO_SPLIT_CONSTS(single);                  // single: 0 = 3xTF32 (fp32 mode), 1 = tf32, 2 = bf16 operands
  extern __shared__ __align__(128) unsigned char wsm[];
  constexpr int HR = TW_M / 2;                        // rows per half-tile slot
  const int slot_bytes = HR * W * 4;                  // multiple of 256
  const int f_bytes = 4 * nch * 128;
  float* g_s = reinterpret_cast<float*>(wsm);         // [2 slots][64 * W]: g, overwritten in place by dS = g * gelu'(s)
  float* s_s = g_s + 2 * HR * W;                      // [2 slots][64 * W]: pre-activation
  unsigned char* f_hi = wsm + 4 * slot_bytes;
  unsigned char* f_lo = f_hi + f_bytes;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(f_lo + f_bytes);
  unsigned long long* raw_full = bars;      // [2 slots] bulk copies of g and s landed
  unsigned long long* conv_done = bars + 2; // [2 slots] converters done: dS is in the g slot
  unsigned long long* a_ready = bars + 4;   // [2 K halves] hi / lo of this half are in TMEM
  unsigned long long* a_free = bars + 6;    // [2 K halves] the MMAs that read this half are done
  unsigned long long* d_full = bars + 8;    // [2] accumulator complete
  unsigned long long* d_free = bars + 10;   // [2] accumulator read back
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 12);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(raw_full + s, 1);
      mbar_init(conv_done + s, TA_CONV_WARPS / 2);
      mbar_init(a_ready + s, TA_CONV_WARPS / 2);
      mbar_init(a_free + s, 1);
      mbar_init(d_full + s, 1);
      mbar_init(d_free + s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TA_MMA_WARP) tmem_alloc(tmem_slot, TA_TM_COLS);
  for (int i = tid; i < f_bytes / 16; i += TA_THREADS) {
    reinterpret_cast<float4*>(f_hi)[i] = __ldg(reinterpret_cast<const float4*>(Fhi) + i);
    reinterpret_cast<float4*>(f_lo)[i] = __ldg(reinterpret_cast<const float4*>(Flo) + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  const int ksteps = nch / 2;                          // K = 8 per MMA
  const int kh0 = (ksteps + 1) / 2;                    // k-steps of the first K half
  const int ntl = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == TA_LOAD_WARP) {
    // ---- producer: per half-tile slot, the dS bulk store of the previous tile, then this tile's two loads ----
    auto store_ds = [&](int it, int h) {              // lane 0 only
      const long r0 = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M + h * HR;
      const long rows = (R - r0 < HR) ? R - r0 : HR;
      if (rows > 0 && ds_out != nullptr) {
        bulk_s2g(ds_out + (size_t)r0 * W, g_s + (size_t)h * HR * W, (unsigned)(rows * W * 4));
        bulk_commit();
      }
    };
    for (int it = 0; it < ntl; ++it) {
      for (int h = 0; h < 2; ++h) {
        if (it > 0) {
          mbar_wait(conv_done + h, (unsigned)(it - 1) & 1u);
          if (lane == 0) {
            store_ds(it - 1, h);
            bulk_wait_read<0>();                      // the slot may be overwritten once the store has read it
          }
          __syncwarp();
        }
        const long r0 = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M + h * HR;
        const long rows = (R - r0 < HR) ? R - r0 : HR;
        if (lane == 0) {
          if (rows > 0) {
            const unsigned bytes = (unsigned)(rows * W * 4);   // multiple of 16 (checked on the host)
            mbar_arrive_expect_tx(raw_full + h, 2 * bytes);
            bulk_g2s(g_s + (size_t)h * HR * W, x + (size_t)r0 * W, bytes, raw_full + h);
            bulk_g2s(s_s + (size_t)h * HR * W, preact + (size_t)r0 * W, bytes, raw_full + h);
          } else {
            mbar_arrive(raw_full + h);
          }
        }
        __syncwarp();
      }
    }
    for (int h = 0; h < 2; ++h) {
      mbar_wait(conv_done + h, (unsigned)(ntl - 1) & 1u);
      if (lane == 0) store_ds(ntl - 1, h);
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp < TA_CONV_WARPS) {
    // ---- converters: thread = row (TMEM lane), one K half ------------------------------------------
    const int quad = warp & 3, kq = warp >> 2, half = kq >> 1;
    const int row = quad * 32 + lane;
    const int slot = quad >> 1, lrow = row - slot * HR;        // half-tile slot of this lane quadrant
    const int hs0 = half ? kh0 : 0, hs1 = half ? ksteps : kh0, hm = (hs0 + hs1 + 1) / 2;   // this half, split in two
    const int ks0 = (kq & 1) ? hm : hs0, ks1 = (kq & 1) ? hs1 : hm;
    const unsigned ta = tmem_base + ((unsigned)(quad * 32) << 16);
    for (int it = 0; it < ntl; ++it) {
      const long row0 = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M;
      const bool live = row0 + row < R;
      mbar_wait(raw_full + slot, (unsigned)it & 1u);
      mbar_wait(a_free + half, ((unsigned)it & 1u) ^ 1u);     // the previous tile's MMAs on this half are done
      tc_fence_after();
      float* __restrict__ rp = g_s + (size_t)slot * HR * W + (size_t)lrow * W;
      const float* __restrict__ sp = s_s + (size_t)slot * HR * W + (size_t)lrow * W;
#pragma unroll 2
      for (int ks = ks0; ks < ks1; ++ks) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const int w = 8 * ks + e;                            // W is even: pairs are all-in or all-out
          float2 v = make_float2(0.f, 0.f);
          if (live && w < W) {
            const float2 gv = *reinterpret_cast<const float2*>(rp + w);
            const float2 sv = *reinterpret_cast<const float2*>(sp + w);
            v = make_float2(gv.x * gelu_fast_grad(sv.x), gv.y * gelu_fast_grad(sv.y));
            *reinterpret_cast<float2*>(rp + w) = v;            // dS, bulk-stored by the producer warp
          }
          split_rm(v.x, hi[e], lo[e], sp_rnd, sp_msk);
          split_rm(v.y, hi[e + 1], lo[e + 1], sp_rnd, sp_msk);
        }
        tmem_st8(ta + (unsigned)(8 * ks), hi);
        if (!single) tmem_st8(ta + TA_TM_LO + (unsigned)(8 * ks), lo);
      }
      tmem_st_wait();
      fence_proxy_async();                                     // dS writes -> visible to the bulk store
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(conv_done + slot);
        mbar_arrive(a_ready + half);
      }
    }
  } else if (warp == TA_MMA_WARP) {
    // ---- MMA issuer (whole warp converged, one elected lane issues) ----------------------------------
    constexpr unsigned idesc = umma_idesc_tf32(TW_M, TW_NQ, 0, 0);
    const unsigned long long d_fh = umma_desc(f_hi, 128, nch * 128), d_fl = umma_desc(f_lo, 128, nch * 128);
    for (int it = 0; it < ntl; ++it) {
      const int d = it & 1;
      mbar_wait(d_free + d, (((unsigned)it >> 1) & 1u) ^ 1u);
      const unsigned td = tmem_base + TA_TM_D + (unsigned)(d * TW_NQ);
      for (int half = 0; half < 2; ++half) {
        mbar_wait(a_ready + half, (unsigned)it & 1u);
        tc_fence_after();
        __syncwarp();
        const int ks0 = half ? kh0 : 0, ks1 = half ? ksteps : kh0;
#pragma unroll 1
        for (int ks = ks0; ks < ks1; ++ks) {
          const unsigned long long fo = (unsigned long long)(ks * (256 >> 4));
          const unsigned ah = tmem_base + (unsigned)(8 * ks), al = ah + TA_TM_LO;
          if (!single) {                                               // tf32 mode: the hi * hi pass alone
            tc_mma_tf32_ts_elect(td, al, d_fh + fo, idesc, ks != 0);   // lo * hi
            tc_mma_tf32_ts_elect(td, ah, d_fl + fo, idesc, 1u);        // hi * lo
          }
          tc_mma_tf32_ts_elect(td, ah, d_fh + fo, idesc, single ? (unsigned)(ks != 0) : 1u);   // hi * hi
        }
        tc_commit_elect(a_free + half);
      }
      tc_commit_elect(d_full + d);
    }
  } else {
    // ---- epilogue: T1 rows to global memory ----------------------------------------------------------
    const int quad = warp & 3;
    for (int it = 0; it < ntl; ++it) {
      const int d = it & 1;
      mbar_wait(d_full + d, ((unsigned)it >> 1) & 1u);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((unsigned)(quad * 32) << 16) + TA_TM_D + (unsigned)(d * TW_NQ), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_free + d);
      const long row = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M + quad * 32 + lane;
      if (row < R) {
        float4* __restrict__ o = reinterpret_cast<float4*>(T1 + row * TQ);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (4 * q < TQ) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TA_MMA_WARP) tmem_dealloc(tmem_base, TA_TM_COLS);
}


}  // namespace

size_t fwd2d_tca_smem_bytes(int W, int nch) { return 2ul * TW_M * W * 4 + 2ul * 4 * nch * 128 + 12 * 8 + 16; }

// plain K1 through TMEM; returns FNO_E_ARG-free "not eligible" as 1 so that the caller can fall back
int launch_fwd2d_tca(const Plan* p, const float* x, float* T1, long planes, cudaStream_t st, bool attr_only) {
  if (attr_only) {
    if (cudaFuncSetAttribute(fwd2d_tca_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(fwd2d_tca)");
    return launch_fwd2d_tcap(p, nullptr, nullptr, nullptr, nullptr, 0, nullptr, true);
  }
  const long R = planes * p->H;
  const long tiles = (R + TW_M - 1) / TW_M;
  const int KP = p->tc_nch * 4;
  const size_t smem = fwd2d_tca_smem_bytes(p->W, p->tc_nch);
  // eligibility: K padding fits the TMEM budget, tiles are 16-byte multiples (bulk copy), ring fits shared memory
  if (KP > TA_MAXK || smem > 227 * 1024 || (R * p->W) % 4 != 0 || (reinterpret_cast<size_t>(x) & 15) != 0 ||
      tiles > 0x7fffffffL)
    return 1;
  const int ctas = (int)(tiles < 148 ? tiles : 148);
  const int TQ = (2 * p->m2 + 3) & ~3;
  fwd2d_tca_kernel<<<ctas, TA_THREADS, smem, st>>>(x, T1, (g_math_mode.load() == FNO_MATH_BF16 ? p->tcF_bf : p->tcF_hi), p->tcF_lo, R, p->W, p->tc_nch, TQ, (int)tiles,
                                                   g_math_mode.load());
  count_launch();
  return check_launch("fwd2d_tca_kernel");
}


int launch_fwd2d_tcap(const Plan* p, const float* g, const float* preact, float* ds_out, float* T1, long planes,
                      cudaStream_t st, bool attr_only) {
  if (attr_only) {
    if (cudaFuncSetAttribute(fwd2d_tcap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(fwd2d_tcap)");
    return FNO_OK;
  }
  const long R = planes * p->H;
  const long tiles = (R + TW_M - 1) / TW_M;
  const int KP = p->tc_nch * 4;
  const size_t smem = fwd2d_tca_smem_bytes(p->W, p->tc_nch);       // same footprint: 4 half-tile slots
  if (KP > TA_MAXK || smem > 227 * 1024 || (R * p->W) % 4 != 0 || ((R % (TW_M / 2)) * p->W) % 4 != 0 ||
      ((reinterpret_cast<size_t>(g) | reinterpret_cast<size_t>(preact) | reinterpret_cast<size_t>(ds_out)) & 15) != 0 ||
      tiles > 0x7fffffffL)
    return 1;
  const int ctas = (int)(tiles < 148 ? tiles : 148);
  const int TQ = (2 * p->m2 + 3) & ~3;
  fwd2d_tcap_kernel<<<ctas, TA_THREADS, smem, st>>>(g, preact, ds_out, T1, (g_math_mode.load() == FNO_MATH_BF16 ? p->tcF_bf : p->tcF_hi), p->tcF_lo, R, p->W, p->tc_nch, TQ,
                                                    (int)tiles, g_math_mode.load());
  count_launch();
  return check_launch("fwd2d_tcap_kernel");
}

}  // namespace fno
