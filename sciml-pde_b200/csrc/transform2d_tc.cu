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
// Staging: A[m = row][k = w] K-major, 16-byte chunks of 4 consecutive w at a 144-byte stride (bank
// skew: consecutive lanes store consecutive chunks of a row conflict-free), 8-row groups at
// nch * 144 bytes.  B = F[q][w] hi / lo lives in the plan, already in the K-major layout.
#include "tc_common.cuh"

namespace fno {
namespace {

constexpr int TW_M = 128;                  // rows per tile = TMEM lanes
constexpr int TW_NQ = 32;                  // accumulator columns (2 * m2 <= 32)
constexpr int TW_LBO = 144;                // byte stride between the 16-byte w chunks of a row group
constexpr int TW_LOADERS = 448;            // 14 loader warps (+ 1 MMA-issue warp): 128 registers per thread
constexpr int TW_THREADS = TW_LOADERS + 32;
constexpr int TW_MAXI = 10;                // rows per loader thread: ceil(128 / (448 / nch)), nch <= 34

template <bool PREMUL>
__global__ void __launch_bounds__(TW_THREADS, 1)
fwd2d_tc_kernel(const float* __restrict__ x, const float* __restrict__ preact, float* __restrict__ ds_out,
                float* __restrict__ T1, const float* __restrict__ Fhi, const float* __restrict__ Flo, long R, int W,
                int nch, int TQ, int total_tiles) {
  extern __shared__ __align__(128) unsigned char wsm[];
  const int a_bytes = 16 * nch * TW_LBO;
  const int f_bytes = 4 * nch * 128;
  unsigned char* a_hi = wsm;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* f_hi = a_lo + a_bytes;
  unsigned char* f_lo = f_hi + f_bytes;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(f_lo + f_bytes);
  unsigned long long* a_ready = bars;
  unsigned long long* d_full = bars + 1;
  unsigned long long* d_free = bars + 2;
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(a_ready, TW_LOADERS / 32);
    mbar_init(d_full, 1);
    mbar_init(d_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TW_LOADERS / 32) tmem_alloc(tmem_slot, 128);   // 3 independent accumulator chains (one per split pass)
  for (int i = tid; i < f_bytes / 16; i += TW_THREADS) {
    reinterpret_cast<float4*>(f_hi)[i] = __ldg(reinterpret_cast<const float4*>(Fhi) + i);
    reinterpret_cast<float4*>(f_lo)[i] = __ldg(reinterpret_cast<const float4*>(Flo) + i);
  }
  // zero A once: chunks / lanes past W are never written again
  for (int i = tid; i < 2 * a_bytes / 16; i += TW_THREADS) reinterpret_cast<float4*>(a_hi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  const int ksteps = nch / 2;
  const int ntl = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == TW_LOADERS / 32) {
    // ---- MMA issuer -------------------------------------------------------------------------------
    constexpr unsigned idesc = umma_idesc_tf32(TW_M, TW_NQ, 0, 0);
    for (int it = 0; it < ntl; ++it) {
      mbar_wait(a_ready, (unsigned)it & 1u);
      mbar_wait(d_free, ((unsigned)it & 1u) ^ 1u);
      tc_fence_after();
      if (lane == 0) {
        // the three split passes (lo*hi, hi*lo, hi*hi) accumulate into three separate TMEM tiles and are
        // issued interleaved: a chain of tiny N = 32 MMAs on ONE accumulator is bound by the MMA
        // pipeline latency, three independent chains overlap it; the epilogue adds the tiles
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          const unsigned long long ah = umma_desc(a_hi + ks * 2 * TW_LBO, TW_LBO, nch * TW_LBO);
          const unsigned long long al = umma_desc(a_lo + ks * 2 * TW_LBO, TW_LBO, nch * TW_LBO);
          const unsigned long long bh = umma_desc(f_hi + ks * 256, 128, nch * 128);
          const unsigned long long bl = umma_desc(f_lo + ks * 256, 128, nch * 128);
          tc_mma_tf32(tmem_base, al, bh, idesc, ks > 0);
          tc_mma_tf32(tmem_base + TW_NQ, ah, bl, idesc, ks > 0);
          tc_mma_tf32(tmem_base + 2 * TW_NQ, ah, bh, idesc, ks > 0);
        }
        tc_commit(d_full);
      }
      __syncwarp();
    }
  } else {
    // ---- loaders (all 16 warps) / epilogue (warps 0-3) ---------------------------------------------
    // loader thread = (chunk cl of a row, row residue rl): it stages chunk cl of rows rl, rl + RP, ... of
    // every tile; consecutive lanes hold consecutive 16-byte chunks of a row -> coalesced 512-byte runs
    // from HBM and conflict-free 144-byte-strided stores into A
    const int RP = TW_LOADERS / nch;               // rows covered per pass
    const int cl = tid % nch, rl = tid / nch;
    const bool loader = rl < RP;
    const int left = W - 4 * cl;
    const int nv = left >= 4 ? 4 : (left > 0 ? left : 0);      // valid floats of the chunk (W is even)
    float2 raw[TW_MAXI][2], sraw[TW_MAXI][2];
    auto load_raw = [&](int it) {
      const long row0 = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M;
      const long rows_left = (it < ntl) ? R - row0 : 0;
#pragma unroll
      for (int u = 0; u < TW_MAXI; ++u) {
        const int r = rl + u * RP;
        const bool in = loader && r < TW_M && r < rows_left;
        const size_t off = (size_t)(row0 + (in ? r : 0)) * W + 4 * cl;
        raw[u][0] = (in && nv >= 2) ? __ldg(reinterpret_cast<const float2*>(x + off)) : make_float2(0.f, 0.f);
        raw[u][1] = (in && nv >= 4) ? __ldg(reinterpret_cast<const float2*>(x + off) + 1) : make_float2(0.f, 0.f);
        if (PREMUL) {
          sraw[u][0] = (in && nv >= 2) ? __ldg(reinterpret_cast<const float2*>(preact + off)) : make_float2(0.f, 0.f);
          sraw[u][1] = (in && nv >= 4) ? __ldg(reinterpret_cast<const float2*>(preact + off) + 1) : make_float2(0.f, 0.f);
        }
      }
    };
    auto store_raw = [&](int it) {
      const long row0 = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M;
      const long rows_left = R - row0;
#pragma unroll
      for (int u = 0; u < TW_MAXI; ++u) {
        const int r = rl + u * RP;
        if (!loader || r >= TW_M) continue;
        float v[4] = {raw[u][0].x, raw[u][0].y, raw[u][1].x, raw[u][1].y};
        if (PREMUL) {
          v[0] *= gelu_fast_grad(sraw[u][0].x); v[1] *= gelu_fast_grad(sraw[u][0].y);
          v[2] *= gelu_fast_grad(sraw[u][1].x); v[3] *= gelu_fast_grad(sraw[u][1].y);
          if (ds_out != nullptr && r < rows_left) {
            float* __restrict__ d = ds_out + (size_t)(row0 + r) * W + 4 * cl;
            if (nv >= 2) *reinterpret_cast<float2*>(d) = make_float2(v[0], v[1]);
            if (nv >= 4) *(reinterpret_cast<float2*>(d) + 1) = make_float2(v[2], v[3]);
          }
        }
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(v[e], hi[e], lo[e]);
        const int so = (r & 7) * 16 + (r >> 3) * nch * TW_LBO + cl * TW_LBO;
        *reinterpret_cast<float4*>(a_hi + so) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(a_lo + so) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    };
    load_raw(0);
    for (int it = 0; it < ntl; ++it) {
      store_raw(it);                       // A is free: the MMAs of tile it-1 completed (d_full waited below)
      load_raw(it + 1);                    // next tile's loads fly during the MMAs and the epilogue
      mbar_wait(d_full, (unsigned)it & 1u);
      if (warp < 4) {
        tc_fence_after();
        float v[32], v2[32];
        tmem_ld32(tmem_base + ((unsigned)(warp * 32) << 16), v);                 // lo*hi
        tmem_ld32(tmem_base + ((unsigned)(warp * 32) << 16) + TW_NQ, v2);        // hi*lo
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[i];
        tmem_ld32(tmem_base + ((unsigned)(warp * 32) << 16) + 2 * TW_NQ, v2);    // hi*hi
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[i];
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_free);
        const long row = ((long)blockIdx.x + (long)it * gridDim.x) * TW_M + warp * 32 + lane;
        if (row < R) {
          float4* __restrict__ o = reinterpret_cast<float4*>(T1 + row * TQ);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (4 * q < TQ) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TW_LOADERS / 32) tmem_dealloc(tmem_base, 128);
}

}  // namespace

size_t fwd2d_tc_smem_bytes(int nch) { return 2ul * 16 * nch * TW_LBO + 2ul * 4 * nch * 128 + 64; }

int launch_fwd2d_tc(const Plan* p, const float* x, const float* preact, float* ds_out, float* T1, long planes,
                    cudaStream_t st, bool attr_only) {
  const size_t smem = fwd2d_tc_smem_bytes(p->tc_nch);
  if (attr_only) {
    // the limit is per function, not per plan: always opt in to the device maximum
    if (cudaFuncSetAttribute(fwd2d_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(fwd2d_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(fwd2d_tc)");
    return FNO_OK;
  }
  const long R = planes * p->H;
  const long tiles = (R + TW_M - 1) / TW_M;
  if (tiles > 0x7fffffffL || R * p->W > 0x7fffffffffL) { set_error("fwd2d_tc: tensor too large"); return FNO_E_ARG; }
  const int ctas = (int)(tiles < 148 ? tiles : 148);
  const int TQ = (2 * p->m2 + 3) & ~3;   // row pitch of T1 in floats (16-byte aligned rows)
  if (preact != nullptr)
    fwd2d_tc_kernel<true><<<ctas, TW_THREADS, smem, st>>>(x, preact, ds_out, T1, p->tcF_hi, p->tcF_lo, R, p->W, p->tc_nch, TQ, (int)tiles);
  else
    fwd2d_tc_kernel<false><<<ctas, TW_THREADS, smem, st>>>(x, nullptr, nullptr, T1, p->tcF_hi, p->tcF_lo, R, p->W, p->tc_nch, TQ, (int)tiles);
  count_launch();
  return check_launch("fwd2d_tc_kernel");
}

}  // namespace fno
