// The 1x1-conv bypass of a wide Fourier layer (nn.Conv2d(width, width, 1), fno/fno.py:131-134,162; BASELINE configs[2]:
// width 64) and its data gradient on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate through the
// 3xTF32 split of head_tc.cu.
//
// At width 64 the product is 64 MACs per loaded float: the FP32 kernel (pointwise2_kernel) is bound by instruction
// issue at 0.24 of the HBM roof (cfg 3).  Here a 128-pixel tile is one GEMM
//     D[128 pixels x 64 out channels] = A[128 x K] B[K x 64],   K = in channels <= 64
// with the activation tile staged K-major in shared memory (thread = pixel: coalesced 128-byte channel-row loads, split
// hi / lo in registers, four channels per 16-byte chunk -- the layout head_tc.cu verified on the hardware, SBO widened to
// 16 chunks), W (or W^T for the data gradient) hi / lo staged once per CTA, the accumulator in TMEM (double-buffered:
// the MMAs of tile i+1 run under the stores of tile i) and an epilogue that only adds the bias and writes coalesced
// channel rows.  What is left is a streaming kernel: 64 KB in + out per tile against 24 MMAs.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace fno {
namespace {

constexpr int PT_M = 128;                       // pixels per tile = TMEM lanes
constexpr int PT_N = 64;                        // out channels (padded) = accumulator columns
constexpr int PT_K = 64;                        // in channels (padded)
constexpr int PT_SBO = (PT_K / 4) * 128;        // 8-row group pitch: 16 chunks of 16 bytes per row, LBO = 128
constexpr int PT_A_BYTES = (PT_M / 8) * PT_SBO; // 32 KB
constexpr int PT_B_BYTES = (PT_N / 8) * PT_SBO; // 16 KB
constexpr int PT_EPI_WARPS = 16;
constexpr int PT_EPI_THREADS = 32 * PT_EPI_WARPS;
constexpr int PT_THREADS = PT_EPI_THREADS + 32;
constexpr int PT_MAXCH = 4;                     // 16-byte channel chunks staged per thread
constexpr int PT_SMEM = 4 * PT_A_BYTES + 2 * PT_B_BYTES + PT_N * 4 + 6 * 8 + 16;

__device__ __forceinline__ int pt_off(int n, int k) { return (n & 7) * 16 + (n >> 3) * PT_SBO + (k >> 2) * 128 + (k & 3) * 4; }

// out[b, n, p] = sum_k Wm[n, k] in[b, k, p] (+ bias[n]);  Wm[n, k] = W[n * ldw + k]  (transpose == 0: the convolution)
//                                                        Wm[n, k] = W[k * ldw + n]  (transpose != 0: its data gradient)
// Persistent CTA, one per SM; 16 loader / epilogue warps (4 TMEM lane quadrants x 4 column quarters) + 1 MMA-issue warp.
__global__ void __launch_bounds__(PT_THREADS, 1)
pointwise_tc_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ out, long N, int Cin, int Cout, int ldw, int transpose, int tiles_per_sample,
                    int total_tiles, int single) {
  FNO_SPLIT_CONSTS(single);                     // single: 0 = 3xTF32 (fp32 mode), 1 = tf32, 2 = bf16 operands
  extern __shared__ __align__(128) unsigned char psm[];
  unsigned char* a_hi = psm;                    // [2 stages][PT_A_BYTES]
  unsigned char* a_lo = a_hi + 2 * PT_A_BYTES;
  unsigned char* w_hi = a_lo + 2 * PT_A_BYTES;
  unsigned char* w_lo = w_hi + PT_B_BYTES;
  float* bs = reinterpret_cast<float*>(w_lo + PT_B_BYTES);            // [PT_N]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(bs + PT_N);
  unsigned long long* a_ready = bars;           // [2]
  unsigned long long* d_full = bars + 2;        // [2]
  unsigned long long* d_free = bars + 4;        // [2]
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 6);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_ready + s, PT_EPI_WARPS);
      mbar_init(d_full + s, 1);
      mbar_init(d_free + s, PT_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PT_EPI_WARPS) tmem_alloc(tmem_slot, 2 * PT_N);
  for (int i = tid; i < PT_N * PT_K; i += PT_THREADS) {
    const int n = i / PT_K, k = i - n * PT_K;
    float hi = 0.f, lo = 0.f;
    if (n < Cout && k < Cin) split_rm(__ldg(W + (transpose ? (size_t)k * ldw + n : (size_t)n * ldw + k)), hi, lo, sp_rnd, sp_msk);
    *reinterpret_cast<float*>(w_hi + pt_off(n, k)) = hi;
    *reinterpret_cast<float*>(w_lo + pt_off(n, k)) = lo;
  }
  for (int i = tid; i < PT_N; i += PT_THREADS) bs[i] = (bias != nullptr && i < Cout) ? __ldg(bias + i) : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  const int ksteps = (Cin + 7) / 8;
  const int ntl = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp == PT_EPI_WARPS) {
    // ---- MMA issuer: the whole warp stays converged, one elected lane issues ---------------------------
    constexpr unsigned idesc = umma_idesc_tf32(PT_M, PT_N, /*A K-major*/ 0, /*B K-major*/ 0);
    const unsigned long long d_a_h = umma_desc(a_hi, 128, PT_SBO), d_a_l = umma_desc(a_lo, 128, PT_SBO);
    const unsigned long long d_w_h = umma_desc(w_hi, 128, PT_SBO), d_w_l = umma_desc(w_lo, 128, PT_SBO);
    for (int it = 0; it < ntl; ++it) {
      const int st = it & 1;
      const unsigned ph = (unsigned)(it >> 1) & 1u;
      mbar_wait(a_ready + st, ph);
      mbar_wait(d_free + st, ph ^ 1u);
      tc_fence_after();
      __syncwarp();
      const unsigned long long so = (unsigned long long)(st * (PT_A_BYTES >> 4));
      const unsigned d = tmem_base + (unsigned)(st * PT_N);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass)              // lo*hi, hi*lo, hi*hi
#pragma unroll
        for (int ks = 0; ks < PT_K / 8; ++ks)           // K = 8 per instruction = two 16-byte chunks: +256 B
          if (ks < ksteps && (pass == 2 || !single))
            tc_mma_tf32_elect(d, (pass == 0 ? d_a_l : d_a_h) + so + (unsigned long long)(ks * (256 >> 4)),
                              (pass == 1 ? d_w_l : d_w_h) + (unsigned long long)(ks * (256 >> 4)), idesc,
                              single ? (unsigned)(ks != 0) : (unsigned)((pass | ks) != 0));
      tc_commit_elect(d_full + st);
    }
  } else {
    // ---- loader / epilogue warps ------------------------------------------------------------------
    const int quad = warp & 3;                     // TMEM lane quadrant of this warp
    const int colq = warp >> 2;                    // out channels [16 * colq, 16 * colq + 16)
    const int m = quad * 32 + lane;                // pixel of the tile owned in the epilogue
    const int pm = tid & (PT_M - 1);               // pixel of the tile staged by this thread
    const int kq = tid >> 7;                       // chunk residue staged by this thread
    const int abase = (pm & 7) * 16 + (pm >> 3) * PT_SBO;
    float raw[PT_MAXCH][4];
    auto load_raw = [&](int it) {                  // this thread's channels of tile `it` (zeros past the end)
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const bool inr = it < ntl;
      const int b = inr ? tile / tiles_per_sample : 0;
      const long p = inr ? (long)(tile - b * tiles_per_sample) * PT_M + pm : 0;
      const bool valid = inr && p < N;
      const float* __restrict__ ip = in + (size_t)b * Cin * N + (valid ? p : 0);
#pragma unroll
      for (int u = 0; u < PT_MAXCH; ++u) {
        const int kc = kq + 4 * u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * kc + e;
          raw[u][e] = (valid && c < Cin) ? __ldg(ip + (size_t)c * N) : 0.f;
        }
      }
    };
    auto store_raw = [&](int st) {
#pragma unroll
      for (int u = 0; u < PT_MAXCH; ++u) {
        const int kc = kq + 4 * u;
        if (kc < 2 * ksteps) {
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_rm(raw[u][e], hi[e], lo[e], sp_rnd, sp_msk);
          *reinterpret_cast<float4*>(a_hi + st * PT_A_BYTES + abase + kc * 128) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          if (!single) *reinterpret_cast<float4*>(a_lo + st * PT_A_BYTES + abase + kc * 128) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready + st);
    };
    // prologue: tile 0 staged, tile 1 in flight
    load_raw(0);
    store_raw(0);
    load_raw(1);
    for (int it = 0; it < ntl; ++it) {
      const int st = it & 1;
      const unsigned ph = (unsigned)(it >> 1) & 1u;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = tile / tiles_per_sample;
      const long p = (long)(tile - b * tiles_per_sample) * PT_M + m;
      // stage tile it+1 (its operand buffer was last read by the MMAs of tile it-1, whose completion this thread
      // observed in the previous epilogue), then put tile it+2's loads in flight
      if (it + 1 < ntl) store_raw(st ^ 1);
      load_raw(it + 2);

      mbar_wait(d_full + st, ph);
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem_base + ((unsigned)(quad * 32) << 16) + (unsigned)(st * PT_N + colq * 16), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_free + st);     // the accumulator may be overwritten by tile it+2's MMAs
      if (p < N) {
        float* __restrict__ op = out + ((size_t)b * Cout + colq * 16) * N + p;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (colq * 16 + i < Cout) op[(size_t)i * N] = v[i] + bs[colq * 16 + i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PT_EPI_WARPS) tmem_dealloc(tmem_base, 2 * PT_N);
}

}  // namespace

bool pointwise_tc_supported(int Cin, int Cout) {
  static const bool off = [] { const char* e = std::getenv("FNO_PW_TC"); return e != nullptr && e[0] == '0'; }();
  const int wmax = Cin > Cout ? Cin : Cout;
  return !off && wmax > 32 && wmax <= PT_K;
}

int launch_pointwise_tc(const float* in, const float* W, const float* bias, float* out, int B, int Co, int Ci, long N,
                        int transpose, cudaStream_t st) {
  const int Cout = transpose ? Ci : Co, Cin = transpose ? Co : Ci;
  const long tps = (N + PT_M - 1) / PT_M;
  const long total = tps * B;
  if (total > 0x7fffffffL) { set_error("fno_pointwise_fwd: too many tiles"); return FNO_E_ARG; }
  static PerDeviceOnce done;
  if (done.need()) {
    if (cudaFuncSetAttribute(pointwise_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(pointwise_tc)");
    done.mark();
  }
  const int ctas = (int)(total < 148 ? total : 148);
  pointwise_tc_kernel<<<ctas, PT_THREADS, PT_SMEM, st>>>(in, W, bias, out, N, Cin, Cout, Ci, transpose, (int)tps, (int)total,
                                                         g_math_mode.load());
  count_launch();
  return check_launch("pointwise_tc_kernel");
}

}  // namespace fno
