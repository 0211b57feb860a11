// K2: per-mode complex channel mixing over the retained spectrum (compl_mul2d / compl_mul3d,
// fno/fno.py:66-68, :255-257) and its autograd backward.  FP32 CUDA-core path: at width 20 this
// is a bandwidth/latency-bound streaming kernel (0.9 MFLOP per sample-layer), so lanes map to
// consecutive modes and every global access is a coalesced 8-byte complex element.  The corner
// parameters are read in the reference's own state_dict layout [Ci, Co, m1, m2(, m3)].
//
// Width 33..64 (BASELINE configs[2]: width 64, 25 flop per byte at batch 32) runs on the tensor cores instead:
// per mode the einsum is a [2Co x Ci] x [Ci x 2B] real GEMM (mix_tc_kernel) and its weight gradient a
// [2Ci x B] x [B x 2Co] one (mix_wgrad_tc_kernel) -- tcgen05.mma.kind::tf32 with a 3xTF32 split in fp32 mode,
// accumulators in TMEM; see the second half of this file.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace fno {
namespace {

struct ModeGeo {
  int nd;      // 2 or 3
  int m1;      // modes of the first (split) axis
  int m2;      // 3-D: modes of the second split axis; 2-D: unused
  int inner;   // modes of the half-spectrum axis
  int M;       // retained modes per plane
  int Mc;      // modes per corner tensor
};

__device__ __forceinline__ void mode_to_corner(const ModeGeo& g, int m, int& corner, int& local) {
  if (g.nd == 2) {
    const int r = m / g.inner;
    const int k = m - r * g.inner;
    corner = r / g.m1;
    local = (r - corner * g.m1) * g.inner + k;
  } else {
    const int k = m % g.inner;
    const int rr = m / g.inner;
    const int r2 = rr % (2 * g.m2);
    const int r1 = rr / (2 * g.m2);
    const int c1 = r1 / g.m1, c2 = r2 / g.m2;
    corner = c1 + 2 * c2;
    local = ((r1 - c1 * g.m1) * g.m2 + (r2 - c2 * g.m2)) * g.inner + k;
  }
}

struct WPtrs {
  const float2* w[4];
};
struct GWPtrs {
  float2* w[4];
};

constexpr int TB = 4;  // batch entries per thread
constexpr int TO = 4;  // output channels per thread

// out[b, oc, m] = sum_s in[b, s, m] * Wc(m)[...]      (CONJ_T = false: forward, s = i, oc = o)
// out[b, oc, m] = sum_s in[b, s, m] * conj(Wc(m)[oc, s])  (CONJ_T = true: data gradient, s = o, oc = i)
template <bool CONJ_T>
__global__ void __launch_bounds__(128)
mix_kernel(const float2* __restrict__ in, float2* __restrict__ out, WPtrs wp, ModeGeo geo, int B, int NS,
           int NO, int Co) {
  const int m = blockIdx.x * 32 + threadIdx.x;
  if (m >= geo.M) return;
  const int o0 = blockIdx.y * TO;
  const int b0 = (blockIdx.z * blockDim.y + threadIdx.y) * TB;
  if (b0 >= B) return;
  int corner, local;
  mode_to_corner(geo, m, corner, local);
  const float2* __restrict__ W = wp.w[corner];
  float2 acc[TB][TO];
#pragma unroll
  for (int bb = 0; bb < TB; ++bb)
#pragma unroll
    for (int oo = 0; oo < TO; ++oo) acc[bb][oo] = make_float2(0.f, 0.f);

  // 4 channels per trip: the 32 independent loads of a trip are in flight together (the kernel is a
  // chain of L2 round trips otherwise -- 25 us for 0.9 MFLOP per sample at cfg 1)
#pragma unroll 4
  for (int s = 0; s < NS; ++s) {
    float2 xv[TB], wv[TO];
#pragma unroll
    for (int bb = 0; bb < TB; ++bb) {
      const int b = b0 + bb;
      xv[bb] = (b < B) ? __ldg(in + ((size_t)b * NS + s) * geo.M + m) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int oo = 0; oo < TO; ++oo) {
      const int oc = o0 + oo;
      float2 v = make_float2(0.f, 0.f);
      if (oc < NO) {
        const size_t widx = CONJ_T ? ((size_t)oc * Co + s) : ((size_t)s * Co + oc);
        v = __ldg(W + widx * geo.Mc + local);
        if (CONJ_T) v.y = -v.y;
      }
      wv[oo] = v;
    }
#pragma unroll
    for (int bb = 0; bb < TB; ++bb)
#pragma unroll
      for (int oo = 0; oo < TO; ++oo) {
        acc[bb][oo].x = fmaf(xv[bb].x, wv[oo].x, acc[bb][oo].x);
        acc[bb][oo].x = fmaf(-xv[bb].y, wv[oo].y, acc[bb][oo].x);
        acc[bb][oo].y = fmaf(xv[bb].x, wv[oo].y, acc[bb][oo].y);
        acc[bb][oo].y = fmaf(xv[bb].y, wv[oo].x, acc[bb][oo].y);
      }
  }
#pragma unroll
  for (int bb = 0; bb < TB; ++bb) {
    const int b = b0 + bb;
    if (b >= B) break;
#pragma unroll
    for (int oo = 0; oo < TO; ++oo) {
      const int oc = o0 + oo;
      if (oc < NO) out[((size_t)b * NO + oc) * geo.M + m] = acc[bb][oo];
    }
  }
}

// gW[i, o, m] = sum_b conj(X[b, i, m]) * gY[b, o, m]; thread <-> (m, 4 i, 4 o); the batch is split
// over blockDim.y slices and reduced through shared memory in a fixed order (deterministic).
constexpr int TI = 4;
constexpr int WG_SLICES = 8;

__global__ void __launch_bounds__(32 * WG_SLICES)
mix_wgrad_kernel(const float2* __restrict__ X, const float2* __restrict__ gY, GWPtrs gw, ModeGeo geo, int B,
                 int Ci, int Co) {
  __shared__ float2 red[WG_SLICES][TI * TO][32];
  const int lane = threadIdx.x;
  const int sl = threadIdx.y;
  const int m = blockIdx.x * 32 + lane;
  const int i0 = blockIdx.y * TI;
  const int o0 = blockIdx.z * TO;
  const bool valid = m < geo.M;
  float2 acc[TI][TO];
#pragma unroll
  for (int ii = 0; ii < TI; ++ii)
#pragma unroll
    for (int oo = 0; oo < TO; ++oo) acc[ii][oo] = make_float2(0.f, 0.f);
  if (valid) {
#pragma unroll 4
    for (int b = sl; b < B; b += WG_SLICES) {
      float2 xv[TI], gv[TO];
#pragma unroll
      for (int ii = 0; ii < TI; ++ii)
        xv[ii] = (i0 + ii < Ci) ? __ldg(X + ((size_t)b * Ci + i0 + ii) * geo.M + m) : make_float2(0.f, 0.f);
#pragma unroll
      for (int oo = 0; oo < TO; ++oo)
        gv[oo] = (o0 + oo < Co) ? __ldg(gY + ((size_t)b * Co + o0 + oo) * geo.M + m) : make_float2(0.f, 0.f);
#pragma unroll
      for (int ii = 0; ii < TI; ++ii)
#pragma unroll
        for (int oo = 0; oo < TO; ++oo) {
          // conj(x) * g = (xr gr + xi gi) + i (xr gi - xi gr)
          acc[ii][oo].x = fmaf(xv[ii].x, gv[oo].x, acc[ii][oo].x);
          acc[ii][oo].x = fmaf(xv[ii].y, gv[oo].y, acc[ii][oo].x);
          acc[ii][oo].y = fmaf(xv[ii].x, gv[oo].y, acc[ii][oo].y);
          acc[ii][oo].y = fmaf(-xv[ii].y, gv[oo].x, acc[ii][oo].y);
        }
    }
  }
#pragma unroll
  for (int ii = 0; ii < TI; ++ii)
#pragma unroll
    for (int oo = 0; oo < TO; ++oo) red[sl][ii * TO + oo][lane] = acc[ii][oo];
  __syncthreads();
  if (!valid) return;
  int corner, local;
  mode_to_corner(geo, m, corner, local);
  float2* __restrict__ G = gw.w[corner];
  constexpr int PER = (TI * TO) / WG_SLICES;  // entries finalised by each slice
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int idx = sl * PER + e;
    float2 s = red[0][idx][lane];
#pragma unroll
    for (int q = 1; q < WG_SLICES; ++q) {
      const float2 v = red[q][idx][lane];
      s.x += v.x;
      s.y += v.y;
    }
    const int i = i0 + idx / TO, o = o0 + idx % TO;
    if (i < Ci && o < Co) G[((size_t)i * Co + o) * geo.Mc + local] = s;
  }
}

// ================================================================================================
// K2 on the 5th-generation tensor cores (tcgen05 + TMEM), width 33..64.
//
// For ONE retained mode the complex contraction  Y[b, o] = sum_i X[b, i] W[i, o]  is a real GEMM whose M rows are
// the output channels twice (rows 0..63: real part of W, rows 64..127: imaginary part), whose N columns are the batch
// twice (columns 0..31: Re X, 32..63: Im X) and whose K is the input channel:
//     D[(o, c), (b, c')] = sum_i W_c[i, o] X_c'[b, i]        M = 128, N = 64, K = Ci <= 64
//     Re Y[b, o] = D[(o,re),(b,re)] - D[(o,im),(b,im)]       Im Y[b, o] = D[(o,re),(b,im)] + D[(o,im),(b,re)]
// so neither operand is duplicated in shared memory (the block-real form [[Wr, -Wi], [Wi, Wr]] would stage W twice) and
// the four partial products are combined in the epilogue: TMEM lane o meets lane 64 + o through shared memory.  The data
// gradient  gX[b, i] = sum_o gY[b, o] conj(W[i, o])  is the same kernel with the roles of i and o swapped and -Im W
// staged (CONJ_T).  The weight gradient  gW[i, o] = sum_b conj(X[b, i]) gY[b, o]  contracts the batch instead:
//     D[(i, c), (o, c')] = sum_b X_c[b, i] G_c'[b, o]        M = 128, N = 128, K = 32 per batch chunk (accumulated)
//     Re gW = D[(i,re),(o,re)] + D[(i,im),(o,im)]            Im gW = D[(i,re),(o,im)] - D[(i,im),(o,re)]
//
// Memory side: the spectra and the corner parameters keep the mode index fastest ([.., m] complex64), so one mode of
// one (row, channel) pair is a lone 8-byte element.  A CTA therefore owns TWO adjacent modes: every global access is a
// 16-byte vector (half a sector; the neighbouring CTA takes the other half from L2), staged straight into the two modes'
// operand buffers -- split hi / lo in registers (3xTF32: fp32-mode accuracy, tc_common.cuh) and stored in the no-swizzle
// K-major core-matrix layout with LBO = 144 B, which makes the 4-byte staging stores of a warp (lane = K index)
// conflict-free.  At cfg 3 (512 modes, B = 32, width 64) that is 256 CTAs of 96 KB loaded / 32 KB stored each and 48
// (forward) or 24 (weight gradient) MMAs; the kernel is bound by its scattered 16-byte accesses, not by the tensor pipe.
// ================================================================================================
constexpr int MT_THREADS = 256;
constexpr int MT_LBO = 144;                       // bytes between the 16-byte K chunks of a row
constexpr int MT_BCH = 32;                        // batch entries per CTA (forward) / per accumulation step (wgrad)
constexpr int MT_SBO = 16 * MT_LBO;               // 8-row group pitch for K = 64 channels (16 chunks)
constexpr int MT_A_BYTES = 16 * MT_SBO;           // 128 rows: (channel, re | im)
constexpr int MT_B_BYTES = 8 * MT_SBO;            // 64 rows: (batch entry, re | im)
constexpr int MT_MODE_BYTES = 2 * MT_A_BYTES + 2 * MT_B_BYTES;     // A hi, A lo, B hi, B lo of one mode
constexpr int MT_SMEM = 2 * MT_MODE_BYTES + 32;
constexpr int MW_SBO = 8 * MT_LBO;                // K = 32 batch entries (8 chunks)
constexpr int MW_OP_BYTES = 16 * MW_SBO;          // 128 rows
constexpr int MW_MODE_BYTES = 4 * MW_OP_BYTES;    // X hi, X lo, G hi, G lo of one mode
constexpr int MW_SMEM = 2 * MW_MODE_BYTES + 32;

__device__ __forceinline__ int mt_off(int n, int k, int sbo) {
  return (n & 7) * 16 + (n >> 3) * sbo + (k >> 2) * MT_LBO + (k & 3) * 4;
}
__device__ __forceinline__ void mt_put(unsigned char* hi_buf, unsigned char* lo_buf, int off, float x, unsigned rnd,
                                       unsigned msk, int single) {
  float hi, lo;
  split_rm(x, hi, lo, rnd, msk);
  *reinterpret_cast<float*>(hi_buf + off) = hi;
  if (!single) *reinterpret_cast<float*>(lo_buf + off) = lo;
}

// out[b, r, m] = sum_k in[b, k, m] * W(m)[k, r]         (CONJ_T = false: forward, k = i, r = o)
// out[b, r, m] = sum_k in[b, k, m] * conj(W(m)[r, k])   (CONJ_T = true: data gradient, k = o, r = i)
// grid = (M / 2 mode pairs, ceil(B / 32)); NK contraction channels, NR result channels, both <= 64
template <bool CONJ_T>
__global__ void __launch_bounds__(MT_THREADS, 1)
mix_tc_kernel(const float2* __restrict__ in, float2* __restrict__ out, WPtrs wp, ModeGeo geo, int B, int NK, int NR,
              int Co, int single) {
  FNO_SPLIT_CONSTS(single);
  extern __shared__ __align__(128) unsigned char msm[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(msm + 2 * MT_MODE_BYTES);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bar + 1);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int m0 = 2 * (int)blockIdx.x;
  const int b0 = (int)blockIdx.y * MT_BCH;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  int corner, local;
  mode_to_corner(geo, m0, corner, local);
  const float2* __restrict__ W = wp.w[corner] + local;

  // ---- stage the weights of both modes: thread <-> (contraction channel k = lane-consecutive, result channel r) ----
  {
    const int k = tid & 63;
#pragma unroll
    for (int it = 0; it < 16; it += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = (tid >> 6) + 4 * (it + u);
        const size_t widx = CONJ_T ? ((size_t)r * Co + k) : ((size_t)k * Co + r);
        v[u] = (k < NK && r < NR) ? __ldg(reinterpret_cast<const float4*>(W + widx * geo.Mc))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = (tid >> 6) + 4 * (it + u);
        const int o_re = mt_off(r, k, MT_SBO), o_im = mt_off(64 + r, k, MT_SBO);
        unsigned char* a0 = msm;
        unsigned char* a1 = msm + MT_MODE_BYTES;
        mt_put(a0, a0 + MT_A_BYTES, o_re, v[u].x, sp_rnd, sp_msk, single);
        mt_put(a0, a0 + MT_A_BYTES, o_im, CONJ_T ? -v[u].y : v[u].y, sp_rnd, sp_msk, single);
        mt_put(a1, a1 + MT_A_BYTES, o_re, v[u].z, sp_rnd, sp_msk, single);
        mt_put(a1, a1 + MT_A_BYTES, o_im, CONJ_T ? -v[u].w : v[u].w, sp_rnd, sp_msk, single);
      }
    }
  }
  // ---- stage the spectrum slice: row b = Re, row 32 + b = Im, K = channel --------------------------------------
  {
    const int k = tid & 63;
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int bb = (tid >> 6) + 4 * u;
      v[u] = (k < NK && b0 + bb < B) ? __ldg(reinterpret_cast<const float4*>(in + ((size_t)(b0 + bb) * NK + k) * geo.M + m0))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int bb = (tid >> 6) + 4 * u;
      const int o_re = mt_off(bb, k, MT_SBO), o_im = mt_off(32 + bb, k, MT_SBO);
      unsigned char* x0 = msm + 2 * MT_A_BYTES;
      unsigned char* x1 = x0 + MT_MODE_BYTES;
      mt_put(x0, x0 + MT_B_BYTES, o_re, v[u].x, sp_rnd, sp_msk, single);
      mt_put(x0, x0 + MT_B_BYTES, o_im, v[u].y, sp_rnd, sp_msk, single);
      mt_put(x1, x1 + MT_B_BYTES, o_re, v[u].z, sp_rnd, sp_msk, single);
      mt_put(x1, x1 + MT_B_BYTES, o_im, v[u].w, sp_rnd, sp_msk, single);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- MMA issue: the warp stays converged, one elected lane issues (tc_common.cuh) -----------------------
    constexpr unsigned idesc = umma_idesc_tf32(128, 64, 0, 0);
    const int ksteps = (NK + 7) >> 3;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      unsigned char* base = msm + j * MT_MODE_BYTES;
      const unsigned long long d_a_h = umma_desc(base, MT_LBO, MT_SBO), d_a_l = umma_desc(base + MT_A_BYTES, MT_LBO, MT_SBO);
      const unsigned long long d_b_h = umma_desc(base + 2 * MT_A_BYTES, MT_LBO, MT_SBO),
                               d_b_l = umma_desc(base + 2 * MT_A_BYTES + MT_B_BYTES, MT_LBO, MT_SBO);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass)                // lo*hi, hi*lo, hi*hi
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)                    // K = 8 per instruction = two 16-byte chunks
          if (ks < ksteps && (pass == 2 || !single))
            tc_mma_tf32_elect(tmem_base + (unsigned)(64 * j),
                              (pass == 0 ? d_a_l : d_a_h) + (unsigned long long)(ks * (2 * MT_LBO >> 4)),
                              (pass == 1 ? d_b_l : d_b_h) + (unsigned long long)(ks * (2 * MT_LBO >> 4)), idesc,
                              single ? (unsigned)(ks != 0) : (unsigned)((pass | ks) != 0));
    }
    tc_commit_elect(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();

  // ---- epilogue: warp <-> (lane quadrant q, batch half hb); quadrants 2, 3 (Im W rows) hand their partial products
  // to quadrants 0, 1 through shared memory (the operand buffers are free: every MMA has completed) ------------
  const int q = warp & 3, hb = warp >> 2;
  const int r = (q & 1) * 32 + lane;                       // result channel of this thread's TMEM lane
  float pv[2][16], qv[2][16];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const unsigned t = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(64 * j + 16 * hb);
    tmem_ld16(t, pv[j]);                                   // . * Re in
    tmem_ld16(t + 32u, qv[j]);                             // . * Im in
  }
  float* ex = reinterpret_cast<float*>(msm);               // [mode][P | Q][b 32][r 64]
  if (q >= 2) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        ex[((j * 2 + 0) * 32 + 16 * hb + e) * 64 + r] = pv[j][e];
        ex[((j * 2 + 1) * 32 + 16 * hb + e) * 64 + r] = qv[j][e];
      }
  }
  tc_fence_before();
  __syncthreads();
  if (q < 2 && r < NR) {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int b = b0 + 16 * hb + e;
      if (b < B) {
        float4 y;
        y.x = pv[0][e] - ex[((0 * 2 + 1) * 32 + 16 * hb + e) * 64 + r];
        y.y = qv[0][e] + ex[((0 * 2 + 0) * 32 + 16 * hb + e) * 64 + r];
        y.z = pv[1][e] - ex[((1 * 2 + 1) * 32 + 16 * hb + e) * 64 + r];
        y.w = qv[1][e] + ex[((1 * 2 + 0) * 32 + 16 * hb + e) * 64 + r];
        *reinterpret_cast<float4*>(out + ((size_t)b * NR + r) * geo.M + m0) = y;
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// gW[i, o, m] = sum_b conj(X[b, i, m]) gY[b, o, m]; grid = M / 2 mode pairs, the batch is walked in chunks of 32
// (K of the GEMM) that accumulate into the same TMEM tile, so the result is deterministic.
__global__ void __launch_bounds__(MT_THREADS, 1)
mix_wgrad_tc_kernel(const float2* __restrict__ X, const float2* __restrict__ gY, GWPtrs gw, ModeGeo geo, int B, int Ci,
                    int Co, int single) {
  FNO_SPLIT_CONSTS(single);
  extern __shared__ __align__(128) unsigned char msm[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(msm + 2 * MW_MODE_BYTES);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bar + 1);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int m0 = 2 * (int)blockIdx.x;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  constexpr unsigned idesc = umma_idesc_tf32(128, 128, 0, 0);

  const int nchunks = (B + MT_BCH - 1) / MT_BCH;
  for (int c = 0; c < nchunks; ++c) {
    // ---- stage chunk c: thread <-> (batch entry = lane = K index, channel); rows ch = Re, 64 + ch = Im -------
    const int b = c * MT_BCH + lane;
#pragma unroll
    for (int t = 0; t < 2; ++t) {                          // t = 0: X (A operand), t = 1: gY (B operand)
      const float2* __restrict__ src = t == 0 ? X : gY;
      const int NC = t == 0 ? Ci : Co;
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int ch = (tid >> 5) + 8 * u;
        v[u] = (b < B && ch < NC) ? __ldg(reinterpret_cast<const float4*>(src + ((size_t)b * NC + ch) * geo.M + m0))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int ch = (tid >> 5) + 8 * u;
        const int o_re = mt_off(ch, lane, MW_SBO), o_im = mt_off(64 + ch, lane, MW_SBO);
        unsigned char* p0 = msm + t * 2 * MW_OP_BYTES;
        unsigned char* p1 = p0 + MW_MODE_BYTES;
        mt_put(p0, p0 + MW_OP_BYTES, o_re, v[u].x, sp_rnd, sp_msk, single);
        mt_put(p0, p0 + MW_OP_BYTES, o_im, v[u].y, sp_rnd, sp_msk, single);
        mt_put(p1, p1 + MW_OP_BYTES, o_re, v[u].z, sp_rnd, sp_msk, single);
        mt_put(p1, p1 + MW_OP_BYTES, o_im, v[u].w, sp_rnd, sp_msk, single);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        unsigned char* base = msm + j * MW_MODE_BYTES;
        const unsigned long long d_a_h = umma_desc(base, MT_LBO, MW_SBO), d_a_l = umma_desc(base + MW_OP_BYTES, MT_LBO, MW_SBO);
        const unsigned long long d_b_h = umma_desc(base + 2 * MW_OP_BYTES, MT_LBO, MW_SBO),
                                 d_b_l = umma_desc(base + 3 * MW_OP_BYTES, MT_LBO, MW_SBO);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (pass == 2 || !single)
              tc_mma_tf32_elect(tmem_base + (unsigned)(128 * j),
                                (pass == 0 ? d_a_l : d_a_h) + (unsigned long long)(ks * (2 * MT_LBO >> 4)),
                                (pass == 1 ? d_b_l : d_b_h) + (unsigned long long)(ks * (2 * MT_LBO >> 4)), idesc,
                                (c != 0 || ks != 0 || (!single && pass != 0)) ? 1u : 0u);
      }
      tc_commit_elect(bar);
    }
    mbar_wait(bar, (unsigned)c & 1u);                      // operands consumed (and, after the last chunk, D complete)
    tc_fence_after();
  }

  // ---- epilogue: quadrants 2, 3 (Im X rows) pass D[64 + i, :] to quadrants 0, 1 through shared memory ----------
  const int q = warp & 3, hw = warp >> 2;                  // hw: output channels [32 hw, 32 hw + 32)
  const int i = (q & 1) * 32 + lane;
  float* ex = reinterpret_cast<float*>(msm);               // [mode][column 128][i 64]
  if (q >= 2) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int s = 0; s < 4; ++s) {                        // columns 32 hw + 16 (s & 1) + 64 (s >> 1) ...
        const int col = 32 * hw + 16 * (s & 1) + 64 * (s >> 1);
        float v[16];
        tmem_ld16(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(128 * j + col), v);
#pragma unroll
        for (int e = 0; e < 16; ++e) ex[(j * 128 + col + e) * 64 + i] = v[e];
      }
  }
  tc_fence_before();
  __syncthreads();
  if (q < 2) {
    int corner, local;
    mode_to_corner(geo, m0, corner, local);
    float2* __restrict__ G = gw.w[corner] + local;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
      const int o0 = 32 * hw + 16 * s;
      float re[2][16], im[2][16];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const unsigned t = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(128 * j + o0);
        tmem_ld16(t, re[j]);                               // Xr . Gr
        tmem_ld16(t + 64u, im[j]);                         // Xr . Gi
      }
      if (i < Ci) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int o = o0 + e;
          if (o < Co) {
            float4 g;
            g.x = re[0][e] + ex[(0 * 128 + 64 + o) * 64 + i];
            g.y = im[0][e] - ex[(0 * 128 + o) * 64 + i];
            g.z = re[1][e] + ex[(1 * 128 + 64 + o) * 64 + i];
            g.w = im[1][e] - ex[(1 * 128 + o) * 64 + i];
            *reinterpret_cast<float4*>(G + ((size_t)i * Co + o) * geo.Mc) = g;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

bool aligned16(const void* p) { return (reinterpret_cast<size_t>(p) & 15u) == 0; }

// envelope of the tensor-core form: wide layers (the FP32 kernel wins at width 20: 0.9 MFLOP per sample), an even
// innermost mode count (a pair of adjacent modes then shares its corner and row) and 16-byte aligned tensors
bool mix_tc_envelope(const ModeGeo& geo, int Ci, int Co) {
  static const bool off = [] { const char* e = std::getenv("FNO_MIX_TC"); return e != nullptr && e[0] == '0'; }();
  const int wmax = Ci > Co ? Ci : Co;
  return !off && wmax > 32 && wmax <= 64 && (geo.inner % 2) == 0;
}

int setup_mix_tc_attrs() {
  static PerDeviceOnce done;
  if (done.need()) {
    if (cudaFuncSetAttribute(mix_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(mix_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(mix_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MW_SMEM) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(mix_tc)");
    done.mark();
  }
  return FNO_OK;
}

ModeGeo make_geo(const Plan* p) {
  ModeGeo g;
  g.nd = p->nd;
  if (p->nd == 2) {
    g.m1 = p->m1;
    g.m2 = 0;
    g.inner = p->m2;
    g.M = 2 * p->m1 * p->m2;
    g.Mc = p->m1 * p->m2;
  } else {
    g.m1 = p->m1x;
    g.m2 = p->m1;
    g.inner = p->m2;
    g.M = 2 * p->m1x * 2 * p->m1 * p->m2;
    g.Mc = p->m1x * p->m1 * p->m2;
  }
  return g;
}

}  // namespace

int launch_mix_fwd(const Plan* p, const float* X, const float* const* w, float* Y, int B, int Ci, int Co,
                   cudaStream_t st) {
  const ModeGeo geo = make_geo(p);
  WPtrs wp;
  const int nc = p->nd == 2 ? 2 : 4;
  for (int c = 0; c < 4; ++c) wp.w[c] = reinterpret_cast<const float2*>(w[c < nc ? c : 0]);
  bool al = aligned16(X) && aligned16(Y);
  for (int c = 0; c < nc; ++c) al = al && aligned16(w[c]);
  if (al && mix_tc_envelope(geo, Ci, Co)) {
    if (int rc = setup_mix_tc_attrs()) return rc;
    dim3 grid(geo.M / 2, (B + MT_BCH - 1) / MT_BCH);
    mix_tc_kernel<false><<<grid, MT_THREADS, MT_SMEM, st>>>(reinterpret_cast<const float2*>(X), reinterpret_cast<float2*>(Y),
                                                             wp, geo, B, Ci, Co, Co, g_math_mode.load());
    count_launch();
    return check_launch("mix_tc_kernel<fwd>");
  }
  dim3 block(32, 4);
  dim3 grid((geo.M + 31) / 32, (Co + TO - 1) / TO, (B + TB * 4 - 1) / (TB * 4));
  mix_kernel<false><<<grid, block, 0, st>>>(reinterpret_cast<const float2*>(X), reinterpret_cast<float2*>(Y), wp,
                                            geo, B, Ci, Co, Co);
  count_launch();
  return check_launch("mix_kernel<fwd>");
}

int launch_mix_bwd_data(const Plan* p, const float* gY, const float* const* w, float* gX, int B, int Ci, int Co,
                        cudaStream_t st) {
  const ModeGeo geo = make_geo(p);
  WPtrs wp;
  const int nc = p->nd == 2 ? 2 : 4;
  for (int c = 0; c < 4; ++c) wp.w[c] = reinterpret_cast<const float2*>(w[c < nc ? c : 0]);
  bool al = aligned16(gY) && aligned16(gX);
  for (int c = 0; c < nc; ++c) al = al && aligned16(w[c]);
  if (al && mix_tc_envelope(geo, Ci, Co)) {
    if (int rc = setup_mix_tc_attrs()) return rc;
    dim3 grid(geo.M / 2, (B + MT_BCH - 1) / MT_BCH);
    mix_tc_kernel<true><<<grid, MT_THREADS, MT_SMEM, st>>>(reinterpret_cast<const float2*>(gY), reinterpret_cast<float2*>(gX),
                                                            wp, geo, B, Co, Ci, Co, g_math_mode.load());
    count_launch();
    return check_launch("mix_tc_kernel<bwd_data>");
  }
  dim3 block(32, 4);
  dim3 grid((geo.M + 31) / 32, (Ci + TO - 1) / TO, (B + TB * 4 - 1) / (TB * 4));
  mix_kernel<true><<<grid, block, 0, st>>>(reinterpret_cast<const float2*>(gY), reinterpret_cast<float2*>(gX), wp,
                                           geo, B, Co, Ci, Co);
  count_launch();
  return check_launch("mix_kernel<bwd_data>");
}

int launch_mix_bwd_weight(const Plan* p, const float* X, const float* gY, float* const* gw, int B, int Ci,
                          int Co, cudaStream_t st) {
  const ModeGeo geo = make_geo(p);
  GWPtrs gp;
  const int nc = p->nd == 2 ? 2 : 4;
  for (int c = 0; c < 4; ++c) gp.w[c] = reinterpret_cast<float2*>(gw[c < nc ? c : 0]);
  bool al = aligned16(X) && aligned16(gY);
  for (int c = 0; c < nc; ++c) al = al && aligned16(gw[c]);
  if (al && mix_tc_envelope(geo, Ci, Co)) {
    if (int rc = setup_mix_tc_attrs()) return rc;
    mix_wgrad_tc_kernel<<<geo.M / 2, MT_THREADS, MW_SMEM, st>>>(reinterpret_cast<const float2*>(X),
                                                                reinterpret_cast<const float2*>(gY), gp, geo, B, Ci, Co,
                                                                g_math_mode.load());
    count_launch();
    return check_launch("mix_wgrad_tc_kernel");
  }
  dim3 block(32, WG_SLICES);
  dim3 grid((geo.M + 31) / 32, (Ci + TI - 1) / TI, (Co + TO - 1) / TO);
  mix_wgrad_kernel<<<grid, block, 0, st>>>(reinterpret_cast<const float2*>(X),
                                           reinterpret_cast<const float2*>(gY), gp, geo, B, Ci, Co);
  count_launch();
  return check_launch("mix_wgrad_kernel");
}

bool mix_tc_supported(const Plan* p, int Ci, int Co) { return mix_tc_envelope(make_geo(p), Ci, Co); }

}  // namespace fno
