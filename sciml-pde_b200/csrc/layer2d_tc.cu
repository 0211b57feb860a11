// K3 on the tensor cores with the 1x1-conv bypass folded into the same GEMM (fno/fno.py:161-164: x1 = conv(x),
// x2 = w(x), x = gelu(x1 + x2)) and, as its adjoint, the data gradient of a Fourier layer (K3 of gX plus W^T dS).
//
// For one row (b, h) of the padded plane the pre-activation of ALL output channels is ONE GEMM
//
//   s[w, co] = sum_q' F[w, q'] T2[b, co, h, q']  +  sum_ci a[b, ci, h, w] Wl[co, ci]  +  bias[co]
//
//   M = 128 (w: TMEM lane), N = channels, K = 2*m2 (contiguous-axis inverse DFT) + C + 1 (bypass + bias)
//
// whose A operand lives in TENSOR memory: the twiddle columns F = [cos | sin](2 pi q w / W) are written once per
// CTA, the activation columns a[b, :, h, w] are loaded straight from HBM by converter warps (lane = w: one
// coalesced 128-byte request per channel row), split hi / lo in registers (3xTF32: fp32-mode accuracy) and
// stored with tcgen05.st; the bias rides on a constant-one column.  The B operand [T2(b, h) ; Wl] sits in
// shared memory: Wl (hi / lo) once per CTA, the 2.3 KB T2 tile of the row re-split per tile by one warp.
// T2 = strided-axis inverse of the mixed spectrum (130 x 24 reals per plane instead of 130 x 130) is produced
// by hinv_tiles_kernel directly in the K-major core-matrix layout of that operand, with the C2R column
// weights, the 1/(HW) scale and the sign of the sine part folded in.  The epilogue reads D[w, co] back with
// tcgen05.ld (thread = w), stores the pre-activation and its exact-erf GELU with one coalesced 128-byte
// store per channel and warp.  The bypass output `lin`, its 173 MB round trip and the separate
// pointwise kernel of round 1 are gone; the activation is read once and each output written once.
//
// The last W - 128 columns of a row (2 for the 130-wide padded planes of cfg 1) do not fit the 128 TMEM lanes; an
// "edge" warp computes them on the FP32 pipes from the same shared-memory operands.
#include "tc_common.cuh"

namespace fno {
namespace {

constexpr int L2_RMAX = 4;        // edge columns (W - 128) handled on the FP32 pipes
constexpr int L2_EPF = 4;         // prefetch depth (tiles) of the edge warp's activation loads
constexpr int L2_NR = 4;          // T2 tiles in flight (bulk copies into a shared-memory ring)
constexpr int L2_PD = 3;          // tiles of activation loads in flight per converter thread

__device__ __forceinline__ void tmem_st4(unsigned taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3]))
               : "memory");
}

// ------------------------------------------------------------------------------------------------------
// strided-axis inverse into B-operand tiles:  Z[h, q] = sum_r Y[r, q] e^{+2 pi i k_r h / H}  (k_r signed),
//   T2[(b, h)][n = c][k = q]      =  sc_q Re Z,      T2[..][c][m2 + q] = -sc_q Im Z
// (sc_q = scale * c2r weight).  Frequencies are folded onto j = |k| and rows onto pairs (t, H - t):
//   P_j = Y[+j] + Y[-j], M_j = Y[+j] - Y[-j]:   Re Z = Y0r + sum_j (Pr cos - Mi sin),  Im Z = Y0i + sum_j (Pi cos + Mr sin)
// thread = (channel of an 8-row group, q); the 4 m1 folded coefficients stay in registers.
// tile layout (floats): (c >> 3) * KQ * 8 + (k >> 2) * 32 + (c & 7) * 4 + (k & 3)   [8 x 16-byte core matrices]
// ------------------------------------------------------------------------------------------------------
template <int M1T>
__global__ void __launch_bounds__(256)
hinv_tiles_kernel(const float2* __restrict__ Y, float* __restrict__ T2g, const float* __restrict__ twH, int H, int W,
                  int m1, int m2, int C, int D1, int KQ, int tile_floats, int TL, int cmode, float scale) {
  constexpr int JP = ((2 * M1T + 1) + 3) & ~3;
  const int q = threadIdx.x % m2, cl = threadIdx.x / m2;
  const int c = blockIdx.z * 8 + cl;
  if (cl >= 8 || c >= C) return;
  const int bd = blockIdx.x;                  // b * D1 + d1
  const int b = bd / D1, d1 = bd - b * D1;
  const size_t plane = ((size_t)b * C + c) * D1 + d1;
  const float2* __restrict__ Yp = Y + plane * (size_t)(2 * m1) * m2 + q;
  float PR[M1T + 1], PI[M1T + 1], MR[M1T + 1], MI[M1T + 1];
  const float2 y0 = __ldg(Yp);
#pragma unroll
  for (int j = 1; j <= M1T; ++j) {
    float2 yp = make_float2(0.f, 0.f), yn = make_float2(0.f, 0.f);
    if (j < m1) yp = __ldg(Yp + (size_t)j * m2);
    if (j <= m1) yn = __ldg(Yp + (size_t)(2 * m1 - j) * m2);
    PR[j] = yp.x + yn.x; PI[j] = yp.y + yn.y;
    MR[j] = yp.x - yn.x; MI[j] = yp.y - yn.y;
  }
  float sc = scale;
  if (cmode && q != 0 && !((W & 1) == 0 && 2 * q == W)) sc *= 2.0f;
  const int NP = H / 2 + 1;
  const int t0 = blockIdx.y * TL, t1 = (t0 + TL < NP) ? t0 + TL : NP;
  const int eoff = (c >> 3) * KQ * 8 + (c & 7) * 4;
  const int o_re = eoff + (q >> 2) * 32 + (q & 3), o_im = eoff + ((m2 + q) >> 2) * 32 + ((m2 + q) & 3);
  float* __restrict__ base = T2g + (size_t)bd * H * tile_floats;
  for (int t = t0; t < t1; ++t) {
    const float4* __restrict__ r4 = reinterpret_cast<const float4*>(twH + (size_t)t * JP);
    float tw[JP];
#pragma unroll
    for (int i = 0; i < JP / 4; ++i) {
      const float4 v = __ldg(r4 + i);
      tw[4 * i] = v.x; tw[4 * i + 1] = v.y; tw[4 * i + 2] = v.z; tw[4 * i + 3] = v.w;
    }
    float er = y0.x, ei = y0.y, odr = 0.f, odi = 0.f;
#pragma unroll
    for (int j = 1; j <= M1T; ++j) {
      er = fmaf(PR[j], tw[j], er);
      ei = fmaf(PI[j], tw[j], ei);
      odr = fmaf(MI[j], tw[M1T + j], odr);
      odi = fmaf(MR[j], tw[M1T + j], odi);
    }
    float* __restrict__ o = base + (size_t)t * tile_floats;
    o[o_re] = sc * (er - odr);
    o[o_im] = -sc * (ei + odi);
    if (t != 0 && 2 * t != H) {
      float* __restrict__ o2 = base + (size_t)(H - t) * tile_floats;
      o2[o_re] = sc * (er + odr);
      o2[o_im] = -sc * (ei - odi);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------------
template <int KA, int NPAD, int CWQ>
struct L2Cfg {
  static constexpr int KC = (((KA + CWQ - 1) / CWQ) + 3) & ~3;   // A columns per converter warp
  static constexpr int EW = NPAD / 16;                           // epilogue warps per lane quadrant
  static constexpr int CONV_WARPS = 4 * CWQ;
  static constexpr int EPI_WARP0 = CONV_WARPS;
  static constexpr int EPI_WARPS = 4 * EW;
  static constexpr int MMA_WARP = CONV_WARPS + EPI_WARPS;
  static constexpr int BPREP_WARP = MMA_WARP + 1;
  static constexpr int EDGE_WARP = MMA_WARP + 2;
  static constexpr int THREADS = 32 * (MMA_WARP + 3);
  static constexpr int F4_PER_LANE = NPAD / 4;                   // float4 of a T2 tile per B-prep lane (KQ <= 32)
  static constexpr int ECH = NPAD / 32;                          // channels per edge-warp lane
  static constexpr unsigned TM_COLS = (2 * 32 + 4 * KA + 2 * NPAD <= 256) ? 256u : 512u;
};

struct L2Args {
  const float* a;       // [B, C, RS, W] activation (forward) / dS (adjoint)
  const float* T2g;     // [B * RS] tiles from hinv_tiles_kernel
  const float* Wl;      // [C, C] bypass weight
  const float* bias;    // [C] or null
  float* s_out;         // optional pre-activation
  float* out;
  const float* twW;     // [2][m2][WP] cos / sin of 2 pi q w / W
  int WP, W, m2, KQ, C, RS;
  int tile_floats;
  long total_tiles;     // B * RS
  int transpose_w, apply_gelu, single;
  FastDiv rs_div;
};

template <int KA, int NPAD, int CWQ>
__global__ void __launch_bounds__(L2Cfg<KA, NPAD, CWQ>::THREADS, 1)
layer2d_tc_kernel(const L2Args p) {
  using Cfg = L2Cfg<KA, NPAD, CWQ>;
  constexpr int KC = Cfg::KC;
  extern __shared__ __align__(128) unsigned char lsm[];
  const int KQ = p.KQ, C = p.C, W = p.W;
  const int t2_tile = NPAD * KQ;                          // floats of one hi (or lo) T2 operand tile
  float* bT2 = reinterpret_cast<float*>(lsm);             // [2 stages][hi | lo][NPAD * KQ]
  float* bW = bT2 + 4 * t2_tile;                          // [hi | lo][NPAD * KA]
  float* eW = bW + 2 * NPAD * KA;                         // [C][C + 1] raw weight (edge warp)
  float* eB = eW + ((C * (C + 1) + 3) & ~3);              // [NPAD] bias
  float* eF = eB + NPAD;                                  // [L2_RMAX][KQ] twiddles of the edge columns
  float* eA = eF + L2_RMAX * KQ;                          // [NPAD][L2_RMAX] edge activations of the current tile
  float* ring = eA + NPAD * L2_RMAX;                      // [L2_NR][tile_floats] raw T2 tiles (bulk-copy destination)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<size_t>(ring + (size_t)L2_NR * p.tile_floats) + 15) & ~size_t(15));
  unsigned long long* a_ready = bars;        // [2] converters wrote A buffer s
  unsigned long long* b_ready = bars + 2;    // [2] T2 operand tile s is split and staged
  unsigned long long* mma_done = bars + 4;   // [2] MMAs of the tile on buffers s complete (A / B free, D full)
  unsigned long long* d_free = bars + 6;     // [2] accumulator s read back
  unsigned long long* edge_done = bars + 8;  // [2] edge warp done with T2 tile s
  unsigned long long* raw_full = bars + 10;  // [L2_NR] bulk copy of a raw T2 tile landed
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 10 + L2_NR);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int Wm = W < 128 ? W : 128;          // columns on the tensor cores
  const int r_edge = W - Wm;                 // columns on the edge warp
  const int m2x2 = 2 * p.m2;
  const size_t cs = (size_t)p.RS * W;        // channel stride
  const long t_begin = (p.total_tiles * (long)blockIdx.x) / (long)gridDim.x;          // total_tiles < 2^32, grid <= 148
  const long t_end = (p.total_tiles * (long)(blockIdx.x + 1)) / (long)gridDim.x;
  const int ntl = (int)(t_end - t_begin);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_ready + s, Cfg::CONV_WARPS);
      mbar_init(b_ready + s, 1);
      mbar_init(mma_done + s, 1);
      mbar_init(d_free + s, Cfg::EPI_WARPS);
      mbar_init(edge_done + s, 1);
    }
    for (int s = 0; s < L2_NR; ++s) mbar_init(raw_full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == Cfg::MMA_WARP) tmem_alloc(tmem_slot, Cfg::TM_COLS);
  // constant operands: Wl (+ bias row) hi / lo in the K-major core-matrix layout, zero T2 tiles (rows >= C stay zero)
  for (int i = tid; i < 4 * t2_tile; i += Cfg::THREADS) bT2[i] = 0.f;
  for (int i = tid; i < NPAD * KA; i += Cfg::THREADS) {
    const int n = i / KA, k = i - n * KA;
    float v = 0.f;
    if (n < C) {
      if (k < C) v = __ldg(p.Wl + (p.transpose_w ? (size_t)k * C + n : (size_t)n * C + k));
      else if (k == C && p.bias != nullptr) v = __ldg(p.bias + n);
    }
    float hi, lo;
    split_tf32(v, hi, lo);
    const int off = (n >> 3) * KA * 8 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
    bW[off] = hi;
    bW[NPAD * KA + off] = lo;
  }
  for (int i = tid; i < C * C; i += Cfg::THREADS) {
    const int n = i / C, k = i - n * C;
    eW[n * (C + 1) + k] = __ldg(p.Wl + (p.transpose_w ? (size_t)k * C + n : (size_t)n * C + k));
  }
  for (int i = tid; i < NPAD; i += Cfg::THREADS) eB[i] = (i < C && p.bias != nullptr) ? __ldg(p.bias + i) : 0.f;
  for (int i = tid; i < L2_RMAX * KQ; i += Cfg::THREADS) {
    const int j = i / KQ, k = i - j * KQ;
    eF[i] = (j < r_edge && k < m2x2) ? __ldg(p.twW + (size_t)k * p.WP + Wm + j) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  const unsigned TM_FHI = 0, TM_FLO = (unsigned)KQ, TM_A = 2u * KQ, TM_D = 2u * KQ + 4u * KA;

  if (warp < Cfg::CONV_WARPS) {
    // ---- converters: lane = w, KC columns of the bypass part of A -----------------------------------
    const int quad = warp & 3, cw = warp >> 2;
    const int w = quad * 32 + lane;
    const bool wv = w < Wm;
    const unsigned ta = tmem_base + ((unsigned)(quad * 32) << 16);
    if (cw == 0) {
      // twiddle columns, once: F[w, q'] = twW[q'][w]
      for (int g = 0; g < KQ / 4; ++g) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 4 * g + e;
          const float v = (wv && k < m2x2) ? __ldg(p.twW + (size_t)k * p.WP + w) : 0.f;
          split_tf32(v, hi[e], lo[e]);
        }
        tmem_st4(ta + TM_FHI + 4u * g, hi);
        tmem_st4(ta + TM_FLO + 4u * g, lo);
      }
      tmem_st_wait();
    }
    const int col0 = cw * KC;
    float pre[L2_PD][KC];
    auto load_tile = [&](int it, float (&r)[KC]) {
      if (it >= ntl) return;
      const long T = t_begin + it;
      const unsigned b = p.rs_div.div((unsigned)T);
      const unsigned row = (unsigned)T - b * (unsigned)p.RS;
      const float* __restrict__ src = p.a + ((size_t)b * C * p.RS + row) * W + w + (size_t)col0 * cs;
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        r[i] = (wv && col0 + i < C) ? __ldg(src) : 0.f;
        src += cs;
      }
    };
    auto convert_tile = [&](int it, float (&r)[KC]) {
      const int s = it & 1;
      const unsigned ph = ((unsigned)it >> 1) & 1u;
      mbar_wait(mma_done + s, ph ^ 1u);          // the MMAs of tile it - 2 no longer read A buffer s
      tc_fence_after();
      const unsigned ah = ta + TM_A + (unsigned)(s * 2 * KA) + (unsigned)col0, al = ah + KA;
#pragma unroll
      for (int g = 0; g < KC / 4; ++g) {
        if (col0 + 4 * g >= KA) break;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = col0 + 4 * g + e;
          const float v = (col == C) ? 1.0f : r[4 * g + e];    // constant-one column carries the bias
          split_tf32(v, hi[e], lo[e]);
        }
        tmem_st4(ah + 4u * g, hi);
        if (!p.single) tmem_st4(al + 4u * g, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready + s);
    };
#pragma unroll
    for (int d = 0; d < L2_PD; ++d) load_tile(d, pre[d]);
    for (int it = 0; it < ntl; it += L2_PD) {
#pragma unroll
      for (int d = 0; d < L2_PD; ++d) {
        if (it + d < ntl) {
          convert_tile(it + d, pre[d]);
          load_tile(it + d + L2_PD, pre[d]);
        }
      }
    }
  } else if (warp < Cfg::MMA_WARP) {
    // ---- epilogue: thread = w, a slice of the output channels ------------------------------------------
    const int ew = warp - Cfg::EPI_WARP0;
    const int quad = ew & 3, e = ew >> 2;
    const int CPW = (C + Cfg::EW - 1) / Cfg::EW;
    const int c0 = e * CPW;
    const int cn = (C - c0 < CPW) ? C - c0 : CPW;
    const int w = quad * 32 + lane;
    const bool wv = w < Wm;
    for (int it = 0; it < ntl; ++it) {
      const int s = it & 1;
      const unsigned ph = ((unsigned)it >> 1) & 1u;
      const long T = t_begin + it;
      const unsigned b = p.rs_div.div((unsigned)T);
      const unsigned row = (unsigned)T - b * (unsigned)p.RS;
      mbar_wait(mma_done + s, ph);
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem_base + ((unsigned)(quad * 32) << 16) + TM_D + (unsigned)(s * NPAD + c0), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_free + s);
      size_t off = (((size_t)b * C + c0) * p.RS + row) * W + w;
      if (wv) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i < cn) {
            const float x = v[i];
            if (p.s_out != nullptr) p.s_out[off] = x;
            p.out[off] = p.apply_gelu ? gelu_fast(x) : x;
            off += cs;
          }
        }
      }
    }
  } else if (warp == Cfg::MMA_WARP) {
    // ---- MMA issuer (whole warp converged, one elected lane issues) ------------------------------------
    constexpr unsigned idesc = umma_idesc_tf32(128, NPAD, 0, 0);
    const unsigned long long d_wh = umma_desc(bW, 128, KA * 32), d_wl = umma_desc(bW + NPAD * KA, 128, KA * 32);
    const int single = p.single;
    for (int it = 0; it < ntl; ++it) {
      const int s = it & 1;
      const unsigned ph = ((unsigned)it >> 1) & 1u;
      const unsigned td = tmem_base + TM_D + (unsigned)(s * NPAD);
      const unsigned long long d_th = umma_desc(bT2 + (size_t)(2 * s) * t2_tile, 128, KQ * 32);
      const unsigned long long d_tl = umma_desc(bT2 + (size_t)(2 * s + 1) * t2_tile, 128, KQ * 32);
      mbar_wait(d_free + s, ph ^ 1u);
      mbar_wait(b_ready + s, ph);
      tc_fence_after();
      __syncwarp();
#pragma unroll 1
      for (int ks = 0; ks < KQ / 8; ++ks) {
        const unsigned long long fo = (unsigned long long)(ks * 16);
        const unsigned fh = tmem_base + TM_FHI + 8u * ks, fl = tmem_base + TM_FLO + 8u * ks;
        if (!single) {
          tc_mma_tf32_ts_elect(td, fl, d_th + fo, idesc, ks != 0);
          tc_mma_tf32_ts_elect(td, fh, d_tl + fo, idesc, 1u);
        }
        tc_mma_tf32_ts_elect(td, fh, d_th + fo, idesc, single ? (unsigned)(ks != 0) : 1u);
      }
      mbar_wait(a_ready + s, ph);
      tc_fence_after();
      __syncwarp();
      const unsigned ab = tmem_base + TM_A + (unsigned)(s * 2 * KA);
#pragma unroll 1
      for (int ks = 0; ks < KA / 8; ++ks) {
        const unsigned long long fo = (unsigned long long)(ks * 16);
        const unsigned ah = ab + 8u * ks, al = ah + KA;
        if (!single) {
          tc_mma_tf32_ts_elect(td, al, d_wh + fo, idesc, 1u);
          tc_mma_tf32_ts_elect(td, ah, d_wl + fo, idesc, 1u);
        }
        tc_mma_tf32_ts_elect(td, ah, d_wh + fo, idesc, 1u);
      }
      tc_commit_elect(mma_done + s);
    }
  } else if (warp == Cfg::BPREP_WARP) {
    // ---- T2 operand tile: global (fp32, operand layout) -> hi / lo in shared memory ----------------------
    constexpr int NF = Cfg::F4_PER_LANE;
    const int NG = (C + 7) >> 3;
    const int nf4 = NG * KQ * 2;                      // float4 per tile
    const unsigned tile_bytes = (unsigned)p.tile_floats * 4u;
    auto issue = [&](int it) {                        // lane 0: one bulk copy per tile
      const int slot = it & (L2_NR - 1);
      mbar_arrive_expect_tx(raw_full + slot, tile_bytes);
      bulk_g2s(ring + (size_t)slot * p.tile_floats, p.T2g + (size_t)(t_begin + it) * p.tile_floats, tile_bytes, raw_full + slot);
    };
    if (lane == 0)
      for (int i = 0; i < L2_NR && i < ntl; ++i) issue(i);
    __syncwarp();
    for (int it = 0; it < ntl; ++it) {
      const int s = it & 1;
      const unsigned ph = ((unsigned)it >> 1) & 1u;
      const int slot = it & (L2_NR - 1);
      mbar_wait(raw_full + slot, ((unsigned)it / L2_NR) & 1u);
      float4 raw[NF];
      const float4* __restrict__ src = reinterpret_cast<const float4*>(ring + (size_t)slot * p.tile_floats);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int f = lane + 32 * i;
        raw[i] = (f < nf4) ? src[f] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      if (lane == 0 && it + L2_NR < ntl) issue(it + L2_NR);          // the slot has been read: refill it
      mbar_wait(mma_done + s, ph ^ 1u);
      if (r_edge > 0) mbar_wait(edge_done + s, ph ^ 1u);
      float4* __restrict__ dh = reinterpret_cast<float4*>(bT2 + (size_t)(2 * s) * t2_tile);
      float4* __restrict__ dl = reinterpret_cast<float4*>(bT2 + (size_t)(2 * s + 1) * t2_tile);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int f = lane + 32 * i;
        if (f < nf4) {
          // float4 f of the tile: row group g, chunk kc, row r  ->  n = 8 g + r, k = 4 kc .. 4 kc + 3
          const int g = f / (KQ * 2), rem = f - g * KQ * 2;
          const int kc = rem >> 3, n = 8 * g + (rem & 7);
          const float in[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = (n < C && 4 * kc + e < m2x2) ? in[e] : 0.f;
            split_tf32(v, hi[e], lo[e]);
          }
          dh[f] = make_float4(hi[0], hi[1], hi[2], hi[3]);
          dl[f] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready + s);
    }
  } else {
    // ---- edge warp: the W - 128 columns that do not fit the TMEM lanes, on the FP32 pipes ------------------
    constexpr int ECH = Cfg::ECH;
    float pa[L2_EPF][ECH][L2_RMAX];
    auto load_tile = [&](int it, float (&r)[ECH][L2_RMAX]) {
      if (it >= ntl || r_edge == 0) return;
      const long T = t_begin + it;
      const unsigned b = p.rs_div.div((unsigned)T);
      const unsigned row = (unsigned)T - b * (unsigned)p.RS;
#pragma unroll
      for (int ch = 0; ch < ECH; ++ch) {
        const int ci = lane + 32 * ch;
        const float* __restrict__ src = p.a + (((size_t)b * C + ci) * p.RS + row) * W + Wm;
#pragma unroll
        for (int j = 0; j < L2_RMAX; ++j) r[ch][j] = (ci < C && j < r_edge) ? __ldg(src + j) : 0.f;
      }
    };
    auto do_tile = [&](int it, float (&r)[ECH][L2_RMAX]) {
      const int s = it & 1;
      const unsigned ph = ((unsigned)it >> 1) & 1u;
      const long T = t_begin + it;
      const unsigned b = p.rs_div.div((unsigned)T);
      const unsigned row = (unsigned)T - b * (unsigned)p.RS;
#pragma unroll
      for (int ch = 0; ch < ECH; ++ch) {
        const int ci = lane + 32 * ch;
        if (ci < C) *reinterpret_cast<float4*>(eA + ci * L2_RMAX) = make_float4(r[ch][0], r[ch][1], r[ch][2], r[ch][3]);
      }
      __syncwarp();
      mbar_wait(b_ready + s, ph);
      const float* __restrict__ th = bT2 + (size_t)(2 * s) * t2_tile;
      const float* __restrict__ tl = th + t2_tile;
#pragma unroll
      for (int ch = 0; ch < ECH; ++ch) {
        const int co = lane + 32 * ch;
        if (co < C) {
          const int ro = (co >> 3) * KQ * 8 + (co & 7) * 4;
          const size_t off = (((size_t)b * C + co) * p.RS + row) * W + Wm;
          for (int j0 = 0; j0 < r_edge; j0 += 2) {             // two edge columns per pass
            float acc0 = eB[co], acc1 = acc0;
            const float* __restrict__ f0 = eF + j0 * KQ;
            const float* __restrict__ f1 = f0 + KQ;
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) {
              if (kc < KQ / 4) {
                const float4 h4 = *reinterpret_cast<const float4*>(th + ro + kc * 32);
                const float4 l4 = *reinterpret_cast<const float4*>(tl + ro + kc * 32);
                const float t0 = h4.x + l4.x, t1 = h4.y + l4.y, t2 = h4.z + l4.z, t3 = h4.w + l4.w;   // = the fp32 value
                const float4 fa = *reinterpret_cast<const float4*>(f0 + kc * 4);
                const float4 fb = *reinterpret_cast<const float4*>(f1 + kc * 4);
                acc0 = fmaf(t0, fa.x, acc0); acc1 = fmaf(t0, fb.x, acc1);
                acc0 = fmaf(t1, fa.y, acc0); acc1 = fmaf(t1, fb.y, acc1);
                acc0 = fmaf(t2, fa.z, acc0); acc1 = fmaf(t2, fb.z, acc1);
                acc0 = fmaf(t3, fa.w, acc0); acc1 = fmaf(t3, fb.w, acc1);
              }
            }
#pragma unroll 4
            for (int ci = 0; ci < C; ++ci) {
              const float wv = eW[co * (C + 1) + ci];
              const float2 av = *reinterpret_cast<const float2*>(eA + ci * L2_RMAX + j0);
              acc0 = fmaf(wv, av.x, acc0);
              acc1 = fmaf(wv, av.y, acc1);
            }
            if (p.s_out != nullptr) {
              p.s_out[off + j0] = acc0;
              if (j0 + 1 < r_edge) p.s_out[off + j0 + 1] = acc1;
            }
            p.out[off + j0] = p.apply_gelu ? gelu_fast(acc0) : acc0;
            if (j0 + 1 < r_edge) p.out[off + j0 + 1] = p.apply_gelu ? gelu_fast(acc1) : acc1;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(edge_done + s);
    };
    if (r_edge > 0) {
#pragma unroll
    for (int d = 0; d < L2_EPF; ++d) load_tile(d, pa[d]);
    for (int it = 0; it < ntl; it += L2_EPF) {
#pragma unroll
      for (int d = 0; d < L2_EPF; ++d) {
        if (it + d < ntl) {
          do_tile(it + d, pa[d]);
          load_tile(it + d + L2_EPF, pa[d]);
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc(tmem_base, Cfg::TM_COLS);
}

template <int KA, int NPAD, int CWQ>
size_t layer2d_smem_bytes(int KQ, int C) {
  return sizeof(float) * ((size_t)4 * NPAD * KQ + 2 * NPAD * KA + (size_t)((C * (C + 1) + 3) & ~3) + NPAD + L2_RMAX * KQ +
                          NPAD * L2_RMAX + (size_t)L2_NR * ((C + 7) / 8) * KQ * 8) + 16 + (10 + L2_NR) * 8 + 16;
}

template <int KA, int NPAD, int CWQ>
int launch_layer2d_t(const L2Args& args, cudaStream_t st) {
  using Cfg = L2Cfg<KA, NPAD, CWQ>;
  const size_t smem = layer2d_smem_bytes<KA, NPAD, CWQ>(args.KQ, args.C);
  if (smem > 48 * 1024) { set_error("layer2d_tc: shared memory %zu", smem); return FNO_E_ARG; }
  const long ctas = args.total_tiles < 148 ? args.total_tiles : 148;
  layer2d_tc_kernel<KA, NPAD, CWQ><<<(unsigned)ctas, Cfg::THREADS, smem, st>>>(args);
  count_launch();
  return check_launch("layer2d_tc_kernel");
}

template <int M1T>
int launch_hinv_t(const Plan* p, const float* Y, float* T2g, int B, int C, int KQ, int tile_floats, int cmode, float scale,
                  cudaStream_t st) {
  const int NP = p->H / 2 + 1;
  int threads = 8 * p->m2;
  threads = (threads + 31) & ~31;
  // row-pair slices: enough CTAs for ~4 waves of small blocks
  const long base = (long)B * p->D1 * ((C + 7) / 8);
  int TS = (int)((4L * 148 * 4 + base - 1) / base);
  if (TS < 1) TS = 1;
  if (TS > NP) TS = NP;
  const int TL = (NP + TS - 1) / TS;
  TS = (NP + TL - 1) / TL;
  dim3 grid((unsigned)(B * p->D1), (unsigned)TS, (unsigned)((C + 7) / 8));
  hinv_tiles_kernel<M1T><<<grid, threads, 0, st>>>(reinterpret_cast<const float2*>(Y), T2g, p->twH, p->H, p->W, p->m1, p->m2,
                                                  C, p->D1, KQ, tile_floats, TL, cmode, scale);
  count_launch();
  return check_launch("hinv_tiles_kernel");
}

}  // namespace

// geometry the tensor-core layer kernel covers: one 128-lane tile (+ <= 4 edge columns) per row, 2 m2 <= 32
// twiddle columns, width + bias <= 32, 8 m2 <= 256 threads for the strided-axis stage
bool layer2d_tc_supported(const Plan* p, int C) {
  if (p == nullptr || C < 1 || C > 31) return false;
  if (p->W > 128 + L2_RMAX || 2 * p->m2 > 32 || 8 * p->m2 > 256) return false;
  if ((long)p->H * p->D1 >= (1L << 31)) return false;
  return true;
}

static int kq_of(const Plan* p) { return (2 * p->m2 + 7) & ~7; }
static int tile_floats_of(const Plan* p, int C) { return ((C + 7) / 8) * kq_of(p) * 8; }

size_t layer2d_tc_workspace_bytes(const Plan* p, int B, int C) {
  if (!layer2d_tc_supported(p, C) || B <= 0) return 0;
  return sizeof(float) * (size_t)B * p->D1 * p->H * tile_floats_of(p, C);
}

// out = act( K3(Y) + Wl a + bias ),  s_out = pre-activation (optional);  transpose_w: Wl^T a (adjoint)
int launch_layer2d_tc(const Plan* p, const float* Y, const float* a, const float* Wl, const float* bias, float* s_out,
                      float* out, float* work, int B, int C, int cmode, float scale, int apply_gelu, int transpose_w,
                      cudaStream_t st) {
  if (!layer2d_tc_supported(p, C)) { set_error("layer2d_tc: geometry not supported"); return FNO_E_ARG; }
  if ((reinterpret_cast<size_t>(work) & 15) != 0) { set_error("layer2d_tc: workspace must be 16-byte aligned"); return FNO_E_ARG; }
  const int KQ = kq_of(p), tile_floats = tile_floats_of(p, C);
  const long total = (long)B * p->D1 * p->H;
  if (total >= (1L << 32)) { set_error("layer2d_tc: too many rows"); return FNO_E_ARG; }
  int rc;
  switch (p->M1T) {
    case 4: rc = launch_hinv_t<4>(p, Y, work, B, C, KQ, tile_floats, cmode, scale, st); break;
    case 8: rc = launch_hinv_t<8>(p, Y, work, B, C, KQ, tile_floats, cmode, scale, st); break;
    case 12: rc = launch_hinv_t<12>(p, Y, work, B, C, KQ, tile_floats, cmode, scale, st); break;
    case 16: rc = launch_hinv_t<16>(p, Y, work, B, C, KQ, tile_floats, cmode, scale, st); break;
    case 24: rc = launch_hinv_t<24>(p, Y, work, B, C, KQ, tile_floats, cmode, scale, st); break;
    case 32: rc = launch_hinv_t<32>(p, Y, work, B, C, KQ, tile_floats, cmode, scale, st); break;
    default: set_error("unsupported padded modes1 %d", p->M1T); return FNO_E_ARG;
  }
  if (rc != FNO_OK) return rc;
  L2Args args;
  args.a = a; args.T2g = work; args.Wl = Wl; args.bias = bias; args.s_out = s_out; args.out = out;
  args.twW = p->twW; args.WP = p->WP; args.W = p->W; args.m2 = p->m2; args.KQ = KQ; args.C = C;
  args.RS = p->D1 * p->H; args.tile_floats = tile_floats; args.total_tiles = total;
  args.transpose_w = transpose_w; args.apply_gelu = apply_gelu;
  args.single = g_math_mode.load() == FNO_MATH_TF32;
  args.rs_div.init((unsigned)args.RS);
  const int KA = (C + 1 + 7) & ~7;
  switch (KA) {
    case 8: return launch_layer2d_t<8, 32, 2>(args, st);
    case 16: return launch_layer2d_t<16, 32, 2>(args, st);
    case 24: return launch_layer2d_t<24, 32, 2>(args, st);
    case 32: return launch_layer2d_t<32, 32, 2>(args, st);
    default: set_error("layer2d_tc: width %d unsupported", C); return FNO_E_ARG;
  }
}

}  // namespace fno
