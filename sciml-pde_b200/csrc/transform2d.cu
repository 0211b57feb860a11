// K1 (pruned forward transform) and K3 (zero-padding inverse transform + fused epilogue) for one
// 2-D plane [H, W] -- FP32 CUDA-core path.
//
// Both kernels share one factorisation (oracle/dft_oracle.py):
//
//   strided axis H ("column pass"):  one thread per (plane, TN adjacent columns).  The 2*m1 kept
//   signed frequencies {-m1..m1-1} are folded onto |k| = j in [0, m1]:  with
//       A[j] = sum_h x[h] cos(2 pi j h / H),   B[j] = sum_h x[h] sin(2 pi j h / H)
//   the column DFT is U[+j] = A[j] - i B[j], U[-j] = A[j] + i B[j]; rows are further folded in
//   pairs (h, H-h): e = x[h] + x[H-h] feeds the cosines, o = x[h] - x[H-h] the sines.  That is
//   2*m1+1 real accumulators per column and (2*m1+1) FMAs per row *pair*.  The row twiddles are
//   warp-uniform float4 broadcasts from shared memory and are shared by the thread's TN columns,
//   so a row pair costs (2*m1+1)/4 LDS.128 + 2 vector loads per TN*(2*m1+1) FFMA.  x is read
//   straight from global memory, exactly once, coalesced (lanes <-> consecutive column groups).
//
//   contiguous axis W: the small [2*m1+1, W] x [W, 2*m2] contraction runs from shared memory as a
//   register-tiled product (2 frequencies j x 4 wavenumbers k2 per thread, the W range split over
//   several threads so that the whole CTA takes part), combined through shared memory.
//
// K3 mirrors it: the W-axis inverse is applied first to the tiny retained spectrum, leaving
// 2*m1+1 real coefficients per column in registers; the H-axis pass then emits two output rows
// (h, H-h) per (2*m1+1) FMAs and fuses "+ addend" (1x1-conv bypass), the optional store of the
// pre-activation, and the exact-erf GELU.
//
// Several planes are flattened into one CTA (thread <-> (g, column group), g < G) so that widths
// like 130 = 2 * 65 do not waste a nearly empty warp per plane.
#include <cstdlib>

#include "common.cuh"

namespace fno {

namespace {

template <int M1T>
struct Geo {
  static constexpr int NC = M1T + 1;      // cosine accumulators j = 0..M1T
  static constexpr int NJ = 2 * M1T + 1;  // + sine accumulators j = 1..M1T
  static constexpr int JP = (NJ + 3) & ~3;
};

template <int TN>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec ld(const float* p) { Vec r; r.v[0] = __ldg(p); return r; }
  __device__ __forceinline__ static Vec ld_plain(const float* p) { Vec r; r.v[0] = *p; return r; }
  __device__ __forceinline__ void st(float* p) const { *p = v[0]; }
};
template <>
struct Vec<2> {
  float v[2];
  __device__ __forceinline__ static Vec ld(const float* p) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    Vec r; r.v[0] = t.x; r.v[1] = t.y; return r;
  }
  __device__ __forceinline__ static Vec ld_plain(const float* p) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    Vec r; r.v[0] = t.x; r.v[1] = t.y; return r;
  }
  __device__ __forceinline__ void st(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
constexpr int PF_DIST = 16;   // row pairs of look-ahead for the L2 prefetches of the column passes

template <int JP>
__device__ __forceinline__ void load_row(float (&tw)[JP], const float* __restrict__ row) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int q = 0; q < JP / 4; ++q) {
    const float4 v = r4[q];
    tw[4 * q + 0] = v.x; tw[4 * q + 1] = v.y; tw[4 * q + 2] = v.z; tw[4 * q + 3] = v.w;
  }
}

// acc[n][0..M1T] += e[n] * cos row, acc[n][M1T+1..2*M1T] += o[n] * sin row
template <int M1T, int TN>
__device__ __forceinline__ void fold_accumulate(float (&acc)[TN][Geo<M1T>::NJ], const float* __restrict__ row,
                                                const Vec<TN>& e, const Vec<TN>& o) {
  float tw[Geo<M1T>::JP];
  load_row<Geo<M1T>::JP>(tw, row);
#pragma unroll
  for (int j = 0; j <= M1T; ++j)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[n][j] = fmaf(e.v[n], tw[j], acc[n][j]);
#pragma unroll
  for (int j = 1; j <= M1T; ++j)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[n][M1T + j] = fmaf(o.v[n], tw[M1T + j], acc[n][M1T + j]);
}

// cosine part only (self-paired rows h = 0 and h = H/2 have no odd component)
template <int M1T, int TN>
__device__ __forceinline__ void fold_accumulate_even(float (&acc)[TN][Geo<M1T>::NJ], const float* __restrict__ row,
                                                     const Vec<TN>& e) {
  float tw[Geo<M1T>::JP];
  load_row<Geo<M1T>::JP>(tw, row);
#pragma unroll
  for (int j = 0; j <= M1T; ++j)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[n][j] = fmaf(e.v[n], tw[j], acc[n][j]);
}

constexpr int S2_JT = 2;   // stage-2 tile: frequencies j per thread
constexpr int S2_KT = 4;   //               wavenumbers k2 per thread
constexpr int S2_ACC = S2_JT * S2_KT * 4;

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
template <int M1T, int TN, bool PREMUL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
fwd2d_kernel(const float* __restrict__ x, const float* __restrict__ preact, float* __restrict__ ds_out,
             float2* __restrict__ X, const float* __restrict__ twH, const float* __restrict__ twW,
             int H, int W, int WP, int m1, int m2, int G, long planes, int cmode, float scale) {
  constexpr int NJ = Geo<M1T>::NJ;
  constexpr int JP = Geo<M1T>::JP;
  const int NP = H / 2 + 1;
  extern __shared__ __align__(16) float smem[];
  float* twH_s = smem;                       // [NP][JP]
  float* twW_s = twH_s + NP * JP;            // [2][m2][WP]
  float* uab = twW_s + 2 * m2 * WP;          // [G][NJ][WP]   (re-used for the stage-2 partial sums)

  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  {
    const float4* s4 = reinterpret_cast<const float4*>(twH);
    float4* d4 = reinterpret_cast<float4*>(twH_s);
    for (int i = tid; i < NP * JP / 4; i += nthr) d4[i] = __ldg(s4 + i);
    s4 = reinterpret_cast<const float4*>(twW);
    d4 = reinterpret_cast<float4*>(twW_s);
    for (int i = tid; i < 2 * m2 * WP / 4; i += nthr) d4[i] = __ldg(s4 + i);
  }
  // zero the pad columns of uab (read by the float4 loops of stage 2)
  const int padw = WP - W;
  if (padw > 0) {
    for (int i = tid; i < G * NJ * padw; i += nthr) {
      const int r = i / padw;
      uab[r * WP + W + (i - r * padw)] = 0.0f;
    }
  }
  __syncthreads();

  // ---- stage 1: column pass over H ----------------------------------------------------------
  const int tpp = W / TN;                    // threads per plane
  const int g = tid / tpp;
  const int w0 = (tid - g * tpp) * TN;
  const long plane = (long)blockIdx.x * G + g;
  if (g < G) {
    float acc[TN][NJ];
#pragma unroll
    for (int n = 0; n < TN; ++n)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[n][j] = 0.0f;
    if (plane < planes) {
      const size_t base = (size_t)plane * H * W + w0;
      const float* __restrict__ xp = x + base;
      const float* __restrict__ sp = PREMUL ? preact + base : nullptr;
      float* __restrict__ dp = (PREMUL && ds_out != nullptr) ? ds_out + base : nullptr;
      auto finish = [&](Vec<TN> v, const Vec<TN>& s, int h) -> Vec<TN> {
        if (PREMUL) {
#pragma unroll
          for (int n = 0; n < TN; ++n) v.v[n] *= gelu_fast_grad(s.v[n]);
          if (dp != nullptr) v.st(dp + h * W);
        }
        return v;
      };
      auto load1 = [&](int h) -> Vec<TN> {
        const Vec<TN> v = Vec<TN>::ld(xp + h * W);
        Vec<TN> s = v;
        if (PREMUL) s = Vec<TN>::ld(sp + h * W);
        return finish(v, s, h);
      };
      fold_accumulate_even<M1T, TN>(acc, twH_s, load1(0));
      const int npairs = (H - 1) / 2;
      // PG row pairs per step: all loads of a step are issued before the FMAs that consume them
      // (memory-level parallelism per thread; the resident warps provide the rest).  A software
      // pipeline (loads of group g+1 before the FMAs of group g) was measured slower (0.109 vs
      // 0.101 ms at cfg 1): the burst form keeps more bytes in flight per thread.
      constexpr int PG = PREMUL ? 2 : 4;
      int t = 1;
      for (; t + PG - 1 <= npairs; t += PG) {
        Vec<TN> v1[PG], v2[PG], s1[PG], s2[PG];
        // pull the rows PF_DIST pairs ahead into L2: the loads below then pay the L2 latency, not HBM's
        if (t + PF_DIST + PG - 1 <= npairs) {
#pragma unroll
          for (int u = 0; u < PG; ++u) {
            prefetch_l2(xp + (t + PF_DIST + u) * W);
            prefetch_l2(xp + (H - t - PF_DIST - u) * W);
            if (PREMUL) {
              prefetch_l2(sp + (t + PF_DIST + u) * W);
              prefetch_l2(sp + (H - t - PF_DIST - u) * W);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < PG; ++u) {
          v1[u] = Vec<TN>::ld(xp + (t + u) * W);
          v2[u] = Vec<TN>::ld(xp + (H - t - u) * W);
          if (PREMUL) {
            s1[u] = Vec<TN>::ld(sp + (t + u) * W);
            s2[u] = Vec<TN>::ld(sp + (H - t - u) * W);
          }
        }
#pragma unroll
        for (int u = 0; u < PG; ++u) {
          const Vec<TN> a = finish(v1[u], s1[u], t + u), b = finish(v2[u], s2[u], H - t - u);
          Vec<TN> e, o;
#pragma unroll
          for (int n = 0; n < TN; ++n) { e.v[n] = a.v[n] + b.v[n]; o.v[n] = a.v[n] - b.v[n]; }
          fold_accumulate<M1T, TN>(acc, twH_s + (t + u) * JP, e, o);
        }
      }
      for (; t <= npairs; ++t) {
        const Vec<TN> a = load1(t), b = load1(H - t);
        Vec<TN> e, o;
#pragma unroll
        for (int n = 0; n < TN; ++n) { e.v[n] = a.v[n] + b.v[n]; o.v[n] = a.v[n] - b.v[n]; }
        fold_accumulate<M1T, TN>(acc, twH_s + t * JP, e, o);
      }
      if ((H & 1) == 0) fold_accumulate_even<M1T, TN>(acc, twH_s + (H / 2) * JP, load1(H / 2));
    }
    float* urow = uab + (size_t)g * NJ * WP + w0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      Vec<TN> v;
#pragma unroll
      for (int n = 0; n < TN; ++n) v.v[n] = acc[n][j];
      v.st(urow + j * WP);
    }
  }
  __syncthreads();

  // ---- stage 2: contraction over W from shared memory --------------------------------------
  // tile = (plane gg, S2_JT frequencies j, S2_KT wavenumbers k2, slice ks of the W range); per (j, k2):
  //   P1 = sum A cos, P2 = sum A sin, P3 = sum B cos, P4 = sum B sin
  //   X[+j, k2] = (P1 - P4) - i (P2 + P3),   X[-j, k2] = (P1 + P4) + i (P3 - P2)
  const int nJG = (m1 + S2_JT) / S2_JT;            // ceil((m1 + 1) / JT)
  const int nKG = (m2 + S2_KT - 1) / S2_KT;
  const int ntiles = G * nJG * nKG;
  const int W4 = WP >> 2;
  int KS = nthr / ntiles;                          // W-range slices per tile (0 if ntiles > nthr)
  if (KS > W4) KS = W4;
  // sliced partial sums are staged in the uab region once it is dead: bound KS by its capacity
  while (KS > 1 && (size_t)ntiles * KS * S2_ACC > (size_t)G * NJ * WP) --KS;
  const bool nyq_even = ((W & 1) == 0);

  // partial sums of one tile over the float4 chunks [q0, q1) of the W range
  auto tile_sums = [&](float (&pacc)[S2_ACC], int tile, int q0, int q1) {
    const int kg = tile % nKG;
    const int jg = (tile / nKG) % nJG;
    const int gg = tile / (nKG * nJG);
#pragma unroll
    for (int i = 0; i < S2_ACC; ++i) pacc[i] = 0.f;
    const float4* A4[S2_JT];
    const float4* B4[S2_JT];
#pragma unroll
    for (int a = 0; a < S2_JT; ++a) {
      int j = jg * S2_JT + a;
      if (j > m1) j = m1;                           // clamp: duplicate work, not stored
      A4[a] = reinterpret_cast<const float4*>(uab + ((size_t)gg * NJ + j) * WP);
      B4[a] = reinterpret_cast<const float4*>(uab + ((size_t)gg * NJ + M1T + (j > 0 ? j : 1)) * WP);
    }
    const float4* C4[S2_KT];
    const float4* S4[S2_KT];
#pragma unroll
    for (int b = 0; b < S2_KT; ++b) {
      int k2 = kg * S2_KT + b;
      if (k2 >= m2) k2 = m2 - 1;
      C4[b] = reinterpret_cast<const float4*>(twW_s + (size_t)k2 * WP);
      S4[b] = reinterpret_cast<const float4*>(twW_s + (size_t)(m2 + k2) * WP);
    }
    for (int q = q0; q < q1; ++q) {
      float4 av[S2_JT], bv[S2_JT];
#pragma unroll
      for (int a = 0; a < S2_JT; ++a) { av[a] = A4[a][q]; bv[a] = B4[a][q]; }
#pragma unroll
      for (int b = 0; b < S2_KT; ++b) {
        const float4 c = C4[b][q], sn = S4[b][q];
#pragma unroll
        for (int a = 0; a < S2_JT; ++a) {
          float* p = pacc + (a * S2_KT + b) * 4;
#define FNO_DOT4(P, U, V) P = fmaf(U.x, V.x, P); P = fmaf(U.y, V.y, P); P = fmaf(U.z, V.z, P); P = fmaf(U.w, V.w, P);
          FNO_DOT4(p[0], av[a], c) FNO_DOT4(p[1], av[a], sn) FNO_DOT4(p[2], bv[a], c) FNO_DOT4(p[3], bv[a], sn)
#undef FNO_DOT4
        }
      }
    }
  };
  // stores the two signed-frequency outputs of (tile, pair pi) from p = (P1, P2, P3, P4)
  auto emit_pair = [&](int tile, int pi, float4 p) {
    const int kg = tile % nKG;
    const int jg = (tile / nKG) % nJG;
    const int gg = tile / (nKG * nJG);
    const int a = pi / S2_KT, b = pi - a * S2_KT;
    const int j = jg * S2_JT + a, k2 = kg * S2_KT + b;
    const long pl = (long)blockIdx.x * G + gg;
    if (j > m1 || k2 >= m2 || pl >= planes) return;
    if (j == 0) { p.z = 0.f; p.w = 0.f; }
    float sc = scale;
    if (cmode && k2 != 0 && !(nyq_even && 2 * k2 == W)) sc *= 2.0f;
    float2* Xp = X + (size_t)pl * (2 * m1) * m2;
    if (j < m1) Xp[(size_t)j * m2 + k2] = make_float2((p.x - p.w) * sc, -(p.y + p.z) * sc);
    if (j >= 1) Xp[(size_t)(2 * m1 - j) * m2 + k2] = make_float2((p.x + p.w) * sc, (p.z - p.y) * sc);
  };

  float pacc[S2_ACC];
  if (KS <= 1) {
    // more tiles than spare threads: each thread owns whole tiles
    for (int tile = tid; tile < ntiles; tile += nthr) {
      tile_sums(pacc, tile, 0, W4);
#pragma unroll
      for (int pi = 0; pi < S2_JT * S2_KT; ++pi)
        emit_pair(tile, pi, make_float4(pacc[4 * pi], pacc[4 * pi + 1], pacc[4 * pi + 2], pacc[4 * pi + 3]));
    }
    return;
  }
  // KS threads per tile, each over a slice of the W range; partial sums combined through shared
  // memory (the uab region, dead after the barrier)
  const bool live = tid < ntiles * KS;
  const int ks = live ? tid / ntiles : 0;
  const int tile = live ? tid - ks * ntiles : 0;
  const int cps = (W4 + KS - 1) / KS;
  if (live) {
    const int q0 = ks * cps;
    tile_sums(pacc, tile, q0, (q0 + cps < W4) ? q0 + cps : W4);
  }
  __syncthreads();
  float* red = uab;                                 // [KS][ntiles][S2_ACC]
  if (live) {
    float4* r4 = reinterpret_cast<float4*>(red + (size_t)tid * S2_ACC);
#pragma unroll
    for (int i = 0; i < S2_ACC / 4; ++i)
      r4[i] = make_float4(pacc[4 * i], pacc[4 * i + 1], pacc[4 * i + 2], pacc[4 * i + 3]);
  }
  __syncthreads();
  if (!live) return;
  for (int pi = ks; pi < S2_JT * S2_KT; pi += KS) {
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sl = 0; sl < KS; ++sl) {
      const float4 v = *reinterpret_cast<const float4*>(red + ((size_t)(sl * ntiles + tile) * S2_ACC) + pi * 4);
      p.x += v.x; p.y += v.y; p.z += v.z; p.w += v.w;
    }
    emit_pair(tile, pi, p);
  }
}

// ------------------------------------------------------------------------------------------
// K1, strided-axis half after the tensor-core W-axis stage (transform2d_tc.cu): per plane the
// [H, 2*m2] matrix T1 = [sum_w x cos | sum_w x sin] is folded along H exactly like stage 1 of the
// direct kernel -- thread = (plane, k2) owns the cosine and the sine column of its wavenumber, so
// the signed-frequency outputs need no exchange between threads:
//   T1[h] = Tc - i Ts;  with A^c_j = sum_h Tc cos(j th), B^c_j = sum_h Tc sin(j th) (same for Ts)
//   X[+j, k2] = (A^c - B^s) - i (B^c + A^s),   X[-j, k2] = (A^c + B^s) + i (B^c - A^s)
// ------------------------------------------------------------------------------------------
constexpr int HP_S = 4;   // row slices per (plane, k2): adjacent lanes, combined with two shuffle steps

template <int M1T>
__global__ void __launch_bounds__(384)
hpass2d_kernel(const float* __restrict__ T1, float2* __restrict__ X, const float* __restrict__ twH, int H, int W,
               int m1, int m2, int G, long planes, int cmode, float scale) {
  constexpr int NJ = Geo<M1T>::NJ;
  constexpr int JP = Geo<M1T>::JP;
  const int NP = H / 2 + 1;
  extern __shared__ __align__(16) float smem[];
  float* twH_s = smem;                       // [NP][JP]
  const int tid = threadIdx.x;
  {
    const float4* s4 = reinterpret_cast<const float4*>(twH);
    float4* d4 = reinterpret_cast<float4*>(twH_s);
    for (int i = tid; i < NP * JP / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  }
  __syncthreads();
  // thread = (plane g, wavenumber k2, row slice sl); the HP_S slices of one (g, k2) are adjacent lanes of a warp
  // (blockDim is a multiple of 32 and HP_S divides 32), idle threads of the last warp keep running for the shuffles
  const int sl = tid & (HP_S - 1);
  const int k2r = (tid / HP_S) % m2;
  const int g = tid / (HP_S * m2);
  const long plane = (long)blockIdx.x * G + g;
  const bool live = g < G && plane < planes;
  const int TQ = (2 * m2 + 3) & ~3;          // row pitch of T1 (transform2d_tc.cu)
  const float* __restrict__ tp = T1 + (size_t)(live ? plane : 0) * H * TQ + k2r;
  auto ld = [&](int h) -> Vec<2> {
    Vec<2> v;
    v.v[0] = __ldg(tp + (size_t)h * TQ);
    v.v[1] = __ldg(tp + (size_t)h * TQ + m2);
    return v;
  };
  float acc[2][NJ];
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[n][j] = 0.0f;
  if (live) {
    if (sl == 0) fold_accumulate_even<M1T, 2>(acc, twH_s, ld(0));
    const int npairs = (H - 1) / 2;
    constexpr int PG = 4;
    int t = 1 + sl;                            // this slice's row pairs: 1 + sl, 1 + sl + HP_S, ...
    for (; t + (PG - 1) * HP_S <= npairs; t += PG * HP_S) {
      Vec<2> v1[PG], v2[PG];
#pragma unroll
      for (int u = 0; u < PG; ++u) { v1[u] = ld(t + u * HP_S); v2[u] = ld(H - t - u * HP_S); }
#pragma unroll
      for (int u = 0; u < PG; ++u) {
        Vec<2> e, o;
#pragma unroll
        for (int n = 0; n < 2; ++n) { e.v[n] = v1[u].v[n] + v2[u].v[n]; o.v[n] = v1[u].v[n] - v2[u].v[n]; }
        fold_accumulate<M1T, 2>(acc, twH_s + (t + u * HP_S) * JP, e, o);
      }
    }
    for (; t <= npairs; t += HP_S) {
      const Vec<2> a = ld(t), b = ld(H - t);
      Vec<2> e, o;
#pragma unroll
      for (int n = 0; n < 2; ++n) { e.v[n] = a.v[n] + b.v[n]; o.v[n] = a.v[n] - b.v[n]; }
      fold_accumulate<M1T, 2>(acc, twH_s + t * JP, e, o);
    }
    if ((H & 1) == 0 && sl == HP_S - 1) fold_accumulate_even<M1T, 2>(acc, twH_s + (H / 2) * JP, ld(H / 2));
  }
  // combine the row slices (adjacent lanes)
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float v = acc[n][j];
#pragma unroll
      for (int off = 1; off < HP_S; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[n][j] = v;
    }
  if (!live || sl != 0) return;
  const int k2 = k2r;
  float sc = scale;
  if (cmode && k2 != 0 && !((W & 1) == 0 && 2 * k2 == W)) sc *= 2.0f;
  float2* Xp = X + (size_t)plane * (2 * m1) * m2;
#pragma unroll
  for (int j = 0; j <= M1T; ++j) {
    if (j > m1) break;
    const float Ac = acc[0][j], As = acc[1][j];
    const float Bc = (j > 0) ? acc[0][M1T + j] : 0.f, Bs = (j > 0) ? acc[1][M1T + j] : 0.f;
    if (j < m1) Xp[(size_t)j * m2 + k2] = make_float2((Ac - Bs) * sc, -(Bc + As) * sc);
    if (j >= 1) Xp[(size_t)(2 * m1 - j) * m2 + k2] = make_float2((Ac + Bs) * sc, (Bc - As) * sc);
  }
}

// ------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------
template <int M1T, int TN, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
inv2d_kernel(const float2* __restrict__ Y, const float* addend, float* __restrict__ s_out, float* out,
             const float* __restrict__ twH, const float* __restrict__ twW, int H, int W, int WP, int m1,
             int m2, int G, long planes, int cmode, float scale, int apply_gelu) {
  constexpr int NC = Geo<M1T>::NC;
  constexpr int JP = Geo<M1T>::JP;
  const int NP = H / 2 + 1;
  extern __shared__ __align__(16) float smem[];
  float* twH_s = smem;                                   // [NP][JP]
  float* twW_s = twH_s + NP * JP;                        // [2][m2][WP]
  float4* yp = reinterpret_cast<float4*>(twW_s + 2 * m2 * WP);  // [G][m2][NC]

  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  {
    const float4* s4 = reinterpret_cast<const float4*>(twH);
    float4* d4 = reinterpret_cast<float4*>(twH_s);
    for (int i = tid; i < NP * JP / 4; i += nthr) d4[i] = __ldg(s4 + i);
    s4 = reinterpret_cast<const float4*>(twW);
    d4 = reinterpret_cast<float4*>(twW_s);
    for (int i = tid; i < 2 * m2 * WP / 4; i += nthr) d4[i] = __ldg(s4 + i);
  }
  // stage 0: fold the +j / -j spectrum rows:  Yp = Y[+j] + Y[-j],  Dm = Y[-j] - Y[+j], pre-scaled
  const bool nyq_even = ((W & 1) == 0);
  for (int i = tid; i < G * m2 * NC; i += nthr) {
    const int j = i % NC;
    const int k2 = (i / NC) % m2;
    const int gg = i / (NC * m2);
    const long pl = (long)blockIdx.x * G + gg;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pl < planes && j <= m1) {
      const float2* Yp = Y + (size_t)pl * (2 * m1) * m2;
      float2 yp_ = make_float2(0.f, 0.f), ym_ = make_float2(0.f, 0.f);
      if (j < m1) yp_ = __ldg(Yp + (size_t)j * m2 + k2);
      if (j >= 1) ym_ = __ldg(Yp + (size_t)(2 * m1 - j) * m2 + k2);
      float sc = scale;
      if (cmode && k2 != 0 && !(nyq_even && 2 * k2 == W)) sc *= 2.0f;
      v = make_float4((yp_.x + ym_.x) * sc, (yp_.y + ym_.y) * sc, (ym_.x - yp_.x) * sc, (ym_.y - yp_.y) * sc);
    }
    yp[i] = v;
  }
  __syncthreads();

  const int tpp = W / TN;
  const int g = tid / tpp;
  const int w0 = (tid - g * tpp) * TN;
  const long plane = (long)blockIdx.x * G + g;
  if (!((g < G) && (plane < planes))) return;

  // stage A: W-axis inverse on the retained spectrum -> per-column coefficients in registers
  //   out[h, w] = sum_j Cc[j] cos(2 pi j h / H) + Ss[j] sin(2 pi j h / H)
  float cc[TN][NC], ss[TN][NC];
#pragma unroll
  for (int n = 0; n < TN; ++n)
#pragma unroll
    for (int j = 0; j < NC; ++j) { cc[n][j] = 0.f; ss[n][j] = 0.f; }
  for (int k2 = 0; k2 < m2; ++k2) {
    const Vec<TN> c = Vec<TN>::ld_plain(twW_s + (size_t)k2 * WP + w0);
    const Vec<TN> s = Vec<TN>::ld_plain(twW_s + (size_t)(m2 + k2) * WP + w0);
    const float4* q4 = yp + ((size_t)g * m2 + k2) * NC;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float4 q = q4[j];
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        cc[n][j] = fmaf(q.x, c.v[n], cc[n][j]);
        cc[n][j] = fmaf(-q.y, s.v[n], cc[n][j]);
        if (j > 0) {
          ss[n][j] = fmaf(q.z, s.v[n], ss[n][j]);
          ss[n][j] = fmaf(q.w, c.v[n], ss[n][j]);
        }
      }
    }
  }

  // stage B: H-axis pass, two rows per step, fused epilogue.  Rows are addressed by ELEMENT OFFSETS that are
  // multiples of W kept in 32-bit registers (one IMAD per row shared by the three arrays); the per-access
  // 64-bit `base + h * W` rebuild cost five integer instructions per load / store (ncu r1_e: LEA + IADD3 + MOV
  // were 27 % of this kernel's instructions, which is FP32-issue-bound).
  const size_t base = (size_t)plane * H * W + w0;
  const bool has_add = addend != nullptr;
  const float* ap = has_add ? addend + base : nullptr;
  float* __restrict__ sop = s_out != nullptr ? s_out + base : nullptr;
  float* op = out + base;
  auto ld_add_o = [&](unsigned off) -> Vec<TN> {
    if (has_add) return Vec<TN>::ld_plain(ap + off);
    Vec<TN> z;
#pragma unroll
    for (int n = 0; n < TN; ++n) z.v[n] = 0.f;
    return z;
  };
  auto ld_add = [&](int h) -> Vec<TN> { return ld_add_o((unsigned)h * (unsigned)W); };
  auto emit_o = [&](unsigned off, Vec<TN> v) {
    if (sop != nullptr) v.st(sop + off);
    if (apply_gelu) {
#pragma unroll
      for (int n = 0; n < TN; ++n) v.v[n] = gelu_fast(v.v[n]);
    }
    v.st(op + off);
  };
  auto emit = [&](int h, Vec<TN> v) { emit_o((unsigned)h * (unsigned)W, v); };
  auto eval = [&](int t, Vec<TN>& e, Vec<TN>& o) {
    float tw[JP];
    load_row<JP>(tw, twH_s + (size_t)t * JP);
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      float e0 = 0.f, e1 = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
      for (int j = 0; j <= M1T; ++j) {
        if (j & 1) e1 = fmaf(cc[n][j], tw[j], e1); else e0 = fmaf(cc[n][j], tw[j], e0);
      }
#pragma unroll
      for (int j = 1; j <= M1T; ++j) {
        if (j & 1) o1 = fmaf(ss[n][j], tw[M1T + j], o1); else o0 = fmaf(ss[n][j], tw[M1T + j], o0);
      }
      e.v[n] = e0 + e1;
      o.v[n] = o0 + o1;
    }
  };
  {
    Vec<TN> e, o;
    const Vec<TN> a0 = ld_add(0);
    eval(0, e, o);
#pragma unroll
    for (int n = 0; n < TN; ++n) e.v[n] += a0.v[n];
    emit(0, e);
  }
  const int npairs = (H - 1) / 2;
  // `addend` may alias `out` (in-place epilogue), so the compiler cannot move a load above an
  // earlier store; the loads of a whole group of PG row pairs are therefore issued by hand before
  // any of the group's stores, and the next group's loads before this group's arithmetic.
  constexpr int PG = (TN == 2) ? 2 : 4;
  int t = 1;
  const unsigned uW = (unsigned)W;
  unsigned ou = uW, od = (unsigned)(H - 1) * uW;      // element offsets of rows t and H - t
  Vec<TN> a1[PG], a2[PG];
  if (t + PG - 1 <= npairs) {
#pragma unroll
    for (int u = 0; u < PG; ++u) { a1[u] = ld_add_o(ou + u * uW); a2[u] = ld_add_o(od - u * uW); }
  }
  for (; t + PG - 1 <= npairs; t += PG, ou += PG * uW, od -= PG * uW) {
    Vec<TN> n1[PG], n2[PG];
    const bool more = (t + 2 * PG - 1 <= npairs);
    if (has_add && t + PF_DIST + PG - 1 <= npairs) {
#pragma unroll
      for (int u = 0; u < PG; ++u) {
        prefetch_l2(ap + (ou + (PF_DIST + u) * uW));
        prefetch_l2(ap + (od - (PF_DIST + u) * uW));
      }
    }
#pragma unroll
    for (int u = 0; u < PG; ++u) {
      if (more) { n1[u] = ld_add_o(ou + (PG + u) * uW); n2[u] = ld_add_o(od - (PG + u) * uW); }
      else { n1[u] = a1[u]; n2[u] = a2[u]; }
    }
#pragma unroll
    for (int u = 0; u < PG; ++u) {
      Vec<TN> e, o, up, dn;
      eval(t + u, e, o);
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        up.v[n] = e.v[n] + o.v[n] + a1[u].v[n];
        dn.v[n] = e.v[n] - o.v[n] + a2[u].v[n];
      }
      emit_o(ou + u * uW, up);
      emit_o(od - u * uW, dn);
    }
#pragma unroll
    for (int u = 0; u < PG; ++u) { a1[u] = n1[u]; a2[u] = n2[u]; }
  }
  for (; t <= npairs; ++t) {
    Vec<TN> e, o, up, dn;
    const Vec<TN> b1 = ld_add(t), b2 = ld_add(H - t);
    eval(t, e, o);
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      up.v[n] = e.v[n] + o.v[n] + b1.v[n];
      dn.v[n] = e.v[n] - o.v[n] + b2.v[n];
    }
    emit(t, up);
    emit(H - t, dn);
  }
  if ((H & 1) == 0) {
    Vec<TN> e, o;
    const Vec<TN> b1 = ld_add(H / 2);
    eval(H / 2, e, o);
#pragma unroll
    for (int n = 0; n < TN; ++n) e.v[n] += b1.v[n];
    emit(H / 2, e);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
constexpr size_t kMaxOptinSmemC = 227 * 1024;
size_t fwd_smem_bytes(const Plan* p, int G) {
  const int NJ = 2 * p->M1T + 1;
  return sizeof(float) * ((size_t)p->NP * p->JP + 2ul * p->m2 * p->WP + (size_t)G * NJ * p->WP);
}
size_t inv_smem_bytes(const Plan* p, int G) {
  return sizeof(float) * ((size_t)p->NP * p->JP + 2ul * p->m2 * p->WP) +
         sizeof(float4) * (size_t)G * p->m2 * (p->M1T + 1);
}

// columns per thread: two when the width is even and the accumulators still fit the register file
inline int pick_tn(const Plan* p) { return ((p->W & 1) == 0 && p->M1T <= 16) ? 2 : 1; }

template <int M1T, int TN, int MAXT, int MINB>
int launch_fwd_t(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, long planes,
                 int cmode, float scale, cudaStream_t st, int threads, size_t smem, bool attr_only, int G_launch) {
  auto k0 = fwd2d_kernel<M1T, TN, false, MAXT, MINB>;
  auto k1 = fwd2d_kernel<M1T, TN, true, MAXT, MINB>;
  if (attr_only) {
    if (cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(fwd2d)");
    return FNO_OK;
  }
  if (MINB == 3) {   // experimental class: opt in to large dynamic shared memory on first use
    static PerDeviceOnce done3;
    if (done3.need()) {
      cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxOptinSmemC);
      cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxOptinSmemC);
      done3.mark();
    }
  }
  const int G = G_launch;
  const unsigned grid = (unsigned)((planes + G - 1) / G);
  if (preact != nullptr)
    k1<<<grid, threads, smem, st>>>(x, preact, ds_out, reinterpret_cast<float2*>(X), p->twH, p->twW, p->H,
                                    p->W, p->WP, p->m1, p->m2, G, planes, cmode, scale);
  else
    k0<<<grid, threads, smem, st>>>(x, nullptr, nullptr, reinterpret_cast<float2*>(X), p->twH, p->twW, p->H,
                                    p->W, p->WP, p->m1, p->m2, G, planes, cmode, scale);
  count_launch();
  return check_launch("fwd2d_kernel");
}

template <int M1T, int TN, int MAXT, int MINB>
int launch_inv_t(const Plan* p, const float* Y, const float* addend, float* s_out, float* out, long planes,
                 int cmode, float scale, int apply_gelu, cudaStream_t st, int threads, size_t smem,
                 bool attr_only, int G_launch) {
  auto k = inv2d_kernel<M1T, TN, MAXT, MINB>;
  if (attr_only) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(inv2d)");
    return FNO_OK;
  }
  if (MINB == 3) {
    static PerDeviceOnce done3;
    if (done3.need()) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxOptinSmemC);
      done3.mark();
    }
  }
  const int G = G_launch;
  const unsigned grid = (unsigned)((planes + G - 1) / G);
  k<<<grid, threads, smem, st>>>(reinterpret_cast<const float2*>(Y), addend, s_out, out, p->twH, p->twW, p->H,
                                 p->W, p->WP, p->m1, p->m2, G, planes, cmode, scale, apply_gelu);
  count_launch();
  return check_launch("inv2d_kernel");
}

inline int round_threads(int n) { return (n + 31) & ~31; }
constexpr size_t kMaxOptinSmem = 227 * 1024;  // sm_100: 232448 B opt-in per CTA
constexpr int kSMs = 148;
// 80-register builds for 3 resident CTAs of <= 224 threads.  Measured at cfg 1 / batch 128: the inverse
// transform gains (115 -> 101 us, 141 -> 136 us with GELU), the forward transform loses to spills
// (101 -> 105 us, 174 -> 212 us with GELU'), so only K3 uses it (FNO_T3=1 forces it for K1 as well).
static const bool kThreeCtasFwd = [] { const char* e = std::getenv("FNO_T3"); return e != nullptr && e[0] == '1'; }();
static const bool kThreeCtasInv = [] { const char* e = std::getenv("FNO_T3"); return e == nullptr || e[0] != '0'; }();

// Planes per CTA for this launch.  The plan's G maximises lane efficiency; with `planes` known the
// choice also has to avoid a nearly empty last wave: a launch of n CTAs on 148 * cps resident slots
// takes ceil(n / slots) rounds of (cps * G) planes per SM, e.g. 2560 planes (cfg 1, batch 128):
// G = 4 -> 2 CTAs/SM, 640 CTAs = 2.16 waves = 3 rounds of 8 planes; G = 3 -> 3 CTAs/SM, 854 CTAs
// = 1.92 waves = 2 rounds of 9 planes.  Only the <= 288-thread class (96 registers) is searched.
inline int pick_planes_per_cta(const Plan* p, long planes, bool fwd, int tpp, int G_plan) {
  if (round_threads(G_plan * tpp) > 288) return G_plan;
  int best = G_plan;
  double best_cost = 1e30;
  for (int G = 1; G <= 16; ++G) {
    const int thr = round_threads(G * tpp);
    if (thr > 288) break;
    const size_t sm = (fwd ? fwd_smem_bytes(p, G) : inv_smem_bytes(p, G)) + 1024;
    if (sm > 113 * 1024) break;
    long cps = 65536 / (96 * thr);
    if ((long)(kMaxOptinSmem / sm) < cps) cps = (long)(kMaxOptinSmem / sm);
    if (2048 / thr < cps) cps = 2048 / thr;
    if (cps < 1) cps = 1;
    const long ctas = (planes + G - 1) / G;
    const long rounds = (ctas + kSMs * cps - 1) / (kSMs * cps);
    // time ~ rounds x (lanes issued per SM per round); idle lanes of a ragged CTA still cost issue slots
    const double cost = (double)rounds * cps * thr * (cps * thr >= 512 ? 1.0 : 1.25);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = G; }
  }
  return best;
}

// launch-bound classes: <= 288 threads with 2 resident CTAs (<= 112 registers), <= 448 threads
// alone on the SM (<= 144 registers), wider CTAs (very wide planes) at 64 registers
template <int M1T, int TN>
int dispatch_fwd_tn(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, long planes,
                    int cmode, float scale, cudaStream_t st, bool attr_only) {
  const int tpp = p->W / TN;
  const int G = attr_only ? p->G_fwd : pick_planes_per_cta(p, planes, true, tpp, p->G_fwd);
  const int threads = round_threads(G * tpp);
  // attr_only: opt every instantiation this plan can reach into the device maximum once; the
  // limit is per kernel function, so it must never be lowered by a later, smaller plan
  const size_t smem = attr_only ? kMaxOptinSmem : fwd_smem_bytes(p, G);
  if (threads <= 224 && !attr_only && kThreeCtasFwd)
    return launch_fwd_t<M1T, TN, 256, 3>(p, x, preact, ds_out, X, planes, cmode, scale, st, threads, smem, attr_only, G);
  if (threads <= 288)
    return launch_fwd_t<M1T, TN, 288, 2>(p, x, preact, ds_out, X, planes, cmode, scale, st, threads, smem, attr_only, G);
  if (threads <= 448)
    return launch_fwd_t<M1T, TN, 448, 1>(p, x, preact, ds_out, X, planes, cmode, scale, st, threads, smem, attr_only, G);
  return launch_fwd_t<M1T, TN, 1024, 1>(p, x, preact, ds_out, X, planes, cmode, scale, st, threads, smem, attr_only, G);
}

template <int M1T, int TN>
int dispatch_inv_tn(const Plan* p, const float* Y, const float* addend, float* s_out, float* out, long planes,
                    int cmode, float scale, int apply_gelu, cudaStream_t st, bool attr_only) {
  const int tpp = p->W / TN;
  const int G = attr_only ? p->G_inv : pick_planes_per_cta(p, planes, false, tpp, p->G_inv);
  const int threads = round_threads(G * tpp);
  const size_t smem = attr_only ? kMaxOptinSmem : inv_smem_bytes(p, G);
  if (threads <= 224 && !attr_only && kThreeCtasInv)
    return launch_inv_t<M1T, TN, 256, 3>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, threads, smem, attr_only, G);
  if (threads <= 288)
    return launch_inv_t<M1T, TN, 288, 2>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, threads, smem, attr_only, G);
  if (threads <= 448)
    return launch_inv_t<M1T, TN, 448, 1>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, threads, smem, attr_only, G);
  return launch_inv_t<M1T, TN, 1024, 1>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, threads, smem, attr_only, G);
}

template <int M1T>
int dispatch_fwd(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, long planes,
                 int cmode, float scale, cudaStream_t st, bool attr_only) {
  if constexpr (M1T <= 16) {
    if (pick_tn(p) == 2)
      return dispatch_fwd_tn<M1T, 2>(p, x, preact, ds_out, X, planes, cmode, scale, st, attr_only);
  }
  return dispatch_fwd_tn<M1T, 1>(p, x, preact, ds_out, X, planes, cmode, scale, st, attr_only);
}

template <int M1T>
int dispatch_inv(const Plan* p, const float* Y, const float* addend, float* s_out, float* out, long planes,
                 int cmode, float scale, int apply_gelu, cudaStream_t st, bool attr_only) {
  if constexpr (M1T <= 16) {
    if (pick_tn(p) == 2)
      return dispatch_inv_tn<M1T, 2>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, attr_only);
  }
  return dispatch_inv_tn<M1T, 1>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, attr_only);
}

}  // namespace

#define FNO_DISPATCH_M1T(FN, ...)                         \
  switch (p->M1T) {                                       \
    case 4: return FN<4>(__VA_ARGS__);                    \
    case 8: return FN<8>(__VA_ARGS__);                    \
    case 12: return FN<12>(__VA_ARGS__);                  \
    case 16: return FN<16>(__VA_ARGS__);                  \
    case 24: return FN<24>(__VA_ARGS__);                  \
    case 32: return FN<32>(__VA_ARGS__);                  \
    default: set_error("unsupported padded modes1 %d", p->M1T); return FNO_E_ARG; \
  }

static bool misaligned8(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
  return ((reinterpret_cast<size_t>(a) | reinterpret_cast<size_t>(b) | reinterpret_cast<size_t>(c) |
           reinterpret_cast<size_t>(d)) & 7) != 0;
}

int launch_fwd2d(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, long planes,
                 int cmode, float scale, cudaStream_t st) {
  if (misaligned8(x, preact, ds_out, X)) { set_error("fwd_transform: tensors must be 8-byte aligned"); return FNO_E_ARG; }
  FNO_DISPATCH_M1T(dispatch_fwd, p, x, preact, ds_out, X, planes, cmode, scale, st, false)
}

template <int M1T>
static int launch_hpass_t(const Plan* p, const float* T1, float* X, long planes, int cmode, float scale, cudaStream_t st) {
  int G = 384 / (p->m2 * HP_S);
  if (G > 2) G = 2;                                   // ~100-thread CTAs, > 1000 of them at a bench-sized batch (measured:
                                                      // 8 planes per CTA, which amortises the 7 KB twiddle table, is slower)
  if (G < 1) { set_error("hpass2d: modes2 %d too large", p->m2); return FNO_E_ARG; }
  const int threads = round_threads(G * p->m2 * HP_S);
  const size_t smem = sizeof(float) * (size_t)p->NP * p->JP;
  const unsigned grid = (unsigned)((planes + G - 1) / G);
  hpass2d_kernel<M1T><<<grid, threads, smem, st>>>(T1, reinterpret_cast<float2*>(X), p->twH, p->H, p->W, p->m1, p->m2, G,
                                                   planes, cmode, scale);
  count_launch();
  return check_launch("hpass2d_kernel");
}

// K1 with a caller-provided workspace: W-axis stage on the tensor cores + H-axis fold on the FP32 pipes
// when the geometry allows (even W <= 136, 2*m2 <= 32) and the batch is large enough to fill the
// persistent grid; otherwise the direct kernel.
int launch_fwd2d_ws(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, float* work,
                    long planes, int cmode, float scale, cudaStream_t st) {
  const bool aligned = ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(preact) | reinterpret_cast<size_t>(ds_out)) & 7) == 0 &&
                       (reinterpret_cast<size_t>(work) & 15) == 0;
  if (work == nullptr || p->tc_nch == 0 || !aligned || planes * p->H < 4096 ||
      sizeof(float) * (size_t)p->NP * p->JP > 48 * 1024)
    return launch_fwd2d(p, x, preact, ds_out, X, planes, cmode, scale, st);
  // A operand through TMEM (transform2d_tc.cu): 36 + 22 us against 101 us for the FP32 kernel at cfg 1; with the GELU'
  // premultiply 108 + 22 against 174 us.  FNO_K1P_TC=0 keeps the FP32 kernel for the premultiply form only.
  static const bool premul_tc = [] { const char* e = std::getenv("FNO_K1P_TC"); return e == nullptr || e[0] != '0'; }();
  int rc = (preact == nullptr) ? launch_fwd2d_tca(p, x, work, planes, st, false)
                               : (premul_tc ? launch_fwd2d_tcap(p, x, preact, ds_out, work, planes, st, false) : 1);
  if (rc == 1) return launch_fwd2d(p, x, preact, ds_out, X, planes, cmode, scale, st);   // outside the tensor-core envelope
  if (rc != FNO_OK) return rc;
  switch (p->M1T) {
    case 4: return launch_hpass_t<4>(p, work, X, planes, cmode, scale, st);
    case 8: return launch_hpass_t<8>(p, work, X, planes, cmode, scale, st);
    case 12: return launch_hpass_t<12>(p, work, X, planes, cmode, scale, st);
    case 16: return launch_hpass_t<16>(p, work, X, planes, cmode, scale, st);
    case 24: return launch_hpass_t<24>(p, work, X, planes, cmode, scale, st);
    case 32: return launch_hpass_t<32>(p, work, X, planes, cmode, scale, st);
    default: set_error("unsupported padded modes1 %d", p->M1T); return FNO_E_ARG;
  }
}

int launch_inv2d(const Plan* p, const float* Y, const float* addend, float* s_out, float* out, long planes,
                 int cmode, float scale, int apply_gelu, cudaStream_t st) {
  if (misaligned8(Y, addend, s_out, out)) { set_error("inv_transform: tensors must be 8-byte aligned"); return FNO_E_ARG; }
  FNO_DISPATCH_M1T(dispatch_inv, p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, false)
}

static int setup_fwd_attr(const Plan* p) {
  FNO_DISPATCH_M1T(dispatch_fwd, p, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f, nullptr, true)
}
static int setup_inv_attr(const Plan* p) {
  FNO_DISPATCH_M1T(dispatch_inv, p, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f, 0, nullptr, true)
}

// Picks planes-per-CTA and raises the dynamic shared-memory limit of the instantiations this
// plan will launch.  Called once from plan creation (not capture-time).
int setup_transform2d_attrs(const Plan* pc) {
  Plan* p = const_cast<Plan*>(pc);
  const int tpp = p->W / pick_tn(p);
  if (round_threads(tpp) > 1024) { set_error("W = %d too large (max %d)", p->W, 1024 * pick_tn(p)); return FNO_E_ARG; }
  auto pick = [&](bool fwd) {
    // planes per CTA: best lane efficiency with <= 288 threads (2 resident CTAs per SM); wider
    // CTAs only if a narrow one would idle more than 15 % of its lanes.  The shared-memory budget
    // keeps two CTAs resident (<= 110 KB each) whenever a single plane allows it.
    int best = 1;
    double best_eff = 0.0;
    for (int limit : {288, 448}) {
      const size_t budget = (limit == 288) ? 110 * 1024 : 200 * 1024;
      for (int G = 1; G <= 32; ++G) {
        const int thr = round_threads(G * tpp);
        if (thr > limit) break;
        const size_t sm = fwd ? fwd_smem_bytes(p, G) : inv_smem_bytes(p, G);
        if (sm > budget && G > 1) break;
        const double eff = (double)(G * tpp) / thr;
        if (eff > best_eff + 0.02) { best_eff = eff; best = G; }
      }
      if (best_eff >= 0.85) break;
    }
    const char* env = std::getenv(fwd ? "FNO_G_FWD" : "FNO_G_INV");  // tuning override
    if (env != nullptr && std::atoi(env) > 0) {
      const int G = std::atoi(env);
      const size_t sm = fwd ? fwd_smem_bytes(p, G) : inv_smem_bytes(p, G);
      if (round_threads(G * tpp) <= 1024 && sm <= 200 * 1024) best = G;
    }
    return best;
  };
  p->G_fwd = pick(true);
  p->G_inv = pick(false);
  if (fwd_smem_bytes(p, p->G_fwd) > kMaxOptinSmem || inv_smem_bytes(p, p->G_inv) > kMaxOptinSmem) {
    set_error("plane %dx%d with modes (%d,%d) needs too much shared memory", p->H, p->W, p->m1, p->m2);
    return FNO_E_ARG;
  }
  int rc = setup_fwd_attr(p);
  if (rc != FNO_OK) return rc;
  return setup_inv_attr(p);
}

}  // namespace fno
