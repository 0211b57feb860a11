// K1 (pruned forward transform) and K3 (zero-padding inverse transform + fused epilogue) for one
// 2-D plane [H, W] -- FP32 CUDA-core path.
//
// Both kernels share one factorisation (oracle/dft_oracle.py):
//
//   strided axis H ("column pass"):  one thread per (plane, w) column.  The 2*m1 kept signed
//   frequencies {-m1..m1-1} are folded onto |k| = j in [0, m1]:  with
//       A[j] = sum_h x[h] cos(2 pi j h / H),   B[j] = sum_h x[h] sin(2 pi j h / H)
//   the column DFT is U[+j] = A[j] - i B[j], U[-j] = A[j] + i B[j]; rows are further folded in
//   pairs (h, H-h): e = x[h] + x[H-h] feeds the cosines, o = x[h] - x[H-h] the sines.  That is
//   2*m1+1 real accumulators per column and (2*m1+1) FMAs per row *pair*, the row twiddles being
//   warp-uniform float4 broadcasts from shared memory.  x is read straight from global memory,
//   exactly once, fully coalesced (lanes <-> consecutive w).
//
//   contiguous axis W: the small [2*m1+1, W] x [W, m2] contraction runs from shared memory.
//
// K3 mirrors it: the W-axis inverse is applied first to the tiny retained spectrum, leaving
// 2*m1+1 real coefficients per column in registers; the H-axis pass then emits two output rows
// (h, H-h) per (2*m1+1) FMAs and fuses "+ addend" (1x1-conv bypass), the optional store of the
// pre-activation, and the exact-erf GELU.
//
// Several planes are flattened into one CTA (thread <-> (g, w), g < G) so that widths like
// 130 = 4*32+2 do not waste a nearly empty warp per plane.
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

// acc[0..M1T] += e * cos row, acc[M1T+1..2*M1T] += o * sin row  (row = JP floats, 16-B aligned)
template <int M1T>
__device__ __forceinline__ void fold_accumulate(float (&acc)[Geo<M1T>::NJ], const float* __restrict__ row,
                                                float e, float o) {
  constexpr int JP = Geo<M1T>::JP;
  float tw[JP];
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int q = 0; q < JP / 4; ++q) {
    const float4 v = r4[q];
    tw[4 * q + 0] = v.x;
    tw[4 * q + 1] = v.y;
    tw[4 * q + 2] = v.z;
    tw[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int j = 0; j <= M1T; ++j) acc[j] = fmaf(e, tw[j], acc[j]);
#pragma unroll
  for (int j = 1; j <= M1T; ++j) acc[M1T + j] = fmaf(o, tw[M1T + j], acc[M1T + j]);
}

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
template <int M1T, bool PREMUL, int MAXT, int MINB>
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
  float* uab = twW_s + 2 * m2 * WP;          // [G][NJ][WP]

  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  for (int i = tid; i < NP * JP; i += nthr) twH_s[i] = twH[i];
  for (int i = tid; i < 2 * m2 * WP; i += nthr) twW_s[i] = twW[i];
  // zero the pad columns of uab (read by the float4 loops of stage 2)
  const int padw = WP - W;
  if (padw > 0) {
    for (int i = tid; i < G * NJ * padw; i += nthr) {
      const int r = i / padw;
      uab[r * WP + W + (i - r * padw)] = 0.0f;
    }
  }
  __syncthreads();

  // ---- column pass over H ------------------------------------------------------------------
  const int g = tid / W;
  const int w = tid - g * W;
  const long plane = (long)blockIdx.x * G + g;
  const bool active = (g < G) && (plane < planes);
  if (active) {
    float acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = 0.0f;
    const size_t base = (size_t)plane * H * W + w;
    const float* __restrict__ xp = x + base;
    auto load = [&](int h) -> float {
      float v = __ldg(xp + (size_t)h * W);
      if (PREMUL) {
        v *= gelu_exact_grad(__ldg(preact + base + (size_t)h * W));
        if (ds_out != nullptr) ds_out[base + (size_t)h * W] = v;
      }
      return v;
    };
    // h = 0 (self-paired)
    fold_accumulate<M1T>(acc, twH_s, load(0), 0.0f);
    const int npairs = (H - 1) / 2;
    // PG row pairs per step: all 2*PG global loads are issued before the FMAs that consume them,
    // which is what keeps enough bytes in flight per SM (the kernel has no other latency hiding
    // than its own memory-level parallelism and the other resident warps)
    constexpr int PG = 4;
    int t = 1;
    for (; t + PG - 1 <= npairs; t += PG) {
      float v1[PG], v2[PG];
#pragma unroll
      for (int u = 0; u < PG; ++u) {
        v1[u] = load(t + u);
        v2[u] = load(H - t - u);
      }
#pragma unroll
      for (int u = 0; u < PG; ++u) fold_accumulate<M1T>(acc, twH_s + (t + u) * JP, v1[u] + v2[u], v1[u] - v2[u]);
    }
    for (; t <= npairs; ++t) {
      const float v1 = load(t);
      const float v2 = load(H - t);
      fold_accumulate<M1T>(acc, twH_s + t * JP, v1 + v2, v1 - v2);
    }
    if ((H & 1) == 0) fold_accumulate<M1T>(acc, twH_s + (H / 2) * JP, load(H / 2), 0.0f);
    float* urow = uab + (size_t)g * NJ * WP + w;
#pragma unroll
    for (int j = 0; j < NJ; ++j) urow[j * WP] = acc[j];
  } else if (g < G) {
    float* urow = uab + (size_t)g * NJ * WP + w;
#pragma unroll
    for (int j = 0; j < NJ; ++j) urow[j * WP] = 0.0f;
  }
  __syncthreads();

  // ---- contraction over W from shared memory -------------------------------------------------
  // item = (g, j, pair of k2): four real sums per k2
  //   P1 = sum A cos, P2 = sum A sin, P3 = sum B cos, P4 = sum B sin
  //   X[+j, k2] = (P1 - P4) - i (P2 + P3),   X[-j, k2] = (P1 + P4) + i (P3 - P2)
  const int K2P = (m2 + 1) >> 1;
  const int nitems = G * (m1 + 1) * K2P;
  const int W4 = WP >> 2;
  const bool nyq_even = ((W & 1) == 0);
  for (int item = tid; item < nitems; item += nthr) {
    const int kp = item % K2P;
    const int j = (item / K2P) % (m1 + 1);
    const int gg = item / (K2P * (m1 + 1));
    const long pl = (long)blockIdx.x * G + gg;
    if (pl >= planes) continue;
    const int k2a = 2 * kp;
    const int k2b = (k2a + 1 < m2) ? k2a + 1 : k2a;  // clamp: duplicate work, not stored
    const float4* A4 = reinterpret_cast<const float4*>(uab + ((size_t)gg * NJ + j) * WP);
    const float4* B4 = reinterpret_cast<const float4*>(uab + ((size_t)gg * NJ + M1T + (j > 0 ? j : 1)) * WP);
    const float bmask = (j > 0) ? 1.0f : 0.0f;
    const float4* Ca = reinterpret_cast<const float4*>(twW_s + (size_t)k2a * WP);
    const float4* Sa = reinterpret_cast<const float4*>(twW_s + (size_t)(m2 + k2a) * WP);
    const float4* Cb = reinterpret_cast<const float4*>(twW_s + (size_t)k2b * WP);
    const float4* Sb = reinterpret_cast<const float4*>(twW_s + (size_t)(m2 + k2b) * WP);
    float p1a = 0.f, p2a = 0.f, p3a = 0.f, p4a = 0.f, p1b = 0.f, p2b = 0.f, p3b = 0.f, p4b = 0.f;
#pragma unroll 2
    for (int q = 0; q < W4; ++q) {
      const float4 a = A4[q];
      const float4 b = B4[q];
      const float4 ca = Ca[q], sa = Sa[q], cb = Cb[q], sb = Sb[q];
#define FNO_DOT4(P, U, V)      \
  P = fmaf(U.x, V.x, P);       \
  P = fmaf(U.y, V.y, P);       \
  P = fmaf(U.z, V.z, P);       \
  P = fmaf(U.w, V.w, P);
      FNO_DOT4(p1a, a, ca) FNO_DOT4(p2a, a, sa) FNO_DOT4(p3a, b, ca) FNO_DOT4(p4a, b, sa)
      FNO_DOT4(p1b, a, cb) FNO_DOT4(p2b, a, sb) FNO_DOT4(p3b, b, cb) FNO_DOT4(p4b, b, sb)
#undef FNO_DOT4
    }
    p3a *= bmask; p4a *= bmask; p3b *= bmask; p4b *= bmask;
    float2* Xp = X + (size_t)pl * (2 * m1) * m2;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int k2 = s ? k2a + 1 : k2a;
      if (k2 >= m2) break;
      const float p1 = s ? p1b : p1a, p2 = s ? p2b : p2a, p3 = s ? p3b : p3a, p4 = s ? p4b : p4a;
      float sc = scale;
      if (cmode && k2 != 0 && !(nyq_even && 2 * k2 == W)) sc *= 2.0f;
      if (j < m1) Xp[(size_t)j * m2 + k2] = make_float2((p1 - p4) * sc, -(p2 + p3) * sc);
      if (j >= 1) Xp[(size_t)(2 * m1 - j) * m2 + k2] = make_float2((p1 + p4) * sc, (p3 - p2) * sc);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------
template <int M1T, int MAXT, int MINB>
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
  for (int i = tid; i < NP * JP; i += nthr) twH_s[i] = twH[i];
  for (int i = tid; i < 2 * m2 * WP; i += nthr) twW_s[i] = twW[i];
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
      if (j < m1) yp_ = Yp[(size_t)j * m2 + k2];
      if (j >= 1) ym_ = Yp[(size_t)(2 * m1 - j) * m2 + k2];
      float sc = scale;
      if (cmode && k2 != 0 && !(nyq_even && 2 * k2 == W)) sc *= 2.0f;
      v = make_float4((yp_.x + ym_.x) * sc, (yp_.y + ym_.y) * sc, (ym_.x - yp_.x) * sc, (ym_.y - yp_.y) * sc);
    }
    yp[i] = v;
  }
  __syncthreads();

  const int g = tid / W;
  const int w = tid - g * W;
  const long plane = (long)blockIdx.x * G + g;
  if (!((g < G) && (plane < planes))) return;

  // stage A: W-axis inverse on the retained spectrum -> per-column coefficients in registers
  //   out[h, w] = sum_j Cc[j] cos(2 pi j h / H) + Ss[j] sin(2 pi j h / H)
  float cc[NC], ss[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) { cc[j] = 0.f; ss[j] = 0.f; }
  for (int k2 = 0; k2 < m2; ++k2) {
    const float c = twW_s[(size_t)k2 * WP + w];
    const float s = twW_s[(size_t)(m2 + k2) * WP + w];
    const float4* q4 = yp + ((size_t)g * m2 + k2) * NC;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float4 q = q4[j];
      cc[j] = fmaf(q.x, c, cc[j]);
      cc[j] = fmaf(-q.y, s, cc[j]);
      if (j > 0) {
        ss[j] = fmaf(q.z, s, ss[j]);
        ss[j] = fmaf(q.w, c, ss[j]);
      }
    }
  }

  // stage B: H-axis pass, two rows per step, fused epilogue
  const size_t base = (size_t)plane * H * W + w;
  const bool has_add = addend != nullptr;
  auto ld_add = [&](int h) -> float { return has_add ? addend[base + (size_t)h * W] : 0.f; };
  auto emit = [&](int h, float v) {
    const size_t idx = base + (size_t)h * W;
    if (s_out != nullptr) s_out[idx] = v;
    if (apply_gelu) v = gelu_exact(v);
    out[idx] = v;
  };
  auto eval = [&](int t, float& e, float& o) {
    const float4* r4 = reinterpret_cast<const float4*>(twH_s + (size_t)t * JP);
    float tw[JP];
#pragma unroll
    for (int q = 0; q < JP / 4; ++q) {
      const float4 v = r4[q];
      tw[4 * q + 0] = v.x; tw[4 * q + 1] = v.y; tw[4 * q + 2] = v.z; tw[4 * q + 3] = v.w;
    }
    float e0 = 0.f, e1 = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int j = 0; j <= M1T; ++j) {
      if (j & 1) e1 = fmaf(cc[j], tw[j], e1); else e0 = fmaf(cc[j], tw[j], e0);
    }
#pragma unroll
    for (int j = 1; j <= M1T; ++j) {
      if (j & 1) o1 = fmaf(ss[j], tw[M1T + j], o1); else o0 = fmaf(ss[j], tw[M1T + j], o0);
    }
    e = e0 + e1;
    o = o0 + o1;
  };
  {
    float e, o;
    const float a0 = ld_add(0);
    eval(0, e, o);
    emit(0, e + a0);
  }
  const int npairs = (H - 1) / 2;
  // `addend` may alias `out` (in-place epilogue), so the compiler cannot move a load above an
  // earlier store; the loads of a whole group of PG row pairs are therefore issued by hand before
  // any of the group's stores, and the next group's loads before this group's arithmetic.
  constexpr int PG = 4;
  int t = 1;
  float a1[PG], a2[PG];
  if (t + PG - 1 <= npairs) {
#pragma unroll
    for (int u = 0; u < PG; ++u) { a1[u] = ld_add(t + u); a2[u] = ld_add(H - t - u); }
  }
  for (; t + PG - 1 <= npairs; t += PG) {
    float n1[PG], n2[PG];
    const bool more = (t + 2 * PG - 1 <= npairs);
#pragma unroll
    for (int u = 0; u < PG; ++u) {
      n1[u] = more ? ld_add(t + PG + u) : 0.f;
      n2[u] = more ? ld_add(H - t - PG - u) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PG; ++u) {
      float e, o;
      eval(t + u, e, o);
      emit(t + u, e + o + a1[u]);
      emit(H - t - u, e - o + a2[u]);
    }
#pragma unroll
    for (int u = 0; u < PG; ++u) { a1[u] = n1[u]; a2[u] = n2[u]; }
  }
  for (; t <= npairs; ++t) {
    float e, o;
    const float b1 = ld_add(t), b2 = ld_add(H - t);
    eval(t, e, o);
    emit(t, e + o + b1);
    emit(H - t, e - o + b2);
  }
  if ((H & 1) == 0) {
    float e, o;
    const float b1 = ld_add(H / 2);
    eval(H / 2, e, o);
    emit(H / 2, e + b1);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
size_t fwd_smem_bytes(const Plan* p, int G) {
  const int NJ = 2 * p->M1T + 1;
  return sizeof(float) * ((size_t)p->NP * p->JP + 2ul * p->m2 * p->WP + (size_t)G * NJ * p->WP);
}
size_t inv_smem_bytes(const Plan* p, int G) {
  return sizeof(float) * ((size_t)p->NP * p->JP + 2ul * p->m2 * p->WP) +
         sizeof(float4) * (size_t)G * p->m2 * (p->M1T + 1);
}

template <int M1T, int MAXT, int MINB>
int launch_fwd_t(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, long planes,
                 int cmode, float scale, cudaStream_t st, int threads, size_t smem, bool attr_only) {
  auto k0 = fwd2d_kernel<M1T, false, MAXT, MINB>;
  auto k1 = fwd2d_kernel<M1T, true, MAXT, MINB>;
  if (attr_only) {
    if (cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(fwd2d)");
    return FNO_OK;
  }
  const int G = p->G_fwd;
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

template <int M1T, int MAXT, int MINB>
int launch_inv_t(const Plan* p, const float* Y, const float* addend, float* s_out, float* out, long planes,
                 int cmode, float scale, int apply_gelu, cudaStream_t st, int threads, size_t smem,
                 bool attr_only) {
  auto k = inv2d_kernel<M1T, MAXT, MINB>;
  if (attr_only) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(inv2d)");
    return FNO_OK;
  }
  const int G = p->G_inv;
  const unsigned grid = (unsigned)((planes + G - 1) / G);
  k<<<grid, threads, smem, st>>>(reinterpret_cast<const float2*>(Y), addend, s_out, out, p->twH, p->twW, p->H,
                                 p->W, p->WP, p->m1, p->m2, G, planes, cmode, scale, apply_gelu);
  count_launch();
  return check_launch("inv2d_kernel");
}

inline int round_threads(int n) { return (n + 31) & ~31; }
constexpr size_t kMaxOptinSmem = 227 * 1024;  // sm_100: 232448 B opt-in per CTA

template <int M1T>
int dispatch_fwd(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, long planes,
                 int cmode, float scale, cudaStream_t st, bool attr_only) {
  const int threads = round_threads(p->G_fwd * p->W);
  // attr_only: opt every instantiation this plan can reach into the device maximum once; the
  // limit is per kernel function, so it must never be lowered by a later, smaller plan
  const size_t smem = attr_only ? kMaxOptinSmem : fwd_smem_bytes(p, p->G_fwd);
  if (threads <= 288)
    return launch_fwd_t<M1T, 288, 3>(p, x, preact, ds_out, X, planes, cmode, scale, st, threads, smem, attr_only);
  if (threads <= 576)
    return launch_fwd_t<M1T, 576, 1>(p, x, preact, ds_out, X, planes, cmode, scale, st, threads, smem, attr_only);
  return launch_fwd_t<M1T, 1024, 1>(p, x, preact, ds_out, X, planes, cmode, scale, st, threads, smem, attr_only);
}

template <int M1T>
int dispatch_inv(const Plan* p, const float* Y, const float* addend, float* s_out, float* out, long planes,
                 int cmode, float scale, int apply_gelu, cudaStream_t st, bool attr_only) {
  const int threads = round_threads(p->G_inv * p->W);
  const size_t smem = attr_only ? kMaxOptinSmem : inv_smem_bytes(p, p->G_inv);
  if (threads <= 288)
    return launch_inv_t<M1T, 288, 3>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, threads, smem, attr_only);
  if (threads <= 576)
    return launch_inv_t<M1T, 576, 1>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, threads, smem, attr_only);
  return launch_inv_t<M1T, 1024, 1>(p, Y, addend, s_out, out, planes, cmode, scale, apply_gelu, st, threads, smem, attr_only);
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

int launch_fwd2d(const Plan* p, const float* x, const float* preact, float* ds_out, float* X, long planes,
                 int cmode, float scale, cudaStream_t st) {
  FNO_DISPATCH_M1T(dispatch_fwd, p, x, preact, ds_out, X, planes, cmode, scale, st, false)
}

int launch_inv2d(const Plan* p, const float* Y, const float* addend, float* s_out, float* out, long planes,
                 int cmode, float scale, int apply_gelu, cudaStream_t st) {
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
  const size_t kMaxSmem = 200 * 1024;
  auto pick = [&](bool fwd) {
    // planes per CTA: best lane efficiency with <= 288 threads (3 CTAs/SM at <= 72 registers);
    // wider CTAs only if a narrow one would idle more than 15 % of its lanes
    int best = 1;
    double best_eff = 0.0;
    for (int limit : {288, 576}) {
      for (int G = 1; G <= 16; ++G) {
        const int thr = round_threads(G * p->W);
        if (thr > limit) break;
        const size_t sm = fwd ? fwd_smem_bytes(p, G) : inv_smem_bytes(p, G);
        if (sm > kMaxSmem) break;
        const double eff = (double)(G * p->W) / thr;
        if (eff > best_eff + 0.02) { best_eff = eff; best = G; }
      }
      if (best_eff >= 0.85) break;
    }
    const char* env = std::getenv(fwd ? "FNO_G_FWD" : "FNO_G_INV");  // tuning override
    if (env != nullptr && std::atoi(env) > 0) {
      const int G = std::atoi(env);
      const size_t sm = fwd ? fwd_smem_bytes(p, G) : inv_smem_bytes(p, G);
      if (round_threads(G * p->W) <= 1024 && sm <= kMaxSmem) best = G;
    }
    return best;
  };
  p->G_fwd = pick(true);
  p->G_inv = pick(false);
  if (round_threads(p->W) > 1024) { set_error("W = %d too large (max 1024)", p->W); return FNO_E_ARG; }
  if (fwd_smem_bytes(p, p->G_fwd) > 227 * 1024 || inv_smem_bytes(p, p->G_inv) > 227 * 1024) {
    set_error("plane %dx%d with modes (%d,%d) needs too much shared memory", p->H, p->W, p->m1, p->m2);
    return FNO_E_ARG;
  }
  int rc = setup_fwd_attr(p);
  if (rc != FNO_OK) return rc;
  return setup_inv_attr(p);
}

}  // namespace fno
