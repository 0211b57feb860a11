// K2: per-mode complex channel mixing over the retained spectrum (compl_mul2d / compl_mul3d,
// fno/fno.py:66-68, :255-257) and its autograd backward.  FP32 CUDA-core path: at width 20 this
// is a bandwidth/latency-bound streaming kernel (0.9 MFLOP per sample-layer), so lanes map to
// consecutive modes and every global access is a coalesced 8-byte complex element.  The corner
// parameters are read in the reference's own state_dict layout [Ci, Co, m1, m2(, m3)].
#include "common.cuh"

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
  dim3 block(32, WG_SLICES);
  dim3 grid((geo.M + 31) / 32, (Ci + TI - 1) / TI, (Co + TO - 1) / TO);
  mix_wgrad_kernel<<<grid, block, 0, st>>>(reinterpret_cast<const float2*>(X),
                                           reinterpret_cast<const float2*>(gY), gp, geo, B, Ci, Co);
  count_launch();
  return check_launch("mix_wgrad_kernel");
}

}  // namespace fno
