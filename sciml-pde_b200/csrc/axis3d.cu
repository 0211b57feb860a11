// Third (outer, strided) axis of the 3-D transforms.  A 3-D volume [D1, D2, D3] is processed as
// D1 independent [D2, D3] planes by the 2-D kernels (transform2d.cu); what remains is a pruned
// complex DFT along D1 applied to the already-reduced per-slice spectra S[planes, D1, Q]
// (Q = 2*m2*m3 complex modes per slice, ~13 % of the volume's bytes at the reference's 3-D
// configuration).  One thread owns one (volume, q) column: consecutive lanes touch consecutive
// complex elements (coalesced 8-byte accesses) and the twiddles are warp-uniform broadcasts.
#include "common.cuh"

namespace fno {
namespace {

// X[p, r, q] = sum_d S[p, d, q] * exp(-2 pi i k_r d / D1)      (R = 2*m1x kept rows)
template <int RT>
__global__ void __launch_bounds__(128)
axis_fwd_kernel(const float2* __restrict__ S, float2* __restrict__ X, const float2* __restrict__ twX, int D1,
                int R, long Q, long planes) {
  extern __shared__ __align__(16) float2 tw_s[];  // [D1][R] (cos, sin) of 2 pi k_r d / D1
  for (int i = threadIdx.x; i < D1 * R; i += blockDim.x) tw_s[i] = twX[i];
  __syncthreads();
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long p = blockIdx.y;
  if (q >= Q || p >= planes) return;
  float2 acc[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = make_float2(0.f, 0.f);
  const float2* __restrict__ sp = S + (size_t)p * D1 * Q + q;
#pragma unroll 2
  for (int d = 0; d < D1; ++d) {
    const float2 v = __ldg(sp + (size_t)d * Q);
    const float2* tw = tw_s + (size_t)d * R;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      if (r < R) {
        const float2 t = tw[r];  // e^{-i th} = c - i s
        acc[r].x = fmaf(v.x, t.x, acc[r].x);
        acc[r].x = fmaf(v.y, t.y, acc[r].x);
        acc[r].y = fmaf(v.y, t.x, acc[r].y);
        acc[r].y = fmaf(-v.x, t.y, acc[r].y);
      }
    }
  }
  float2* __restrict__ xp = X + (size_t)p * R * Q + q;
#pragma unroll
  for (int r = 0; r < RT; ++r)
    if (r < R) xp[(size_t)r * Q] = acc[r];
}

// Z[p, d, q] = sum_r Y[p, r, q] * exp(+2 pi i k_r d / D1)
template <int RT>
__global__ void __launch_bounds__(128)
axis_inv_kernel(const float2* __restrict__ Y, float2* __restrict__ Z, const float2* __restrict__ twX, int D1,
                int R, long Q, long planes) {
  extern __shared__ __align__(16) float2 tw_s[];
  for (int i = threadIdx.x; i < D1 * R; i += blockDim.x) tw_s[i] = twX[i];
  __syncthreads();
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long p = blockIdx.y;
  if (q >= Q || p >= planes) return;
  float2 y[RT];
  const float2* __restrict__ yp = Y + (size_t)p * R * Q + q;
#pragma unroll
  for (int r = 0; r < RT; ++r) y[r] = (r < R) ? __ldg(yp + (size_t)r * Q) : make_float2(0.f, 0.f);
  float2* __restrict__ zp = Z + (size_t)p * D1 * Q + q;
  for (int d = 0; d < D1; ++d) {
    const float2* tw = tw_s + (size_t)d * R;
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      if (r < R) {
        const float2 t = tw[r];  // e^{+i th} = c + i s
        float2& a = (r & 1) ? a1 : a0;
        a.x = fmaf(y[r].x, t.x, a.x);
        a.x = fmaf(-y[r].y, t.y, a.x);
        a.y = fmaf(y[r].y, t.x, a.y);
        a.y = fmaf(y[r].x, t.y, a.y);
      }
    }
    zp[(size_t)d * Q] = make_float2(a0.x + a1.x, a0.y + a1.y);
  }
}

}  // namespace

#define FNO_DISPATCH_RT(K, ...)                                                   \
  if (R <= 8) K<8><<<grid, 128, smem, st>>>(__VA_ARGS__);                         \
  else if (R <= 16) K<16><<<grid, 128, smem, st>>>(__VA_ARGS__);                  \
  else if (R <= 24) K<24><<<grid, 128, smem, st>>>(__VA_ARGS__);                  \
  else if (R <= 32) K<32><<<grid, 128, smem, st>>>(__VA_ARGS__);                  \
  else { set_error("3-D modes1 %d too large (max 16)", R / 2); return FNO_E_ARG; }

int launch_axis_fwd(const Plan* p, const float* S, float* X, long planes, long Q, cudaStream_t st) {
  const int R = 2 * p->m1x;
  dim3 grid((unsigned)((Q + 127) / 128), (unsigned)planes);
  const size_t smem = sizeof(float2) * (size_t)p->D1 * R;
  if (planes > 65535) { set_error("axis_fwd: %ld volumes > 65535", planes); return FNO_E_ARG; }
  FNO_DISPATCH_RT(axis_fwd_kernel, reinterpret_cast<const float2*>(S), reinterpret_cast<float2*>(X),
                  reinterpret_cast<const float2*>(p->twX), p->D1, R, Q, planes)
  count_launch();
  return check_launch("axis_fwd_kernel");
}

int launch_axis_inv(const Plan* p, const float* Y, float* Z, long planes, long Q, cudaStream_t st) {
  const int R = 2 * p->m1x;
  dim3 grid((unsigned)((Q + 127) / 128), (unsigned)planes);
  const size_t smem = sizeof(float2) * (size_t)p->D1 * R;
  if (planes > 65535) { set_error("axis_inv: %ld volumes > 65535", planes); return FNO_E_ARG; }
  FNO_DISPATCH_RT(axis_inv_kernel, reinterpret_cast<const float2*>(Y), reinterpret_cast<float2*>(Z),
                  reinterpret_cast<const float2*>(p->twX), p->D1, R, Q, planes)
  count_launch();
  return check_launch("axis_inv_kernel");
}

}  // namespace fno
