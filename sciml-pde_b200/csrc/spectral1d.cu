// SpectralConv1d: the 1-D member of the spectral-convolution family north_star names.  The reference tree has no 1-D
// layer, so this follows the 2-D layer's conventions (fno/fno.py:35-92) one dimension down:
//     x_ft = rfft(x);  out_ft[:, :, :m] = einsum("bix,iox->box", x_ft[:, :, :m], weights1);  y = irfft(out_ft, n = N)
// Kernels (FP32 CUDA cores: a row is N <= 4096 floats and m <= 64 modes, the work is bandwidth-trivial):
//   dft1d_fwd_kernel   X[r, k] = scale * c_k * sum_n x[r, n] exp(-2 pi i k n / N),  k < m      (pruned rfft; with cmode = 1
//                      and scale = 1/N also the backward of the inverse, c_k = C2R column weights)
//   dft1d_inv_kernel   y[r, n] = scale * sum_k c_k Re(Y[r, k] exp(+2 pi i k n / N))            (zero-padded irfft; with
//                      cmode = 0, scale = 1 the backward of the forward transform)
//   mix1d_*_kernel     per-mode complex channel mixing and its two gradients
// C2R semantics as in the 2-D kernels: c_0 = 1, c_k = 2, c_{N/2} = 1 for even N; Im(Y_0) and Im(Y_{N/2}) are ignored.
#include "common.cuh"

namespace fno {
namespace {

constexpr int S1_KC = 8;           // modes per pass of the forward kernel
constexpr int S1_MAXN = 4096;

__device__ __forceinline__ void fill_twiddles(float2* tw, int N) {
  for (int r = threadIdx.x; r < N; r += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)r / (float)N, &s, &c);
    tw[r] = make_float2(c, s);
  }
}

__device__ __forceinline__ float c2r_weight(int k, int N, int cmode) {
  if (!cmode) return 1.0f;
  return (k == 0 || 2 * k == N) ? 1.0f : 2.0f;
}

// one warp per row; lane strides over n, S1_KC modes per pass
__global__ void __launch_bounds__(256)
dft1d_fwd_kernel(const float* __restrict__ x, float2* __restrict__ X, long rows, int N, int m, int cmode, float scale) {
  extern __shared__ float2 tw[];
  fill_twiddles(tw, N);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (long r = (long)blockIdx.x * wpb + warp; r < rows; r += (long)gridDim.x * wpb) {
    const float* xr = x + (size_t)r * N;
    for (int k0 = 0; k0 < m; k0 += S1_KC) {
      float re[S1_KC], im[S1_KC];
      int ph[S1_KC], st[S1_KC];
#pragma unroll
      for (int j = 0; j < S1_KC; ++j) {
        re[j] = im[j] = 0.f;
        const int k = k0 + j;
        ph[j] = (int)(((long)k * lane) % N);
        st[j] = (int)(((long)k * 32) % N);
      }
      for (int n = lane; n < N; n += 32) {
        const float v = __ldg(xr + n);
#pragma unroll
        for (int j = 0; j < S1_KC; ++j) {
          const float2 w = tw[ph[j]];
          re[j] = fmaf(v, w.x, re[j]);
          im[j] = fmaf(-v, w.y, im[j]);
          ph[j] += st[j];
          if (ph[j] >= N) ph[j] -= N;
        }
      }
#pragma unroll
      for (int j = 0; j < S1_KC; ++j) {
        float a = re[j], b = im[j];
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        const int k = k0 + j;
        if (lane == 0 && k < m) {
          const float c = scale * c2r_weight(k, N, cmode);
          X[(size_t)r * m + k] = make_float2(a * c, b * c);
        }
      }
    }
  }
}

// block = one row chunk; thread per output sample
__global__ void __launch_bounds__(256)
dft1d_inv_kernel(const float2* __restrict__ Y, const float* __restrict__ addend, float* __restrict__ y, long rows, int N,
                 int m, int cmode, float scale) {
  extern __shared__ float2 sm[];
  float2* tw = sm;                 // [N]
  float2* Yr = sm + N;             // [m], weights folded in
  fill_twiddles(tw, N);
  for (long r = blockIdx.x; r < rows; r += gridDim.x) {
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
      const float c = scale * c2r_weight(k, N, cmode);
      float2 v = __ldg(Y + (size_t)r * m + k);
      if (cmode && (k == 0 || 2 * k == N)) v.y = 0.f;
      Yr[k] = make_float2(v.x * c, v.y * c);
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float acc = 0.f;
      int ph = 0;
      for (int k = 0; k < m; ++k) {
        const float2 w = tw[ph], v = Yr[k];
        acc = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc));
        ph += n;
        if (ph >= N) ph -= N;
      }
      const size_t o = (size_t)r * N + n;
      y[o] = addend ? acc + __ldg(addend + o) : acc;
    }
  }
}

// Y[b, o, k] = sum_i X[b, i, k] W[i, o, k]
__global__ void __launch_bounds__(256)
mix1d_fwd_kernel(const float2* __restrict__ X, const float2* __restrict__ W, float2* __restrict__ Y, int B, int Ci, int Co,
                 int m) {
  const long total = (long)B * Co * m;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % m);
    const long bo = idx / m;
    const int o = (int)(bo % Co), b = (int)(bo / Co);
    float re = 0.f, im = 0.f;
    for (int i = 0; i < Ci; ++i) {
      const float2 xv = __ldg(X + ((size_t)b * Ci + i) * m + k), wv = __ldg(W + ((size_t)i * Co + o) * m + k);
      re = fmaf(xv.x, wv.x, fmaf(-xv.y, wv.y, re));
      im = fmaf(xv.x, wv.y, fmaf(xv.y, wv.x, im));
    }
    Y[idx] = make_float2(re, im);
  }
}

// gX[b, i, k] = sum_o gY[b, o, k] conj(W[i, o, k])
__global__ void __launch_bounds__(256)
mix1d_dgrad_kernel(const float2* __restrict__ gY, const float2* __restrict__ W, float2* __restrict__ gX, int B, int Ci,
                   int Co, int m) {
  const long total = (long)B * Ci * m;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % m);
    const long bi = idx / m;
    const int i = (int)(bi % Ci), b = (int)(bi / Ci);
    float re = 0.f, im = 0.f;
    for (int o = 0; o < Co; ++o) {
      const float2 g = __ldg(gY + ((size_t)b * Co + o) * m + k), wv = __ldg(W + ((size_t)i * Co + o) * m + k);
      re = fmaf(g.x, wv.x, fmaf(g.y, wv.y, re));
      im = fmaf(g.y, wv.x, fmaf(-g.x, wv.y, im));
    }
    gX[idx] = make_float2(re, im);
  }
}

// gW[i, o, k] = sum_b conj(X[b, i, k]) gY[b, o, k]
__global__ void __launch_bounds__(256)
mix1d_wgrad_kernel(const float2* __restrict__ X, const float2* __restrict__ gY, float2* __restrict__ gW, int B, int Ci,
                   int Co, int m) {
  const long total = (long)Ci * Co * m;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % m);
    const long io = idx / m;
    const int o = (int)(io % Co), i = (int)(io / Co);
    float re = 0.f, im = 0.f;
    for (int b = 0; b < B; ++b) {
      const float2 xv = __ldg(X + ((size_t)b * Ci + i) * m + k), g = __ldg(gY + ((size_t)b * Co + o) * m + k);
      re = fmaf(xv.x, g.x, fmaf(xv.y, g.y, re));
      im = fmaf(xv.x, g.y, fmaf(-xv.y, g.x, im));
    }
    gW[idx] = make_float2(re, im);
  }
}

bool bad_geo(int N, int m) { return N < 2 || N > S1_MAXN || m < 1 || m > N / 2 + 1; }

unsigned blocks_for(long total, int per_block) {
  long b = (total + per_block - 1) / per_block;
  if (b > 148 * 8) b = 148 * 8;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace fno

using namespace fno;

extern "C" int fno_sc1d_fwd_transform(const float* x, float* X, long rows, int N, int m, int cmode, float scale,
                                      fno_stream_t stream) {
  if (!x || !X || rows <= 0 || bad_geo(N, m)) {
    set_error("fno_sc1d_fwd_transform: bad argument (2 <= N <= %d, 1 <= m <= N/2+1)", S1_MAXN);
    return FNO_E_ARG;
  }
  dft1d_fwd_kernel<<<blocks_for(rows, 8), 256, sizeof(float2) * N, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<float2*>(X), rows, N, m, cmode, scale);
  count_launch();
  return check_launch("dft1d_fwd_kernel");
}

extern "C" int fno_sc1d_inv_transform(const float* Y, const float* addend, float* y, long rows, int N, int m, int cmode,
                                      float scale, fno_stream_t stream) {
  if (!Y || !y || rows <= 0 || bad_geo(N, m)) {
    set_error("fno_sc1d_inv_transform: bad argument (2 <= N <= %d, 1 <= m <= N/2+1)", S1_MAXN);
    return FNO_E_ARG;
  }
  dft1d_inv_kernel<<<blocks_for(rows, 1), 256, sizeof(float2) * (N + m), static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(Y), addend, y, rows, N, m, cmode, scale);
  count_launch();
  return check_launch("dft1d_inv_kernel");
}

extern "C" int fno_mix1d_fwd(const float* X, const float* W, float* Y, int B, int Ci, int Co, int m, fno_stream_t stream) {
  if (!X || !W || !Y || B <= 0 || Ci <= 0 || Co <= 0 || m <= 0) { set_error("fno_mix1d_fwd: bad argument"); return FNO_E_ARG; }
  mix1d_fwd_kernel<<<blocks_for((long)B * Co * m, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(X), reinterpret_cast<const float2*>(W), reinterpret_cast<float2*>(Y), B, Ci, Co, m);
  count_launch();
  return check_launch("mix1d_fwd_kernel");
}

extern "C" int fno_mix1d_bwd(const float* X, const float* gY, const float* W, float* gX, float* gW, int B, int Ci, int Co,
                             int m, fno_stream_t stream) {
  if (!gY || B <= 0 || Ci <= 0 || Co <= 0 || m <= 0 || (gX && !W) || (gW && !X)) {
    set_error("fno_mix1d_bwd: bad argument");
    return FNO_E_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gX) {
    mix1d_dgrad_kernel<<<blocks_for((long)B * Ci * m, 256), 256, 0, st>>>(
        reinterpret_cast<const float2*>(gY), reinterpret_cast<const float2*>(W), reinterpret_cast<float2*>(gX), B, Ci, Co, m);
    count_launch();
  }
  if (gW) {
    mix1d_wgrad_kernel<<<blocks_for((long)Ci * Co * m, 256), 256, 0, st>>>(
        reinterpret_cast<const float2*>(X), reinterpret_cast<const float2*>(gY), reinterpret_cast<float2*>(gW), B, Ci, Co, m);
    count_launch();
  }
  return check_launch("mix1d_bwd");
}
