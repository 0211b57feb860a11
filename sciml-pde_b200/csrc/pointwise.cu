// 1x1-conv bypass of a Fourier layer (nn.Conv2d/3d(width, width, 1), fno/fno.py:131-134,162) on
// channel-first activations [B, C, N]: a per-pixel C x C matrix product.  FP32 CUDA-core path.
//
//   forward / data gradient: each thread owns VEC consecutive pixels and OT output channels; the
//     weight tile sits transposed in shared memory so that the OT weights of one input channel
//     are warp-uniform float4 broadcasts; every activation element is read exactly once per
//     output-channel tile with a coalesced 128-bit load.
//   weight gradient: gW[o,i] = sum_{b,p} ds[b,o,p] a[b,i,p] is a [C x K][K x C] product with a
//     huge K = B*N and a tiny output, so it is a split-K reduction: a CTA stages a K-slab of both
//     operands in shared memory, each warp owns a T x T output tile with lanes striding K, and
//     per-CTA partials are combined by a second fixed-order pass (deterministic, no atomics).
#include "common.cuh"

namespace fno {
namespace {

template <int OT, int VEC, bool TRANSPOSE>
__global__ void __launch_bounds__(256)
pointwise_kernel(const float* __restrict__ in, const float* __restrict__ Wm, const float* __restrict__ bias,
                 float* __restrict__ out, int Cin, int Cout, int Co, int Ci, long N) {
  extern __shared__ __align__(16) float ws[];  // [Cin][OT] weights, then [OT] bias
  const int o0 = blockIdx.y * OT;
  const int b = blockIdx.z;
  for (int idx = threadIdx.x; idx < Cin * OT; idx += blockDim.x) {
    const int s = idx / OT, oo = idx - s * OT;
    const int oc = o0 + oo;
    float v = 0.f;
    if (oc < Cout) v = TRANSPOSE ? Wm[(size_t)s * Ci + oc] : Wm[(size_t)oc * Ci + s];
    ws[idx] = v;
  }
  float* bs = ws + Cin * OT;
  for (int idx = threadIdx.x; idx < OT; idx += blockDim.x)
    bs[idx] = (bias != nullptr && o0 + idx < Cout) ? bias[o0 + idx] : 0.f;
  __syncthreads();

  const long p0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (p0 >= N) return;
  float acc[OT][VEC];
#pragma unroll
  for (int oo = 0; oo < OT; ++oo)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[oo][v] = bs[oo];
  const float* __restrict__ ip = in + (size_t)b * Cin * N + p0;
#pragma unroll 4
  for (int s = 0; s < Cin; ++s) {
    float xv[VEC];
    if constexpr (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(ip + (size_t)s * N));
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else {
      xv[0] = __ldg(ip + (size_t)s * N);
    }
    const float* wrow = ws + s * OT;
#pragma unroll
    for (int oo = 0; oo < OT; ++oo) {
      const float wv = wrow[oo];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[oo][v] = fmaf(wv, xv[v], acc[oo][v]);
    }
  }
  float* __restrict__ op = out + (size_t)b * Cout * N + p0;
#pragma unroll
  for (int oo = 0; oo < OT; ++oo) {
    if (o0 + oo >= Cout) break;
    if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(op + (size_t)(o0 + oo) * N) =
          make_float4(acc[oo][0], acc[oo][1], acc[oo][2], acc[oo][3]);
    } else {
      op[(size_t)(o0 + oo) * N] = acc[oo][0];
    }
  }
}

template <int OT, int VEC>
int launch_pw(const float* in, const float* Wm, const float* bias, float* out, int B, int Co, int Ci, long N,
              int transpose, cudaStream_t st) {
  const int Cin = transpose ? Co : Ci;
  const int Cout = transpose ? Ci : Co;
  const int threads = 256;
  const long nvec = (N + VEC - 1) / VEC;
  dim3 grid((unsigned)((nvec + threads - 1) / threads), (Cout + OT - 1) / OT, B);
  const size_t smem = sizeof(float) * ((size_t)Cin * OT + OT);
  if (transpose)
    pointwise_kernel<OT, VEC, true><<<grid, threads, smem, st>>>(in, Wm, bias, out, Cin, Cout, Co, Ci, N);
  else
    pointwise_kernel<OT, VEC, false><<<grid, threads, smem, st>>>(in, Wm, bias, out, Cin, Cout, Co, Ci, N);
  count_launch();
  return check_launch("pointwise_kernel");
}

// ------------------------------------------------------------------------------------------
// weight / bias gradient
// ------------------------------------------------------------------------------------------
constexpr int KT = 128;      // pixels per shared-memory slab
constexpr int WG_T = 10;     // output tile edge per warp (register tile T x T)
constexpr int WG_MAXW = 16;  // warps per CTA (max tiles per CTA)

// part[(b * chunks + chunk)][Co][Ci + 1]  (last column: bias gradient)
template <int T>
__global__ void __launch_bounds__(32 * WG_MAXW)
wgrad_partial_kernel(const float* __restrict__ ds, const float* __restrict__ a, float* __restrict__ part,
                     int Co, int Ci, long N, long chunk_len, int tiles_i, int ntiles) {
  extern __shared__ __align__(16) float sm[];  // [Co][KT] ds slab, [Ci][KT] a slab
  float* ds_s = sm;
  float* a_s = sm + (size_t)Co * KT;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int b = blockIdx.y;
  const long k_begin = (long)blockIdx.x * chunk_len;
  long k_end = k_begin + chunk_len;
  if (k_end > N) k_end = N;
  const int tile = blockIdx.z * nwarps + warp;
  const bool has_tile = tile < ntiles;
  const int to = has_tile ? (tile / tiles_i) * T : 0;
  const int ti = has_tile ? (tile % tiles_i) * T : 0;
  const bool bias_tile = has_tile && (ti == 0);

  float acc[T][T];
  float accb[T];
#pragma unroll
  for (int r = 0; r < T; ++r) {
    accb[r] = 0.f;
#pragma unroll
    for (int c = 0; c < T; ++c) acc[r][c] = 0.f;
  }
  const float* __restrict__ dsb = ds + (size_t)b * Co * N;
  const float* __restrict__ ab = a + (size_t)b * Ci * N;
  for (long k0 = k_begin; k0 < k_end; k0 += KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < Co * KT; idx += blockDim.x) {
      const int c = idx / KT, kk = idx - c * KT;
      const long k = k0 + kk;
      ds_s[idx] = (k < k_end) ? __ldg(dsb + (size_t)c * N + k) : 0.f;
    }
    for (int idx = threadIdx.x; idx < Ci * KT; idx += blockDim.x) {
      const int c = idx / KT, kk = idx - c * KT;
      const long k = k0 + kk;
      a_s[idx] = (k < k_end) ? __ldg(ab + (size_t)c * N + k) : 0.f;
    }
    __syncthreads();
    if (has_tile) {
#pragma unroll
      for (int kk = 0; kk < KT; kk += 32) {
        float dv[T], av[T];
#pragma unroll
        for (int r = 0; r < T; ++r) dv[r] = (to + r < Co) ? ds_s[(to + r) * KT + kk + lane] : 0.f;
#pragma unroll
        for (int c = 0; c < T; ++c) av[c] = (ti + c < Ci) ? a_s[(ti + c) * KT + kk + lane] : 0.f;
#pragma unroll
        for (int r = 0; r < T; ++r) {
          accb[r] += dv[r];
#pragma unroll
          for (int c = 0; c < T; ++c) acc[r][c] = fmaf(dv[r], av[c], acc[r][c]);
        }
      }
    }
  }
  if (!has_tile) return;
  float* __restrict__ pp = part + ((size_t)b * gridDim.x + blockIdx.x) * Co * (Ci + 1);
#pragma unroll
  for (int r = 0; r < T; ++r) {
#pragma unroll
    for (int c = 0; c < T; ++c) {
      float v = acc[r][c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0 && to + r < Co && ti + c < Ci) pp[(size_t)(to + r) * (Ci + 1) + ti + c] = v;
    }
    if (bias_tile) {
      float v = accb[r];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0 && to + r < Co) pp[(size_t)(to + r) * (Ci + 1) + Ci] = v;
    }
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gW, float* __restrict__ gb,
                                    int nparts, int Co, int Ci) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = Co * (Ci + 1);
  if (idx >= total) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * total + idx];
  const int o = idx / (Ci + 1), i = idx - o * (Ci + 1);
  if (i < Ci) {
    if (gW != nullptr) gW[(size_t)o * Ci + i] = s;
  } else if (gb != nullptr) {
    gb[o] = s;
  }
}

int wgrad_chunks(long N) {
  // enough CTAs per sample to fill the machine at small batch, but >= 4 slabs of work each
  long c = (N + 4 * KT - 1) / (4 * KT);
  if (c > 8) c = 8;
  if (c < 1) c = 1;
  return (int)c;
}

}  // namespace
}  // namespace fno

using namespace fno;

extern "C" int fno_pointwise_fwd(const float* in, const float* W, const float* bias, float* out, int B, int Co,
                                 int Ci, long N, int transpose, fno_stream_t stream) {
  if (!in || !W || !out || B <= 0 || Co <= 0 || Ci <= 0 || N <= 0) {
    set_error("fno_pointwise_fwd: bad argument");
    return FNO_E_ARG;
  }
  if (B > 65535) { set_error("fno_pointwise_fwd: batch %d > 65535", B); return FNO_E_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Cout = transpose ? Ci : Co;
  const bool vec4 = (N % 4 == 0) && ((reinterpret_cast<size_t>(in) | reinterpret_cast<size_t>(out)) % 16 == 0);
  if (vec4) {
    if (Cout % 20 == 0) return launch_pw<20, 4>(in, W, bias, out, B, Co, Ci, N, transpose, st);
    if (Cout % 16 == 0) return launch_pw<16, 4>(in, W, bias, out, B, Co, Ci, N, transpose, st);
    if (Cout % 8 == 0) return launch_pw<8, 4>(in, W, bias, out, B, Co, Ci, N, transpose, st);
    return launch_pw<4, 4>(in, W, bias, out, B, Co, Ci, N, transpose, st);
  }
  if (Cout % 20 == 0) return launch_pw<20, 1>(in, W, bias, out, B, Co, Ci, N, transpose, st);
  if (Cout % 16 == 0) return launch_pw<16, 1>(in, W, bias, out, B, Co, Ci, N, transpose, st);
  if (Cout % 8 == 0) return launch_pw<8, 1>(in, W, bias, out, B, Co, Ci, N, transpose, st);
  return launch_pw<4, 1>(in, W, bias, out, B, Co, Ci, N, transpose, st);
}

extern "C" size_t fno_pointwise_wgrad_workspace_bytes(int B, int Co, int Ci, long N) {
  if (B <= 0 || Co <= 0 || Ci <= 0 || N <= 0) return 0;
  return sizeof(float) * (size_t)B * wgrad_chunks(N) * Co * (Ci + 1);
}

extern "C" int fno_pointwise_wgrad(const float* ds, const float* a, float* gW, float* gb, void* work, int B, int Co,
                                   int Ci, long N, fno_stream_t stream) {
  if (!ds || !a || !work || B <= 0 || Co <= 0 || Ci <= 0 || N <= 0) {
    set_error("fno_pointwise_wgrad: bad argument");
    return FNO_E_ARG;
  }
  if (B > 65535) { set_error("fno_pointwise_wgrad: batch %d > 65535", B); return FNO_E_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = wgrad_chunks(N);
  long chunk_len = (N + chunks - 1) / chunks;
  chunk_len = ((chunk_len + KT - 1) / KT) * KT;
  constexpr int T = WG_T;
  const int tiles_o = (Co + T - 1) / T, tiles_i = (Ci + T - 1) / T;
  const int ntiles = tiles_o * tiles_i;
  const int warps = ntiles < WG_MAXW ? ntiles : WG_MAXW;
  const int zgroups = (ntiles + warps - 1) / warps;
  const size_t smem = sizeof(float) * (size_t)(Co + Ci) * KT;
  if (smem > 200 * 1024) { set_error("fno_pointwise_wgrad: width %d too large", Co + Ci); return FNO_E_ARG; }
  static std::atomic<size_t> attr_set{0};
  if (smem > 48 * 1024 && attr_set.load() < smem) {
    if (cudaFuncSetAttribute(wgrad_partial_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return check_launch("cudaFuncSetAttribute(wgrad)");
    attr_set.store(smem);
  }
  float* part = static_cast<float*>(work);
  dim3 grid(chunks, B, zgroups);
  wgrad_partial_kernel<T><<<grid, 32 * warps, smem, st>>>(ds, a, part, Co, Ci, N, chunk_len, tiles_i, ntiles);
  count_launch();
  int rc = check_launch("wgrad_partial_kernel");
  if (rc != FNO_OK) return rc;
  const int total = Co * (Ci + 1);
  wgrad_reduce_kernel<<<(total + 127) / 128, 128, 0, st>>>(part, gW, gb, B * chunks, Co, Ci);
  count_launch();
  return check_launch("wgrad_reduce_kernel");
}
