// Microbenchmark: FP32 FMA issue throughput on B200 -- scalar FFMA vs packed fma.rn.f32x2 (FFMA2),
// and FFMA interleaved with LDS.128 broadcasts (the inner loop shape of the transform kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(float2& d, float2 a, float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
               "l"(*reinterpret_cast<unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  __shared__ float4 tw[64];
  if (threadIdx.x < 64) tw[threadIdx.x] = make_float4(s, s * 0.5f, s * 0.25f, s * 0.125f);
  __syncthreads();
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float2 a2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a2[i] = make_float2(a[2 * i], a[2 * i + 1]);
  const float b = s, c = s * 0.5f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 32 independent FFMA
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], b, c);
    } else if (MODE == 1) {  // 16 independent FFMA2 (= 32 FMAs)
#pragma unroll
      for (int i = 0; i < 16; ++i) fma2(a2[i], make_float2(b, c), make_float2(c, b));
    } else if (MODE == 2) {  // 32 FFMA + 8 LDS.128 broadcast
      float4 t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = tw[(it + q) & 63];
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], reinterpret_cast<float*>(t)[i], c);
    } else {  // 16 FFMA2 + 8 LDS.128 broadcast
      float4 t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = tw[(it + q) & 63];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        fma2(a2[i], make_float2(reinterpret_cast<float*>(t)[2 * i], reinterpret_cast<float*>(t)[2 * i + 1]), make_float2(c, b));
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) r += a[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* out) {
  const int iters = 20000, grid = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, 100, 1.0001f);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)grid * 256 * iters * 32;
  printf("%-28s %8.3f ms  %7.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", name, ms, fma / ms / 1e9,
         fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("FFMA x32", out);
  run<1>("FFMA2 x16", out);
  run<2>("FFMA x32 + 8 LDS.128", out);
  run<3>("FFMA2 x16 + 8 LDS.128", out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
