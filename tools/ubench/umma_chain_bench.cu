// How fast do small tcgen05.mma.kind::tf32 instructions retire when they accumulate into ONE TMEM tile
// (dependent chain) versus several independent tiles issued round-robin?  Times `n` MMAs (K = 8 each)
// from first issue to the commit's mbarrier arrival with clock64, for the shapes csrc/head_bwd_tc.cu uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_chain_bench umma_chain_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long umma_desc(const void* smem, unsigned lbo, unsigned sbo) {
  const unsigned long long addr = (smem_u32(smem) & 0x3FFFFu) >> 4;
  return addr | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) | ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr unsigned idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void mma_elect(unsigned d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int M, int N, int CHAINS>
__global__ void __launch_bounds__(64, 1) bench(long long* out, int n) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 196608);
  unsigned* slot = reinterpret_cast<unsigned*>(sm + 196608 + 16);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tb = *slot;
  if (warp == 1) {
    constexpr unsigned idesc = idesc_tf32(M, N);
    // K-major operands, K = 128 resident: A [M][128] (SBO 4096), B [N][128]
    const unsigned long long a0 = umma_desc(sm, 128, 4096), b0 = umma_desc(sm + 65536, 128, 4096);
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < n; i += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
          mma_elect(tb + (unsigned)((u % CHAINS) * N), a0 + (unsigned long long)(u * 16), b0 + (unsigned long long)(u * 16), idesc, 1);
      }
      const long long t1 = clock64();
      asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
      mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      if (lane == 0) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int M, int N, int CHAINS>
void run(long long* d, int n) {
  cudaFuncSetAttribute(bench<M, N, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  bench<M, N, CHAINS><<<1, 64, 200000>>>(d, n);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[6];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("M %3d N %3d chains %d  n %4d: %s  issue %6lld clk, complete %6lld clk -> %.1f clk / MMA (last of 3 reps)\n", M, N, CHAINS, n,
         cudaGetErrorString(e), h[4], h[5], (double)h[5] / n);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int n = 256;
  run<64, 24, 1>(d, n); run<64, 24, 2>(d, n); run<64, 24, 3>(d, n); run<64, 24, 6>(d, n);
  run<128, 32, 1>(d, n); run<128, 32, 3>(d, n); run<128, 32, 6>(d, n);
  run<128, 64, 1>(d, n); run<128, 64, 2>(d, n); run<128, 64, 3>(d, n);
  run<128, 128, 1>(d, n); run<128, 128, 2>(d, n);
  run<64, 48, 1>(d, n); run<64, 48, 2>(d, n); run<64, 48, 4>(d, n);
  run<128, 256, 1>(d, n);
  return 0;
}
