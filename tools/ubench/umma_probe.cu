// Probe of the tcgen05.mma kind::tf32 operand layouts used by csrc/head_tc.cu (no-swizzle canonical
// layouts, A MN-major / B K-major and the transposed re-use).  One CTA, one 128x128x8 (or x24) tile with
// exactly representable inputs; prints the mismatch count against the expected product for several
// descriptor variants.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long umma_desc(const void* smem, unsigned lbo, unsigned sbo, unsigned layout = 0) {
  const unsigned long long addr = (smem_u32(smem) & 0x3FFFFu) >> 4;
  return addr | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) | ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((unsigned long long)layout << 61);
}
__host__ __device__ constexpr unsigned idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

// variant bits: 1 = swap A lbo/sbo, 2 = swap B lbo/sbo, 4 = A K-major test (A stored K-major like B)
__global__ void __launch_bounds__(160, 1) probe(float* D, int variant, int ksteps, int N) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* A = sm;                 // 16 KB
  unsigned char* Bm = sm + 20480;        // 16 KB
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 36864);
  unsigned* slot = reinterpret_cast<unsigned*>(sm + 36864 + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool a_kmajor = (variant & 4) != 0 || (variant & 16) != 0;
  const bool a_sw128 = (variant & 8) != 0;
  const bool b_mn = (variant & 16) != 0;
  for (int i = tid; i < 9216; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  __syncthreads();
  // A[m][k] = one-hot at k == (m % (8*ksteps));  B[n][k] = 8n + k  (k < 8*ksteps)
  const int K = 8 * ksteps;
  for (int i = tid; i < 128 * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    const float v = (k == (m % K)) ? 1.f : 0.f;
    int off;
    if (a_sw128) off = ((((m >> 2) & 7) ^ (k & 7)) * 16) + (m & 3) * 4 + (k & 7) * 128 + (m >> 5) * 1024 + (k >> 3) * 4096;
    else if (!a_kmajor) off = (m & 3) * 4 + (m >> 2) * 128 + (k & 7) * 16 + (k >> 3) * 4096;
    else if (variant & 64) off = (m & 7) * 16 + (m >> 3) * (8 * 144) + (k >> 2) * 144 + (k & 3) * 4;
    else off = (m & 7) * 16 + (m >> 3) * 1024 + (k >> 2) * 128 + (k & 3) * 4;
    *reinterpret_cast<float*>(A + off) = v;
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    int off = (n & 7) * 16 + (n >> 3) * 1024 + (k >> 2) * 128 + (k & 3) * 4;
    if (b_mn) off = (n & 3) * 4 + (n >> 2) * 128 + (k & 7) * 16 + (k >> 3) * 4096;
    *reinterpret_cast<float*>(Bm + off) = (float)(K * n + k);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tb = *slot;
  if (warp == 4 && lane == 0) {
    const unsigned idesc = idesc_tf32(128, N, a_kmajor ? 0 : 1, b_mn ? 1 : 0);
    for (int ks = 0; ks < ksteps; ++ks) {
      unsigned a_lbo = 4096, a_sbo = 128, b_lbo = 128, b_sbo = 1024;
      const unsigned char* Ap = A + ks * 4096;
      if (a_kmajor) { a_lbo = 128; a_sbo = 1024; Ap = A + ks * 256; }
      if (variant & 64) { a_lbo = 144; a_sbo = 8 * 144; Ap = A + ks * 288; }
      unsigned a_layout = 0;
      if (a_sw128) { a_lbo = 1024; a_sbo = 4096; a_layout = 2; }
      if (b_mn) { b_lbo = 4096; b_sbo = 128; }
      if (variant & 1) { unsigned t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (variant & 2) { unsigned t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      const unsigned long long ad = umma_desc(Ap, a_lbo, a_sbo, a_layout),
                               bd = umma_desc(b_mn ? Bm + ks * 4096 : Bm + ks * 256, b_lbo, b_sbo);
      const unsigned acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  if (warp < 4) {
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
      unsigned r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tb + ((unsigned)(warp * 32) << 16) + (unsigned)c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tb) : "memory");
}

int main() {
  float* D;
  cudaMalloc(&D, 128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  std::vector<float> h(128 * 128);
  const int cases[][3] = {{68, 1, 32}, {68, 3, 32}, {68, 3, 128},{4, 1, 128}, {4, 3, 128}, {4, 3, 32}, {8, 1, 128}, {9, 1, 128}, {8, 3, 128}, {16, 1, 128}, {18, 1, 128}, {16, 3, 32}, {0, 1, 128}};
  for (auto& c : cases) {
    const int variant = c[0], ks = c[1], N = c[2], K = 8 * ks;
    cudaMemset(D, 0xff, 128 * 128 * 4);
    probe<<<1, 160, 40000>>>(D, variant, ks, N);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), D, 128 * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) bad += (h[m * N + n] != (float)(K * n + (m % K)));
    printf("variant %d ksteps %d N %d: %s, mismatches %d / %d;  D[0][0..3] = %g %g %g %g | D[1][0..1] = %g %g | D[5][2] = %g (want %d) | D[127][%d] = %g (want %d)\n",
           variant, ks, N, cudaGetErrorString(e), bad, 128 * N, h[0], h[1], h[2], h[3], h[N], h[N + 1], h[5 * N + 2],
           K * 2 + 5 % K, N - 1, h[127 * N + N - 1], K * (N - 1) + 127 % K);
    if (e != cudaSuccess) break;
  }
  return 0;
}
