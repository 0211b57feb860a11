// (variant: A operand read from TENSOR MEMORY -- tcgen05.mma [d], [a_tmem], b_desc -- instead of shared memory)
// Probe of the TMEM accumulator layout of tcgen05.mma.cta_group::1.kind::tf32 with M = 64 (and M = 128
// as the control): which TMEM lane holds row m of D?  A[m][k] = (k == m), B[n][k] = 64 n + k, K = 64, so
// D[m][n] = 64 n + m identifies its own row and column.  All 128 lanes x N columns are pre-filled with
// -1 (tcgen05.st), then dumped after the MMAs.  Used by csrc/head_bwd_tc.cu (dh = dpre W1 runs with
// 64 pixels on M).   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_m64_probe umma_m64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
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

constexpr int K = 64;
// a_lbo: byte stride between the 16-byte K chunks of A (128 or 144)
__global__ void __launch_bounds__(160, 1) probe(float* D, int M, int N, int a_lbo) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* A = sm;                 // <= 128 rows: 16 groups x 16 chunks x 144 B = 36 864 B
  unsigned char* Bm = sm + 36864;        // <= 64 rows: 8 groups x 2048 B = 16 384 B
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 36864 + 16384);
  unsigned* slot = reinterpret_cast<unsigned*>(sm + 36864 + 16384 + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (36864 + 16384) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  __syncthreads();
  const int a_sbo = (K / 4) * a_lbo;
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    const int off = (m & 7) * 16 + (m >> 3) * a_sbo + (k >> 2) * a_lbo + (k & 3) * 4;
    *reinterpret_cast<float*>(A + off) = (k == (m % K)) ? (m < K ? 1.f : 2.f) : 0.f;   // rows >= 64 (M = 128) carry a factor 2
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const int off = (n & 7) * 16 + (n >> 3) * 2048 + (k >> 2) * 128 + (k & 3) * 4;
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
  if (warp < 4) {   // A[m][k] -> TMEM lane m, column 64 + k
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 8) {
      unsigned r[8];
      for (int i = 0; i < 8; ++i) r[i] = __float_as_uint((m < M && (c0 + i) == (m % K)) ? (m < K ? 1.f : 2.f) : 0.f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tb + ((unsigned)(warp * 32) << 16) + 64u + (unsigned)c0),
                   "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
  }
  if (warp < 4) {   // sentinel fill: -1 in every lane, columns [0, 64)
    const unsigned s = __float_as_uint(-1.f);
    for (int c0 = 0; c0 < 64; c0 += 8) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tb + ((unsigned)(warp * 32) << 16) + (unsigned)c0), "r"(s) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 4 && lane == 0) {
    const unsigned idesc = idesc_tf32(M, N);
    for (int ks = 0; ks < K / 8; ++ks) {
      const unsigned long long bd = umma_desc(Bm + ks * 256, 128, 2048);
      const unsigned acc = ks > 0;
      const unsigned at = tb + 64u + (unsigned)(ks * 8);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tb), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  if (warp < 4) {
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 8) {
      unsigned r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tb + ((unsigned)(warp * 32) << 16) + (unsigned)c0)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i) D[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tb) : "memory");
}

int main() {
  float* D;
  cudaMalloc(&D, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  std::vector<float> h(128 * 64);
  const int cases[][3] = {{128, 32, 128}, {64, 24, 128}, {128, 16, 128}};
  for (auto& c : cases) {
    const int M = c[0], N = c[1], lbo = c[2];
    cudaMemset(D, 0, 128 * 64 * 4);
    probe<<<1, 160, 60000>>>(D, M, N, lbo);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), D, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    printf("M %d N %d a_lbo %d: %s\n", M, N, lbo, cudaGetErrorString(e));
    if (e != cudaSuccess) break;
    // per lane: which row does column 0 hold, and do the other columns agree (value = 64 n + m [x2 for m >= 64])
    int written_cols_max = 0;
    for (int l = 0; l < 128; ++l) {
      const float v0 = h[l * 64], v1 = h[l * 64 + 1];
      int row = -1, ok = 1, ncols = 0;
      if (v0 >= 0.f) {
        const float f = (v1 - v0 == 128.f) ? 2.f : 1.f;     // rows >= 64 carry a factor 2
        row = (int)(v0 / f) + (f == 2.f ? 64 : 0);
        for (int n = 0; n < 64; ++n) {
          const float v = h[l * 64 + n];
          if (v < 0.f) continue;
          ++ncols;
          if (v != f * (float)(64 * n + (row % 64))) ok = 0;
        }
      }
      if (ncols > written_cols_max) written_cols_max = ncols;
      printf("%s%d:%d%s", (l % 16 == 0) ? "\n  lane->row " : " ", l, row, ok ? "" : "!");
    }
    printf("\n  columns written per live lane: %d\n", written_cols_max);
  }
  return 0;
}
