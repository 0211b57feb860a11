// Step tail of the reference training loop, device-side and sync-free:
//
//   loss   nrmse(out, target).mean()  (fno/train.py:34-40, :266-267) and its gradient w.r.t. `out`
//          in one pass pair (the torch form is ~10 elementwise / reduce kernels each way);
//   update total_norm = || [ ||g_p|| ]_p ||; clip_value = max(5, 0.1 total_norm);
//          clip_grad_norm_(params, clip_value); Adam(lr, betas, eps, weight_decay as coupled L2);
//          CosineAnnealingLR stepped every iteration (fno/train.py:251-259, :273-278) -- the
//          reference evaluates max(5, 0.1 * total_norm) on the host (a device->host sync per
//          step); here the clip coefficient, the learning rate of the step and the bias
//          corrections are computed on the device from a device-resident step counter, so the
//          whole step can live in one CUDA graph.
//
// Tensors are addressed through a chunk table (tensor pointers + offsets), multi-tensor style:
// parameters stay in the module's own storage (state_dict contract); complex parameters are
// treated as interleaved real pairs, which is exactly what torch's Adam does (view_as_real).
#include "common.cuh"

namespace fno {
namespace {

constexpr int LOSS_BLOCKS = 32;     // partial blocks per sample
constexpr int LOSS_THREADS = 256;
constexpr int LVMAX = 8;

// part[b][blk][v][2] = sum (out - y)^2, sum y^2 over the block's share of the P pixels
__global__ void __launch_bounds__(LOSS_THREADS)
nrmse_partial_kernel(const float* __restrict__ out, const float* __restrict__ tgt, float* __restrict__ part, long P,
                     int V) {
  const int b = blockIdx.y;
  const float* __restrict__ ob = out + (size_t)b * P * V;
  const float* __restrict__ tb = tgt + (size_t)b * P * V;
  float s1[LVMAX], s2[LVMAX];
#pragma unroll
  for (int v = 0; v < LVMAX; ++v) { s1[v] = 0.f; s2[v] = 0.f; }
  const long per = (P + gridDim.x - 1) / gridDim.x;
  const long p0 = (long)blockIdx.x * per;
  long p1 = p0 + per;
  if (p1 > P) p1 = P;
  for (long p = p0 + threadIdx.x; p < p1; p += LOSS_THREADS) {
#pragma unroll
    for (int v = 0; v < LVMAX; ++v) {
      if (v < V) {
        const float y = __ldg(tb + p * V + v);
        const float d = __ldg(ob + p * V + v) - y;
        s1[v] = fmaf(d, d, s1[v]);
        s2[v] = fmaf(y, y, s2[v]);
      }
    }
  }
  __shared__ float red[LOSS_THREADS / 32][LVMAX][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < LVMAX; ++v) {
    float a = s1[v], c = s2[v];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      c += __shfl_xor_sync(0xffffffffu, c, off);
    }
    if (lane == 0) { red[warp][v][0] = a; red[warp][v][1] = c; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * V) {
    const int v = threadIdx.x >> 1, k = threadIdx.x & 1;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += red[w][v][k];
    part[(((size_t)b * gridDim.x + blockIdx.x) * V + v) * 2 + k] = s;
  }
}

// coef[b][v] = 2 / (P * den * B * V);  loss = mean_{b,v} (num / den), den = 1e-7 + mean y^2
__global__ void __launch_bounds__(256)
nrmse_final_kernel(const float* __restrict__ part, float* __restrict__ coef, float* __restrict__ loss, long P, int V,
                   int B, int nblk) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < B * V; i += blockDim.x) {
    const int b = i / V, v = i - b * V;
    double num = 0.0, sq = 0.0;
    for (int k = 0; k < nblk; ++k) {
      num += (double)part[(((size_t)b * nblk + k) * V + v) * 2 + 0];
      sq += (double)part[(((size_t)b * nblk + k) * V + v) * 2 + 1];
    }
    const double den = 1e-7 + sq / (double)P;
    acc += (num / (double)P) / den;
    coef[i] = (float)(2.0 / ((double)P * den * (double)B * (double)V));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(red[0] / ((double)B * (double)V));
}

// dout[b,p,v] = gscale * coef[b][v] * (out - y)      (gscale = upstream gradient of the scalar loss)
__global__ void __launch_bounds__(256)
nrmse_grad_kernel(const float* __restrict__ out, const float* __restrict__ tgt, const float* __restrict__ coef,
                  const float* __restrict__ gscale, float* __restrict__ dout, long PV, int V, long total) {
  const float gs = gscale != nullptr ? __ldg(gscale) : 1.0f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / PV;
    const int v = (int)((i - b * PV) % V);
    dout[i] = gs * __ldg(coef + b * V + v) * (__ldg(out + i) - __ldg(tgt + i));
  }
}

// ------------------------------------------------------------------------------------------
// fused clip + Adam
// ------------------------------------------------------------------------------------------
constexpr int OPT_CHUNK = 4096;     // floats per chunk (one CTA)
constexpr int OPT_THREADS = 256;

struct OptChunk {                   // 40 bytes; built once on the host, lives in device memory
  float* p;
  const float* g;
  float* m;
  float* v;
  int n;
  float lr0;                        // base learning rate of the chunk's parameter group (0: hparams[0])
};

__global__ void __launch_bounds__(OPT_THREADS)
sqnorm_partial_kernel(const OptChunk* __restrict__ chunks, float* __restrict__ part) {
  const OptChunk c = chunks[blockIdx.x];
  float s = 0.f;
  for (int i = threadIdx.x; i < c.n; i += OPT_THREADS) {
    const float g = __ldg(c.g + i);
    s = fmaf(g, g, s);
  }
  __shared__ float red[OPT_THREADS / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
    part[blockIdx.x] = t;
  }
}

// state[0] = step count (as float, exact up to 2^24), [1] = total_norm, [2] = clip coefficient,
// [3] = lr of this step, [4] = 1 - beta1^t, [5] = 1 - beta2^t, [6] = clip_value, [7] = clipped norm
// hp: lr0, eta_min, T_max (<= 0: constant lr), beta1, beta2, eps, weight_decay, clip_floor, clip_frac,
//     sched_extra (scheduler steps taken in addition to one per optimizer step: the reference steps its
//     CosineAnnealingLR once more per epoch, fno/train.py:340), 1 - beta1, 1 - beta2 (rounded from double)
__global__ void __launch_bounds__(256)
opt_prepare_kernel(const float* __restrict__ part, int nparts, float* __restrict__ state, const float* __restrict__ hp) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += (double)part[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const double total = sqrt(red[0]);
  const double clip_value = fmax((double)hp[7], (double)hp[8] * total);   // max(5, 0.1 * total_norm)
  double coef = clip_value / (total + 1e-6);                              // clip_grad_norm_
  if (coef > 1.0) coef = 1.0;
  const double t_prev = (double)state[0];      // scheduler steps taken so far = optimizer steps so far
  const double t = t_prev + 1.0;
  double lr = (double)hp[0];
  if (hp[2] > 0.f)
    lr = (double)hp[1] + ((double)hp[0] - (double)hp[1]) * 0.5 * (1.0 + cos(M_PI * (t_prev + (double)hp[9]) / (double)hp[2]));
  state[0] = (float)t;
  state[1] = (float)total;
  state[2] = (float)coef;
  state[3] = (float)lr;
  state[4] = (float)(1.0 - pow((double)hp[3], t));
  state[5] = (float)(1.0 - pow((double)hp[4], t));
  state[6] = (float)clip_value;
  state[7] = (float)(total * coef);
}

__global__ void __launch_bounds__(OPT_THREADS)
adam_apply_kernel(const OptChunk* __restrict__ chunks, const float* __restrict__ state, const float* __restrict__ hp) {
  const OptChunk c = chunks[blockIdx.x];
  const float coef = state[2], bc1 = state[4], bc2 = state[5];
  const float b1 = hp[3], b2 = hp[4], eps = hp[5], wd = hp[6];
  // 1 - beta as torch has it: evaluated in double from the Python floats, THEN rounded (1.0f - 0.999f is off by 1.3e-5)
  const float omb1 = hp[10], omb2 = hp[11];
  // parameter groups differ in their base lr only (fno_aux/fno_train_aux.py:175-179); the schedule factor is shared:
  // lr_g = eta_min + (lr0_g - eta_min) * (lr - eta_min) / (lr0 - eta_min)
  float lr = state[3];
  if (c.lr0 > 0.f && c.lr0 != hp[0]) lr = (hp[0] != hp[1]) ? hp[1] + (c.lr0 - hp[1]) * ((lr - hp[1]) / (hp[0] - hp[1])) : c.lr0;
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (int i = threadIdx.x; i < c.n; i += OPT_THREADS) {
    const float p = c.p[i];
    float g = __ldg(c.g + i) * coef;
    g = fmaf(wd, p, g);                               // Adam(weight_decay): coupled L2
    const float m = fmaf(b1, c.m[i], omb1 * g);
    const float v = fmaf(b2, c.v[i], omb2 * g * g);
    c.m[i] = m;
    c.v[i] = v;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    c.p[i] = p - step_size * (m / denom);
  }
}

}  // namespace
}  // namespace fno

using namespace fno;

extern "C" size_t fno_nrmse_workspace_bytes(int B, int V) {
  if (B <= 0 || V <= 0) return 0;
  return sizeof(float) * (2ul * (size_t)B * LOSS_BLOCKS * V + (size_t)B * V);
}

extern "C" int fno_nrmse_fwd(const float* out, const float* target, float* loss, void* work, int B, long P, int V,
                             fno_stream_t stream) {
  if (!out || !target || !loss || !work || B <= 0 || B > 65535 || P <= 0 || V < 1 || V > LVMAX) {
    set_error("fno_nrmse_fwd: bad argument");
    return FNO_E_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(work);
  float* coef = part + 2ul * (size_t)B * LOSS_BLOCKS * V;
  nrmse_partial_kernel<<<dim3(LOSS_BLOCKS, B), LOSS_THREADS, 0, st>>>(out, target, part, P, V);
  count_launch();
  int rc = check_launch("nrmse_partial_kernel");
  if (rc != FNO_OK) return rc;
  nrmse_final_kernel<<<1, 256, 0, st>>>(part, coef, loss, P, V, B, LOSS_BLOCKS);
  count_launch();
  return check_launch("nrmse_final_kernel");
}

extern "C" int fno_nrmse_bwd(const float* out, const float* target, const void* work, const float* gscale,
                             float* dout, int B, long P, int V, fno_stream_t stream) {
  if (!out || !target || !work || !dout || B <= 0 || P <= 0 || V < 1 || V > LVMAX) {
    set_error("fno_nrmse_bwd: bad argument");
    return FNO_E_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* coef = static_cast<const float*>(work) + 2ul * (size_t)B * LOSS_BLOCKS * V;
  const long total = (long)B * P * V;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  nrmse_grad_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, target, coef, gscale, dout, P * V, V, total);
  count_launch();
  return check_launch("nrmse_grad_kernel");
}

extern "C" int fno_opt_chunk_floats(void) { return OPT_CHUNK; }
extern "C" size_t fno_opt_chunk_bytes(void) { return sizeof(OptChunk); }

extern "C" int fno_clip_adam_step(const void* chunks, int nchunks, float* partials, float* state, const float* hparams,
                                  fno_stream_t stream) {
  if (!chunks || nchunks <= 0 || !partials || !state || !hparams) {
    set_error("fno_clip_adam_step: bad argument");
    return FNO_E_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const OptChunk* ch = static_cast<const OptChunk*>(chunks);
  sqnorm_partial_kernel<<<nchunks, OPT_THREADS, 0, st>>>(ch, partials);
  count_launch();
  int rc = check_launch("sqnorm_partial_kernel");
  if (rc != FNO_OK) return rc;
  opt_prepare_kernel<<<1, 256, 0, st>>>(partials, nchunks, state, hparams);
  count_launch();
  rc = check_launch("opt_prepare_kernel");
  if (rc != FNO_OK) return rc;
  adam_apply_kernel<<<nchunks, OPT_THREADS, 0, st>>>(ch, state, hparams);
  count_launch();
  return check_launch("adam_apply_kernel");
}
