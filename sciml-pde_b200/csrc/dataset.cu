// Device-resident windowed dataset.  The reference loaders re-open the HDF5 file and read
// the WHOLE trajectory for every item (fno/utils_2d_rd_baseline.py:74-86), slice the window on the host
// and rebuild the grid per sample (:97-102).  Here the trajectories live on the GPU in a time-inner layout
//     traj [n_traj, pixels, T, V]
// so that the window of item (trajectory i, start s) is, per pixel, ONE contiguous run of
// (initial_step + rollout) * V floats starting at ((i * pixels + p) * T + s) * V, and a batch is
// gathered by a single streaming copy kernel straight into the layouts the lift kernel and the loss
// read:  xx [B, pixels, initial_step, V],  yy [B, pixels, rollout, V].
#include "common.cuh"

namespace fno {
namespace {

__global__ void __launch_bounds__(256)
window_gather_kernel(const float* __restrict__ traj, const long long* __restrict__ traj_idx, const int* __restrict__ t_start,
                     float* __restrict__ xx, float* __restrict__ yy, long long total_x, long long total, int npix, int T,
                     int V, int XE, int YE, FastDiv by_xe, FastDiv by_ye, FastDiv by_pix) {
  // one thread per OUTPUT element (writes fully coalesced; reads are contiguous runs of XE / YE floats)
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const bool is_x = o < total_x;
    const unsigned r = (unsigned)(is_x ? o : o - total_x);          // < 2^32 (checked on the host)
    const unsigned pixel = is_x ? by_xe.div(r) : by_ye.div(r);      // b * npix + p
    const unsigned e = r - pixel * (unsigned)(is_x ? XE : YE);
    const unsigned b = by_pix.div(pixel);
    const unsigned p = pixel - b * (unsigned)npix;
    const long long ti = __ldg(traj_idx + b);
    const int ts = __ldg(t_start + b) + (is_x ? 0 : XE / V);
    const float v = __ldg(traj + ((size_t)(ti * npix + p) * T + ts) * V + e);
    if (is_x) xx[o] = v;
    else yy[o - total_x] = v;
  }
}

}  // namespace
}  // namespace fno

using namespace fno;

extern "C" int fno_window_gather(const float* traj, const long long* traj_idx, const int* t_start, float* xx, float* yy,
                                 int B, long npix, int T, int V, int initial_step, int rollout, fno_stream_t stream) {
  if (!traj || !traj_idx || !t_start || !xx || !yy || B <= 0 || npix <= 0 || T <= 0 || V <= 0 || initial_step <= 0 ||
      rollout <= 0 || initial_step + rollout > T) {
    set_error("fno_window_gather: bad argument");
    return FNO_E_ARG;
  }
  const long long total_x = (long long)B * npix * initial_step * V;
  const long long total_y = (long long)B * npix * rollout * V;
  if (total_x >= (1ll << 32) || total_y >= (1ll << 32) || npix >= (1l << 31)) {
    set_error("fno_window_gather: batch too large (element counts must fit 32 bits)");
    return FNO_E_ARG;
  }
  FastDiv by_xe, by_ye, by_pix;
  by_xe.init((unsigned)(initial_step * V));
  by_ye.init((unsigned)(rollout * V));
  by_pix.init((unsigned)npix);
  const long long total = total_x + total_y;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  window_gather_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      traj, traj_idx, t_start, xx, yy, total_x, total, (int)npix, T, V, initial_step * V, rollout * V, by_xe, by_ye, by_pix);
  count_launch();
  return check_launch("window_gather_kernel");
}
