// ssim.cu -- the SSIM term of the reference's training loss on the device (src/utils/losses.py:72-95):
//   ssim_loss = 1 - mean over (image, channel 0..1, valid 11x11 window) of S(scaled prediction, scaled target)
// and its gradient with respect to the raw prediction.  The arithmetic lives in ssim_core.h (shared with the host
// harness the tests check against torch autograd); the kernels below only distribute it:
//   forward : one thread per window -> S summed per block (double) + the three partial-derivative maps a, b, c
//   backward: one thread per prediction pixel gathers G^T * (a, b, c) over the <= 121 windows that contain it
// 250 x 250 tiles, B = 16: 1.8 M windows x 121 taps x 5 moments = 1.1 GFLOP forward, 0.7 GFLOP backward -- fp32 FMA
// work of ~0.1 ms next to an 8 ms step, so the direct (non-separable) form is kept for its simplicity.
#ifndef MAU_KERNEL_ENV            // a test harness may supply the execution environment instead (it then defines this,
#include "ops.h"                  // MAU_LAUNCH, MAU_CUDA, MAU_LAUNCHED, fail and ceil_div before including this file)
#define MAU_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif
#include "ssim_core.h"

namespace mau {
namespace {

struct Window {
  float g[mau_ssim::kWin];
};

// grid (ceil(Hv*Wv / 256), B*2), block 256
__global__ void __launch_bounds__(256) ssim_forward_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int C,
                                                           int H, int W, Window win, float* __restrict__ a,
                                                           float* __restrict__ b, float* __restrict__ c,
                                                           double* __restrict__ acc) {
  __shared__ double red[256];
  const int tid = threadIdx.x, plane2 = blockIdx.y, bi = plane2 >> 1, ch = plane2 & 1;
  const int Hv = H - mau_ssim::kWin + 1, Wv = W - mau_ssim::kWin + 1;
  const float* x = pred + ((long long)bi * C + ch) * H * W;
  const float* y = tgt + ((long long)bi * C + ch) * H * W;
  const int o = blockIdx.x * blockDim.x + tid;
  double s = 0.0;
  if (o < Hv * Wv) {
    const int i = o / Wv, j = o - i * Wv;
    const mau_ssim::Point p = mau_ssim::window(x, y, W, i, j, ch, win.g);
    const long long at = (long long)plane2 * Hv * Wv + o;
    a[at] = p.a;
    b[at] = p.b;
    c[at] = p.c;
    s = (double)p.s;
  }
  red[tid] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) red[tid] += red[tid + off];
    __syncthreads();
  }
  if (tid == 0) atomicAdd(acc, red[0]);
}

// grid (ceil(H*W / 256), B*C), block 256
__global__ void __launch_bounds__(256) ssim_backward_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int C,
                                                            int H, int W, Window win, const float* __restrict__ a,
                                                            const float* __restrict__ b, const float* __restrict__ c,
                                                            float coef, float* __restrict__ grad) {
  const int plane = blockIdx.y, bi = plane / C, ch = plane - bi * C;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  float* gp = grad + (long long)plane * H * W;
  if (ch >= 2) {   // the reference stacks channels 0 and 1 only
    gp[p] = 0.f;
    return;
  }
  const int Hv = H - mau_ssim::kWin + 1, Wv = W - mau_ssim::kWin + 1;
  const long long base = ((long long)bi * 2 + ch) * Hv * Wv;
  const float* x = pred + (long long)plane * H * W;
  const float* y = tgt + (long long)plane * H * W;
  const int yy = p / W, xx = p - yy * W;
  gp[p] = coef * mau_ssim::gather_grad(a + base, b + base, c + base, Hv, Wv, x, y, W, yy, xx, ch, win.g);
}

__global__ void ssim_finalize_kernel(const double* __restrict__ acc, double inv_n, float* __restrict__ loss) {
  loss[0] = (float)(1.0 - acc[0] * inv_n);
}

}  // namespace

long long ssim_work_floats(int B, int H, int W) {
  const long long Hv = H - mau_ssim::kWin + 1, Wv = W - mau_ssim::kWin + 1;
  return Hv > 0 && Wv > 0 ? 3ll * B * 2 * Hv * Wv : 0;
}

int op_ssim_loss(const float* pred, const float* tgt, int B, int C, int H, int W, float* loss, float* grad, float* work,
                 double* acc, cudaStream_t st) {
  using mau_ssim::kWin;
  if (B < 1 || C < 2) return fail("ssim_loss: needs B >= 1 and the two target channels (NDVI, temperature), got B=%d C=%d", B, C);
  if (H < kWin || W < kWin) return fail("ssim_loss: Kernel size can't be greater than actual input size (%d x %d < %d)", H, W, kWin);
  if (std::min(H, W) >= 384) return fail("ssim_loss: tiles of %d x %d are average-pooled by piq before SSIM; not implemented", H, W);
  if ((long long)B * C > 65535) return fail("ssim_loss: B*C = %lld exceeds the grid limit 65535", (long long)B * C);
  const int Hv = H - kWin + 1, Wv = W - kWin + 1;
  const long long nwin_plane = (long long)Hv * Wv, nwin = (long long)B * 2 * nwin_plane;
  Window win;
  mau_ssim::gaussian_window(win.g);
  float *a = work, *b = work + nwin, *c = work + 2 * nwin;
  MAU_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  MAU_LAUNCH(ssim_forward_kernel, dim3((unsigned)ceil_div((int)nwin_plane, 256), (unsigned)(B * 2), 1), dim3(256), st, pred, tgt, C, H, W, win, a,
             b, c, acc);
  MAU_LAUNCHED();
  MAU_LAUNCH(ssim_finalize_kernel, dim3(1), dim3(1), st, acc, 1.0 / (double)nwin, loss);
  MAU_LAUNCHED();
  if (grad) {
    MAU_LAUNCH(ssim_backward_kernel, dim3((unsigned)ceil_div(H * W, 256), (unsigned)(B * C), 1), dim3(256), st, pred, tgt, C, H, W, win, a, b, c,
               (float)(-1.0 / (double)nwin), grad);
    MAU_LAUNCHED();
  }
  return 0;
}

}  // namespace mau
