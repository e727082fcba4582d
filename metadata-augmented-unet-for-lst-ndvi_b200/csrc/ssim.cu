// ssim.cu -- the SSIM term of the reference's training loss on the device (src/utils/losses.py:72-95):
//   ssim_loss = 1 - mean over (image, channel 0..1, valid 11x11 window) of S(scaled prediction, scaled target)
// and its gradient with respect to the raw prediction.  The per-window arithmetic lives in ssim_core.h (shared with
// the host harness the tests check against torch autograd).  The Gaussian window is an outer product, so both kernels
// filter separably on a shared-memory tile (the first version evaluated all 121 taps per window / per pixel and cost
// 0.54 ms of an 8.4 ms training step at 16 x 2 x 250 x 250; profiles/r02_ssim.md):
//   forward : block = 16 x 32 windows.  (26 x 42) scaled x / y values -> shared memory; horizontal 11-tap pass for the
//             five moments (x, y, x^2, y^2, xy) on 26 rows; vertical 11-tap pass + S and the three partial-derivative
//             maps a, b, c per window; S summed per block in double.
//   backward: block = 16 x 32 pixels.  d(sum S)/d pixel = G^T a + 2 x G^T b + y G^T c: the (26 x 42) neighbourhood of
//             a, b, c (zero outside the valid windows) -> shared memory, the same two passes, then the chain rule of
//             the channel scaling.  An upstream scalar (autograd's d L / d ssim_loss) is read from device memory.
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

constexpr int kTH = 16, kTW = 32;                       // outputs per block (windows / pixels)
constexpr int kIH = kTH + mau_ssim::kWin - 1;           // 26 input rows
constexpr int kIW = kTW + mau_ssim::kWin - 1;           // 42 input columns

// grid (ceil(Wv / 32), ceil(Hv / 16), B * 2), block 256
__global__ void __launch_bounds__(256) ssim_forward_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int C,
                                                           int H, int W, int prescaled, Window win, float* __restrict__ a,
                                                           float* __restrict__ b, float* __restrict__ c,
                                                           double* __restrict__ acc) {
  __shared__ float sx[kIH][kIW], sy[kIH][kIW];
  __shared__ float hm[5][kIH][kTW];
  __shared__ double red[256];
  const int tid = threadIdx.x, plane2 = blockIdx.z, bi = plane2 >> 1, ch = plane2 & 1;
  const int Hv = H - mau_ssim::kWin + 1, Wv = W - mau_ssim::kWin + 1;
  const int i0 = blockIdx.y * kTH, j0 = blockIdx.x * kTW;
  const float* x = pred + ((long long)bi * C + ch) * H * W;
  const float* y = tgt + ((long long)bi * C + ch) * H * W;
  const int sch = prescaled ? mau_ssim::kPrescaled : ch;
  for (int e = tid; e < kIH * kIW; e += 256) {
    const int r = e / kIW, q = e - r * kIW;
    const int gi = i0 + r, gj = j0 + q;
    const bool in = gi < H && gj < W;
    sx[r][q] = in ? mau_ssim::scale_value(x[(long long)gi * W + gj], sch) : 0.f;
    sy[r][q] = in ? mau_ssim::scale_value(y[(long long)gi * W + gj], sch) : 0.f;
  }
  __syncthreads();
  for (int e = tid; e < kIH * kTW; e += 256) {          // horizontal pass: five moments per (row, output column)
    const int r = e / kTW, q = e - r * kTW;
    float mx = 0.f, my = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
    for (int k = 0; k < mau_ssim::kWin; ++k) {
      const float w = win.g[k], xv = sx[r][q + k], yv = sy[r][q + k];
      mx += w * xv;
      my += w * yv;
      xx += w * xv * xv;
      yy += w * yv * yv;
      xy += w * xv * yv;
    }
    hm[0][r][q] = mx; hm[1][r][q] = my; hm[2][r][q] = xx; hm[3][r][q] = yy; hm[4][r][q] = xy;
  }
  __syncthreads();
  double s = 0.0;
  for (int e = tid; e < kTH * kTW; e += 256) {          // vertical pass + the window's SSIM and derivative maps
    const int r = e / kTW, q = e - r * kTW;
    const int wi = i0 + r, wj = j0 + q;
    if (wi < Hv && wj < Wv) {
      float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < mau_ssim::kWin; ++k) {
        const float w = win.g[k];
#pragma unroll
        for (int t = 0; t < 5; ++t) m[t] += w * hm[t][r + k][q];
      }
      const mau_ssim::Point p = mau_ssim::point_from_moments(m[0], m[1], m[2], m[3], m[4]);
      const long long at = (long long)plane2 * Hv * Wv + (long long)wi * Wv + wj;
      a[at] = p.a;
      b[at] = p.b;
      c[at] = p.c;
      s += (double)p.s;
    }
  }
  red[tid] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) red[tid] += red[tid + off];
    __syncthreads();
  }
  if (tid == 0) atomicAdd(acc, red[0]);
}

// grid (ceil(W / 32), ceil(H / 16), B * C), block 256.  up (nullable): device scalar multiplied into the result.
__global__ void __launch_bounds__(256) ssim_backward_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int C,
                                                            int H, int W, int prescaled, Window win, const float* __restrict__ a,
                                                            const float* __restrict__ b, const float* __restrict__ c,
                                                            float coef, const float* __restrict__ up, float* __restrict__ grad) {
  __shared__ float sm[3][kIH][kIW];
  __shared__ float hz[3][kIH][kTW];
  const int tid = threadIdx.x, plane = blockIdx.z, bi = plane / C, ch = plane - bi * C;
  const int y0 = blockIdx.y * kTH, x0 = blockIdx.x * kTW;
  float* gp = grad + (long long)plane * H * W;
  if (ch >= 2) {   // the reference stacks channels 0 and 1 only (uniform per block: no barrier is skipped by part of a block)
    for (int e = tid; e < kTH * kTW; e += 256) {
      const int yy = y0 + e / kTW, xx = x0 + e % kTW;
      if (yy < H && xx < W) gp[(long long)yy * W + xx] = 0.f;
    }
    return;
  }
  const int Hv = H - mau_ssim::kWin + 1, Wv = W - mau_ssim::kWin + 1;
  const long long base = ((long long)bi * 2 + ch) * Hv * Wv;
  // windows that contain pixel (yy, xx): rows yy - 10 .. yy, columns xx - 10 .. xx  ->  tile origin (y0 - 10, x0 - 10)
  for (int e = tid; e < kIH * kIW; e += 256) {
    const int r = e / kIW, q = e - r * kIW;
    const int wi = y0 - (mau_ssim::kWin - 1) + r, wj = x0 - (mau_ssim::kWin - 1) + q;
    const bool in = wi >= 0 && wi < Hv && wj >= 0 && wj < Wv;
    const long long o = base + (long long)wi * Wv + wj;
    sm[0][r][q] = in ? a[o] : 0.f;
    sm[1][r][q] = in ? b[o] : 0.f;
    sm[2][r][q] = in ? c[o] : 0.f;
  }
  __syncthreads();
  // sum_k g[k] * v[xx - k] over the tile = sum_k g[10 - k] * sm[..][q + k]: the window is symmetric, g[10 - k] == g[k]
  for (int e = tid; e < kIH * kTW; e += 256) {
    const int r = e / kTW, q = e - r * kTW;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < mau_ssim::kWin; ++k) {
      const float w = win.g[mau_ssim::kWin - 1 - k];
      t0 += w * sm[0][r][q + k];
      t1 += w * sm[1][r][q + k];
      t2 += w * sm[2][r][q + k];
    }
    hz[0][r][q] = t0; hz[1][r][q] = t1; hz[2][r][q] = t2;
  }
  __syncthreads();
  const float scale = up ? coef * up[0] : coef;
  const float* x = pred + (long long)plane * H * W;
  const float* y = tgt + (long long)plane * H * W;
  for (int e = tid; e < kTH * kTW; e += 256) {
    const int r = e / kTW, q = e - r * kTW;
    const int yy = y0 + r, xx = x0 + q;
    if (yy < H && xx < W) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int k = 0; k < mau_ssim::kWin; ++k) {
        const float w = win.g[mau_ssim::kWin - 1 - k];
        t0 += w * hz[0][r + k][q];
        t1 += w * hz[1][r + k][q];
        t2 += w * hz[2][r + k][q];
      }
      const long long o = (long long)yy * W + xx;
      gp[o] = scale * mau_ssim::combine_grad(t0, t1, t2, x[o], y[o], prescaled ? mau_ssim::kPrescaled : ch);
    }
  }
}

// piq's down-sampling of large tiles: out[b, ch, i, j] = mean over the f x f block of the SCALED raw values (ch < 2);
// xs / ys: [B, 2, Hp, Wp].  grid (ceil(Hp * Wp / 256), B * 2), block 256
__global__ void __launch_bounds__(256) ssim_pool_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int C, int H,
                                                        int W, int f, int Hp, int Wp, float* __restrict__ xs, float* __restrict__ ys) {
  const int plane2 = blockIdx.y, bi = plane2 >> 1, ch = plane2 & 1;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= Hp * Wp) return;
  const int i = o / Wp, j = o - i * Wp;
  const float* x = pred + ((long long)bi * C + ch) * H * W;
  const float* y = tgt + ((long long)bi * C + ch) * H * W;
  float sx = 0.f, sy = 0.f;
  for (int di = 0; di < f; ++di)
    for (int dj = 0; dj < f; ++dj) {
      const long long at = (long long)(i * f + di) * W + j * f + dj;
      sx += mau_ssim::scale_value(x[at], ch);
      sy += mau_ssim::scale_value(y[at], ch);
    }
  const float inv = 1.f / (float)(f * f);
  xs[(long long)plane2 * Hp * Wp + o] = sx * inv;
  ys[(long long)plane2 * Hp * Wp + o] = sy * inv;
}

// chain rule of the pooling and of the channel scaling: grad[b, ch, h, w] = up * gp[b, ch, h / f, w / f] / f^2 * slope
// inside the pooled region, 0 outside it (rows / columns avg_pool2d drops) and for channels >= 2.
// grid (ceil(H * W / 256), B * C), block 256
__global__ void __launch_bounds__(256) ssim_unpool_kernel(const float* __restrict__ pred, int C, int H, int W, int f, int Hp, int Wp,
                                                          const float* __restrict__ gpool, const float* __restrict__ up,
                                                          float* __restrict__ grad) {
  const int plane = blockIdx.y, bi = plane / C, ch = plane - bi * C;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int h = p / W, w = p - h * W;
  float g = 0.f;
  if (ch < 2 && h < Hp * f && w < Wp * f) {
    const float scale = (up ? up[0] : 1.f) / (float)(f * f);
    g = scale * gpool[((long long)(bi * 2 + ch) * Hp + h / f) * Wp + w / f] *
        mau_ssim::scale_slope(pred[(long long)plane * H * W + p], ch);
  }
  grad[(long long)plane * H * W + p] = g;
}

__global__ void ssim_finalize_kernel(const double* __restrict__ acc, double inv_n, float* __restrict__ loss) {
  loss[0] = (float)(1.0 - acc[0] * inv_n);
}

int check_shape(int B, int C, int H, int W) {
  using mau_ssim::kWin;
  if (B < 1 || C < 2) return fail("ssim_loss: needs B >= 1 and the two target channels (NDVI, temperature), got B=%d C=%d", B, C);
  if (H < kWin || W < kWin) return fail("ssim_loss: Kernel size can't be greater than actual input size (%d x %d < %d)", H, W, kWin);
  if ((long long)B * C > 65535) return fail("ssim_loss: B*C = %lld exceeds the grid limit 65535", (long long)B * C);
  return 0;
}

}  // namespace

static int g_force_pool = 0;      // tests: exercise the pooled path on small tiles (0 = piq's rule)
void ssim_debug_force_pool(int f) { g_force_pool = f; }
static int pool_f(int H, int W) { return g_force_pool > 0 ? g_force_pool : mau_ssim::pool_factor(H, W); }

// a, b, c maps of the (pooled) planes; with pooling also the pooled scaled planes xs, ys and the pooled gradient
long long ssim_work_floats(int B, int H, int W) {
  const int f = pool_f(H, W);
  const long long Hp = H / f, Wp = W / f;
  const long long Hv = Hp - mau_ssim::kWin + 1, Wv = Wp - mau_ssim::kWin + 1;
  if (Hv <= 0 || Wv <= 0) return 0;
  return 3ll * B * 2 * Hv * Wv + (f > 1 ? 3ll * B * 2 * Hp * Wp : 0);
}

// loss[0] = 1 - mean SSIM; work receives the derivative maps a, b, c the backward needs
int op_ssim_forward(const float* pred, const float* tgt, int B, int C, int H, int W, float* loss, float* work, double* acc,
                    cudaStream_t st) {
  using mau_ssim::kWin;
  if (int rc = check_shape(B, C, H, W)) return rc;
  const int f = pool_f(H, W);
  const int Hp = H / f, Wp = W / f;
  if (Hp < kWin || Wp < kWin) return fail("ssim_loss: Kernel size can't be greater than actual input size (%d x %d pooled by %d)", H, W, f);
  const int Hv = Hp - kWin + 1, Wv = Wp - kWin + 1;
  const long long nwin = (long long)B * 2 * Hv * Wv;
  Window win;
  mau_ssim::gaussian_window(win.g);
  float *a = work, *b = work + nwin, *c = work + 2 * nwin;
  const float *px = pred, *py = tgt;
  int Cc = C, pres = 0;
  if (f > 1) {
    float* xs = work + 3 * nwin;
    float* ys = xs + (long long)B * 2 * Hp * Wp;
    MAU_LAUNCH(ssim_pool_kernel, dim3((unsigned)ceil_div(Hp * Wp, 256), (unsigned)(B * 2), 1), dim3(256), st, pred, tgt, C, H, W, f, Hp, Wp, xs, ys);
    MAU_LAUNCHED();
    px = xs; py = ys; Cc = 2; pres = 1;
  }
  MAU_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  MAU_LAUNCH(ssim_forward_kernel, dim3((unsigned)ceil_div(Wv, kTW), (unsigned)ceil_div(Hv, kTH), (unsigned)(B * 2)), dim3(256), st, px, py,
             Cc, Hp, Wp, pres, win, a, b, c, acc);
  MAU_LAUNCHED();
  MAU_LAUNCH(ssim_finalize_kernel, dim3(1), dim3(1), st, acc, 1.0 / (double)nwin, loss);
  MAU_LAUNCHED();
  return 0;
}

// grad = upstream * d loss / d pred from the maps op_ssim_forward left in work (upstream: nullable device scalar)
int op_ssim_backward(const float* pred, const float* tgt, int B, int C, int H, int W, const float* work, const float* upstream,
                     float* grad, cudaStream_t st) {
  using mau_ssim::kWin;
  if (int rc = check_shape(B, C, H, W)) return rc;
  const int f = pool_f(H, W);
  const int Hp = H / f, Wp = W / f;
  if (Hp < kWin || Wp < kWin) return fail("ssim_loss: Kernel size can't be greater than actual input size (%d x %d pooled by %d)", H, W, f);
  const int Hv = Hp - kWin + 1, Wv = Wp - kWin + 1;
  const long long nwin = (long long)B * 2 * Hv * Wv;
  Window win;
  mau_ssim::gaussian_window(win.g);
  const float *a = work, *b = work + nwin, *c = work + 2 * nwin;
  const float coef = (float)(-1.0 / (double)nwin);
  if (f == 1) {
    MAU_LAUNCH(ssim_backward_kernel, dim3((unsigned)ceil_div(W, kTW), (unsigned)ceil_div(H, kTH), (unsigned)(B * C)), dim3(256), st, pred, tgt, C,
               H, W, 0, win, a, b, c, coef, upstream, grad);
    MAU_LAUNCHED();
    return 0;
  }
  const float* xs = work + 3 * nwin;
  const float* ys = xs + (long long)B * 2 * Hp * Wp;
  float* gpool = const_cast<float*>(ys) + (long long)B * 2 * Hp * Wp;
  MAU_LAUNCH(ssim_backward_kernel, dim3((unsigned)ceil_div(Wp, kTW), (unsigned)ceil_div(Hp, kTH), (unsigned)(B * 2)), dim3(256), st, xs, ys, 2,
             Hp, Wp, 1, win, a, b, c, coef, static_cast<const float*>(nullptr), gpool);
  MAU_LAUNCHED();
  MAU_LAUNCH(ssim_unpool_kernel, dim3((unsigned)ceil_div(H * W, 256), (unsigned)(B * C), 1), dim3(256), st, pred, C, H, W, f, Hp, Wp,
             static_cast<const float*>(gpool), upstream, grad);
  MAU_LAUNCHED();
  return 0;
}

int op_ssim_loss(const float* pred, const float* tgt, int B, int C, int H, int W, float* loss, float* grad, float* work,
                 double* acc, cudaStream_t st) {
  if (int rc = op_ssim_forward(pred, tgt, B, C, H, W, loss, work, acc, st)) return rc;
  if (grad) return op_ssim_backward(pred, tgt, B, C, H, W, work, nullptr, grad, st);
  return 0;
}

}  // namespace mau
