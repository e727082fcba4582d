// loss.cu -- training loss terms that sit directly on the model output (fp32 NCHW):
//   pixel   = mean |p - t|            (F.l1_loss,  reference src/utils/losses.py:67)
//           | mean (p - t)^2          (F.mse_loss, reference src/utils/losses.py:33)
//   gradient= mean| |dy p| - |dy t| | + mean| |dx p| - |dx t| |   (src/utils/losses.py:5-25)
//   total   = pixel + lambda * gradient
// One pass produces the three sums and (optionally) d total / d p in gather form (no atomics on
// the gradient).  SSIM (piq, third party) is not part of this kernel.
#include "ops.h"
#include "vec.cuh"
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

namespace mau {
namespace {

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

// partials: [gridDim.x][3] doubles
// grad[i] = wp * d pixel / d p_i + wg * d gradient / d p_i with wp = wp_h + [*g_total] + [*g_pixel] and
// wg = wg_h + [*g_total] * lambda_total + [*g_grad]: the device scalars are the upstream gradients autograd hands to the
// backward of {total, pixel, gradient} (total = pixel + lambda_total * gradient), folded in here instead of a separate
// element-wise multiply.  partials may be null (backward-only launch).
__global__ void __launch_bounds__(256) loss_kernel(int kind, const float* __restrict__ p, const float* __restrict__ t,
                                                   int planes, int H, int W, float inv_n, float inv_ny, float inv_nx,
                                                   float wp_h, float wg_h, float lambda_total,
                                                   const float* __restrict__ g_total, const float* __restrict__ g_pixel,
                                                   const float* __restrict__ g_grad, double* __restrict__ partials,
                                                   float* __restrict__ grad) {
  float wp = wp_h, lambda = wg_h;
  if (g_total) { const float g = *g_total; wp += g; lambda += g * lambda_total; }
  if (g_pixel) wp += *g_pixel;
  if (g_grad) lambda += *g_grad;
  inv_n *= wp;        // only the gradient uses inv_n / lambda below; the sums are normalised by loss_finalize_kernel
  // fp32 partial sums per thread (a thread sees at most a few hundred elements), fp64 only across threads:
  // B200 issues FP64 adds at a small fraction of the FP32 rate and this loop was bound by them
  float f_pix = 0.f, f_dy = 0.f, f_dx = 0.f;
  // a block works on whole image rows (h known per row, w = thread offset): no division per element
  const long long rows = (long long)planes * H;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x)
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    const int h = (int)(row % H);
    const long long i = row * W + w;
    const float pc = p[i], tc = t[i];
    const float d = pc - tc;
    float g;
    if (kind == 0) { f_pix += fabsf(d); g = sgn(d) * inv_n; }
    else           { f_pix = fmaf(d, d, f_pix); g = 2.f * d * inv_n; }
    if (!partials && lambda == 0.f) {       // backward-only launch without a gradient term (criterion "l1" / "mse"): no neighbours
      grad[i] = g;
      continue;
    }
    if (h + 1 < H) {   // pair (h, h+1): this pixel is the upper one
      const float a = p[i + W] - pc, b = t[i + W] - tc;
      const float u = fabsf(a) - fabsf(b);
      f_dy += fabsf(u);
      g -= lambda * inv_ny * sgn(u) * sgn(a);
    }
    if (h > 0) {       // pair (h-1, h): this pixel is the lower one
      const float a = pc - p[i - W], b = tc - t[i - W];
      g += lambda * inv_ny * sgn(fabsf(a) - fabsf(b)) * sgn(a);
    }
    if (w + 1 < W) {
      const float a = p[i + 1] - pc, b = t[i + 1] - tc;
      const float u = fabsf(a) - fabsf(b);
      f_dx += fabsf(u);
      g -= lambda * inv_nx * sgn(u) * sgn(a);
    }
    if (w > 0) {
      const float a = pc - p[i - 1], b = tc - t[i - 1];
      g += lambda * inv_nx * sgn(fabsf(a) - fabsf(b)) * sgn(a);
    }
    if (grad) grad[i] = g;
  }
  if (!partials) return;
  __shared__ double red[3][8];
  double s_pix = (double)f_pix, s_dy = (double)f_dy, s_dx = (double)f_dx;
  s_pix = warp_sum(s_pix); s_dy = warp_sum(s_dy); s_dx = warp_sum(s_dx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s_pix; red[1][warp] = s_dy; red[2][warp] = s_dx; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    partials[blockIdx.x * 3 + threadIdx.x] = s;
  }
}
__global__ void loss_finalize_kernel(const double* partials, int nblocks, double inv_n, double inv_ny, double inv_nx,
                                     float lambda, float* losses) {
  __shared__ double red[3][32];
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
    for (int q = 0; q < 3; ++q) s[q] += partials[i * 3 + q];
  for (int q = 0; q < 3; ++q) s[q] = warp_sum(s[q]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) for (int q = 0; q < 3; ++q) red[q][warp] = s[q];
  __syncthreads();
  if (threadIdx.x == 0) {
    double a[3] = {0.0, 0.0, 0.0};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) for (int q = 0; q < 3; ++q) a[q] += red[q][w];
    const double pixel = a[0] * inv_n;
    const double gradl = a[1] * inv_ny + a[2] * inv_nx;
    losses[0] = (float)(pixel + (double)lambda * gradl);
    losses[1] = (float)pixel;
    losses[2] = (float)gradl;
    losses[3] = 0.f;
  }
}

}  // namespace

// per-(device, stream) partial-sum scratch, allocated once: a stream-ordered cudaMallocAsync/cudaFreeAsync
// pair per call made the pool trim and re-grow at every synchronisation
constexpr int kMaxLossBlocks = 148 * 8;
static double* loss_scratch(cudaStream_t st) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, double*> pool;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  double*& p = pool[{dev, st}];
  if (!p && cudaMalloc(reinterpret_cast<void**>(&p), sizeof(double) * 3 * kMaxLossBlocks) != cudaSuccess) p = nullptr;
  return p;
}

int op_loss(int kind, const float* pred, const float* tgt, int B, int C, int H, int W, float lambda_grad,
            float* losses, float* grad, cudaStream_t st) {
  if (kind != 0 && kind != 1) return fail("loss: kind must be 0 (L1) or 1 (MSE)");
  const long long n = (long long)B * C * H * W;
  const long long ny = (long long)B * C * (H - 1) * W, nx = (long long)B * C * H * (W - 1);
  if (n <= 0 || ny <= 0 || nx <= 0) return fail("loss: empty tensor");
  int blocks = (int)std::min<long long>((long long)B * C * H, kMaxLossBlocks);
  double* partials = loss_scratch(st);
  if (!partials) return fail("loss: scratch allocation failed");
  loss_kernel<<<blocks, 256, 0, st>>>(kind, pred, tgt, B * C, H, W, 1.f / (float)n, 1.f / (float)ny, 1.f / (float)nx,
                                      1.f, lambda_grad, 0.f, nullptr, nullptr, nullptr, partials, grad);
  MAU_LAUNCHED();
  loss_finalize_kernel<<<1, 256, 0, st>>>(partials, blocks, 1.0 / (double)n, 1.0 / (double)ny, 1.0 / (double)nx,
                                          lambda_grad, losses);
  MAU_LAUNCHED();
  return 0;
}

int op_loss_backward(int kind, const float* pred, const float* tgt, int B, int C, int H, int W, float lambda_total,
                     const float* g_total, const float* g_pixel, const float* g_grad, float* grad, cudaStream_t st) {
  if (kind != 0 && kind != 1) return fail("loss: kind must be 0 (L1) or 1 (MSE)");
  if (!grad) return fail("loss backward: null gradient");
  const long long n = (long long)B * C * H * W;
  const long long ny = (long long)B * C * (H - 1) * W, nx = (long long)B * C * H * (W - 1);
  if (n <= 0 || ny <= 0 || nx <= 0) return fail("loss: empty tensor");
  const int blocks = (int)std::min<long long>((long long)B * C * H, kMaxLossBlocks);
  loss_kernel<<<blocks, 256, 0, st>>>(kind, pred, tgt, B * C, H, W, 1.f / (float)n, 1.f / (float)ny, 1.f / (float)nx,
                                      0.f, 0.f, lambda_total, g_total, g_pixel, g_grad, nullptr, grad);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
