// ssim_core.h -- arithmetic of the SSIM loss term, shared by the CUDA kernels (ssim.cu) and by the host harness
// the tests compile from this same header (oracle/ssim_host.cpp), so the formulas are checked on the CPU against
// torch autograd before they ever run on a GPU.
//
// Reference call site: src/utils/losses.py:72-90
//     targets_scaled = stack([(t[:,0] + 1) / 2, clamp(t[:,1], 0, 1)]);  outputs_scaled likewise
//     ssim_vals = piq.ssim(outputs_scaled, targets_scaled, data_range=1.0, reduction='none')
//     ssim_loss = 1 - mean(ssim_vals)
// piq (third party, unpinned in requirements.txt:9, absent from the build image -- parity UNPINNED) publishes the
// algorithm of Wang et al. 2004 as: an 11 x 11 Gaussian window (sigma 1.5, normalised), 'valid' filtering
// (no padding), k1 = 0.01, k2 = 0.03, images average-pooled by f = round(min(H,W)/256) first when f > 1 (min(H,W) >= 384),
//     mu_x = G*x, mu_y = G*y, s_xx = G*x^2 - mu_x^2, s_yy = G*y^2 - mu_y^2, s_xy = G*xy - mu_x mu_y
//     S = (2 mu_x mu_y + c1)(2 s_xy + c2) / ((mu_x^2 + mu_y^2 + c1)(s_xx + s_yy + c2)),   c1 = k1^2, c2 = k2^2
// per-image value = mean of S over the valid window positions and over the two channels.
#ifndef MAU_SSIM_CORE_H_
#define MAU_SSIM_CORE_H_

#include <math.h>

#ifdef __CUDACC__
#define MAU_HD __host__ __device__ __forceinline__
#else
#define MAU_HD inline
#endif

namespace mau_ssim {

constexpr int kWin = 11;                       // piq default kernel_size
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// normalised 1-D Gaussian, sigma = 1.5; the 2-D window is its outer product (piq.functional.gaussian_filter)
inline void gaussian_window(float g[kWin]) {
  double s = 0.0, t[kWin];
  for (int i = 0; i < kWin; ++i) {
    const double c = i - (kWin - 1) / 2.0;
    t[i] = exp(-(c * c) / (2.0 * 1.5 * 1.5));
    s += t[i];
  }
  for (int i = 0; i < kWin; ++i) g[i] = (float)(t[i] / s);
}

// the reference scales the NDVI channel to [0,1] and clamps the (normalised) temperature channel to [0,1];
// ch == kPrescaled: the plane already holds scaled values (the average-pooled planes of large tiles)
constexpr int kPrescaled = 2;
MAU_HD float scale_value(float v, int ch) {
  return ch == 0 ? (v + 1.0f) * 0.5f : (ch == 1 ? fminf(fmaxf(v, 0.0f), 1.0f) : v);
}
// d scale / d v  (torch.clamp passes the gradient where 0 <= v <= 1, bounds included)
MAU_HD float scale_slope(float v, int ch) {
  return ch == 0 ? 0.5f : (ch == 1 ? ((v >= 0.0f && v <= 1.0f) ? 1.0f : 0.0f) : 1.0f);
}
// piq.ssim(downsample=True): both images are average-pooled by f = max(1, round(min(H, W) / 256)) first
// (Python's round: half to even -- nearbyint in the default rounding mode)
inline int pool_factor(int H, int W) {
  const int f = (int)nearbyint((double)(H < W ? H : W) / 256.0);
  return f < 1 ? 1 : f;
}

struct Point {
  float s;        // SSIM of the window
  float a, b, c;  // dS/d mu_x, dS/d (G*x^2), dS/d (G*xy)   (x = scaled prediction)
};

// SSIM of one window from its five Gaussian moments mx = G*x, my = G*y, q = G*x^2, qy = G*y^2, r = G*xy
MAU_HD Point point_from_moments(float mx, float my, float q, float qy, float r) {
  const float sxx = q - mx * mx, syy = qy - my * my, sxy = r - mx * my;
  const float A1 = 2.f * mx * my + kC1, A2 = 2.f * sxy + kC2;
  const float B1 = mx * mx + my * my + kC1, B2 = sxx + syy + kC2;
  const float inv = 1.f / (B1 * B2);
  Point p;
  p.s = A1 * A2 * inv;
  p.a = 2.f * my * (A2 - A1) * inv - 2.f * mx * p.s * (1.f / B1 - 1.f / B2);
  p.b = -p.s / B2;
  p.c = 2.f * A1 * inv;
  return p;
}

// d (sum of S) / d raw prediction at one pixel from the three Gaussian-filtered derivative maps at that pixel
MAU_HD float combine_grad(float sa, float sb, float sc, float x_raw, float y_raw, int ch) {
  const float xv = scale_value(x_raw, ch), yv = scale_value(y_raw, ch);
  return (sa + 2.f * xv * sb + yv * sc) * scale_slope(x_raw, ch);
}

// SSIM of the window whose top-left corner is (i, j) of one raw (unscaled) prediction / target plane of width W
// (direct 121-tap form: the reference point the tiled, separable kernels of ssim.cu are checked against)
MAU_HD Point window(const float* x, const float* y, int W, int i, int j, int ch, const float* g) {
  float mx = 0.f, my = 0.f, q = 0.f, qy = 0.f, r = 0.f;
  for (int ki = 0; ki < kWin; ++ki) {
    const float* xr = x + (long long)(i + ki) * W + j;
    const float* yr = y + (long long)(i + ki) * W + j;
    for (int kj = 0; kj < kWin; ++kj) {
      const float w = g[ki] * g[kj];
      const float xv = scale_value(xr[kj], ch), yv = scale_value(yr[kj], ch);
      mx += w * xv;
      my += w * yv;
      q += w * xv * xv;
      qy += w * yv * yv;
      r += w * xv * yv;
    }
  }
  return point_from_moments(mx, my, q, qy, r);
}
// d (sum of S over all windows) / d raw prediction at pixel (yy, xx): gather over the windows that contain it.
// a, b, c are the [Hv, Wv] maps written by the forward pass (Hv = H - 10, Wv = W - 10).
MAU_HD float gather_grad(const float* a, const float* b, const float* c, int Hv, int Wv, const float* x, const float* y,
                         int W, int yy, int xx, int ch, const float* g) {
  float sa = 0.f, sb = 0.f, sc = 0.f;
  for (int ki = 0; ki < kWin; ++ki) {
    const int wi = yy - ki;
    if (wi < 0 || wi >= Hv) continue;
    for (int kj = 0; kj < kWin; ++kj) {
      const int wj = xx - kj;
      if (wj < 0 || wj >= Wv) continue;
      const float w = g[ki] * g[kj];
      const long long o = (long long)wi * Wv + wj;
      sa += w * a[o];
      sb += w * b[o];
      sc += w * c[o];
    }
  }
  return combine_grad(sa, sb, sc, x[(long long)yy * W + xx], y[(long long)yy * W + xx], ch);
}

}  // namespace mau_ssim
#endif  // MAU_SSIM_CORE_H_
