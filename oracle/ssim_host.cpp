// oracle/ssim_host.cpp -- TEST INFRASTRUCTURE ONLY: a serial host driver around the product's own SSIM arithmetic
// (metadata-augmented-unet-for-lst-ndvi_b200/csrc/ssim_core.h, the header the CUDA kernels include), so that
// tests/test_oracle.py can check those formulas -- forward value and analytic gradient -- against torch autograd of
// oracle/ssim_oracle.py on the CPU.  It mirrors the two kernels of ssim.cu one output element at a time.
// Built by the test with:  g++ -O2 -shared -fPIC oracle/ssim_host.cpp -o <tmp>/libssim_host.so
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/ssim_core.h"

#include <vector>

extern "C" int ssim_host(const float* pred, const float* tgt, int B, int C, int H, int W, double* loss, float* grad) {
  using namespace mau_ssim;
  if (C < 2 || H < kWin || W < kWin) return 1;
  float g[kWin];
  gaussian_window(g);
  const int Hv = H - kWin + 1, Wv = W - kWin + 1;
  const long long nwin = (long long)B * 2 * Hv * Wv;
  std::vector<float> a(nwin), b(nwin), c(nwin);
  double acc = 0.0;
  for (int bi = 0; bi < B; ++bi)
    for (int ch = 0; ch < 2; ++ch) {
      const float* x = pred + ((long long)bi * C + ch) * H * W;
      const float* y = tgt + ((long long)bi * C + ch) * H * W;
      const long long base = ((long long)bi * 2 + ch) * Hv * Wv;
      for (int i = 0; i < Hv; ++i)
        for (int j = 0; j < Wv; ++j) {
          Point p = window(x, y, W, i, j, ch, g);
          acc += p.s;
          a[base + (long long)i * Wv + j] = p.a;
          b[base + (long long)i * Wv + j] = p.b;
          c[base + (long long)i * Wv + j] = p.c;
        }
    }
  *loss = 1.0 - acc / (double)nwin;
  if (grad) {
    const float coef = (float)(-1.0 / (double)nwin);
    for (int bi = 0; bi < B; ++bi)
      for (int ch = 0; ch < C; ++ch) {
        float* gp = grad + ((long long)bi * C + ch) * H * W;
        if (ch >= 2) {
          for (long long k = 0; k < (long long)H * W; ++k) gp[k] = 0.f;
          continue;
        }
        const float* x = pred + ((long long)bi * C + ch) * H * W;
        const float* y = tgt + ((long long)bi * C + ch) * H * W;
        const long long base = ((long long)bi * 2 + ch) * Hv * Wv;
        for (int yy = 0; yy < H; ++yy)
          for (int xx = 0; xx < W; ++xx)
            gp[(long long)yy * W + xx] = coef * gather_grad(a.data() + base, b.data() + base, c.data() + base, Hv, Wv, x, y, W, yy, xx, ch, g);
      }
  }
  return 0;
}

extern "C" int ssim_pool_factor(int H, int W) { return mau_ssim::pool_factor(H, W); }
