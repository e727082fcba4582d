// conv_ffma.h -- host interface of the FFMA convolution path (conv_ffma.cu)
#pragma once
#include <algorithm>
#include "common.h"

namespace mau {

struct ConvFfmaParams {
  const void* x = nullptr; int x_cs = 0, x_C = 0;
  int H = 0, W = 0, tiles_w = 0, tiles_h = 0;
  int nseg = 0; int seg_start[4] = {0, 0, 0, 0}; int seg_chunks[4] = {0, 0, 0, 0};  // chunks of 8 channels
  const float* w = nullptr; int Kp = 0, n_rows = 0;       // packed fp32 [9][Kp][n_rows]
  void* y = nullptr; int y_cs = 0, y_c0 = 0, Cout = 0;
  const float* scale = nullptr; const float* shift = nullptr; int relu = 0, accumulate = 0;
};

int conv_ffma_prepare(ConvFfmaParams* p, const View& xbuf, int nseg, const int* seg_start, const int* seg_len,
                      const float* wpacked, int Kp, int n_rows, const View& y, const float* scale,
                      const float* shift, int relu, int accumulate);
int conv_ffma_launch(int dtype, const ConvFfmaParams& p, int B, cudaStream_t st);
int conv_ffma_pack_fwd(const float* w, int Cout, int Cin, const int* kmap_dev, int Kp, float* out, cudaStream_t st);
int conv_ffma_pack_dgrad(const float* w, int Cout, int Cin, int ci0, int N, int Kp, float* out, cudaStream_t st);
// dW[co][ci_w0 + ci][t] (=|+=) sum dy * x  (slow reference path: fp32 mode and cross-checks)
int wgrad_ffma_launch(int dtype, const View& x_seg, const View& dy, int ci_w0, int Cin_w, float* dw, int accumulate,
                      cudaStream_t st);

}  // namespace mau
