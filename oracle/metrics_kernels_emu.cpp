// oracle/metrics_kernels_emu.cpp -- TEST INFRASTRUCTURE ONLY: csrc/metrics.cu (evaluation-metric kernels, both verified on a
// B200) compiled for the CPU on oracle/cuda_emu.h.  Its purpose is to validate the emulation shim itself: kernels known to be
// correct on hardware must give the oracle's answers under emulation too (tests/test_oracle.py), which is what lends weight to
// the same check of kernels that have not run on hardware yet (oracle/ssim_kernels_emu.cpp).
#include "cuda_emu.h"
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/metrics.cu"

extern "C" int emu_eval_metrics(const float* maps, int maps_c, const float* pred, const float* tgt, int B, int C, int H, int W,
                                float temp_mean, float temp_std, long long* dw_map, double* sums) {
  return mau::op_eval_metrics(maps, maps_c, pred, tgt, B, C, H, W, temp_mean, temp_std, dw_map, sums, nullptr);
}
extern "C" int emu_laplacian_sums(const float* pred, const float* tgt, int B, int C, int H, int W, float temp_mean, float temp_std,
                                  double* out) {
  return mau::op_laplacian_sums(pred, tgt, B, C, H, W, temp_mean, temp_std, out, nullptr);
}
