// oracle/ssim_kernels_emu.cpp -- TEST INFRASTRUCTURE ONLY: the product's SSIM kernel file compiled for the CPU on top of
// oracle/cuda_emu.h, exported with the signature of the C ABI entry point (include/mau_b200.h: mau_ssim_loss) so that
// tests/test_oracle.py can run the kernels' own code -- grid / block indexing, guards, block reduction, launch order --
// against the torch restatement.  g++ -std=c++20 -O1 -shared -fPIC -pthread oracle/ssim_kernels_emu.cpp
#include "cuda_emu.h"
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/ssim.cu"

extern "C" long long emu_ssim_work_floats(int B, int H, int W) { return mau::ssim_work_floats(B, H, W); }
extern "C" int emu_ssim_loss(const float* pred, const float* tgt, int B, int C, int H, int W, float* loss, float* grad, float* work,
                             double* acc) {
  return mau::op_ssim_loss(pred, tgt, B, C, H, W, loss, grad, work, acc, nullptr);
}
extern "C" void emu_ssim_force_pool(int f) { mau::ssim_debug_force_pool(f); }
