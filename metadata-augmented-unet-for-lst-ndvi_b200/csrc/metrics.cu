// metrics.cu -- evaluation metrics of reference test/evaluate.py:210-275 on the device:
//   dw_map = argmax_c(maps[b,c] * c), c in 0..8  (ties -> lowest index, int64, bit-exact)
//   per (sample, channel): count / sum|p-g| / sum (p-g)^2 overall and per Dynamic-World class,
//   temperature channel (index 1) un-normalised first (test/evaluate.py:33-36).
// The host divides: MAE = sum|d|/count, RMSE = sqrt(sum d^2 / count).
#include "ops.h"
#include "vec.cuh"

namespace mau {
namespace {

constexpr int kMaxC = 4;

// grid (chunks, B), block 256; shared histogram [C][10][3] doubles
__global__ void __launch_bounds__(256) eval_metrics_kernel(const float* __restrict__ maps, int maps_c,
                                                           const float* __restrict__ pred,
                                                           const float* __restrict__ tgt, int C, int HW,
                                                           float temp_mean, float temp_std,
                                                           long long* __restrict__ dw_map, double* __restrict__ sums) {
  __shared__ double hist[kMaxC * 10 * 3];
  for (int i = threadIdx.x; i < C * 30; i += blockDim.x) hist[i] = 0.0;
  __syncthreads();
  const int b = blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int cls = 0;
    float best = maps[((long long)b * maps_c) * HW + p] * 0.f;
    for (int c = 1; c < 9; ++c) {
      const float v = maps[((long long)b * maps_c + c) * HW + p] * (float)c;
      if (v > best) { best = v; cls = c; }
    }
    dw_map[(long long)b * HW + p] = cls;
    for (int ch = 0; ch < C; ++ch) {
      float pv = pred[((long long)b * C + ch) * HW + p];
      float gv = tgt[((long long)b * C + ch) * HW + p];
      if (ch == 1 && temp_std != 0.f) {
        pv = __fadd_rn(__fmul_rn(pv, temp_std), temp_mean);   // numpy: two rounded fp32 ops, no FMA
        gv = __fadd_rn(__fmul_rn(gv, temp_std), temp_mean);
      }
      const float d = pv - gv;
      const double ad = fabsf(d), sq = (double)d * (double)d;
      double* h0 = hist + (ch * 10) * 3;
      atomicAdd(h0 + 0, 1.0); atomicAdd(h0 + 1, ad); atomicAdd(h0 + 2, sq);
      double* hk = hist + (ch * 10 + 1 + cls) * 3;
      atomicAdd(hk + 0, 1.0); atomicAdd(hk + 1, ad); atomicAdd(hk + 2, sq);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 30; i += blockDim.x)
    if (hist[i] != 0.0) atomicAdd(&sums[(long long)b * C * 30 + i], hist[i]);
}

}  // namespace

int op_eval_metrics(const float* maps, int maps_c, const float* pred, const float* tgt, int B, int C, int H, int W,
                    float temp_mean, float temp_std, long long* dw_map, double* sums, cudaStream_t st) {
  if (maps_c < 9) return fail("eval_metrics: maps need >= 9 Dynamic World channels, got %d", maps_c);
  if (C < 1 || C > kMaxC) return fail("eval_metrics: 1..%d target channels supported, got %d", kMaxC, C);
  const int HW = H * W;
  MAU_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)B * C * 30, st));
  dim3 grid((unsigned)std::max(1, std::min(ceil_div(HW, 256 * 4), 64)), (unsigned)B, 1);
  eval_metrics_kernel<<<grid, 256, 0, st>>>(maps, maps_c, pred, tgt, C, HW, temp_mean, temp_std, dw_map, sums);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
