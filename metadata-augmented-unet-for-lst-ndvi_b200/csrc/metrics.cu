// metrics.cu -- evaluation metrics of reference test/evaluate.py:210-275 on the device:
//   dw_map = argmax_c(maps[b,c] * c), c in 0..8  (ties -> lowest index, int64, bit-exact)
//   per (sample, channel): count / sum|p-g| / sum (p-g)^2 overall and per Dynamic-World class,
//   temperature channel (index 1) un-normalised first (test/evaluate.py:33-36).
// The host divides: MAE = sum|d|/count, RMSE = sqrt(sum d^2 / count).
//   laplacian_sums: per (sample, channel) sum and sum of squares of scipy.ndimage.laplace(x) (mode 'reflect') for
//   the prediction and the target (test/evaluate.py:241-242, np.var(laplace(.))); the host forms the variance.
#ifndef MAU_KERNEL_ENV            // a test harness may supply the execution environment instead (oracle/cuda_emu.h)
#include "ops.h"
#include "vec.cuh"
#define MAU_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif

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

// scipy.ndimage.laplace on one fp32 plane, default mode 'reflect' (index -1 -> 0, index n -> n-1): the second
// difference [1,-2,1] along each axis is evaluated in double and rounded to fp32 (correlate1d writes an fp32 output
// array), the two axes are then added in fp32 (`output += tmp`).
__device__ __forceinline__ float laplace_at(const float* __restrict__ a, int y, int x, int H, int W, bool unnorm,
                                            float mean, float stdv) {
  const int yu = y > 0 ? y - 1 : 0, yd = y < H - 1 ? y + 1 : H - 1;
  const int xl = x > 0 ? x - 1 : 0, xr = x < W - 1 ? x + 1 : W - 1;
  float c = a[(long long)y * W + x], u = a[(long long)yu * W + x], d = a[(long long)yd * W + x];
  float l = a[(long long)y * W + xl], r = a[(long long)y * W + xr];
  if (unnorm) {   // numpy: two rounded fp32 ops per element, no FMA (test/evaluate.py:33-34)
    c = __fadd_rn(__fmul_rn(c, stdv), mean);
    u = __fadd_rn(__fmul_rn(u, stdv), mean);
    d = __fadd_rn(__fmul_rn(d, stdv), mean);
    l = __fadd_rn(__fmul_rn(l, stdv), mean);
    r = __fadd_rn(__fmul_rn(r, stdv), mean);
  }
  const float d2y = (float)(((double)u + (double)d) - 2.0 * (double)c);
  const float d2x = (float)(((double)l + (double)r) - 2.0 * (double)c);
  return __fadd_rn(d2y, d2x);
}

// grid (chunks, B*C), block 256: out[plane][4] += {sum lap(pred), sum lap(pred)^2, sum lap(tgt), sum lap(tgt)^2}
__global__ void __launch_bounds__(256) laplacian_sums_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                             int C, int H, int W, float temp_mean, float temp_std,
                                                             double* __restrict__ out) {
  __shared__ double red[4][256];
  const int plane = blockIdx.y, tid = threadIdx.x, HW = H * W;
  const bool unnorm = (plane % C) == 1 && temp_std != 0.f;
  const float* P = pred + (long long)plane * HW;
  const float* G = tgt + (long long)plane * HW;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int p = blockIdx.x * blockDim.x + tid; p < HW; p += gridDim.x * blockDim.x) {
    const int y = p / W, x = p - y * W;
    const double lp = (double)laplace_at(P, y, x, H, W, unnorm, temp_mean, temp_std);
    const double lg = (double)laplace_at(G, y, x, H, W, unnorm, temp_mean, temp_std);
    s0 += lp; s1 += lp * lp; s2 += lg; s3 += lg * lg;
  }
  red[0][tid] = s0; red[1][tid] = s1; red[2][tid] = s2; red[3][tid] = s3;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) {
      red[0][tid] += red[0][tid + off]; red[1][tid] += red[1][tid + off];
      red[2][tid] += red[2][tid + off]; red[3][tid] += red[3][tid + off];
    }
    __syncthreads();
  }
  if (tid < 4) atomicAdd(&out[(long long)plane * 4 + tid], red[tid][0]);
}

}  // namespace

int op_laplacian_sums(const float* pred, const float* tgt, int B, int C, int H, int W, float temp_mean, float temp_std,
                      double* out, cudaStream_t st) {
  if (B < 1 || C < 1 || H < 1 || W < 1) return fail("laplacian_sums: empty input");
  if ((long long)B * C > 65535) return fail("laplacian_sums: B*C = %lld exceeds the grid limit 65535", (long long)B * C);
  MAU_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)B * C * 4, st));
  dim3 grid((unsigned)std::max(1, std::min(ceil_div(H * W, 256 * 4), 32)), (unsigned)(B * C), 1);
  MAU_LAUNCH(laplacian_sums_kernel, grid, dim3(256), st, pred, tgt, C, H, W, temp_mean, temp_std, out);
  MAU_LAUNCHED();
  return 0;
}

int op_eval_metrics(const float* maps, int maps_c, const float* pred, const float* tgt, int B, int C, int H, int W,
                    float temp_mean, float temp_std, long long* dw_map, double* sums, cudaStream_t st) {
  if (maps_c < 9) return fail("eval_metrics: maps need >= 9 Dynamic World channels, got %d", maps_c);
  if (C < 1 || C > kMaxC) return fail("eval_metrics: 1..%d target channels supported, got %d", kMaxC, C);
  const int HW = H * W;
  MAU_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)B * C * 30, st));
  dim3 grid((unsigned)std::max(1, std::min(ceil_div(HW, 256 * 4), 64)), (unsigned)B, 1);
  MAU_LAUNCH(eval_metrics_kernel, grid, dim3(256), st, maps, maps_c, pred, tgt, C, HW, temp_mean, temp_std, dw_map, sums);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
