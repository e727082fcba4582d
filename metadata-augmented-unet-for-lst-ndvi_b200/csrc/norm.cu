// norm.cu -- BatchNorm2d of the reference's VGGBlock (src/model.py:13,15) as bandwidth kernels:
// eval-mode folding into the conv epilogue, training-mode batch statistics (biased variance for
// normalisation, unbiased for the running update, momentum 0.1), the fused normalise+ReLU pass and
// the two-pass backward.  Per-channel reductions: 8 channels per thread, pixel lanes reduced through
// shared memory, one double-precision atomic per channel per block.
#include "ops.h"
#include "vec.cuh"
#include <algorithm>

namespace mau {
namespace {

inline DView dv(const View& v) { return DView{v.ptr, v.B, v.H, v.W, v.cs, v.c0, v.C}; }
template <typename T>
__device__ __forceinline__ T* at(const DView& v, long long pix, int c) {
  return static_cast<T*>(v.ptr) + pix * v.cs + v.c0 + c;
}

__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                    const float* conv_bias, int C, float eps, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = s;
  shift[c] = ((conv_bias ? conv_bias[c] : 0.f) - rm[c]) * s + beta[c];
}

// block = 256 threads = G channel groups x L pixel lanes; block b handles pixels [b*per, (b+1)*per).
// Every thread keeps kU pixels' worth of 16-byte loads in flight (NT tensors each) before it touches any of
// them: these kernels have no reuse, so achieved bandwidth is bytes-in-flight / latency.
constexpr int kU = 4;
template <typename T, int NQ, int NT, typename F, int U = (NT == 1 ? 8 : kU)>
__device__ __forceinline__ void channel_reduce(const DView& ref, const DView& second, long long npix, double* out,
                                               F&& body) {
  extern __shared__ float red[];  // [NQ][256][8]
  using Raw = typename V8<T>::Raw;
  const int G = ref.C / 8;
  const int L = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  float acc[NQ][8];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
  if (pl < L) {
    const int c = gi * 8;
    for (long long p = p0 + pl; p < p1; p += (long long)U * L) {
      Raw r0[U], r1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pu = p + (long long)u * L;
        if (pu < p1) {
          r0[u] = V8<T>::load_raw(at<T>(ref, pu, c));
          if (NT > 1) r1[u] = V8<T>::load_raw(at<T>(second, pu, c));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pu = p + (long long)u * L;
        if (pu < p1) {
          float a[8], b[8];
          V8<T>::unpack(r0[u], a);
          if (NT > 1) V8<T>::unpack(r1[u], b);
          body(pu, c, a, b, acc);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[(q * 256 + threadIdx.x) * 8 + k] = acc[q][k];
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * ref.C; i += 256) {
    const int q = i / ref.C, c = i - q * ref.C;
    double s = 0.0;
    for (int l = 0; l < L; ++l) s += (double)red[(q * 256 + l * G + (c >> 3)) * 8 + (c & 7)];
    atomicAdd(&out[q * ref.C + c], s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(DView z, double* sums) {
  const long long npix = (long long)z.B * z.H * z.W;
  channel_reduce<T, 2, 1>(z, z, npix, sums, [&](long long, int, const float (&v)[8], const float (&)[8], float (&acc)[2][8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[0][k] += v[k]; acc[1][k] = fmaf(v[k], v[k], acc[1][k]); }
  });
}

__global__ void bn_finalize_train_kernel(const double* sums, long long count, const float* gamma, const float* beta,
                                         int C, float eps, float momentum, float* rm, float* rv, float* scale,
                                         float* shift, float* save_mean, float* save_rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = (double)count;
  const double mean = sums[c] / n;
  double var = sums[C + c] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float s = gamma[c] * rstd;
  scale[c] = s;
  shift[c] = beta[c] - (float)mean * s;
  save_mean[c] = (float)mean;
  save_rstd[c] = rstd;
  const double unbiased = count > 1 ? var * n / (n - 1.0) : var;
  rm[c] = (1.f - momentum) * rm[c] + momentum * (float)mean;
  rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unbiased;
}

// block = G channel groups x L pixel lanes (like channel_reduce); per-channel coefficients live in
// registers for the whole pixel loop; kU loads in flight per thread
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(DView z, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, DView y) {
  using Raw = typename V8<T>::Raw;
  const int G = z.C / 8;
  const int L = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  if (pl >= L) return;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale[gi * 8 + k]; sh[k] = shift[gi * 8 + k]; }
  const long long npix = (long long)z.B * z.H * z.W;
  const long long stride = (long long)gridDim.x * L;
  for (long long p = (long long)blockIdx.x * L + pl; p < npix; p += kU * stride) {
    Raw r[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p + u * stride < npix) r[u] = V8<T>::load_raw(at<T>(z, p + u * stride, gi * 8));
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p + u * stride < npix) {
        float v[8];
        V8<T>::unpack(r[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f);
        V8<T>::store(at<T>(y, p + u * stride, gi * 8), v);
      }
  }
}

// mean / biased variance in double (the subtraction cancels), inverse standard deviation in fp32 like ATen's
// batch_norm (FP64 divide / sqrt per thread were the dominant cost of the small late layers)
__device__ __forceinline__ void bn_coeffs(const double* __restrict__ sums, int C, int c, double inv_n, float eps, float& mean,
                                          float& var, float& rstd) {
  const double m = sums[c] * inv_n;
  double v = sums[C + c] * inv_n - m * m;
  if (v < 0.0) v = 0.0;
  mean = (float)m;
  var = (float)v;
  rstd = rsqrtf(var + eps);
}

// training forward, one launch: every thread derives mean / rstd / scale / shift of its 8 channels from the
// (already all-reduced, if SyncBN) double sums; block 0 also writes the saved statistics for the backward and
// performs the running-stat update (momentum, unbiased variance) -- then the normalise + ReLU pass.
template <typename T>
__global__ void __launch_bounds__(256) bn_finalize_apply_relu_kernel(DView z, const double* __restrict__ sums, long long count, double inv_n,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, float eps, float momentum,
                                                                     float* __restrict__ rm, float* __restrict__ rv,
                                                                     float* __restrict__ scale_out, float* __restrict__ shift_out,
                                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                     DView y) {
  using Raw = typename V8<T>::Raw;
  const int C = z.C;
  if (blockIdx.x == 0) {
    const float unbias = count > 1 ? (float)((double)count / ((double)count - 1.0)) : 1.f;
    for (int c = threadIdx.x; c < C; c += 256) {
      float mean, var, rstd;
      bn_coeffs(sums, C, c, inv_n, eps, mean, var, rstd);
      const float sc = gamma[c] * rstd;
      scale_out[c] = sc;
      shift_out[c] = beta[c] - mean * sc;
      mean_out[c] = mean;
      rstd_out[c] = rstd;
      rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
      rv[c] = (1.f - momentum) * rv[c] + momentum * (var * unbias);
    }
  }
  const int G = C / 8;
  const int L = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  if (pl >= L) return;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {       // same arithmetic as block 0 above, so every block uses identical coefficients
    const int c = gi * 8 + k;
    float mean, var, rstd;
    bn_coeffs(sums, C, c, inv_n, eps, mean, var, rstd);
    sc[k] = gamma[c] * rstd;
    sh[k] = beta[c] - mean * sc[k];
  }
  const long long npix = (long long)z.B * z.H * z.W;
  const long long stride = (long long)gridDim.x * L;
  for (long long p = (long long)blockIdx.x * L + pl; p < npix; p += kU * stride) {
    Raw r[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p + u * stride < npix) r[u] = V8<T>::load_raw(at<T>(z, p + u * stride, gi * 8));
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p + u * stride < npix) {
        float v[8];
        V8<T>::unpack(r[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f);
        V8<T>::store(at<T>(y, p + u * stride, gi * 8), v);
      }
  }
}

// MAU_BN_BWD_MIN_BLOCKS: resident blocks per SM the two backward kernels are compiled for (register cap 65536 / 256 / n);
// undefined = no cap (a second __launch_bounds__ argument of 1 is NOT the same: it changes ptxas' allocation)
#ifdef MAU_BN_BWD_MIN_BLOCKS
#define MAU_BN_BWD_BOUNDS __launch_bounds__(256, MAU_BN_BWD_MIN_BLOCKS)
#else
#define MAU_BN_BWD_BOUNDS __launch_bounds__(256)
#endif
template <typename T>
__global__ void MAU_BN_BWD_BOUNDS bn_bwd_reduce_kernel(DView gy, DView z, const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, double* sums) {
  const long long npix = (long long)z.B * z.H * z.W;
  const int gi8 = (threadIdx.x % (z.C / 8)) * 8;
  float cm[8], cr[8], sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { cm[k] = mean[gi8 + k]; cr[k] = rstd[gi8 + k]; sc[k] = scale[gi8 + k]; sh[k] = shift[gi8 + k]; }
  channel_reduce<T, 2, 2>(gy, z, npix, sums, [&](long long, int, const float (&g)[8], const float (&zz)[8], float (&acc)[2][8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // ReLU mask recomputed exactly as the forward did (y = relu(fma(z, scale, shift))): no read of y
      const float gt = fmaf(zz[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
      const float xh = (zz[k] - cm[k]) * cr[k];
      acc[0][k] += gt;
      acc[1][k] = fmaf(gt, xh, acc[1][k]);
    }
  });
}

// dz = gamma*rstd*(g~ - mean(g~) - xhat*mean(g~*xhat)).  Pure element-wise pass (kU pixels in flight per thread;
// dz may alias z: a thread only ever touches its own pixels).  The gradient of the convolution bias, sum(dz), is
// identically zero under batch-statistics BatchNorm (the mean subtraction removes any constant), so it is not
// reduced here: the reference's value is fp32 rounding noise around 0, ours is exactly 0.
template <typename T>
__global__ void MAU_BN_BWD_BOUNDS bn_bwd_apply_kernel(DView gy, DView z, const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd,
                                                           const double* __restrict__ sums, long long count, DView dz,
                                                           const double* __restrict__ param_sums, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, float* __restrict__ dbias) {
  using Raw = typename V8<T>::Raw;
  const float inv_n = 1.f / (float)count;
  const int C = z.C;
  if (blockIdx.x == 0 && param_sums) {      // dbeta = sum g~, dgamma = sum g~*xhat (this rank's sums), dbias = 0
    for (int c = threadIdx.x; c < C; c += 256) {
      if (dbeta) dbeta[c] = (float)param_sums[c];
      if (dgamma) dgamma[c] = (float)param_sums[C + c];
      if (dbias) dbias[c] = 0.f;
    }
  }
  const int G = C / 8;
  const int L = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  if (pl >= L) return;
  const int gi8 = gi * 8;
  float ca[8], cm[8], cr[8], m1[8], m2[8], sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = scale[gi8 + k]; sh[k] = shift[gi8 + k];
    cr[k] = rstd[gi8 + k]; cm[k] = mean[gi8 + k]; ca[k] = gamma[gi8 + k] * cr[k];
    m1[k] = (float)sums[gi8 + k] * inv_n; m2[k] = (float)sums[C + gi8 + k] * inv_n;
  }
  const long long npix = (long long)z.B * z.H * z.W;
  const long long stride = (long long)gridDim.x * L;
  for (long long p = (long long)blockIdx.x * L + pl; p < npix; p += kU * stride) {
    Raw rg[kU], rz[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p + u * stride < npix) {
        rg[u] = V8<T>::load_raw(at<T>(gy, p + u * stride, gi8));
        rz[u] = V8<T>::load_raw(at<T>(z, p + u * stride, gi8));
      }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p + u * stride < npix) {
        float g[8], zz[8], o[8];
        V8<T>::unpack(rg[u], g);
        V8<T>::unpack(rz[u], zz);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float gt = fmaf(zz[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
          const float xh = (zz[k] - cm[k]) * cr[k];
          o[k] = ca[k] * (gt - m1[k] - xh * m2[k]);
        }
        V8<T>::store(at<T>(dz, p + u * stride, gi8), o);
      }
  }
}

__global__ void bn_bwd_finalize_kernel(const double* sums, const double* dbias_sums, int C, float* dgamma,
                                       float* dbeta, float* dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = (float)sums[c];
  if (dgamma) dgamma[c] = (float)sums[C + c];
  if (dbias) dbias[c] = dbias_sums ? (float)dbias_sums[c] : 0.f;
}

__global__ void bump_counters_kernel(long long* const* counters, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && counters[i]) *counters[i] += 1;
}

inline bool vec_ok(const View& v) { return v.cs % 8 == 0 && v.c0 % 8 == 0 && v.C % 8 == 0 && v.C <= 2048; }
// Grid-stride element-wise kernels with equal work per block: a grid of 3.3 x the resident blocks ends in a tail wave
// that leaves most SMs idle (bn_bwd_apply at 125^2: 977 blocks on 296 slots), so grids above one wave are rounded down
// to whole waves.  Resident blocks per SM come from the occupancy calculator, once per kernel.
template <typename K>
int resident_per_sm(K kernel, size_t dyn_smem = 0) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, dyn_smem) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
  return n;
}
inline long long whole_waves(long long blocks, int per_sm) {
  const long long slots = 148ll * per_sm;
  return (blocks > slots && whole_waves_enabled()) ? blocks / slots * slots : blocks;
}
inline int reduce_blocks(long long npix, int C) {
  // enough blocks to fill the machine, but at least ~64 pixels per lane-row of a block
  const int L = 256 / (C / 8) > 0 ? 256 / (C / 8) : 1;
  long long b = npix / ((long long)L * 16);
  if (b > 148 * 4) b = 148 * 4;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int op_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                    const float* conv_bias, int C, float eps, float* scale, float* shift, cudaStream_t st) {
  bn_fold_eval_kernel<<<ceil_div(C, 128), 128, 0, st>>>(gamma, beta, rm, rv, conv_bias, C, eps, scale, shift);
  MAU_LAUNCHED();
  return 0;
}

int op_bn_stats(int dt, const View& z, double* sums, cudaStream_t st) {
  if (!vec_ok(z)) return fail("bn_stats: bad view (C=%d cs=%d c0=%d)", z.C, z.cs, z.c0);
  // sums layout [2][C]; with a single slab (C <= 2048) the kernel's [q*C + c] indexing matches
  const size_t smem = 2 * 256 * 8 * sizeof(float);
  const int blocks = std::min(reduce_blocks(z.pixels(), z.C), 148 * 3);      // 80 registers: 3 resident blocks per SM, one wave
  if (dt == DT_BF16) bn_stats_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>(DView{z.ptr, z.B, z.H, z.W, z.cs, z.c0, z.C}, sums);
  else               bn_stats_kernel<float><<<blocks, 256, smem, st>>>(DView{z.ptr, z.B, z.H, z.W, z.cs, z.c0, z.C}, sums);
  MAU_LAUNCHED();
  return 0;
}
int op_bn_finalize_train(const double* sums, long long count, const float* gamma, const float* beta, int C,
                         float eps, float momentum, float* running_mean, float* running_var, float* scale,
                         float* shift, float* save_mean, float* save_rstd, cudaStream_t st) {
  bn_finalize_train_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, count, gamma, beta, C, eps, momentum,
                                                            running_mean, running_var, scale, shift, save_mean,
                                                            save_rstd);
  MAU_LAUNCHED();
  return 0;
}
int op_bn_apply_relu(int dt, const View& z, const float* scale, const float* shift, const View& y,
                     cudaStream_t st) {
  if (!vec_ok(z) || !vec_ok(y) || z.C != y.C || z.pixels() != y.pixels()) return fail("bn_apply: bad views");
  const int L = std::max(1, 256 / (z.C / 8));
  long long b = (z.pixels() + L - 1) / L;
  if (b > 148 * 8) b = 148 * 8;
  if (dt == DT_BF16) bn_apply_relu_kernel<__nv_bfloat16><<<(int)b, 256, 0, st>>>(dv(z), scale, shift, dv(y));
  else               bn_apply_relu_kernel<float><<<(int)b, 256, 0, st>>>(dv(z), scale, shift, dv(y));
  MAU_LAUNCHED();
  return 0;
}
int op_bn_finalize_apply_relu(int dt, const View& z, const double* sums, long long count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                              float* scale, float* shift, float* save_mean, float* save_rstd, const View& y,
                              cudaStream_t st) {
  if (!vec_ok(z) || !vec_ok(y) || z.C != y.C || z.pixels() != y.pixels()) return fail("bn_finalize_apply: bad views");
  const int L = std::max(1, 256 / (z.C / 8));
  long long b = (z.pixels() + (long long)L * 16 - 1) / ((long long)L * 16);     // >= 16 pixels per thread amortise the coefficient set-up
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  static const int occ_bf16 = resident_per_sm(bn_finalize_apply_relu_kernel<__nv_bfloat16>);
  static const int occ_f32 = resident_per_sm(bn_finalize_apply_relu_kernel<float>);
  b = whole_waves(b, dt == DT_BF16 ? occ_bf16 : occ_f32);
  if (dt == DT_BF16)
    bn_finalize_apply_relu_kernel<__nv_bfloat16><<<(int)b, 256, 0, st>>>(dv(z), sums, count, 1.0 / (double)count, gamma, beta, eps, momentum, running_mean,
                                                                       running_var, scale, shift, save_mean, save_rstd, dv(y));
  else
    bn_finalize_apply_relu_kernel<float><<<(int)b, 256, 0, st>>>(dv(z), sums, count, 1.0 / (double)count, gamma, beta, eps, momentum, running_mean,
                                                               running_var, scale, shift, save_mean, save_rstd, dv(y));
  MAU_LAUNCHED();
  return 0;
}
int op_bn_bwd_reduce(int dt, const View& gy, const View& z, const float* scale, const float* shift, const float* mean,
                     const float* rstd, double* sums, cudaStream_t st) {
  if (!vec_ok(z) || !vec_ok(gy)) return fail("bn_bwd_reduce: bad views");
  const size_t smem = 2 * 256 * 8 * sizeof(float);
  static const int occ_bf16 = resident_per_sm(bn_bwd_reduce_kernel<__nv_bfloat16>, smem);
  static const int occ_f32 = resident_per_sm(bn_bwd_reduce_kernel<float>, smem);
  const int blocks = std::min(reduce_blocks(z.pixels(), z.C), 148 * (dt == DT_BF16 ? occ_bf16 : occ_f32));      // one wave of resident blocks
  if (dt == DT_BF16) bn_bwd_reduce_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>(dv(gy), dv(z), scale, shift, mean, rstd, sums);
  else               bn_bwd_reduce_kernel<float><<<blocks, 256, smem, st>>>(dv(gy), dv(z), scale, shift, mean, rstd, sums);
  MAU_LAUNCHED();
  return 0;
}
int op_bn_bwd_apply(int dt, const View& gy, const View& z, const float* scale, const float* shift, const float* gamma,
                    const float* mean, const float* rstd, const double* sums, long long count, const View& dz_out,
                    const double* param_sums, float* dgamma, float* dbeta, float* dbias, cudaStream_t st) {
  if (!vec_ok(z) || !vec_ok(gy) || !vec_ok(dz_out)) return fail("bn_bwd_apply: bad views");
  const int L = std::max(1, 256 / (z.C / 8));
  long long b = (z.pixels() + (long long)L * 16 - 1) / ((long long)L * 16);     // >= 16 pixels per thread amortise the coefficient set-up
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  static const int occ_bf16 = resident_per_sm(bn_bwd_apply_kernel<__nv_bfloat16>);
  static const int occ_f32 = resident_per_sm(bn_bwd_apply_kernel<float>);
  b = whole_waves(b, dt == DT_BF16 ? occ_bf16 : occ_f32);
  if (dt == DT_BF16)
    bn_bwd_apply_kernel<__nv_bfloat16><<<(int)b, 256, 0, st>>>(dv(gy), dv(z), scale, shift, gamma, mean, rstd, sums, count, dv(dz_out), param_sums, dgamma, dbeta, dbias);
  else
    bn_bwd_apply_kernel<float><<<(int)b, 256, 0, st>>>(dv(gy), dv(z), scale, shift, gamma, mean, rstd, sums, count, dv(dz_out), param_sums, dgamma, dbeta, dbias);
  MAU_LAUNCHED();
  return 0;
}
int op_bn_bwd_finalize(const double* sums, const double* dbias_sums, int C, float* dgamma, float* dbeta,
                       float* dbias, cudaStream_t st) {
  bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, dbias_sums, C, dgamma, dbeta, dbias);
  MAU_LAUNCHED();
  return 0;
}
int op_bump_counters(long long* const* counters_dev, int n, cudaStream_t st) {
  bump_counters_kernel<<<ceil_div(n, 128), 128, 0, st>>>(counters_dev, n);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
