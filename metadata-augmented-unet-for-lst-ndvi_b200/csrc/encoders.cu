// encoders.cu -- the two embedding encoders, fp32 throughout (metadata carries raw years ~2019,
// which bf16 cannot represent):
//   MetadataEncoder  Linear(F,32)-ReLU-Linear(32,D)            reference src/model.py:38-48
//   TemporalEncoder  LSTM(1,Hd) last hidden state -> Linear    reference src/model.py:23-34
// The LSTM is one persistent CTA per sample: thread g owns gate row g of W_hh in registers, the
// hidden state is broadcast from shared memory, two barriers per time step, T strictly sequential
// steps (it runs over the zero padding exactly like the reference, which never packs the series).
#include "ops.h"
#include "vec.cuh"

namespace mau {
namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------ metadata MLP
__global__ void mlp_fwd_kernel(const float* __restrict__ md, int F, const float* __restrict__ w0,
                               const float* __restrict__ b0, const float* __restrict__ w2,
                               const float* __restrict__ b2, int D, float* __restrict__ hidden,
                               float* __restrict__ out, int out_stride) {
  __shared__ float h[32];
  const int b = blockIdx.x;
  if (threadIdx.x < 32) {
    float a = b0[threadIdx.x];
    for (int f = 0; f < F; ++f) a = fmaf(md[b * F + f], w0[threadIdx.x * F + f], a);
    a = fmaxf(a, 0.f);
    h[threadIdx.x] = a;
    hidden[b * 32 + threadIdx.x] = a;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = b2[d];
#pragma unroll
    for (int j = 0; j < 32; ++j) a = fmaf(h[j], w2[d * 32 + j], a);
    out[(long long)b * out_stride + d] = a;
  }
}
// single block; every output gradient element is owned by one thread, batch reduced serially
__global__ void mlp_bwd_kernel(const float* __restrict__ md, int B, int F, const float* __restrict__ w0,
                               const float* __restrict__ w2, int D, const float* __restrict__ hidden,
                               const float* __restrict__ gout, int gs, float* dw0, float* db0, float* dw2,
                               float* db2) {
  for (int i = threadIdx.x; i < D * 32 + D; i += blockDim.x) {
    if (i < D * 32) {
      const int d = i / 32, j = i % 32;
      float a = 0.f;
      for (int b = 0; b < B; ++b) a = fmaf(gout[(long long)b * gs + d], hidden[b * 32 + j], a);
      dw2[i] = a;
    } else {
      const int d = i - D * 32;
      float a = 0.f;
      for (int b = 0; b < B; ++b) a += gout[(long long)b * gs + d];
      db2[d] = a;
    }
  }
  for (int i = threadIdx.x; i < 32 * F + 32; i += blockDim.x) {
    const int j = i < 32 * F ? i / F : i - 32 * F;
    const int f = i < 32 * F ? i % F : -1;
    float a = 0.f;
    for (int b = 0; b < B; ++b) {
      if (hidden[b * 32 + j] <= 0.f) continue;
      float dh = 0.f;
      for (int d = 0; d < D; ++d) dh = fmaf(gout[(long long)b * gs + d], w2[d * 32 + j], dh);
      a += f >= 0 ? dh * md[b * F + f] : dh;
    }
    if (f >= 0) dw0[j * F + f] = a; else db0[j] = a;
  }
}

// ------------------------------------------------------------------ small dense layer
__global__ void linear_fwd_kernel(const float* __restrict__ x, int B, int K, const float* __restrict__ w,
                                  const float* __restrict__ bias, int N, float* __restrict__ y, int ys) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * N; i += gridDim.x * blockDim.x) {
    const int b = i / N, n = i % N;
    float a = bias[n];
    for (int k = 0; k < K; ++k) a = fmaf(x[b * K + k], w[n * K + k], a);
    y[(long long)b * ys + n] = a;
  }
}
__global__ void linear_bwd_kernel(const float* __restrict__ x, int B, int K, const float* __restrict__ w, int N,
                                  const float* __restrict__ gy, int gs, float* gx, float* dw, float* db) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < B * K; i += nth) {
    const int b = i / K, k = i % K;
    float a = 0.f;
    for (int n = 0; n < N; ++n) a = fmaf(gy[(long long)b * gs + n], w[n * K + k], a);
    gx[i] = a;
  }
  for (int i = tid; i < N * K; i += nth) {
    const int n = i / K, k = i % K;
    float a = 0.f;
    for (int b = 0; b < B; ++b) a = fmaf(gy[(long long)b * gs + n], x[b * K + k], a);
    dw[i] = a;
  }
  for (int n = tid; n < N; n += nth) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += gy[(long long)b * gs + n];
    db[n] = a;
  }
}

// ------------------------------------------------------------------ LSTM forward
// grid = B, block = 4*HD.  save layout per (b,t): [i f g o](4*HD) ++ c(HD)
template <int HD>
__global__ void __launch_bounds__(4 * HD) lstm_fwd_kernel(const float* __restrict__ series, int T,
                                                          const float* __restrict__ w_ih,
                                                          const float* __restrict__ w_hh,
                                                          const float* __restrict__ b_ih,
                                                          const float* __restrict__ b_hh, float* __restrict__ h_last,
                                                          float* __restrict__ save) {
  extern __shared__ float sm[];
  float* xs = sm;              // [T] (padded to a multiple of 4 so hs stays 16-byte aligned)
  float* hs = xs + ((T + 3) & ~3);   // [HD]
  float* gs = hs + HD;         // [4*HD]
  const int b = blockIdx.x, g = threadIdx.x;
  float wreg[HD];
#pragma unroll
  for (int k = 0; k < HD; ++k) wreg[k] = w_hh[g * HD + k];
  const float wi = w_ih[g];
  const float bias = b_ih[g] + b_hh[g];
  for (int t = g; t < T; t += 4 * HD) xs[t] = series[(long long)b * T + t];
  if (g < HD) hs[g] = 0.f;
  float c = 0.f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float a = fmaf(wi, xs[t], bias);
#pragma unroll
    for (int k = 0; k < HD; k += 4) {
      const float4 h4 = *reinterpret_cast<const float4*>(hs + k);
      a = fmaf(wreg[k], h4.x, a);
      a = fmaf(wreg[k + 1], h4.y, a);
      a = fmaf(wreg[k + 2], h4.z, a);
      a = fmaf(wreg[k + 3], h4.w, a);
    }
    const float act = (g >= 2 * HD && g < 3 * HD) ? tanhf(a) : sigmoidf_(a);
    gs[g] = act;
    if (save) save[((long long)b * T + t) * 5 * HD + g] = act;
    __syncthreads();
    if (g < HD) {
      c = gs[HD + g] * c + gs[g] * gs[2 * HD + g];
      hs[g] = gs[3 * HD + g] * tanhf(c);
      if (save) save[((long long)b * T + t) * 5 * HD + 4 * HD + g] = c;
    }
    __syncthreads();
  }
  if (g < HD) h_last[b * HD + g] = hs[g];
}

// ------------------------------------------------------------------ LSTM backward (BPTT)
// grid = B, block = 4*HD; W_hh row-major in shared memory; dW_hh row g accumulated in registers.
template <int HD>
__global__ void __launch_bounds__(4 * HD) lstm_bwd_kernel(const float* __restrict__ series, int T,
                                                          const float* __restrict__ w_hh,
                                                          const float* __restrict__ save,
                                                          const float* __restrict__ dh_last, float* dw_ih,
                                                          float* dw_hh, float* db_ih, float* db_hh) {
  extern __shared__ float sm[];
  float* ws = sm;                    // [4*HD][HD]
  float* xs = ws + 4 * HD * HD;      // [T] (padded to a multiple of 4)
  float* hp = xs + ((T + 3) & ~3);   // [HD] h_{t-1}
  float* da = hp + HD;               // [4*HD] pre-activation gate grads
  float* part = da + 4 * HD;         // [4][HD] partial dh_prev
  float* dh = part + 4 * HD;         // [HD]
  float* dc = dh + HD;               // [HD]
  const int b = blockIdx.x, g = threadIdx.x;
  const int blk = g / HD, j = g % HD;   // gate block (0 i, 1 f, 2 g, 3 o), unit
  for (int i = g; i < 4 * HD * HD; i += 4 * HD) ws[i] = w_hh[i];
  for (int t = g; t < T; t += 4 * HD) xs[t] = series[(long long)b * T + t];
  if (g < HD) { dh[g] = dh_last[b * HD + g]; dc[g] = 0.f; }
  float dwreg[HD];
#pragma unroll
  for (int k = 0; k < HD; ++k) dwreg[k] = 0.f;
  float dwi = 0.f, dbs = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    const float* sv = save + ((long long)b * T + t) * 5 * HD;
    const float* svp = sv - 5 * HD;  // valid only when t > 0
    if (g < HD) hp[g] = t > 0 ? svp[3 * HD + g] * tanhf(svp[4 * HD + g]) : 0.f;
    // every thread recomputes the cell-level quantities of its unit j (cheap, avoids a barrier)
    const float gi = sv[j], gf = sv[HD + j], gg = sv[2 * HD + j], go = sv[3 * HD + j];
    const float ct = sv[4 * HD + j];
    const float cprev = t > 0 ? svp[4 * HD + j] : 0.f;
    const float tc = tanhf(ct);
    const float dhj = dh[j];
    const float dcj = dc[j] + dhj * go * (1.f - tc * tc);
    float dpre;
    if (blk == 0) dpre = dcj * gg * gi * (1.f - gi);
    else if (blk == 1) dpre = dcj * cprev * gf * (1.f - gf);
    else if (blk == 2) dpre = dcj * gi * (1.f - gg * gg);
    else dpre = dhj * tc * go * (1.f - go);
    __syncthreads();                 // all reads of dh/dc done; hp visible
    da[g] = dpre;
    if (blk == 1) dc[j] = dcj * gf;  // dc_{t-1}
    dwi = fmaf(dpre, xs[t], dwi);
    dbs += dpre;
#pragma unroll
    for (int k = 0; k < HD; k += 4) {
      const float4 h4 = *reinterpret_cast<const float4*>(hp + k);
      dwreg[k] = fmaf(dpre, h4.x, dwreg[k]);
      dwreg[k + 1] = fmaf(dpre, h4.y, dwreg[k + 1]);
      dwreg[k + 2] = fmaf(dpre, h4.z, dwreg[k + 2]);
      dwreg[k + 3] = fmaf(dpre, h4.w, dwreg[k + 3]);
    }
    __syncthreads();                 // da complete
    // dh_{t-1}[j] = sum_g W_hh[g][j] * da[g]; thread (blk, j) sums its gate block
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < HD; ++r) s = fmaf(ws[(blk * HD + r) * HD + j], da[blk * HD + r], s);
    part[blk * HD + j] = s;
    __syncthreads();
    if (g < HD) dh[g] = part[g] + part[HD + g] + part[2 * HD + g] + part[3 * HD + g];
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < HD; ++k) atomicAdd(&dw_hh[g * HD + k], dwreg[k]);
  atomicAdd(&dw_ih[g], dwi);
  atomicAdd(&db_ih[g], dbs);
  atomicAdd(&db_hh[g], dbs);
}

template <int HD>
int lstm_fwd_inst(const float* series, int B, int T, const float* w_ih, const float* w_hh, const float* b_ih,
                  const float* b_hh, float* h_last, float* save, cudaStream_t st) {
  const int Tp = round_up(T, 4);
  const size_t smem = (Tp + HD + 4 * HD) * sizeof(float);
  auto k = lstm_fwd_kernel<HD>;
  if (smem > 48 * 1024) MAU_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lstm_fwd_kernel<HD><<<B, 4 * HD, smem, st>>>(series, T, w_ih, w_hh, b_ih, b_hh, h_last, save);
  MAU_LAUNCHED();
  return 0;
}
template <int HD>
int lstm_bwd_inst(const float* series, int B, int T, const float* w_hh, const float* save, const float* dh_last,
                  float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, cudaStream_t st) {
  const size_t smem = ((size_t)4 * HD * HD + round_up(T, 4) + HD + 4 * HD + 4 * HD + 2 * HD) * sizeof(float);
  auto k = lstm_bwd_kernel<HD>;
  MAU_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lstm_bwd_kernel<HD><<<B, 4 * HD, smem, st>>>(series, T, w_hh, save, dh_last, dw_ih, dw_hh, db_ih, db_hh);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace

int op_mlp_fwd(const float* md, int B, int F, const float* w0, const float* b0, const float* w2, const float* b2,
               int D, float* hidden, float* out, int out_stride, cudaStream_t st) {
  mlp_fwd_kernel<<<B, 128, 0, st>>>(md, F, w0, b0, w2, b2, D, hidden, out, out_stride);
  MAU_LAUNCHED();
  return 0;
}
int op_mlp_bwd(const float* md, int B, int F, const float* w0, const float* w2, int D, const float* hidden,
               const float* gout, int gout_stride, float* dw0, float* db0, float* dw2, float* db2,
               cudaStream_t st) {
  mlp_bwd_kernel<<<1, 512, 0, st>>>(md, B, F, w0, w2, D, hidden, gout, gout_stride, dw0, db0, dw2, db2);
  MAU_LAUNCHED();
  return 0;
}
int op_linear_fwd(const float* x, int B, int K, const float* w, const float* b, int N, float* y, int y_stride,
                  cudaStream_t st) {
  linear_fwd_kernel<<<ceil_div(B * N, 128), 128, 0, st>>>(x, B, K, w, b, N, y, y_stride);
  MAU_LAUNCHED();
  return 0;
}
int op_linear_bwd(const float* x, int B, int K, const float* w, int N, const float* gy, int gy_stride, float* gx,
                  float* dw, float* db, cudaStream_t st) {
  linear_bwd_kernel<<<32, 256, 0, st>>>(x, B, K, w, N, gy, gy_stride, gx, dw, db);
  MAU_LAUNCHED();
  return 0;
}

size_t lstm_save_floats(int B, int T, int Hd) { return (size_t)B * T * 5 * Hd; }
size_t lstm_bwd_scratch_floats(int, int) { return 0; }

#define MAU_LSTM_SWITCH(CALL)                                                               \
  switch (Hd) {                                                                             \
    case 16: return CALL(16);                                                               \
    case 32: return CALL(32);                                                               \
    case 64: return CALL(64);                                                               \
    case 96: return CALL(96);                                                               \
    default: return fail("lstm: hidden size %d not instantiated (16, 32, 64, 96)", Hd); \
  }

int op_lstm_fwd(const float* series, int B, int T, int Hd, const float* w_ih, const float* w_hh, const float* b_ih,
                const float* b_hh, float* h_last, float* save, cudaStream_t st) {
  if (T < 1) return fail("lstm: empty series");
#define CALL(H) lstm_fwd_inst<H>(series, B, T, w_ih, w_hh, b_ih, b_hh, h_last, save, st)
  MAU_LSTM_SWITCH(CALL)
#undef CALL
}
int op_lstm_bwd(const float* series, int B, int T, int Hd, const float* w_hh, const float* save,
                const float* dh_last, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, float* scratch,
                cudaStream_t st) {
  (void)scratch;
#define CALL(H) lstm_bwd_inst<H>(series, B, T, w_hh, save, dh_last, dw_ih, dw_hh, db_ih, db_hh, st)
  MAU_LSTM_SWITCH(CALL)
#undef CALL
}

}  // namespace mau
