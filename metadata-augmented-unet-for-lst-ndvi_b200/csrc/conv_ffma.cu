// conv_ffma.cu -- 3x3 / pad 1 convolution on the fp32 FMA pipe (NHWC, fp32 accumulate).
//
// This is the MAU_PRECISION_FP32 path (1e-5 parity mode: tensor-core TF32 cannot hold 1e-5 through
// 19 layers) and the on-device cross-check for the tcgen05 kernels.  Same operator contract as
// conv_tc.cu: replaces nn.Conv2d(.,.,3,padding=1) of reference src/model.py:12,14 (forward, and the
// data / weight gradients autograd derives from it).
#include "conv_ffma.h"

namespace mau {
namespace {

template <typename T> __device__ __forceinline__ float ld(const T* p);
template <> __device__ __forceinline__ float ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st(T* p, float v);
template <> __device__ __forceinline__ void st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int FT_H = 8, FT_W = 16, FT_N = 64, FT_K = 8;
constexpr int XS_W = FT_W + 2 + 1;  // +1 pad against bank conflicts

// grid: (tiles_w * tiles_h * B, ceil(Cout / 64)); block 256
template <typename T>
__global__ void __launch_bounds__(256) conv3x3_ffma_kernel(ConvFfmaParams p) {
  __shared__ float xs[FT_K][FT_H + 2][XS_W];
  __shared__ __align__(16) float ws[9][FT_K][FT_N];
  const int tiles = p.tiles_w * p.tiles_h;
  const int b = blockIdx.x / tiles;
  const int tr = blockIdx.x - b * tiles;
  const int h0 = (tr / p.tiles_w) * FT_H;
  const int w0 = (tr % p.tiles_w) * FT_W;
  const int n0 = blockIdx.y * FT_N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ph = lane >> 2, pw = (lane & 3) * 4;  // this thread's 4 pixels: row ph, cols pw..pw+3
  const int co = warp * 8;                        // this warp's 8 output channels (relative to n0)
  const T* x = static_cast<const T*>(p.x);

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int kp = 0;
  for (int sg = 0; sg < p.nseg; ++sg) {
    for (int kc = 0; kc < p.seg_chunks[sg]; ++kc, kp += FT_K) {
      const int cA = p.seg_start[sg] + kc * FT_K;
      __syncthreads();
      // stage the halo tile: 10 x 18 pixels x 8 channels
      for (int i = threadIdx.x; i < (FT_H + 2) * (FT_W + 2); i += 256) {
        const int hh = i / (FT_W + 2), ww = i - hh * (FT_W + 2);
        const int h = h0 + hh - 1, w = w0 + ww - 1;
        const bool in = (h >= 0 && h < p.H && w >= 0 && w < p.W);
        const T* px = x + (((long long)b * p.H + (in ? h : 0)) * p.W + (in ? w : 0)) * p.x_cs;
#pragma unroll
        for (int c = 0; c < FT_K; ++c) xs[c][hh][ww] = (in && cA + c < p.x_C) ? ld<T>(px + cA + c) : 0.f;
      }
      // stage the weights: [9][8][64] from packed [9][Kp][Cout_pad]
      for (int i = threadIdx.x; i < 9 * FT_K * FT_N; i += 256) {
        const int n = i & (FT_N - 1);
        const int c = (i >> 6) & (FT_K - 1);
        const int t = i / (FT_K * FT_N);
        ws[t][c][n] = (n0 + n < p.n_rows) ? p.w[((long long)t * p.Kp + kp + c) * p.n_rows + n0 + n] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int r = t / 3, s = t % 3;
#pragma unroll
        for (int c = 0; c < FT_K; ++c) {
          float xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = xs[c][ph + r][pw + i + s];
          const float4 wa = *reinterpret_cast<const float4*>(&ws[t][c][co]);
          const float4 wb = *reinterpret_cast<const float4*>(&ws[t][c][co + 4]);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
        }
      }
    }
  }
  // epilogue
  T* y = static_cast<T*>(p.y);
  const int h = h0 + ph;
  if (h >= p.H) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int w = w0 + pw + i;
    if (w >= p.W) continue;
    T* py = y + (((long long)b * p.H + h) * p.W + w) * p.y_cs + p.y_c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + co + j;
      if (c < p.Cout) {
        float v = acc[i][j] * (p.scale ? p.scale[c] : 1.f) + (p.shift ? p.shift[c] : 0.f);
        if (p.relu) v = fmaxf(v, 0.f);
        if (p.accumulate) v += ld<T>(py + c);
        st<T>(py + c, v);
      }
    }
  }
}

// OIHW fp32 -> [9][Kp][n_rows] fp32
__global__ void pack_w_ffma_fwd_kernel(const float* __restrict__ w, int Cout, int Cin, const int* __restrict__ kmap,
                                       int Kp, float* __restrict__ out) {
  const long long total = 9LL * Kp * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Cout);
    const long long r = i / Cout;
    const int kp = (int)(r % Kp);
    const int t = (int)(r / Kp);
    const int ci = kmap[kp];
    out[i] = ci >= 0 ? w[((long long)n * Cin + ci) * 9 + t] : 0.f;
  }
}
// dgrad pack: out[t][kp = co][n = ci - ci0] = W[co][ci][8 - t]
__global__ void pack_w_ffma_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int ci0, int N, int Kp,
                                         float* __restrict__ out) {
  const long long total = 9LL * Kp * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long r = i / N;
    const int kp = (int)(r % Kp);
    const int t = (int)(r / Kp);
    out[i] = kp < Cout ? w[((long long)kp * Cin + ci0 + n) * 9 + (8 - t)] : 0.f;
  }
}

// weight gradient, direct: one block per (co, ci-group); reduction over all pixels.
// dW[co][ci][t] (+)= sum_{b,h,w} dy[b,h,w,co] * x[b,h+r-1,w+s-1,ci]
template <typename T>
__global__ void __launch_bounds__(256) wgrad_ffma_kernel(const T* __restrict__ x, int x_cs, int x_c0,
                                                         const T* __restrict__ dy, int dy_cs, int dy_c0, int B,
                                                         int H, int W, int Cin, int Cout, int ci_w0, int Cin_w,
                                                         float* __restrict__ dw, int accumulate) {
  // block handles output channel co = blockIdx.x and 8 input channels starting at blockIdx.y*8
  const int co = blockIdx.x;
  const int ci0 = blockIdx.y * 8;
  float acc[8][9];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[i][t] = 0.f;
  const long long npix = (long long)B * H * W;
  for (long long pidx = threadIdx.x; pidx < npix; pidx += 256) {
    const int w = (int)(pidx % W);
    const long long r1 = pidx / W;
    const int h = (int)(r1 % H);
    const int b = (int)(r1 / H);
    const float g = ld<T>(dy + pidx * dy_cs + dy_c0 + co);
    if (g == 0.f) continue;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const T* px = x + (((long long)b * H + hh) * W + ww) * x_cs + x_c0 + ci0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (ci0 + i < Cin) acc[i][t] = fmaf(g, ld<T>(px + i), acc[i][t]);
    }
  }
  __shared__ float red[8][72];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v = acc[i][t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][i * 9 + t] = v;
    }
  __syncthreads();
  if (threadIdx.x < 72) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    const int i = threadIdx.x / 9, t = threadIdx.x % 9;
    if (ci0 + i < Cin) {
      float* d = dw + ((long long)co * Cin_w + ci_w0 + ci0 + i) * 9 + t;
      *d = accumulate ? *d + v : v;
    }
  }
}

}  // namespace

int conv_ffma_launch(int dtype, const ConvFfmaParams& p, int B, cudaStream_t st) {
  dim3 grid((unsigned)(B * p.tiles_w * p.tiles_h), (unsigned)ceil_div(p.Cout, FT_N), 1);
  if (dtype == DT_F32) conv3x3_ffma_kernel<float><<<grid, 256, 0, st>>>(p);
  else                 conv3x3_ffma_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  MAU_LAUNCHED();
  return 0;
}

int conv_ffma_prepare(ConvFfmaParams* p, const View& xbuf, int nseg, const int* seg_start, const int* seg_len,
                      const float* wpacked, int Kp, int n_rows, const View& y, const float* scale,
                      const float* shift, int relu, int accumulate) {
  if (nseg < 1 || nseg > 4) return fail("conv_ffma: 1..4 segments supported");
  *p = ConvFfmaParams();
  p->x = xbuf.ptr; p->x_cs = xbuf.cs; p->x_C = xbuf.C;
  p->H = y.H; p->W = y.W;
  p->tiles_w = ceil_div(y.W, FT_W); p->tiles_h = ceil_div(y.H, FT_H);
  p->nseg = nseg;
  int chunks = 0;
  for (int i = 0; i < nseg; ++i) {
    p->seg_start[i] = seg_start[i];
    p->seg_chunks[i] = ceil_div(round_up(seg_len[i], 64), FT_K);   // same 64-padded K layout as conv_tc
    chunks += p->seg_chunks[i];
  }
  if (chunks * FT_K != Kp) return fail("conv_ffma: packed K mismatch");
  p->w = wpacked; p->Kp = Kp; p->n_rows = n_rows;
  p->y = y.ptr; p->y_cs = y.cs; p->y_c0 = y.c0; p->Cout = y.C;
  p->scale = scale; p->shift = shift; p->relu = relu; p->accumulate = accumulate;
  return 0;
}

int conv_ffma_pack_fwd(const float* w, int Cout, int Cin, const int* kmap_dev, int Kp, float* out, cudaStream_t st) {
  const long long total = 9LL * Kp * Cout;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pack_w_ffma_fwd_kernel<<<blocks, 256, 0, st>>>(w, Cout, Cin, kmap_dev, Kp, out);
  MAU_LAUNCHED();
  return 0;
}
int conv_ffma_pack_dgrad(const float* w, int Cout, int Cin, int ci0, int N, int Kp, float* out, cudaStream_t st) {
  const long long total = 9LL * Kp * N;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pack_w_ffma_dgrad_kernel<<<blocks, 256, 0, st>>>(w, Cout, Cin, ci0, N, Kp, out);
  MAU_LAUNCHED();
  return 0;
}

int wgrad_ffma_launch(int dtype, const View& x_seg, const View& dy, int ci_w0, int Cin_w, float* dw, int accumulate,
                      cudaStream_t st) {
  dim3 grid((unsigned)dy.C, (unsigned)ceil_div(x_seg.C, 8), 1);
  if (dtype == DT_F32)
    wgrad_ffma_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x_seg.ptr), x_seg.cs, x_seg.c0,
                                                   static_cast<const float*>(dy.ptr), dy.cs, dy.c0, dy.B, dy.H, dy.W,
                                                   x_seg.C, dy.C, ci_w0, Cin_w, dw, accumulate);
  else
    wgrad_ffma_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x_seg.ptr), x_seg.cs, x_seg.c0, static_cast<const __nv_bfloat16*>(dy.ptr),
        dy.cs, dy.c0, dy.B, dy.H, dy.W, x_seg.C, dy.C, ci_w0, Cin_w, dw, accumulate);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
