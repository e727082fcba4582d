// embgrad.cu -- backward of a 3x3 convolution w.r.t. an input segment that is CONSTANT over space
// (the embedding planes the U-Net++ concatenates into every decoder node: reference src/model.py:98-108,136-177).
//
// The reference materialises `emb[:, :, None, None].expand(B, 128, H, W)`, so autograd computes a full data
// gradient for those 128 planes and then sums it over space, and a full weight gradient against the planes.
// Both collapse algebraically.  With S[b][t][co] = sum of dz[b, p, co] over the pixels p whose tap-t neighbour
// p + d(t) lies inside the image (the others see zero padding):
//     d emb[b][c]        = sum_{t, co} W[co][c][t] * S[b][t][co]
//     dW[co][c][t]       = sum_b       S[b][t][co] * emb[b][c]
// and S follows from nine numbers per (image, channel): the total, the four border rows / columns and the four
// corners (inclusion-exclusion).  Replaces a dgrad and a wgrad launch per decoder node (5.9 % of the U-Net++
// backward FLOPs) plus the spatial reduction of the embedding gradient by one pass over dz and two tiny kernels.
// The forward convolution over the planes is left dense, so forward numerics are untouched.
#include "ops.h"
#include "vec.cuh"
#include <algorithm>

namespace mau {
namespace {

template <typename T>
__device__ __forceinline__ const T* at(const DView& v, long long pix, int c) {
  return static_cast<const T*>(v.ptr) + pix * v.cs + v.c0 + c;
}
inline DView dv(const View& v) { return DView{v.ptr, v.B, v.H, v.W, v.cs, v.c0, v.C}; }

// sums9[b][q][C]: q = 0 total, 1 row 0, 2 row H-1, 3 col 0, 4 col W-1, 5..8 corners (0,0) (0,W-1) (H-1,0) (H-1,W-1)
// grid (chunks, B): per-image totals, 4 pixels' loads in flight per thread, one float atomic per channel per block
template <typename T>
__global__ void __launch_bounds__(256) image_total_kernel(DView z, float* __restrict__ sums9) {
  extern __shared__ float red[];   // [256][8]
  using Raw = typename V8<T>::Raw;
  const int G = z.C / 8, L = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  const int b = blockIdx.y;
  const int HW = z.H * z.W;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (pl < L)
    for (int p = p0 + pl; p < p1; p += 4 * L) {
      Raw r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (p + u * L < p1) r[u] = V8<T>::load_raw(at<T>(z, (long long)b * HW + p + u * L, gi * 8));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (p + u * L < p1) {
          float v[8];
          V8<T>::unpack(r[u], v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += v[k];
        }
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < z.C; c += 256) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += red[(l * G + (c >> 3)) * 8 + (c & 7)];
    atomicAdd(&sums9[((long long)b * 9 + 0) * z.C + c], s);
  }
}

// grid (4, B): one block per (image, side): row 0, row H-1, column 0, column W-1; G channel groups x L pixel lanes,
// reduced through shared memory; the row blocks also copy the two corners of their row
template <typename T>
__global__ void __launch_bounds__(256) image_border_kernel(DView z, float* __restrict__ sums9) {
  __shared__ float red[256 * 8];
  using Raw = typename V8<T>::Raw;
  const int side = blockIdx.x, b = blockIdx.y, H = z.H, W = z.W, C = z.C;
  const int G = C / 8, L = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  const long long base = (long long)b * H * W;
  const int n = side < 2 ? W : H;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (pl < L)
    for (int i = pl; i < n; i += L) {
      const long long pix = side == 0 ? i : (side == 1 ? (long long)(H - 1) * W + i : (side == 2 ? (long long)i * W : (long long)i * W + W - 1));
      float v[8];
      V8<T>::unpack(V8<T>::load_raw(at<T>(z, base + pix, gi * 8)), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  float* o = sums9 + (long long)b * 9 * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += red[(l * G + (c >> 3)) * 8 + (c & 7)];
    o[(1 + side) * C + c] = s;
    if (side < 2) {
      const long long row = side == 0 ? 0 : (long long)(H - 1) * W;
      const T* p0 = at<T>(z, base + row, c);
      const T* p1 = at<T>(z, base + row + W - 1, c);
      float v0, v1;
      if constexpr (sizeof(T) == 2) { v0 = __bfloat162float(*p0); v1 = __bfloat162float(*p1); } else { v0 = *p0; v1 = *p1; }
      o[(5 + 2 * side) * C + c] = v0;
      o[(6 + 2 * side) * C + c] = v1;
    }
  }
}

// S[t] of (image, channel) from the nine sums; t = r * 3 + s, the neighbour is p + (r - 1, s - 1)
__device__ __forceinline__ void tap_sums(const float* __restrict__ q, int C, float (&S)[9]) {
  const float T = q[0], r0 = q[1 * C], rl = q[2 * C], c0 = q[3 * C], cl = q[4 * C];
  const float k00 = q[5 * C], k0l = q[6 * C], kl0 = q[7 * C], kll = q[8 * C];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      float v = T;
      if (r == 0) v -= r0;          // row 0 has no neighbour above
      if (r == 2) v -= rl;
      if (s == 0) v -= c0;
      if (s == 2) v -= cl;
      if (r == 0 && s == 0) v += k00;
      if (r == 0 && s == 2) v += k0l;
      if (r == 2 && s == 0) v += kl0;
      if (r == 2 && s == 2) v += kll;
      S[r * 3 + s] = v;
    }
}

// blocks [0, nbw): dW[co][ci0 + c][t] = sum_b S[b][t][co] * emb[b][c]   (one thread per (co, c))
// blocks [nbw, ..): demb[b][c] += sum_{t, co} W[co][ci0 + c][t] * S[b][t][co]   (one block per (b, 32 output channels))
constexpr int kCoChunk = 32;
template <bool ROUND_BF16>
__global__ void __launch_bounds__(256) emb_grad_kernel(const float* __restrict__ sums9, int B, int Cout, int Cin, int ci0,
                                                       int E, const float* __restrict__ emb, int emb_stride,
                                                       const float* __restrict__ w, float* __restrict__ dw,
                                                       float* __restrict__ demb, int nbw) {
  auto rnd = [](float v) -> float { return ROUND_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : v; };
  if ((int)blockIdx.x < nbw) {
    if (!dw) return;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Cout * E) return;
    const int c = i % E, co = i / E;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int b = 0; b < B; ++b) {
      float S[9];
      tap_sums(sums9 + (long long)b * 9 * Cout + co, Cout, S);
      const float e = rnd(emb[(long long)b * emb_stride + c]);      // the planes the forward convolved were stored rounded
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[t] = fmaf(S[t], e, acc[t]);
    }
    float* o = dw + ((long long)co * Cin + ci0 + c) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) o[t] = acc[t];
  } else {
    // one block per (image, chunk of kCoChunk output channels): partial sums over the chunk, one atomic per (b, c)
    const int idx = blockIdx.x - nbw;
    const int chunks = (Cout + kCoChunk - 1) / kCoChunk;
    const int b = idx / chunks, co0 = (idx % chunks) * kCoChunk;
    const int nco = min(kCoChunk, Cout - co0);
    __shared__ float sS[9 * kCoChunk];
    for (int i = threadIdx.x; i < nco; i += 256) {
      float S[9];
      tap_sums(sums9 + (long long)b * 9 * Cout + co0 + i, Cout, S);
#pragma unroll
      for (int t = 0; t < 9; ++t) sS[t * kCoChunk + i] = S[t];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < E; c += 256) {
      float a = 0.f;
      for (int i = 0; i < nco; ++i) {
        const float* wp = w + ((long long)(co0 + i) * Cin + ci0 + c) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(rnd(wp[t]), sS[t * kCoChunk + i], a);
      }
      atomicAdd(&demb[(long long)b * emb_stride + c], a);
    }
  }
}

}  // namespace

size_t emb_grad_scratch_floats(int B, int Cout) { return (size_t)B * 9 * Cout; }

int op_emb_segment_grad(int dt, const View& dz, const float* w_oihw, int Cin, int ci0, int E, const float* emb,
                        int emb_stride, float* dw_oihw, float* demb, float* scratch, cudaStream_t st) {
  if (dz.C % 8 || dz.cs % 8 || dz.c0 % 8 || dz.C > 2048) return fail("emb_segment_grad: bad dz view");
  if (dz.H < 2 || dz.W < 2 || dz.B > 65535) return fail("emb_segment_grad: unsupported geometry");
  const int C = dz.C;
  MAU_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * emb_grad_scratch_floats(dz.B, C), st));
  const int L = std::max(1, 256 / (C / 8));
  const int chunks = std::max(1, std::min(ceil_div(dz.H * dz.W, L * 16), ceil_div(148 * 4, dz.B)));
  const dim3 grid((unsigned)chunks, (unsigned)dz.B, 1);
  if (dt == DT_BF16) {
    image_total_kernel<__nv_bfloat16><<<grid, 256, 256 * 8 * sizeof(float), st>>>(dv(dz), scratch);
    MAU_LAUNCHED();
    image_border_kernel<__nv_bfloat16><<<dim3(4, (unsigned)dz.B, 1), 256, 0, st>>>(dv(dz), scratch);
    MAU_LAUNCHED();
  } else {
    image_total_kernel<float><<<grid, 256, 256 * 8 * sizeof(float), st>>>(dv(dz), scratch);
    MAU_LAUNCHED();
    image_border_kernel<float><<<dim3(4, (unsigned)dz.B, 1), 256, 0, st>>>(dv(dz), scratch);
    MAU_LAUNCHED();
  }
  const int nbw = ceil_div(C * E, 256);
  const int nbe = dz.B * ceil_div(C, kCoChunk);
  if (dt == DT_BF16)
    emb_grad_kernel<true><<<nbw + nbe, 256, 0, st>>>(scratch, dz.B, C, Cin, ci0, E, emb, emb_stride, w_oihw, dw_oihw, demb, nbw);
  else
    emb_grad_kernel<false><<<nbw + nbe, 256, 0, st>>>(scratch, dz.B, C, Cin, ci0, E, emb, emb_stride, w_oihw, dw_oihw, demb, nbw);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
