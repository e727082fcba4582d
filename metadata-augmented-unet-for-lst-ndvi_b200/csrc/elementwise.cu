// elementwise.cu -- bandwidth-bound NHWC kernels: layout transposes, 2x2 max-pool (+backward),
// bilinear align_corners resize (+backward), embedding broadcast / reduction, slice copies and the
// 1x1 output head.  All of them move 8 channels (16 B bf16 / 32 B fp32) per thread access so that a
// warp touches whole 128-byte lines; grids are sized in multiples of the SM count and grid-stride.
#include "ops.h"
#include "vec.cuh"
#include <algorithm>

namespace mau {
namespace {

constexpr int kSMs = 148;
inline int grid_for(long long work_items, int threads = 256, int max_waves = 8) {
  long long b = (work_items + threads - 1) / threads;
  long long cap = (long long)kSMs * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
inline DView dv(const View& v) { return DView{v.ptr, v.B, v.H, v.W, v.cs, v.c0, v.C}; }

template <typename T>
__device__ __forceinline__ T* at(const DView& v, long long pix, int c) {
  return static_cast<T*>(v.ptr) + pix * v.cs + v.c0 + c;
}

// ------------------------------------------------------------------ layout
// x [B][C][P] fp32 -> y [B][P][cs] (T), P = H*W.  Tiles of 64 pixels x 32 channels through shared
// memory: coalesced 128-byte reads along P, 8-channel vector stores along C (pad channels up to the
// view's 8-multiple are written as zeros).
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, int B, int C, int P, DView y) {
  // tile = 128 pixels x 32 channels; plane rows are read as float4 (16 B) when P % 4 == 0
  __shared__ float tile[32][129];
  const int p_tiles = (P + 127) / 128, c_tiles = (C + 31) / 32;
  const long long total = (long long)B * p_tiles * c_tiles;
  const int C8 = (C + 7) & ~7;
  const bool vec4 = (P & 3) == 0;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = (int)(t % c_tiles);
    const long long r = t / c_tiles;
    const int pt = (int)(r % p_tiles);
    const int b = (int)(r / p_tiles);
    __syncthreads();
    if (vec4) {
      for (int i = threadIdx.x; i < 32 * 32; i += 256) {          // 32 channels x 32 float4
        const int j = i >> 5, q = i & 31;
        const int c = ct * 32 + j, p = pt * 128 + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C && p < P) v = *reinterpret_cast<const float4*>(x + ((long long)b * C + c) * P + p);
        tile[j][q * 4] = v.x; tile[j][q * 4 + 1] = v.y; tile[j][q * 4 + 2] = v.z; tile[j][q * 4 + 3] = v.w;
      }
    } else {
      for (int i = threadIdx.x; i < 32 * 128; i += 256) {
        const int j = i >> 7, px = i & 127;
        const int c = ct * 32 + j, p = pt * 128 + px;
        tile[j][px] = (c < C && p < P) ? x[((long long)b * C + c) * P + p] : 0.f;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * 4; i += 256) {            // 128 pixels x 4 groups of 8 channels
      const int px = i >> 2, cg = i & 3;
      const int p = pt * 128 + px, c = ct * 32 + cg * 8;
      if (p < P && c < C8 && c + 8 <= y.cs - y.c0) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = tile[cg * 8 + k][px];
        V8<T>::store(at<T>(y, (long long)b * P + p, c), v);
      }
    }
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(DView x, float* __restrict__ y, int P) {
  __shared__ float tile[32][33];
  const int C = x.C, B = x.B;
  const int p_tiles = (P + 31) / 32, c_tiles = (C + 31) / 32;
  const long long total = (long long)B * p_tiles * c_tiles;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = (int)(t % c_tiles);
    const long long r = t / c_tiles;
    const int pt = (int)(r % p_tiles);
    const int b = (int)(r / p_tiles);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int p = pt * 32 + j, c = ct * 32 + tx;
      float v = 0.f;
      if (p < P && c < C) {
        const T* src = at<T>(x, (long long)b * P + p, c);
        if constexpr (sizeof(T) == 2) v = __bfloat162float(*src); else v = *src;
      }
      tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = ct * 32 + j, p = pt * 32 + tx;
      if (c < C && p < P) y[((long long)b * C + c) * P + p] = tile[tx][j];
    }
  }
}

// ------------------------------------------------------------------ max-pool
template <typename T>
__global__ void maxpool_kernel(DView x, DView y) {
  const int G = y.C / 8;
  const long long total = (long long)y.B * y.H * y.W * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int ow = (int)(r % y.W); r /= y.W;
    const int oh = (int)(r % y.H);
    const int b = (int)(r / y.H);
    const long long p00 = ((long long)b * x.H + 2 * oh) * x.W + 2 * ow;
    float a[8], bb[8], c[8], d[8], o[8];
    V8<T>::load(at<T>(x, p00, g * 8), a);
    V8<T>::load(at<T>(x, p00 + 1, g * 8), bb);
    V8<T>::load(at<T>(x, p00 + x.W, g * 8), c);
    V8<T>::load(at<T>(x, p00 + x.W + 1, g * 8), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaxf(a[k], bb[k]), fmaxf(c[k], d[k]));
    V8<T>::store(at<T>(y, ((long long)b * y.H + oh) * y.W + ow, g * 8), o);
  }
}

// one thread per 2x2 window (windows also cover the odd trailing row/col, which get no gradient)
template <typename T>
__global__ void maxpool_bwd_kernel(DView x, DView gy, DView add, int has_add, DView gx) {
  const int G = x.C / 8;
  const int WH = (x.H + 1) / 2, WW = (x.W + 1) / 2;
  const long long total = (long long)x.B * WH * WW * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int ow = (int)(r % WW); r /= WW;
    const int oh = (int)(r % WH);
    const int b = (int)(r / WH);
    const bool full = (2 * oh + 1 < x.H) && (2 * ow + 1 < x.W);   // a real pooling window
    float gv[8];
    float xv[4][8];
    int best[8];
    if (full) {
      V8<T>::load(at<T>(gy, ((long long)b * gy.H + oh) * gy.W + ow, g * 8), gv);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        V8<T>::load(at<T>(x, ((long long)b * x.H + 2 * oh + (q >> 1)) * x.W + 2 * ow + (q & 1), g * 8), xv[q]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int bi = 0; float bv = xv[0][k];
#pragma unroll
        for (int q = 1; q < 4; ++q) if (xv[q][k] > bv) { bv = xv[q][k]; bi = q; }
        best[k] = bi;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int h = 2 * oh + (q >> 1), w = 2 * ow + (q & 1);
      if (h >= x.H || w >= x.W) continue;
      const long long pix = ((long long)b * x.H + h) * x.W + w;
      float o[8];
      if (has_add) V8<T>::load(at<T>(add, pix, g * 8), o);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
      }
      if (full) {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (best[k] == q) o[k] += gv[k];
      }
      V8<T>::store(at<T>(gx, pix, g * 8), o);
    }
  }
}

// ------------------------------------------------------------------ bilinear
// grid: x over (ow, channel group) of one output row, y = b * Hout + oh.  Source index and lambda are
// recomputed with the same fp32 operations ATen uses (scale * dst, truncation), so no table loads sit
// in front of the data loads.
#include "bilinear_index.cuh"
// One thread = one (output column, 8-channel group) for kRows consecutive output rows: all 4*kRows source
// vectors are requested before the first is used (bytes in flight per thread, not occupancy, is what
// buys bandwidth here); neighbouring rows share source rows, which L1 serves.
constexpr int kRows = 4;
template <typename T>
__global__ void __launch_bounds__(256) bilinear_kernel(DView x, DView y, float sy, float sx, FastDiv divG, FastDiv divQ) {
  using Raw = typename V8<T>::Raw;
  const unsigned G = y.C / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)y.W * G) return;
  unsigned ow, g, b, q;
  divG.divmod(idx, ow, g);
  divQ.divmod(blockIdx.y, b, q);          // q = group of kRows output rows
  int x0, x1; float lx;
  src_index(sx, (int)ow, x.W, x0, x1, lx);
  const float hx = 1.f - lx;
  const long long base = (long long)b * x.H;
  const int oh0 = (int)q * kRows;
  Raw r[kRows][4];
  float ly[kRows];
#pragma unroll
  for (int i = 0; i < kRows; ++i) {
    const int oh = oh0 + i;
    if (oh < y.H) {
      int y0, y1;
      src_index(sy, oh, x.H, y0, y1, ly[i]);
      r[i][0] = V8<T>::load_raw(at<T>(x, (base + y0) * x.W + x0, g * 8));
      r[i][1] = V8<T>::load_raw(at<T>(x, (base + y0) * x.W + x1, g * 8));
      r[i][2] = V8<T>::load_raw(at<T>(x, (base + y1) * x.W + x0, g * 8));
      r[i][3] = V8<T>::load_raw(at<T>(x, (base + y1) * x.W + x1, g * 8));
    }
  }
#pragma unroll
  for (int i = 0; i < kRows; ++i) {
    const int oh = oh0 + i;
    if (oh < y.H) {
      float a[8], bb[8], c[8], d[8], o[8];
      V8<T>::unpack(r[i][0], a); V8<T>::unpack(r[i][1], bb); V8<T>::unpack(r[i][2], c); V8<T>::unpack(r[i][3], d);
      const float hy = 1.f - ly[i];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = hy * (hx * a[k] + lx * bb[k]) + ly[i] * (hx * c[k] + lx * d[k]);
      V8<T>::store(at<T>(y, ((long long)b * y.H + oh) * y.W + ow, g * 8), o);
    }
  }
}
// Streaming form for up-sampling (Hin <= Hout): one thread owns one (output column, 8-channel group) and walks
// down a strip of output rows keeping the two horizontally interpolated source rows it sits between in
// registers.  Every source vector is loaded and unpacked once per strip and every horizontal lerp is computed
// once (the per-row kernel above recomputes them for each of the ~2 output rows that share a source row pair):
// ~2.4x fewer instructions, which is what bounded that kernel (ncu: 75 % issue-slot utilisation at 42 % of HBM
// peak).  The next source row is requested one step before it is needed.
template <typename T>
__global__ void __launch_bounds__(256) bilinear_stream_kernel(DView x, DView y, float sy, float sx, FastDiv divG,
                                                              int rows_per_strip) {
  using Raw = typename V8<T>::Raw;
  const unsigned G = y.C / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)y.W * G) return;
  unsigned ow, g;
  divG.divmod(idx, ow, g);
  const int b = blockIdx.z;
  const int oh_begin = blockIdx.y * rows_per_strip;
  const int oh_end = min(y.H, oh_begin + rows_per_strip);
  int x0, x1; float lx;
  src_index(sx, (int)ow, x.W, x0, x1, lx);
  const float hx = 1.f - lx;
  const T* xb = static_cast<const T*>(x.ptr) + (long long)b * x.H * x.W * x.cs + x.c0 + g * 8;
  T* yb = static_cast<T*>(y.ptr) + ((long long)b * y.H * y.W + ow) * y.cs + y.c0 + g * 8;
  const long long xrow = (long long)x.W * x.cs, yrow = (long long)y.W * y.cs;
  const long long o0 = (long long)x0 * x.cs, o1 = (long long)x1 * x.cs;
  auto hlerp = [&](const Raw& ra, const Raw& rb, float (&out)[8]) {
    float a[8], bb[8];
    V8<T>::unpack(ra, a); V8<T>::unpack(rb, bb);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = hx * a[k] + lx * bb[k];
  };
  int y0, y1; float ly;
  src_index(sy, oh_begin, x.H, y0, y1, ly);
  float cur[8], nxt[8];
  {
    const Raw a0 = V8<T>::load_raw(xb + y0 * xrow + o0), a1 = V8<T>::load_raw(xb + y0 * xrow + o1);
    const Raw b0 = V8<T>::load_raw(xb + y1 * xrow + o0), b1 = V8<T>::load_raw(xb + y1 * xrow + o1);
    hlerp(a0, a1, cur);
    hlerp(b0, b1, nxt);
  }
  int cy = y0;
  Raw pa, pb;                                   // source row cy + 2, requested ahead of time
  bool have = cy + 2 < x.H;
  if (have) { pa = V8<T>::load_raw(xb + (cy + 2) * xrow + o0); pb = V8<T>::load_raw(xb + (cy + 2) * xrow + o1); }
  for (int oh = oh_begin; oh < oh_end; ++oh) {
    src_index(sy, oh, x.H, y0, y1, ly);         // block-uniform
    if (y0 != cy) {                             // moved down by one source row (scale <= 1)
#pragma unroll
      for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
      cy = y0;
      if (have) hlerp(pa, pb, nxt);             // row cy + 1; at the last source row nxt keeps cur (weight ly = 0)
      have = cy + 2 < x.H;
      if (have) { pa = V8<T>::load_raw(xb + (cy + 2) * xrow + o0); pb = V8<T>::load_raw(xb + (cy + 2) * xrow + o1); }
    }
    const float hy = 1.f - ly;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = hy * cur[k] + ly * nxt[k];
    V8<T>::store(yb + oh * yrow, o);
  }
}

// gather form: gx[ih,iw] = sum_{(oh,wy) in rows(ih)} sum_{(ow,wx) in cols(iw)} wy*wx*gy[oh,ow]
// Fast path (every source column receives at most kMaxE contributions -- always true when up-sampling):
// the column list is read once, then for every contributing output row the <= kMaxE gy vectors are
// requested as one batch.
constexpr int kMaxE = kBilinearMaxFan;
template <typename T>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(DView gy, DView gx, BilinearTables t, int accumulate) {
  using Raw = typename V8<T>::Raw;
  const int G = gx.C / 8;
  const long long total = (long long)gx.B * gx.H * gx.W * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int iw = (int)(r % gx.W); r /= gx.W;
    const int ih = (int)(r % gx.H);
    const int b = (int)(r / gx.H);
    const int ra = t.ty_off[ih], rb = t.ty_off[ih + 1];
    const int ca = t.tx_off[iw], nc = t.tx_off[iw + 1] - ca;
    int ow[kMaxE];
    float wx[kMaxE];
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) {
      ow[e] = e < nc ? t.tx_idx[ca + e] : 0;
      wx[e] = e < nc ? t.tx_w[ca + e] : 0.f;
    }
    float o[8];
    const long long opix = ((long long)b * gx.H + ih) * gx.W + iw;
    if (accumulate) V8<T>::load(at<T>(gx, opix, g * 8), o);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    }
    for (int a = ra; a < rb; ++a) {        // same summation order as the general kernel: rows outer, columns inner
      const long long rowbase = ((long long)b * gy.H + t.ty_idx[a]) * gy.W;
      const float wy = t.ty_w[a];
      Raw rv[kMaxE];
#pragma unroll
      for (int c = 0; c < kMaxE; ++c)
        if (c < nc) rv[c] = V8<T>::load_raw(at<T>(gy, rowbase + ow[c], g * 8));
#pragma unroll
      for (int c = 0; c < kMaxE; ++c)
        if (c < nc) {
          float v[8];
          V8<T>::unpack(rv[c], v);
          const float wgt = wy * wx[c];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = fmaf(wgt, v[k], o[k]);
        }
    }
    V8<T>::store(at<T>(gx, opix, g * 8), o);
  }
}
// Streaming form of the backward for up-sampling (Hin <= Hout): bilinear_bwd_lean.cuh
#include "bilinear_bwd_lean.cuh"

// general form (any number of contributions per source index, e.g. down-sampling)
template <typename T>
__global__ void bilinear_bwd_general_kernel(DView gy, DView gx, BilinearTables t, int accumulate) {
  const int G = gx.C / 8;
  const long long total = (long long)gx.B * gx.H * gx.W * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int iw = (int)(r % gx.W); r /= gx.W;
    const int ih = (int)(r % gx.H);
    const int b = (int)(r / gx.H);
    float o[8];
    const long long opix = ((long long)b * gx.H + ih) * gx.W + iw;
    if (accumulate) V8<T>::load(at<T>(gx, opix, g * 8), o);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    }
    for (int a = t.ty_off[ih]; a < t.ty_off[ih + 1]; ++a) {
      const int oh = t.ty_idx[a];
      const float wy = t.ty_w[a];
      for (int c = t.tx_off[iw]; c < t.tx_off[iw + 1]; ++c) {
        const int ow = t.tx_idx[c];
        const float wgt = wy * t.tx_w[c];
        float v[8];
        V8<T>::load(at<T>(gy, ((long long)b * gy.H + oh) * gy.W + ow, g * 8), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(wgt, v[k], o[k]);
      }
    }
    V8<T>::store(at<T>(gx, opix, g * 8), o);
  }
}

// ------------------------------------------------------------------ embeddings
// grid (chunks, B); block 256 = G channel groups x L pixel lanes: a thread converts its 8 embedding values once
// and then only stores (16 B per pixel it owns)
template <typename T>
__global__ void __launch_bounds__(256) embed_broadcast_kernel(const float* __restrict__ emb, int stride, DView y) {
  using Raw = typename V8<T>::Raw;
  const int G = y.C / 8;
  const int L = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  if (pl >= L) return;
  const int b = blockIdx.y;
  const int HW = y.H * y.W;
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = emb[(long long)b * stride + gi * 8 + k];
  T* base = static_cast<T*>(y.ptr) + (long long)b * HW * y.cs + y.c0 + gi * 8;
  Raw r;
  V8<T>::store(reinterpret_cast<T*>(&r), o);           // pack once
  for (int p = blockIdx.x * L + pl; p < HW; p += gridDim.x * L) *reinterpret_cast<Raw*>(base + (long long)p * y.cs) = r;
}
// grid (chunks, B); block 256 = (C/8 groups) x (256 / groups pixel lanes)
template <typename T>
__global__ void embed_reduce_kernel(DView g, float* __restrict__ demb, int stride, int chunks) {
  extern __shared__ float red[];   // [256][8]
  const int G = g.C / 8;
  const int lanes = 256 / G;
  const int gi = threadIdx.x % G, pl = threadIdx.x / G;
  const int b = blockIdx.y;
  const long long HW = (long long)g.H * g.W;
  const long long per = (HW + chunks - 1) / chunks;
  const long long p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (pl < lanes)
    for (long long p = p0 + pl; p < p1; p += lanes) {
      float v[8];
      V8<T>::load(at<T>(g, b * HW + p, gi * 8), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < g.C) {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(l * G + c / 8) * 8 + (c & 7)];
    atomicAdd(&demb[(long long)b * stride + c], s);
  }
}

// dst[b, p, :] = src[0, p, :]: one 16-byte load, B stores
template <typename T>
__global__ void broadcast_batch_kernel(DView src, DView dst) {
  using Raw = typename V8<T>::Raw;
  const int G = src.C / 8;
  const long long HW = (long long)src.H * src.W;
  const long long total = HW * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long pix = i / G;
    const Raw r = V8<T>::load_raw(at<T>(src, pix, g * 8));
    for (int b = 0; b < dst.B; ++b) *reinterpret_cast<Raw*>(at<T>(dst, b * HW + pix, g * 8)) = r;
  }
}

template <typename T>
__global__ void copy_slice_kernel(DView src, DView dst, int accumulate) {
  const int G = src.C / 8;
  const long long total = (long long)src.B * src.H * src.W * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long pix = i / G;
    float v[8];
    V8<T>::load(at<T>(src, pix, g * 8), v);
    if (accumulate) {
      float o[8];
      V8<T>::load(at<T>(dst, pix, g * 8), o);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += o[k];
    }
    V8<T>::store(at<T>(dst, pix, g * 8), v);
  }
}

// ------------------------------------------------------------------ 1x1 head
constexpr int kMaxOC = 8;
constexpr int kHU = 4;
// One thread per pixel: 8-channel vectors of the pixel are loaded four at a time, weights are float4 broadcasts
// from shared memory, no cross-lane reduction, NCHW stores coalesced across the warp.  (The first, 8-lanes-per-pixel
// version spent 160 instructions per lane on shuffles, predicates and a divergent tanh: issue-bound at 44 % of HBM
// peak; this one is bound by L1 tag throughput at the same level -- see DESIGN.md 6.)
template <typename T, int OCT>
__global__ void __launch_bounds__(256) head_pix_kernel(DView x, const float* __restrict__ w, const float* __restrict__ bias,
                                                       int OC, int apply_tanh, float* __restrict__ out) {
  using Raw = typename V8<T>::Raw;
  extern __shared__ float sw[];            // [OCT][C]
  const int C = x.C;
  for (int i = threadIdx.x; i < OCT * C; i += 256) sw[i] = (i / C) < OC ? w[i] : 0.f;
  __syncthreads();
  const int P = x.H * x.W;
  const int b = blockIdx.y;
  const T* xb = static_cast<const T*>(x.ptr) + (long long)b * P * x.cs + x.c0;
  float* ob = out + (long long)b * OC * P;
  const int G = C / 8;
  for (int pix = blockIdx.x * 256 + threadIdx.x; pix < P; pix += gridDim.x * 256) {
    const T* px = xb + (long long)pix * x.cs;
    float acc[OCT];
#pragma unroll
    for (int o = 0; o < OCT; ++o) acc[o] = 0.f;
    for (int g0 = 0; g0 < G; g0 += 4) {
      Raw r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (g0 + u < G) r[u] = V8<T>::load_raw(px + (g0 + u) * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (g0 + u < G) {
          float v[8];
          V8<T>::unpack(r[u], v);
#pragma unroll
          for (int o = 0; o < OCT; ++o) {
            const float4 wa = *reinterpret_cast<const float4*>(sw + o * C + (g0 + u) * 8);
            const float4 wb = *reinterpret_cast<const float4*>(sw + o * C + (g0 + u) * 8 + 4);
            acc[o] = fmaf(v[0], wa.x, acc[o]); acc[o] = fmaf(v[1], wa.y, acc[o]);
            acc[o] = fmaf(v[2], wa.z, acc[o]); acc[o] = fmaf(v[3], wa.w, acc[o]);
            acc[o] = fmaf(v[4], wb.x, acc[o]); acc[o] = fmaf(v[5], wb.y, acc[o]);
            acc[o] = fmaf(v[6], wb.z, acc[o]); acc[o] = fmaf(v[7], wb.w, acc[o]);
          }
        }
    }
#pragma unroll
    for (int o = 0; o < OCT; ++o)
      if (o < OC) {
        float v = acc[o] + bias[o];
        if (apply_tanh && o == 0) v = tanhf(v);
        ob[(long long)o * P + pix] = v;
      }
  }
}

// backward of the head: per pixel g_o = gout_o * (o==0 && tanh ? 1 - out_0^2 : 1);
// gx[c] = sum_o g_o W[o][c];  dW[o][c] += g_o x[c];  db[o] += g_o
#ifdef MAU_HEAD_BWD_OCC4      // bf16 / two outputs compiled for four resident blocks per SM (64 registers)
#define MAU_HEAD_BWD_BOUNDS __launch_bounds__(256, (sizeof(T) == 2 && OCT == 2) ? 4 : 1)
#else
#define MAU_HEAD_BWD_BOUNDS __launch_bounds__(256)
#endif
template <typename T, int OCT>
__global__ void MAU_HEAD_BWD_BOUNDS head_bwd_kernel(DView x, const float* __restrict__ w, int OC, int apply_tanh,
                                                       const float* __restrict__ out, const float* __restrict__ gout,
                                                       DView gx, float* __restrict__ dw, float* __restrict__ db) {
  using Raw = typename V8<T>::Raw;
  extern __shared__ float sm[];      // sdw [OC][C] | sdb [OC]
  float* sdw = sm;
  float* sdb = sdw + OC * x.C;
  for (int i = threadIdx.x; i < OC * x.C; i += blockDim.x) sdw[i] = 0.f;
  if (threadIdx.x < OC) sdb[threadIdx.x] = 0.f;
  __syncthreads();
  const int G = x.C / 8;
  const int P = x.H * x.W;
  const int b = blockIdx.y;
  const int per_block = blockDim.x / G;
  const int sub = threadIdx.x % G, slot = threadIdx.x / G;
  const T* xb = static_cast<const T*>(x.ptr) + (long long)b * P * x.cs + x.c0 + sub * 8;
  T* gxb = static_cast<T*>(gx.ptr) + (long long)b * P * gx.cs + gx.c0 + sub * 8;
  const float* goutb = gout + (long long)b * OC * P;
  const float* outb = out + (long long)b * OC * P;
  float wr[OCT][8], dwl[OCT][8], dbl[OCT];
#pragma unroll
  for (int o = 0; o < OCT; ++o) {
    dbl[o] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { dwl[o][k] = 0.f; wr[o][k] = o < OC ? w[o * x.C + sub * 8 + k] : 0.f; }
  }
  const int stride = gridDim.x * per_block;
  for (int base = blockIdx.x * per_block; base < P; base += kHU * stride) {
    Raw r[kHU];
    float go[kHU][OCT];
#pragma unroll
    for (int u = 0; u < kHU; ++u) {
      const int pix = base + u * stride + slot;
      if (pix < P) {
        r[u] = V8<T>::load_raw(xb + (long long)pix * x.cs);
#pragma unroll
        for (int o = 0; o < OCT; ++o) go[u][o] = o < OC ? goutb[(long long)o * P + pix] : 0.f;
        if (apply_tanh) {
          const float y = outb[pix];
          go[u][0] *= (1.f - y * y);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kHU; ++u) {
      const int pix = base + u * stride + slot;
      if (pix < P) {
        float v[8], rr[8];
        V8<T>::unpack(r[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) rr[k] = 0.f;
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            rr[k] = fmaf(go[u][o], wr[o][k], rr[k]);
            dwl[o][k] = fmaf(go[u][o], v[k], dwl[o][k]);
          }
          if (sub == 0) dbl[o] += go[u][o];
        }
        V8<T>::store(gxb + (long long)pix * gx.cs, rr);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < OCT; ++o)
    if (o < OC) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&sdw[o * x.C + sub * 8 + k], dwl[o][k]);
      if (sub == 0) atomicAdd(&sdb[o], dbl[o]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < OC * x.C; i += blockDim.x) atomicAdd(&dw[i], sdw[i]);
  if (threadIdx.x < OC) atomicAdd(&db[threadIdx.x], sdb[threadIdx.x]);
}

inline bool vec_ok(const View& v) { return v.cs % 8 == 0 && v.c0 % 8 == 0 && v.C % 8 == 0; }

}  // namespace

#define MAU_DISPATCH(dt, KERNEL, GRID, BLOCK, SMEM, ST, ...)                            \
  do {                                                                                  \
    if ((dt) == DT_BF16) KERNEL<__nv_bfloat16><<<GRID, BLOCK, SMEM, ST>>>(__VA_ARGS__); \
    else KERNEL<float><<<GRID, BLOCK, SMEM, ST>>>(__VA_ARGS__);                         \
    MAU_LAUNCHED();                                                                     \
  } while (0)

int op_nchw_to_nhwc(int dt, const float* x, int B, int C, int H, int W, const View& y, cudaStream_t st) {
  const int P = H * W;
  if (y.cs % 8 || y.c0 % 8) return fail("nchw_to_nhwc: destination must be 8-channel aligned");
  const long long tiles = (long long)B * ceil_div(P, 128) * ceil_div(C, 32);
  MAU_DISPATCH(dt, nchw_to_nhwc_kernel, grid_for(tiles, 1, 16), 256, 0, st, x, B, C, P, dv(y));
  return 0;
}
int op_nhwc_to_nchw(int dt, const View& x, float* y, cudaStream_t st) {
  const int P = x.H * x.W;
  const long long tiles = (long long)x.B * ceil_div(P, 32) * ceil_div(x.C, 32);
  MAU_DISPATCH(dt, nhwc_to_nchw_kernel, grid_for(tiles, 1, 16), 256, 0, st, dv(x), y, P);
  return 0;
}
int op_maxpool(int dt, const View& x, const View& y, cudaStream_t st) {
  if (!vec_ok(x) || !vec_ok(y) || x.C != y.C || y.H != x.H / 2 || y.W != x.W / 2)
    return fail("maxpool: bad views (C=%d/%d, %dx%d -> %dx%d)", x.C, y.C, x.H, x.W, y.H, y.W);
  MAU_DISPATCH(dt, maxpool_kernel, grid_for(y.pixels() * (y.C / 8)), 256, 0, st, dv(x), dv(y));
  return 0;
}
int op_maxpool_bwd(int dt, const View& x, const View& gy, const View* addend, const View& gx, cudaStream_t st) {
  if (!vec_ok(x) || !vec_ok(gy) || !vec_ok(gx) || (addend && !vec_ok(*addend))) return fail("maxpool_bwd: bad views");
  const long long items = (long long)x.B * ((x.H + 1) / 2) * ((x.W + 1) / 2) * (x.C / 8);
  MAU_DISPATCH(dt, maxpool_bwd_kernel, grid_for(items), 256, 0, st, dv(x), dv(gy), addend ? dv(*addend) : dv(x),
               addend ? 1 : 0, dv(gx));
  return 0;
}

int op_bilinear(int dt, const View& x, const View& y, const BilinearTables& t, cudaStream_t st) {
  if (!vec_ok(x) || !vec_ok(y) || x.C != y.C || t.Hin != x.H || t.Win != x.W || t.Hout != y.H || t.Wout != y.W)
    return fail("bilinear: bad views/tables");
  const float sy = y.H > 1 ? (float)(x.H - 1) / (float)(y.H - 1) : 0.f;     // area_pixel_compute_scale, align_corners
  const float sx = y.W > 1 ? (float)(x.W - 1) / (float)(y.W - 1) : 0.f;
  const int colblocks = ceil_div(y.W * (y.C / 8), 256);
  if (x.H <= y.H && y.B <= 65535 && x.H >= 2) {      // up-sampling rows: streaming kernel
    int strip = 32;
    while (strip > 8 && (long long)colblocks * y.B * ceil_div(y.H, strip) < 148 * 8) strip >>= 1;
    const dim3 grid((unsigned)colblocks, (unsigned)ceil_div(y.H, strip), (unsigned)y.B);
    MAU_DISPATCH(dt, bilinear_stream_kernel, grid, 256, 0, st, dv(x), dv(y), sy, sx, FastDiv((unsigned)(y.C / 8)), strip);
    return 0;
  }
  const int quads = ceil_div(y.H, kRows);
  if ((long long)y.B * quads > 65535) return fail("bilinear: B*H too large for the row grid");
  const dim3 grid((unsigned)colblocks, (unsigned)(y.B * quads), 1);
  MAU_DISPATCH(dt, bilinear_kernel, grid, 256, 0, st, dv(x), dv(y), sy, sx, FastDiv((unsigned)(y.C / 8)), FastDiv((unsigned)quads));
  return 0;
}
int op_bilinear_bwd(int dt, const View& gy, const View& gx, const BilinearTables& t, int accumulate,
                    cudaStream_t st) {
  if (!vec_ok(gx) || !vec_ok(gy) || gx.C != gy.C || t.Hin != gx.H || t.Win != gx.W || t.Hout != gy.H ||
      t.Wout != gy.W)
    return fail("bilinear_bwd: bad views/tables");
  if (!t.no_stream && t.max_fan_w <= kMaxE && gx.H <= gy.H && gx.H >= 2 && gx.B <= 65535 &&
      (long long)gy.H * gy.W * gy.cs < (1ll << 31)) {
    const float sy = gy.H > 1 ? (float)(gx.H - 1) / (float)(gy.H - 1) : 0.f;
    const int colblocks = ceil_div(gx.W * (gx.C / 8), 256);
    int strip = 16;
    while (strip > 4 && (long long)colblocks * gx.B * ceil_div(gx.H, strip) < 148 * 8) strip >>= 1;
    const dim3 grid((unsigned)colblocks, (unsigned)ceil_div(gx.H, strip), (unsigned)gx.B);
    MAU_DISPATCH(dt, bilinear_bwd_lean_kernel, grid, 256, 0, st, dv(gy), dv(gx), t, sy, FastDiv((unsigned)(gx.C / 8)), strip,
                 accumulate);
    return 0;
  }
  if (t.max_fan_w <= kMaxE)
    MAU_DISPATCH(dt, bilinear_bwd_kernel, grid_for(gx.pixels() * (gx.C / 8)), 256, 0, st, dv(gy), dv(gx), t, accumulate);
  else
    MAU_DISPATCH(dt, bilinear_bwd_general_kernel, grid_for(gx.pixels() * (gx.C / 8)), 256, 0, st, dv(gy), dv(gx), t,
                 accumulate);
  return 0;
}
int op_embed_broadcast(int dt, const float* emb, int emb_stride, const View& y, cudaStream_t st) {
  if (!vec_ok(y) || y.C > 2048 || y.B > 65535) return fail("embed_broadcast: bad view");
  const int L = std::max(1, 256 / (y.C / 8));
  const int per_img = std::max(1, std::min(ceil_div(y.H * y.W, L * 4), ceil_div(148 * 8, y.B)));
  const dim3 grid((unsigned)per_img, (unsigned)y.B, 1);
  MAU_DISPATCH(dt, embed_broadcast_kernel, grid, 256, 0, st, emb, emb_stride, dv(y));
  return 0;
}
int op_embed_reduce(int dt, const View& g, float* demb, int emb_stride, int accumulate, cudaStream_t st) {
  if (!vec_ok(g) || g.C > 256) return fail("embed_reduce: bad view (C=%d)", g.C);
  if (!accumulate) MAU_CUDA(cudaMemsetAsync(demb, 0, sizeof(float) * (size_t)g.B * emb_stride, st));
  const long long HW = (long long)g.H * g.W;
  int chunks = (int)std::min<long long>(std::max<long long>(1, HW / 512), 148 * 4 / std::max(1, g.B) + 1);
  dim3 grid((unsigned)chunks, (unsigned)g.B, 1);
  MAU_DISPATCH(dt, embed_reduce_kernel, grid, 256, 256 * 8 * sizeof(float), st, dv(g), demb, emb_stride, chunks);
  return 0;
}
int op_broadcast_batch(int dt, const View& src, const View& dst, cudaStream_t st) {
  if (!vec_ok(src) || !vec_ok(dst) || src.C != dst.C || src.B != 1 || src.H != dst.H || src.W != dst.W)
    return fail("broadcast_batch: bad views");
  MAU_DISPATCH(dt, broadcast_batch_kernel, grid_for((long long)src.H * src.W * (src.C / 8)), 256, 0, st, dv(src), dv(dst));
  return 0;
}
int op_copy_slice(int dt, const View& src, const View& dst, int accumulate, cudaStream_t st) {
  if (!vec_ok(src) || !vec_ok(dst) || src.C != dst.C || src.pixels() != dst.pixels()) return fail("copy_slice: bad views");
  MAU_DISPATCH(dt, copy_slice_kernel, grid_for(src.pixels() * (src.C / 8)), 256, 0, st, dv(src), dv(dst), accumulate);
  return 0;
}
static bool head_ok(const View& x, int OC) {
  const int G = x.C / 8;
  return x.C % 8 == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0 && OC >= 1 && OC <= kMaxOC && x.cs % 8 == 0 &&
         x.c0 % 8 == 0;
}
#define MAU_HEAD_DISPATCH(KERNEL, GRID, SMEM, ...)                                                        \
  do {                                                                                                    \
    if (dt == DT_BF16) {                                                                                  \
      if (OC <= 2) KERNEL<__nv_bfloat16, 2><<<GRID, 256, SMEM, st>>>(__VA_ARGS__);                        \
      else if (OC <= 4) KERNEL<__nv_bfloat16, 4><<<GRID, 256, SMEM, st>>>(__VA_ARGS__);                   \
      else KERNEL<__nv_bfloat16, 8><<<GRID, 256, SMEM, st>>>(__VA_ARGS__);                                \
    } else {                                                                                              \
      if (OC <= 2) KERNEL<float, 2><<<GRID, 256, SMEM, st>>>(__VA_ARGS__);                                \
      else if (OC <= 4) KERNEL<float, 4><<<GRID, 256, SMEM, st>>>(__VA_ARGS__);                           \
      else KERNEL<float, 8><<<GRID, 256, SMEM, st>>>(__VA_ARGS__);                                        \
    }                                                                                                     \
    MAU_LAUNCHED();                                                                                       \
  } while (0)

template <typename K>
static int head_occupancy(K kernel, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, smem) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
  return n;
}
#define MAU_HEAD_OCCUPANCY(KERNEL, OUT, SMEM)                                                             \
  do {                                                                                                    \
    if (dt == DT_BF16) OUT = OC <= 2 ? head_occupancy(KERNEL<__nv_bfloat16, 2>, SMEM)                     \
                           : (OC <= 4 ? head_occupancy(KERNEL<__nv_bfloat16, 4>, SMEM) : head_occupancy(KERNEL<__nv_bfloat16, 8>, SMEM)); \
    else OUT = OC <= 2 ? head_occupancy(KERNEL<float, 2>, SMEM)                                           \
             : (OC <= 4 ? head_occupancy(KERNEL<float, 4>, SMEM) : head_occupancy(KERNEL<float, 8>, SMEM)); \
  } while (0)

int op_head(int dt, const View& x, const float* w, const float* bias, int OC, int apply_tanh, float* out_nchw,
            cudaStream_t st) {
  if (!head_ok(x, OC)) return fail("head: needs C = 8*2^k <= 256 and out_channels <= 8 (C=%d, OC=%d)", x.C, OC);
  const int per_block = 256 / (x.C / 8);
  if (x.B > 65535) return fail("head: batch too large for the per-image grid");
  (void)per_block;
  const int per_img = std::max(1, std::min(ceil_div(x.H * x.W, 256), ceil_div(148 * 8, x.B)));
  const dim3 grid((unsigned)per_img, (unsigned)x.B, 1);
  const size_t smem = sizeof(float) * (OC <= 2 ? 2 : (OC <= 4 ? 4 : 8)) * x.C;
  MAU_HEAD_DISPATCH(head_pix_kernel, grid, smem, dv(x), w, bias, OC, apply_tanh, out_nchw);
  return 0;
}
int op_head_bwd(int dt, const View& x, const float* w, int OC, int apply_tanh, const float* out_nchw,
                const float* gout_nchw, const View& gx, float* dw, float* db, cudaStream_t st) {
  if (!head_ok(x, OC) || !vec_ok(gx)) return fail("head_bwd: unsupported shape");
  const int per_block = 256 / (x.C / 8);
  const size_t smem = (OC * x.C + OC) * sizeof(float);
  if (x.B > 65535) return fail("head_bwd: batch too large for the per-image grid");
  // one wave of equal blocks: blocks-per-image x batch must not exceed what is resident (80 registers -> 3 blocks per SM
  // for bf16 / 2 outputs; 896 blocks on 444 slots ran as three waves, the last one on 8 SMs)
  int occ = 0;
  MAU_HEAD_OCCUPANCY(head_bwd_kernel, occ, smem);
  const int cap = whole_waves_enabled() ? 148 * occ / x.B : ceil_div(148 * 6, x.B);
  const int per_img = std::max(1, std::min(ceil_div(ceil_div(x.H * x.W, kHU), per_block), cap));
  const dim3 grid((unsigned)per_img, (unsigned)x.B, 1);
  MAU_HEAD_DISPATCH(head_bwd_kernel, grid, smem, dv(x), w, OC, apply_tanh, out_nchw, gout_nchw, dv(gx), dw, db);
  return 0;
}

}  // namespace mau
