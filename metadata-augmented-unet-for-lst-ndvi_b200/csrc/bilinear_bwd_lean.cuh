// bilinear_bwd_lean.cuh -- backward of the align_corners bilinear up-sampling (reference src/model.py:12-17: autograd of
// F.interpolate), streaming form with a lean inner loop.  Included by elementwise.cu inside namespace mau::<anonymous>
// after V8 / DView / BilinearTables / FastDiv / src_index / kMaxE (and by oracle/bilinear_bwd_emu.cpp, which supplies CPU
// stand-ins for those, to check the indexing without a GPU).
//
//   gx[ih, iw] = sum_oh wy(ih, oh) * sum_ow wx(iw, ow) * gy[oh, ow]
//
// One thread owns one (input column, 8-channel group) and walks down the OUTPUT rows that touch its strip of input rows:
// for every output row the column gather hg = sum_c wx_c * gy[oh, ow_c] is formed once and scattered with the two
// vertical weights into two sliding accumulators (input rows y0 and y0 + 1), so every gy vector is read twice in total
// instead of four times, with block-uniform control flow and no per-row table walks.  The first version of this kernel
// was issue-bound (ncu: 63 % issue-slot utilisation at 51 % of HBM peak, ~350 instructions per thread and output row of
// which ~100 are loads, unpacks and FMAs); this one keeps the data flow and trims the bookkeeping (96.4 -> 84.2 us at
// 16 x 250^2 x 128, 56.0 -> 51.4 us at 16 x 124^2 x 256 before the occupancy change below; profiles/r02_training_step.md):
//   * the number of column contributions NC is a template parameter chosen per WARP (2 / 4 / 6: the widest lane decides);
//     lanes with fewer contributions repeat their first one with weight 0, so the loop body has no per-lane predicates
//     (the old body kept every load and FMA group under `c < nc`: 36 branches and 24 convergence barriers per two rows);
//   * one 64-bit pointer per contribution, set up once; inside the loop a load address is pointer + 32-bit row offset
//     (one IMAD.WIDE instead of a 64-bit multiply-add chain per load);
//   * the first term of every sum is a multiply, not an add to a zeroed register.
// A rows-first variant (one thread per OUTPUT column, each gy vector loaded once, columns meeting in shared memory with
// one barrier per finished input row) halved the instruction count again but was no faster alone (96.4 us) and 0.1 ms
// slower inside the training step (the barrier serialises the load latency of all warps of a block); it was dropped.
template <typename T, int NC>
__device__ __forceinline__ void bilinear_bwd_lean_rows(const DView& gy, const DView& gx, const BilinearTables& t, float sy,
                                                       int b, int iw, int g, int ih_b, int ih_e, int accumulate) {
  using Raw = typename V8<T>::Raw;
  const int ca = t.tx_off[iw], nc = t.tx_off[iw + 1] - ca;
  const T* col[NC];                               // gy(b, row 0, ow_c, this channel group)
  float wx[NC];
  const T* gyb = static_cast<const T*>(gy.ptr) + (long long)b * gy.H * gy.W * gy.cs + gy.c0 + g * 8;
#pragma unroll
  for (int e = 0; e < NC; ++e) {
    col[e] = gyb + (e < nc ? t.tx_idx[ca + e] : (nc > 0 ? t.tx_idx[ca] : 0)) * gy.cs;
    wx[e] = e < nc ? t.tx_w[ca + e] : 0.f;
  }
  T* gxb = static_cast<T*>(gx.ptr) + ((long long)b * gx.H * gx.W + iw) * gx.cs + gx.c0 + g * 8;
  const int gyrow = gy.W * gy.cs, gxrow = gx.W * gx.cs;      // one image is < 2^31 elements (checked on the host)
  // output rows touching input rows [ih_b, ih_e): the row lists are sorted by output row
  const int oh_first = t.ty_idx[t.ty_off[ih_b]];
  const int oh_last = t.ty_idx[t.ty_off[ih_e] - 1];
  float acc0[8], acc1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
  int r = ih_b;                                   // input row acc0 belongs to (acc1: r + 1)
  auto finish_row = [&]() {                       // input row r is complete: store it, slide the window
    T* dst = gxb + r * gxrow;
    if (accumulate) {
      float old[8];
      V8<T>::load(dst, old);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc0[k] += old[k];
    }
    V8<T>::store(dst, acc0);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc0[k] = acc1[k]; acc1[k] = 0.f; }
    ++r;
  };
  auto load_row = [&](Raw (&buf)[NC], int oh) {
    const int ro = min(oh, oh_last) * gyrow;      // rows past the strip repeat its last row (never consumed)
#pragma unroll
    for (int c = 0; c < NC; ++c) buf[c] = V8<T>::load_raw(col[c] + ro);
  };
  auto consume = [&](const Raw (&buf)[NC], int oh) {
    int y0, y1; float ly;
    src_index(sy, oh, gx.H, y0, y1, ly);          // block-uniform
    while (y0 > r && r < ih_e) finish_row();
    float hg[8], v[8];
    V8<T>::unpack(buf[0], v);
#pragma unroll
    for (int k = 0; k < 8; ++k) hg[k] = wx[0] * v[k];
#pragma unroll
    for (int c = 1; c < NC; ++c) {
      V8<T>::unpack(buf[c], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) hg[k] = fmaf(wx[c], v[k], hg[k]);
    }
    if (y0 == r) {
      const float w0 = 1.f - ly;
      if (y1 != y0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc0[k] = fmaf(w0, hg[k], acc0[k]); acc1[k] = fmaf(ly, hg[k], acc1[k]); }
      } else {                                    // last input row: both neighbours are this row (ly is 0 there)
        const float w = w0 + ly;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc0[k] = fmaf(w, hg[k], acc0[k]);
      }
    } else if (y1 == r && y0 == r - 1) {          // first rows of the strip: only the lower neighbour is ours
#pragma unroll
      for (int k = 0; k < 8; ++k) acc0[k] = fmaf(ly, hg[k], acc0[k]);
    }
  };
  // two statically named row buffers: the loads of row oh + 1 are in flight while row oh is consumed
  Raw bufA[NC], bufB[NC];
  load_row(bufA, oh_first);
  for (int oh = oh_first; oh <= oh_last; oh += 2) {
    load_row(bufB, oh + 1);
    consume(bufA, oh);
    load_row(bufA, oh + 2);
    if (oh + 1 <= oh_last) consume(bufB, oh + 1);
  }
  while (r < ih_e) finish_row();                  // the last one or two rows of the strip (zeros beyond: every row is written)
}

template <typename T>
__device__ __forceinline__ void bilinear_bwd_lean_body(const DView& gy, const DView& gx, const BilinearTables& t, float sy,
                                                       const FastDiv& divG, int rows_per_strip, int accumulate) {
  const unsigned G = gx.C / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)gx.W * G) return;
  unsigned iw, g;
  divG.divmod(idx, iw, g);
  const int b = blockIdx.z;
  const int ih_b = blockIdx.y * rows_per_strip;
  const int ih_e = min(gx.H, ih_b + rows_per_strip);
  const int nc = t.tx_off[iw + 1] - t.tx_off[iw];
  // the widest lane of the warp picks the unrolling (lanes past the row end have already returned)
  const unsigned lanes = __activemask();
  if (__any_sync(lanes, nc > 4)) bilinear_bwd_lean_rows<T, 6>(gy, gx, t, sy, b, (int)iw, (int)g, ih_b, ih_e, accumulate);
  else if (__any_sync(lanes, nc > 2)) bilinear_bwd_lean_rows<T, 4>(gy, gx, t, sy, b, (int)iw, (int)g, ih_b, ih_e, accumulate);
  else bilinear_bwd_lean_rows<T, 2>(gy, gx, t, sy, b, (int)iw, (int)g, ih_b, ih_e, accumulate);
}
// Three resident blocks per SM (80 registers): the 2- and 4-contribution paths fit, the rare 6-contribution path (the two
// or three warps of a row whose source column is fed by five output columns) spills 80 bytes.  Against the 127-register
// build (two resident blocks): 84.0 -> 75.5 us at 16 x 250^2 x 128, 49.8 -> 41.1 us at 16 x 124^2 x 256 -- with the issue
// slots freed by the lean loop the kernel was latency-bound (ncu: 52 % issue, 23 % warps active, 45 % DRAM).
template <typename T>
__global__ void __launch_bounds__(256, 3) bilinear_bwd_lean_kernel(DView gy, DView gx, BilinearTables t, float sy, FastDiv divG,
                                                                  int rows_per_strip, int accumulate) {
  bilinear_bwd_lean_body<T>(gy, gx, t, sy, divG, rows_per_strip, accumulate);
}
