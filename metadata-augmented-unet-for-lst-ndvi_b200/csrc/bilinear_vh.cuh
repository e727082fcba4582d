// bilinear_vh.cuh -- backward of the align_corners bilinear up-sampling, rows first and columns second.
// Included by elementwise.cu inside namespace mau::<anonymous> after V8 / DView / BilinearTables / src_index / kMaxE
// (and by oracle/bilinear_vh_emu.cpp, which supplies CPU stand-ins for those, to check the indexing without a GPU).
//
//   gx[ih, iw] = sum_oh wy(ih, oh) * sum_ow wx(iw, ow) * gy[oh, ow]      (reference src/model.py:12-17, F.interpolate backward)
//
// The streaming kernel it replaces gave one thread an INPUT column: every gy vector was fetched, unpacked and weighted by
// the two input columns it feeds (~300 instructions per thread and output row, 118 registers, 63 % issue-slot
// utilisation at 51 % of HBM peak).  Here one thread owns an OUTPUT column and an 8-channel group and walks down the
// output rows of its strip: each gy vector is loaded exactly once (fully coalesced, 128 bytes per pixel and CTA),
// unpacked once and folded into the two sliding row accumulators with its two vertical weights.  When an input row is
// complete the CTA's 32 output columns meet in shared memory and one thread per (input column, group) forms the
// <= kMaxE-term horizontal sum and stores it.  Buffers alternate, so one barrier per finished input row is enough.
constexpr int kVhCols = kBilinearVhCols;   // output columns per CTA (the input-column tile is chosen on the host to fit)

template <typename T>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 4 : 2) bilinear_bwd_vh_kernel(DView gy, DView gx, BilinearTables t, float sy,
                                                              int rows_per_strip, int iw_tile, int cg_shift,
                                                              int accumulate) {
  using Raw = typename V8<T>::Raw;
  __shared__ float4 sv[2][2][kVhCols * 8];       // [buffer][channels 0-3 | 4-7][column * cg + group]
  const int tid = threadIdx.x;
  const int cg = 1 << cg_shift;                  // channel groups per CTA (blockDim.x == 32 * cg)
  const int chunks = (gx.C / 8) >> cg_shift;
  const int tile = blockIdx.x / chunks, chunk = blockIdx.x - tile * chunks;
  const int iw_a = tile * iw_tile, iw_b = min(gx.W, iw_a + iw_tile);
  const int oc_a = t.tx_idx[t.tx_off[iw_a]];     // output columns feeding this tile: [oc_a, oc_b], at most kVhCols
  const int oc_b = t.tx_idx[t.tx_off[iw_b] - 1];
  const int b = blockIdx.z;
  const int ih_b = blockIdx.y * rows_per_strip, ih_e = min(gx.H, ih_b + rows_per_strip);
  const int g = tid & (cg - 1);
  const int ch = ((chunk << cg_shift) + g) * 8;
  // role 1: producer of output column ow
  const int ow = oc_a + (tid >> cg_shift);
  const bool live = ow <= oc_b;
  const T* gyp = static_cast<const T*>(gy.ptr) + ((long long)b * gy.H * gy.W + (live ? ow : oc_a)) * gy.cs + gy.c0 + ch;
  const int gyrow = gy.W * gy.cs, gxrow = gx.W * gx.cs;   // one image is < 2^31 elements (checked on the host)
  // role 2: owner of input column iw
  const int iw = iw_a + (tid >> cg_shift);
  const bool owner = iw < iw_b;
  int soff[kMaxE];
  float wx[kMaxE];
  int nc = 0;
  {
    const int ca = owner ? t.tx_off[iw] : 0;
    nc = owner ? t.tx_off[iw + 1] - ca : 0;
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) {            // absent entries repeat the first one with weight 0: no predicates below
      const int idx = e < nc ? t.tx_idx[ca + e] : (nc > 0 ? t.tx_idx[ca] : oc_a);
      soff[e] = ((idx - oc_a) << cg_shift) + g;
      wx[e] = e < nc ? t.tx_w[ca + e] : 0.f;
    }
  }
  const bool wide = __any_sync(0xffffffffu, nc > 4);
  T* gxp = static_cast<T*>(gx.ptr) + ((long long)b * gx.H * gx.W + (owner ? iw : iw_a)) * gx.cs + gx.c0 + ch;
  // output rows touching input rows [ih_b, ih_e): the row lists are sorted by output row
  const int oh_first = t.ty_idx[t.ty_off[ih_b]];
  const int oh_last = t.ty_idx[t.ty_off[ih_e] - 1];
  float acc0[8], acc1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
  int r = ih_b;                                  // input row acc0 belongs to (acc1: r + 1)
  int flip = 0;
  auto finish_row = [&]() {                      // input row r is complete: columns meet, owners store, window slides
    sv[flip][0][tid] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
    sv[flip][1][tid] = make_float4(acc0[4], acc0[5], acc0[6], acc0[7]);
    __syncthreads();
    if (owner) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
      auto term = [&](int e) {
        const float4 lo = sv[flip][0][soff[e]], hi = sv[flip][1][soff[e]];
        const float w = wx[e];
        o[0] = fmaf(w, lo.x, o[0]); o[1] = fmaf(w, lo.y, o[1]); o[2] = fmaf(w, lo.z, o[2]); o[3] = fmaf(w, lo.w, o[3]);
        o[4] = fmaf(w, hi.x, o[4]); o[5] = fmaf(w, hi.y, o[5]); o[6] = fmaf(w, hi.z, o[6]); o[7] = fmaf(w, hi.w, o[7]);
      };
      term(0); term(1); term(2); term(3);
      if (wide) { term(4); term(5); }
      T* dst = gxp + r * gxrow;
      if (accumulate) {
        float old[8];
        V8<T>::load(dst, old);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += old[k];
      }
      V8<T>::store(dst, o);
    }
    flip ^= 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc0[k] = acc1[k]; acc1[k] = 0.f; }
    ++r;
  };
  auto consume = [&](const Raw& raw, int oh) {
    int y0, y1; float ly;
    src_index(sy, oh, gx.H, y0, y1, ly);         // block-uniform
    while (y0 > r && r < ih_e) finish_row();
    float v[8];
    V8<T>::unpack(raw, v);
    const float w0 = 1.f - ly;
    if (y0 == r) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc0[k] = fmaf(w0, v[k], acc0[k]);
      if (y1 != y0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc1[k] = fmaf(ly, v[k], acc1[k]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc0[k] = fmaf(ly, v[k], acc0[k]);
      }
    } else if (y1 == r && y0 == r - 1) {         // first rows of the strip: only the lower neighbour is ours
#pragma unroll
      for (int k = 0; k < 8; ++k) acc0[k] = fmaf(ly, v[k], acc0[k]);
    }
  };
  // two statically named pairs of row buffers: rows oh+2, oh+3 are in flight while rows oh, oh+1 are consumed
  Raw a0, a1, b0, b1;
  a0 = V8<T>::load_raw(gyp + oh_first * gyrow);
  a1 = V8<T>::load_raw(gyp + min(oh_first + 1, oh_last) * gyrow);
  for (int oh = oh_first; oh <= oh_last; oh += 4) {
    b0 = V8<T>::load_raw(gyp + min(oh + 2, oh_last) * gyrow);
    b1 = V8<T>::load_raw(gyp + min(oh + 3, oh_last) * gyrow);
    consume(a0, oh);
    if (oh + 1 <= oh_last) consume(a1, oh + 1);
    a0 = V8<T>::load_raw(gyp + min(oh + 4, oh_last) * gyrow);
    a1 = V8<T>::load_raw(gyp + min(oh + 5, oh_last) * gyrow);
    if (oh + 2 <= oh_last) consume(b0, oh + 2);
    if (oh + 3 <= oh_last) consume(b1, oh + 3);
  }
  while (r < ih_e) finish_row();                 // the last one or two rows of the strip (zeros beyond: every row is written)
}
