// bilinear_index.cuh -- source index and lambda of one output coordinate, recomputed with the same fp32 operations ATen
// uses (scale * dst, truncation), so that no table load sits in front of the data loads.  Included inside namespace
// mau::<anonymous> by elementwise.cu (and by oracle/bilinear_bwd_emu.cpp).
__device__ __forceinline__ void src_index(float scale, int o, int in_size, int& i0, int& i1, float& l1) {
  const float real = __fmul_rn(scale, (float)o);
  i0 = min((int)real, in_size - 1);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
}
