// oracle/bilinear_vh_emu.cpp -- TEST INFRASTRUCTURE ONLY: the product's rows-first bilinear backward kernel
// (csrc/bilinear_vh.cuh, fp32 instantiation) and its host-side tiling (csrc/bilinear_tables.h) compiled for the CPU on
// top of oracle/cuda_emu.h, so that tests/test_oracle.py can check the kernel's own indexing, barriers and strip / tile
// edges against torch autograd without a GPU.  The 8-channel vector type is replaced by a plain fp32 stand-in; the bf16
// pack / unpack helpers of csrc/vec.cuh are not exercised here (they are shared with every other kernel and covered on
// the device).  g++ -std=c++20 -O1 -shared -fPIC -pthread oracle/bilinear_vh_emu.cpp
#include "cuda_emu.h"
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/bilinear_tables.h"

struct float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline bool __any_sync(unsigned, bool) { return true; }   // conservative: every warp takes the 6-term path (absent terms weigh 0)
using std::min;

namespace mau {
struct DView {
  void* ptr;
  int B, H, W, cs, c0, C;
};
template <typename T> struct V8;
template <> struct V8<float> {
  struct Raw {
    float f[8];
  };
  static Raw load_raw(const float* p) {
    Raw r;
    memcpy(r.f, p, sizeof(r.f));
    return r;
  }
  static void unpack(const Raw& r, float (&f)[8]) { memcpy(f, r.f, sizeof(r.f)); }
  static void load(const float* p, float (&f)[8]) { memcpy(f, p, sizeof(f)); }
  static void store(float* p, const float (&f)[8]) { memcpy(p, f, sizeof(f)); }
};
namespace {
constexpr int kMaxE = kBilinearMaxFan;
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/bilinear_index.cuh"
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/bilinear_vh.cuh"
}  // namespace
}  // namespace mau

// gy [B, Hout, Wout, cs] (channels c0 .. c0 + C of every pixel), gx [B, Hin, Win, C]; returns the tile width used (0: the
// shape is not served by this kernel)
extern "C" int emu_bilinear_bwd_vh(const float* gy, int B, int Hin, int Win, int C, int Hout, int Wout, int gy_cs, int gy_c0,
                                   float* gx, int accumulate, int strip) {
  using namespace mau;
  BilinearHost hy, hx;
  bilinear_axis_tables(Hin, Hout, &hy);
  bilinear_axis_tables(Win, Wout, &hx);
  BilinearTables t;
  t.Hin = Hin; t.Win = Win; t.Hout = Hout; t.Wout = Wout;
  t.max_fan_w = hx.max_fan;
  t.vh_tile = bilinear_vh_tile(hx);
  if (t.vh_tile <= 0 || Hin > Hout || Hin < 2 || C % 8) return 0;
  t.ty_off = hy.t_off.data(); t.ty_idx = hy.t_idx.data(); t.ty_w = hy.t_w.data();
  t.tx_off = hx.t_off.data(); t.tx_idx = hx.t_idx.data(); t.tx_w = hx.t_w.data();
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const int G = C / 8;
  int cg_shift = 0;
  while (cg_shift < 3 && G % (2 << cg_shift) == 0) ++cg_shift;
  const int tiles = (Win + t.vh_tile - 1) / t.vh_tile, chunks = G >> cg_shift;
  const dim3 grid((unsigned)(tiles * chunks), (unsigned)((Hin + strip - 1) / strip), (unsigned)B);
  DView vgy{const_cast<float*>(gy), B, Hout, Wout, gy_cs, gy_c0, C}, vgx{gx, B, Hin, Win, C, 0, C};
  emu_launch(bilinear_bwd_vh_kernel<float>, grid, dim3(32u << cg_shift), vgy, vgx, t, sy, strip, t.vh_tile, cg_shift, accumulate);
  return t.vh_tile;
}
