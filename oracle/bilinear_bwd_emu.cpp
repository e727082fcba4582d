// oracle/bilinear_bwd_emu.cpp -- TEST INFRASTRUCTURE ONLY: the product's streaming bilinear backward kernel
// (csrc/bilinear_bwd_lean.cuh, fp32 instantiation) with the tables of csrc/bilinear_tables.h, compiled for the CPU on top
// of oracle/cuda_emu.h, so that tests/test_oracle.py can check the kernel's own indexing -- strip edges, the sliding row
// window, the 2 / 4 / 6-contribution variants, rows loaded past the strip -- against torch autograd without a GPU.  The
// 8-channel vector type is replaced by a plain fp32 stand-in; the bf16 pack / unpack helpers of csrc/vec.cuh are not
// exercised here (every other kernel shares them; covered on the device).  A warp vote is answered per thread (each
// emulated thread then takes the narrowest variant that fits ITS column, a refinement of what a warp does).
// g++ -std=c++20 -O1 -shared -fPIC -pthread oracle/bilinear_bwd_emu.cpp
#include "cuda_emu.h"
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/bilinear_tables.h"

inline float __fsub_rn(float a, float b) { return a - b; }
inline unsigned __activemask() { return 0xffffffffu; }
inline bool __any_sync(unsigned, bool p) { return p; }
using std::min;

namespace mau {
struct DView {
  void* ptr;
  int B, H, W, cs, c0, C;
};
struct FastDiv {
  unsigned d = 1;
  explicit FastDiv(unsigned div) : d(div ? div : 1) {}
  void divmod(unsigned n, unsigned& q, unsigned& r) const { q = n / d; r = n - q * d; }
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
#include "../metadata-augmented-unet-for-lst-ndvi_b200/csrc/bilinear_bwd_lean.cuh"
}  // namespace
}  // namespace mau

// gy [B, Hout, Wout, cs] (channels c0 .. c0 + C of every pixel), gx [B, Hin, Win, C]; returns the largest number of
// contributions a source column receives (0: the shape is not served by this kernel -- the same test as op_bilinear_bwd)
extern "C" int emu_bilinear_bwd_lean(const float* gy, int B, int Hin, int Win, int C, int Hout, int Wout, int gy_cs, int gy_c0,
                                     float* gx, int accumulate, int strip) {
  using namespace mau;
  BilinearHost hy, hx;
  bilinear_axis_tables(Hin, Hout, &hy);
  bilinear_axis_tables(Win, Wout, &hx);
  if (hx.max_fan > kMaxE || Hin > Hout || Hin < 2 || C % 8) return 0;
  BilinearTables t;
  t.Hin = Hin; t.Win = Win; t.Hout = Hout; t.Wout = Wout;
  t.max_fan_w = hx.max_fan;
  t.ty_off = hy.t_off.data(); t.ty_idx = hy.t_idx.data(); t.ty_w = hy.t_w.data();
  t.tx_off = hx.t_off.data(); t.tx_idx = hx.t_idx.data(); t.tx_w = hx.t_w.data();
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const int G = C / 8;
  const dim3 grid((unsigned)((Win * G + 255) / 256), (unsigned)((Hin + strip - 1) / strip), (unsigned)B);
  DView vgy{const_cast<float*>(gy), B, Hout, Wout, gy_cs, gy_c0, C}, vgx{gx, B, Hin, Win, C, 0, C};
  emu_launch(bilinear_bwd_lean_kernel<float>, grid, dim3(256), vgy, vgx, t, sy, FastDiv((unsigned)G), strip, accumulate);
  return hx.max_fan;
}
