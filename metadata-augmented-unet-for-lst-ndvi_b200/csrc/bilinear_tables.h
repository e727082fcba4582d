// bilinear_tables.h -- index / weight tables of the align_corners bilinear resize (reference src/model.py:12-17 and
// 121 / 219 / 245: F.interpolate(..., mode="bilinear", align_corners=True)).
// Plain C++ (no CUDA headers): shared by elementwise.cu, plan.cu, api.cu and by oracle/bilinear_bwd_emu.cpp, which runs
// bilinear_bwd_lean.cuh on the CPU with exactly these tables.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

namespace mau {

constexpr int kBilinearMaxFan = 6;    // contributions per source index the batched backward kernels handle

struct BilinearTables {          // device arrays owned by the plan
  int Hin = 0, Win = 0, Hout = 0, Wout = 0;
  int* y0 = nullptr; int* y1 = nullptr; float* ly = nullptr;   // [Hout]
  int* x0 = nullptr; int* x1 = nullptr; float* lx = nullptr;   // [Wout]
  // transposed (gather) form for the backward: CSR over input rows / cols
  int* ty_off = nullptr; int* ty_idx = nullptr; float* ty_w = nullptr;
  int* tx_off = nullptr; int* tx_idx = nullptr; float* tx_w = nullptr;
  int max_fan_w = 1 << 30;     // most contributions any source column receives (selects the batched backward)
  bool no_stream = false;      // tests: skip the streaming backward kernel (exercises the per-pixel gather kernels)
};
struct BilinearHost {            // host mirror used to build the tables
  std::vector<int> i0, i1; std::vector<float> l;
  std::vector<int> t_off, t_idx; std::vector<float> t_w;
  int max_fan = 0;
};

// area_pixel_compute_scale / compute_source_index_and_lambda of ATen for align_corners=True, in fp32
inline void bilinear_axis_tables(int in, int out, BilinearHost* h) {
  h->i0.resize(out); h->i1.resize(out); h->l.resize(out);
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  std::vector<std::vector<std::pair<int, float>>> inv(in);
  for (int o = 0; o < out; ++o) {
    const float real = scale * (float)o;
    int i0 = (int)real;
    if (i0 > in - 1) i0 = in - 1;
    const int i1 = i0 + (i0 < in - 1 ? 1 : 0);
    float l1 = real - (float)i0;
    l1 = l1 < 0.f ? 0.f : (l1 > 1.f ? 1.f : l1);
    h->i0[o] = i0; h->i1[o] = i1; h->l[o] = l1;
    inv[i0].push_back({o, 1.f - l1});
    inv[i1].push_back({o, l1});
  }
  h->t_off.assign(in + 1, 0);
  h->t_idx.clear(); h->t_w.clear();
  h->max_fan = 0;
  for (int i = 0; i < in; ++i) {
    h->max_fan = std::max(h->max_fan, (int)inv[i].size());
    for (auto& e : inv[i]) { h->t_idx.push_back(e.first); h->t_w.push_back(e.second); }
    h->t_off[i + 1] = (int)h->t_idx.size();
  }
}

}  // namespace mau
