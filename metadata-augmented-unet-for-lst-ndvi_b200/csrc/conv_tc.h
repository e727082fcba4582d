// conv_tc.h -- host interface of the tcgen05 implicit-GEMM 3x3 convolution (conv_tc.cu) and the
// tcgen05 weight-gradient kernel (wgrad_tc.cu)
#pragma once
#include <cuda.h>
#include "common.h"

namespace mau {

enum ConvMode : int { MODE_TAP = 0, MODE_ROW3 = 1, MODE_HALO = 2 };

struct ConvTcParams {
  int tiles_w = 0, tiles_h = 0;  // spatial tiles per image
  int TW = 16, TH = 8;           // tile = TH x TW = 128 output pixels
  int total_ptiles = 0;          // B * tiles_w * tiles_h
  int Cout = 0;                  // valid output channels (N extent of the view)
  int nseg = 0;                  // channel segments of the input buffer feeding K
  int seg_start[4] = {0, 0, 0, 0};
  int seg_chunks[4] = {0, 0, 0, 0};  // 64-channel chunks per segment
  int relu = 0;
  int accumulate = 0;            // TMA reduce-add instead of store
  int halo_base_offset = 0;      // MODE_HALO: fill the descriptor base_offset field from the address
  const float* scale = nullptr;  // per output channel (nullptr -> 1)
  const float* shift = nullptr;  // per output channel (nullptr -> 0)
  // optional fused 1x1 head (eval, single output-channel tile): out[b,o,h,w] = act(sum_c y[c]*head_w[o][c] + head_b[o])
  const float* head_w = nullptr; const float* head_b = nullptr; float* head_out = nullptr;
  int head_oc = 0, head_tanh = 0, H = 0, W = 0;
  int store_y = 1;               // 0: skip the activation store (only the fused head output is needed)
  int bt = 0, bt_col0 = 0;       // data gradient: B read MN-major from the forward weight pack, first packed column
  // training forward: per-channel sum / sum of squares of the STORED (bf16-rounded) output, accumulated by the two
  // otherwise idle warps from the staging tile while the epilogue warps convert the next one: double[2 * Cout], zeroed
  // by the caller (replaces the separate bn_stats pass over z)
  double* stats = nullptr;
};

struct ConvTcOp {
  CUtensorMap tmA, tmB, tmY;
  ConvTcParams p;
  dim3 grid;
  int bn = 0;
  int mode = 0;
  int mt = 1;      // v2: 128-pixel sub-tiles per work item (share each B tile)
  int nbuf = 2;    // v2: TMEM accumulator buffers
  int col3 = 0;    // v3: Cout == 64, the three horizontal taps folded into N = 192 (conv3x3_tc_col3_kernel)
  int bres = 0;    // v2: weights resident in shared memory for the whole launch (K = 64, one 64-wide N tile)
};

int conv_tc_pick_bn(int Cout);
// xbuf: the *whole* input buffer (c0 = 0, C = valid channels); segments select the K channels.
// wpacked: bf16 [9][n_rows][Kp]; y: output view (its C is the GEMM N extent).
int conv_tc_prepare(ConvTcOp* op, const View& xbuf, int nseg, const int* seg_start, const int* seg_len,
                    const void* wpacked, int Kp, int n_rows, const View& y, int mode, const float* scale,
                    const float* shift, int relu, int accumulate, int halo_base_offset);
// data gradient w.r.t. one input segment without a transposed weight pack (halo kernel only)
int conv_tc_prepare_dgrad(ConvTcOp* op, const View& dz, const void* wpack_fwd, int Kp_fwd, int col0, const View& gx,
                          int accumulate);
int conv_tc_launch(const ConvTcOp& op, cudaStream_t st);
int conv_tc_pack_fwd(const float* w_oihw, int Cout, int Cin, const int* kmap_dev, int Kp, void* out,
                     cudaStream_t st);
int conv_tc_pack_dgrad(const float* w_oihw, int Cout, int Cin, int ci0, int N, int Kp, void* out,
                       cudaStream_t st);

// ---- weight gradient (wgrad_tc.cu): dW[co][ci][tap] = sum_pixels dY[p][co] * X[p + d(tap)][ci]
struct WgradTcOp {
  CUtensorMap tmDy, tmX, tmWs;
  int B = 0, H = 0, W = 0;
  int Cout = 0, Cin = 0;         // extents of this launch (Cin = one input segment)
  int ci_w0 = 0;                 // first weight input-channel index of this segment
  int Cin_w = 0;                 // Cin of the full OIHW weight tensor
  int m_tiles = 0, n_tiles = 0, splits = 0, tiles_per_split = 0;
  int bn = 0;
  float* ws = nullptr;           // v2: fp32 workspace [9][M dim][N dim] (nullptr selects the v1 atomics kernel)
  int swap = 0;                  // v2: 1 = M side is the input-channel side (workspace [9][Cin][Cout]); 2 / 3 = tap-pair kernel
                                 // with dY / X as the shifted 64-channel operand (workspace layouts of 0 / 1)
  dim3 grid;
};
// orientation that wastes less of the 128 x BN tiles for this layer (all segments share one workspace layout);
// 2 / 3 select the tap-pair kernel when the output / input side has at most 64 channels
int wgrad_tc_pick_swap(int Cout, int nseg, const int* seg_len);
size_t wgrad_tc_workspace_floats(int Cout, int Cin_w, int swap);
// ws != nullptr: v2 (persistent split-K/stream-K kernel, TMA reduce-add into ws, then wgrad_tc_finalize);
// ws == nullptr: v1 (fp32 atomics straight into the zeroed OIHW gradient)
int wgrad_tc_prepare(WgradTcOp* op, const View& x_seg, const View& dy, int ci_w0, int Cin_w, float* ws, int swap);
int wgrad_tc_launch(const WgradTcOp& op, float* dw_oihw, cudaStream_t st);
// ws [9][..][..] (zeroed before the segment launches) -> dw OIHW (overwrites)
int wgrad_tc_finalize(const float* ws, int swap, int Cout, int Cin_w, float* dw_oihw, cudaStream_t st);

}  // namespace mau
