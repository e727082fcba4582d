// wgrad_tc.cu -- weight gradient of the 3x3 convolution on the sm_100a tensor cores.
//
// Replaces: the dL/dW that autograd derives for nn.Conv2d(cin, cout, 3, padding=1)
// (reference src/model.py:12,14; backward triggered at src/train.py:252).
//
//   dW[co][ci][r][s] = sum_{b,h,w} dY[b,h,w,co] * X[b,h+r-1,w+s-1,ci]
//   GEMM view per filter tap: M = 128 output channels, N = 64 input channels, K = pixels.
//
// Both operands are NHWC bf16, i.e. the reduction dimension (pixels) is the *slow* one: the TMA
// boxes {64 ch, 8 w, 8 h} land in SMEM as [64 pixel rows][128 B of channels] with 128-byte swizzle,
// which is exactly the MN-major canonical UMMA layout (8-row K groups at SBO = 1024 B, 64-channel
// MN groups at LBO = one box).  The X box of tap (r,s) is the dY box shifted by (r-1, s-1); the TMA
// unit's out-of-bounds zero fill is the convolution padding.  One CTA owns one filter row r (three
// fp32 accumulators of 128x64 in TMEM, 192 columns), one (co, ci) tile and one slice of the pixel
// range (split-K); partial results are combined with fp32 atomics straight into the OIHW gradient.
#include "conv_tc.h"
#include "ptx.cuh"
#include "tma.h"

namespace mau {
using namespace ptx;

namespace {

constexpr int kThreads = 256;
constexpr int kDyBox = 8192;      // {64 co, 8, 8} bf16
constexpr int kHaloBox = 10240;   // {64 ci, 10, 8} bf16: the 8 input rows one filter row needs, with a 1-pixel W halo

struct WgradParams {
  int B, H, W;
  int tiles_w, tiles_h;       // 8x8 pixel tiles per image
  int Cout, Cin;              // extents of the dY view / X segment view
  int ci_w0, Cin_w;           // placement inside the OIHW weight tensor
  int tiles_per_split, total_tiles;
};

// BN = input channels per CTA (64 or 128); one CTA = 128 output channels x BN input channels x one filter row.
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads) wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmDy,
                                                               const __grid_constant__ CUtensorMap tmX,
                                                               const WgradParams p, float* __restrict__ dw) {
  constexpr int kStageBytes = 2 * kDyBox + (BN / 64) * kHaloBox;
  constexpr int kTmemCols = BN == 64 ? 256 : 512;        // 3 accumulators of BN columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = blockIdx.x * BN;
  const int co0 = blockIdx.y * 128;
  const int r = blockIdx.z % 3;
  const int split = blockIdx.z / 3;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.total_tiles, t_begin + p.tiles_per_split);
  const int tiles_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmDy);
    prefetch_tensormap(&tmX);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer: whole warp converged, one elected lane issues
    int stage = 0; uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / tiles_img;
      const int rem = t - b * tiles_img;
      const int h0 = (rem / p.tiles_w) * 8, w0 = (rem % p.tiles_w) * 8;
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[stage], kStageBytes);
        uint8_t* s = smem + stage * kStageBytes;
        tma_load_4d(s, &tmDy, &full[stage], co0, w0, h0, b);
        tma_load_4d(s + kDyBox, &tmDy, &full[stage], co0 + 64, w0, h0, b);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(s + 2 * kDyBox + j * kHaloBox, &tmX, &full[stage], ci0 + 64 * j, w0 - 1, h0 + r - 1, b);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp converged, one elected lane issues.  Both operands are MN-major:
    //   A = dY^T: 64-channel groups at LBO = one dY box, 8-pixel K groups at SBO = 1024 B
    //   B = X window of tap s: start shifted by s pixels, 8-pixel K groups at SBO = 10 * 128 B (halo row pitch)
    constexpr uint32_t idesc = idesc_bf16_f32(128, BN, /*a MN-major*/ 1, /*b MN-major*/ 1);
    const uint64_t descA0 = smem_desc_sw128(0, kDyBox, 1024, 0);
    const uint64_t descB0 = smem_desc_sw128(0, kHaloBox, 1280, 0);
    int stage = 0; uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * kStageBytes);
      const uint64_t da = descA0 + (uint64_t)(sa >> 4);
      const uint64_t db = descB0 + (uint64_t)((sa + 2 * kDyBox) >> 4);
      const uint32_t first = t > t_begin ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // 16 pixels = two tile rows per instruction
            umma_bf16(tmem_base + sx * BN, da + (uint64_t)((kk * 2048) >> 4),
                      db + (uint64_t)((kk * 2560 + sx * 128) >> 4), idesc, kk ? 1u : first);
        }
        umma_commit(&empty[stage]);
        if (t == t_end - 1) umma_commit(tmem_full);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4) {
    if (t_end > t_begin) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
#pragma unroll 1
      for (int sx = 0; sx < 3; ++sx) {
#pragma unroll 1
        for (int part = 0; part < BN / 32; ++part) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + sx * BN + part * 32, v);
          tmem_ld_wait();
          if (co < p.Cout) {
            float* row = dw + ((long long)co * p.Cin_w + p.ci_w0 + ci0 + part * 32) * 9 + r * 3 + sx;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ci0 + part * 32 + j < p.Cin) atomicAdd(row + j * 9, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES>
int launch_wgrad(const WgradTcOp& op, const WgradParams& p, float* dw, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)STAGES * (2 * kDyBox + (BN / 64) * kHaloBox) + 8 * (2 * STAGES + 1) + 16;
  static_assert(smem <= 232448, "shared memory budget exceeded");
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = wgrad3x3_tc_kernel<BN, STAGES>;
  if (!attr_done[dev & 15]) {
    MAU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev & 15] = true;
  }
  kern<<<op.grid, kThreads, smem, st>>>(op.tmDy, op.tmX, p, dw);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace

int wgrad_tc_prepare(WgradTcOp* op, const View& x_seg, const View& dy, int ci_w0, int Cin_w) {
  if (x_seg.B != dy.B || x_seg.H != dy.H || x_seg.W != dy.W) return fail("wgrad_tc: geometry mismatch");
  if (x_seg.cs % 8 || x_seg.c0 % 8 || dy.cs % 8 || dy.c0 % 8) return fail("wgrad_tc: views must be 8-channel aligned");
  op->B = dy.B; op->H = dy.H; op->W = dy.W;
  op->Cout = dy.C; op->Cin = x_seg.C; op->ci_w0 = ci_w0; op->Cin_w = Cin_w;
  op->bn = x_seg.C <= 64 ? 64 : 128;
  op->n_tiles = ceil_div(x_seg.C, op->bn);
  op->m_tiles = ceil_div(dy.C, 128);
  const int total_tiles = dy.B * ceil_div(dy.H, 8) * ceil_div(dy.W, 8);
  int want = ceil_div(148 * 2, op->n_tiles * op->m_tiles * 3);
  if (want < 1) want = 1;
  if (want > total_tiles) want = total_tiles;
  op->tiles_per_split = ceil_div(total_tiles, want);
  op->splits = ceil_div(total_tiles, op->tiles_per_split);
  op->grid = dim3((unsigned)op->n_tiles, (unsigned)op->m_tiles, (unsigned)(3 * op->splits));
  MAU_TRY(make_nhwc_map(&op->tmDy, DT_BF16, dy, 64, 8, 8));
  MAU_TRY(make_nhwc_map(&op->tmX, DT_BF16, x_seg, 64, 10, 8));
  return 0;
}

int wgrad_tc_launch(const WgradTcOp& op, float* dw_oihw, cudaStream_t st) {
  WgradParams p;
  p.B = op.B; p.H = op.H; p.W = op.W;
  p.tiles_w = ceil_div(op.W, 8); p.tiles_h = ceil_div(op.H, 8);
  p.Cout = op.Cout; p.Cin = op.Cin; p.ci_w0 = op.ci_w0; p.Cin_w = op.Cin_w;
  p.tiles_per_split = op.tiles_per_split;
  p.total_tiles = op.B * p.tiles_w * p.tiles_h;
  if (op.bn == 64) return launch_wgrad<64, 6>(op, p, dw_oihw, st);
  return launch_wgrad<128, 5>(op, p, dw_oihw, st);
}

}  // namespace mau
