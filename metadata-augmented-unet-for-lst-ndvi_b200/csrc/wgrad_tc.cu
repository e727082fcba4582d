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
#include <algorithm>
#include <cstdlib>

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


// ==========================================================================================
// v2: persistent split-K / stream-K weight gradient with a TMA reduce-add epilogue.
//
//   pixel tile  = 16 wide x 4 high (64 pixels = one pipeline stage, K = 64): dY boxes {64 ch, 16, 4},
//                 X halo boxes {64 ch, 18, 4} at (w0-1, h0+r-1).  One tcgen05.mma (K = 16) consumes one
//                 tile row: the two 8-pixel K groups are 1024 B apart in both boxes; the horizontal tap s
//                 is a start offset of s pixels into the halo row.
//   work unit   = (128-wide M tile, BN-wide N tile, filter row r): three fp32 accumulators (one per
//                 horizontal tap) of 128 x BN in TMEM.  M is the output-channel side and N the input-
//                 channel side, or the other way round (SWAP) when that wastes less of the 128-row tile
//                 (Cout = 64 layers): both operands are MN-major straight out of NHWC, so the roles are
//                 interchangeable.
//   scheduling  = the (unit, pixel-tile) space is cut into equal contiguous ranges, one per CTA.  With
//                 few units every unit is split S ways (classic split-K: the CTAs of one split walk the
//                 same pixels for different units at the same time, so dY / X are shared through L2);
//                 with many units (small images, wide layers -- the activations fit L2) the ranges are
//                 stream-K slices of the unit-major list over all SMs, a CTA may finish one unit and
//                 start the next.  No wave quantisation either way.
//   epilogue    = TMEM -> registers -> 128-byte-swizzled fp32 staging tile (128 x 32) -> TMA reduce-add
//                 into a zeroed fp32 workspace [9][M dim][N dim]; wgrad_finalize_kernel transposes the
//                 workspace into the caller's OIHW gradient.
// ==========================================================================================
constexpr int kDyBox2 = 8192;     // {64, 16, 4} bf16
constexpr int kXBox2 = 9216;      // {64, 18, 4} bf16
constexpr int kStgTile = 16384;   // 128 rows x 32 fp32

struct WgradV2Params {
  int tiles_w, tiles_h;      // 16x4 pixel tiles per image
  int T;                     // pixel tiles in the whole batch
  int m_tiles, n_tiles;      // units = m_tiles * n_tiles * 3
  long long atoms;           // units * T
  int ws_m0, ws_n0;          // workspace coordinates of this segment's first M / N channel
};

template <int BN, bool SWAP, int STAGES>
struct WgradV2Smem {
  static constexpr int NDY = SWAP ? BN / 64 : 2;
  static constexpr int NX = SWAP ? 2 : BN / 64;
  static constexpr int kStage = NDY * kDyBox2 + NX * kXBox2;
  static constexpr size_t kBytes = 1024 + (size_t)STAGES * kStage + 2 * kStgTile + 8 * (2 * STAGES + 2) + 16;
};

struct UnitCoord { int m, n, r; };
__device__ __forceinline__ UnitCoord decode_unit(int u, int n_tiles) {
  UnitCoord c;
  c.r = u % 3;
  const int q = u / 3;
  c.n = q % n_tiles;
  c.m = q / n_tiles;
  return c;
}

template <int BN, bool SWAP, int STAGES>
__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_tc_v2_kernel(const __grid_constant__ CUtensorMap tmDy,
                                                                    const __grid_constant__ CUtensorMap tmX,
                                                                    const __grid_constant__ CUtensorMap tmWs,
                                                                    const WgradV2Params p) {
  using S = WgradV2Smem<BN, SWAP, STAGES>;
  constexpr int kTmemCols = BN == 64 ? 256 : 512;        // 3 accumulators of BN columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStg = smem + STAGES * S::kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 2 * kStgTile);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long a0 = (long long)blockIdx.x * p.atoms / gridDim.x;
  const long long a1 = (long long)(blockIdx.x + 1) * p.atoms / gridDim.x;
  const int tiles_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmDy); prefetch_tensormap(&tmX); prefetch_tensormap(&tmWs);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    int stage = 0; uint32_t phase = 0;
    for (long long a = a0; a < a1;) {
      const int u = (int)(a / p.T);
      const int t0 = (int)(a - (long long)u * p.T);
      const int t1 = (int)min((long long)p.T, (long long)t0 + (a1 - a));
      const UnitCoord uc = decode_unit(u, p.n_tiles);
      const int co0 = SWAP ? uc.n * BN : uc.m * 128;
      const int ci0 = SWAP ? uc.m * 128 : uc.n * BN;
      for (int t = t0; t < t1; ++t) {
        const int b = t / tiles_img;
        const int rem = t - b * tiles_img;
        const int th = rem / p.tiles_w;
        const int h0 = th * 4, w0 = (rem - th * p.tiles_w) * 16;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], S::kStage);
          uint8_t* s = smem + stage * S::kStage;
#pragma unroll
          for (int j = 0; j < S::NDY; ++j) tma_load_4d(s + j * kDyBox2, &tmDy, &full[stage], co0 + 64 * j, w0, h0, b);
#pragma unroll
          for (int j = 0; j < S::NX; ++j)
            tma_load_4d(s + S::NDY * kDyBox2 + j * kXBox2, &tmX, &full[stage], ci0 + 64 * j, w0 - 1, h0 + uc.r - 1, b);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      a += t1 - t0;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // both operands MN-major: 64-channel groups at LBO = one box, 8-pixel K groups at SBO = 1024 B
    constexpr uint32_t idesc = idesc_bf16_f32(128, BN, 1, 1);
    const uint64_t descDy0 = smem_desc_sw128(0, kDyBox2, 1024, 0);
    const uint64_t descX0 = smem_desc_sw128(0, kXBox2, 1024, 0);
    int stage = 0; uint32_t phase = 0, ephase = 0;
    for (long long a = a0; a < a1;) {
      const int u = (int)(a / p.T);
      const int t0 = (int)(a - (long long)u * p.T);
      const int t1 = (int)min((long long)p.T, (long long)t0 + (a1 - a));
      mbar_wait(tmem_empty, ephase ^ 1);       // the epilogue has drained the previous unit's accumulators
      ephase ^= 1;
      tc_fence_after();
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::kStage);
        const uint64_t dDy = descDy0 + (uint64_t)(sa >> 4);
        const uint64_t dX = descX0 + (uint64_t)((sa + S::NDY * kDyBox2) >> 4);
        const uint32_t first = t > t0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int sx = 0; sx < 3; ++sx) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {     // one tile row (16 pixels) per instruction
              const uint64_t ddy = dDy + (uint64_t)((h * 2048) >> 4);
              const uint64_t dx = dX + (uint64_t)(((h * 18 + sx) * 128) >> 4);
              umma_bf16(tmem_base + sx * BN, SWAP ? dx : ddy, SWAP ? ddy : dx, idesc, h ? 1u : first);
            }
          }
          umma_commit(&empty[stage]);
          if (t == t1 - 1) umma_commit(tmem_full);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      a += t1 - t0;
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int et = threadIdx.x - 128;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t fphase = 0;
    int sb = 0;
    for (long long a = a0; a < a1;) {
      const int u = (int)(a / p.T);
      const int t0 = (int)(a - (long long)u * p.T);
      const int t1 = (int)min((long long)p.T, (long long)t0 + (a1 - a));
      const UnitCoord uc = decode_unit(u, p.n_tiles);
      mbar_wait(tmem_full, fphase);
      fphase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int sx = 0; sx < 3; ++sx) {
#pragma unroll 1
        for (int part = 0; part < BN / 32; ++part) {
          uint8_t* stg = sStg + sb * kStgTile;
          const bool last = sx == 2 && part == BN / 32 - 1;
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // staging tile sb is free again
          named_bar_sync(1, 128);
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + sx * BN + part * 32, v);
          tmem_ld_wait();
          uint8_t* rowp = stg + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (last) tc_fence_before();
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (et == 0) {
            if (last) mbar_arrive(tmem_empty);
            tma_reduce_add_3d(&tmWs, stg, p.ws_n0 + uc.n * BN + part * 32, p.ws_m0 + uc.m * 128, uc.r * 3 + sx);
            tma_commit_group();
          }
          sb ^= 1;
        }
      }
      a += t1 - t0;
    }
    if (et == 0) tma_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ==========================================================================================
// v3 ("pair"): layers whose narrow side has <= 64 channels (every convolution at 250x250: Cout = 64; conv1_0.conv1:
// Cin = 64).  The v2 kernel puts that side on a 128-row M tile of which half is padding (452 TFLOP/s on the 64 -> 64
// layers).  Here the narrow side carries the filter-tap shift and TWO taps share one MMA: the M = 128 rows are
// [64 channels of tap u | 64 channels of tap u+1].  Both operands are MN-major, so the second 64-row group of the A
// operand is simply "the same halo box, LBO bytes further": LBO = 128 B (next pixel) for horizontally adjacent taps,
// 2048 B for the (row end, next row start) pair.  Nine taps = five accumulators (the last one half used) instead of
// nine, one work unit covers all three filter rows (one {64, 18, 6} halo box per 16 x 4 pixel tile instead of three
// {64, 18, 4} boxes), and the other operand can be up to 96 channels wide (5 x 96 = 480 TMEM columns).
//   shifted operand S = dY (PAIR_DY: Cout <= 64, window u at ((u/3) * 18 + u % 3) pixels, tap = 8 - u) or
//                       X  (PAIR_X:  Cin  <= 64, same windows, tap = u);  the other operand O is the plain tile box.
//   dW[co][ci][tap] = sum_q dY[q - d(tap)][co] * X[q][ci] = sum_p dY[p][co] * X[p + d(tap)][ci]
// Epilogue: accumulator j holds rows [tap a | tap b]; each half is TMA-reduce-added (box {32, 64}) into the fp32
// workspace [9][narrow side][wide side], which wgrad_finalize_kernel transposes into OIHW.
// ==========================================================================================
constexpr int kPairHalo = 14336;       // {64, 18, 6} bf16 = 13824 B, padded to a multiple of 1024
constexpr int kPairOBox = 8192;        // {64, 16, 4} bf16

struct WgradPairParams {
  int tiles_w, tiles_h, T;   // 16 x 4 pixel tiles
  int n_tiles;               // units (BN-wide tiles of the wide side)
  long long atoms;           // n_tiles * T
  int ws_n0;                 // workspace column of this segment's first wide-side channel
  int ws_m0;                 // workspace row of this segment's first narrow-side channel
  int tap_reversed;          // 1: window u is tap 8 - u (dY carries the shift)
};

template <int BN, int STAGES>
struct WgradPairSmem {
  static constexpr int NO = BN <= 64 ? 1 : 2;
  static constexpr int kStage = kPairHalo + NO * kPairOBox;
  static constexpr size_t kBytes = 1024 + (size_t)STAGES * kStage + 2 * kStgTile + 8 * (2 * STAGES + 2) + 16;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_tc_pair_kernel(const __grid_constant__ CUtensorMap tmS,
                                                                      const __grid_constant__ CUtensorMap tmO,
                                                                      const __grid_constant__ CUtensorMap tmWs,
                                                                      const WgradPairParams p) {
  using S = WgradPairSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStg = smem + STAGES * S::kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 2 * kStgTile);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long a0 = (long long)blockIdx.x * p.atoms / gridDim.x;
  const long long a1 = (long long)(blockIdx.x + 1) * p.atoms / gridDim.x;
  const int tiles_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmS); prefetch_tensormap(&tmO); prefetch_tensormap(&tmWs);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    int stage = 0; uint32_t phase = 0;
    for (long long a = a0; a < a1;) {
      const int u = (int)(a / p.T);
      const int t0 = (int)(a - (long long)u * p.T);
      const int t1 = (int)min((long long)p.T, (long long)t0 + (a1 - a));
      const int n0 = u * BN;
      for (int t = t0; t < t1; ++t) {
        const int b = t / tiles_img;
        const int rem = t - b * tiles_img;
        const int th = rem / p.tiles_w;
        const int h0 = th * 4, w0 = (rem - th * p.tiles_w) * 16;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], 13824 + S::NO * kPairOBox);
          uint8_t* s = smem + stage * S::kStage;
          tma_load_4d(s, &tmS, &full[stage], 0, w0 - 1, h0 - 1, b);
#pragma unroll
          for (int j = 0; j < S::NO; ++j) tma_load_4d(s + kPairHalo + j * kPairOBox, &tmO, &full[stage], n0 + 64 * j, w0, h0, b);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      a += t1 - t0;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = idesc_bf16_f32(128, BN, 1, 1);
    // A = shifted operand: two 64-channel M groups LBO apart (the tap pair), 8-pixel K groups at SBO = 1024 B
    const uint64_t descA_near = smem_desc_sw128(0, 128, 1024, 0);      // pair = horizontally adjacent windows
    const uint64_t descA_wrap = smem_desc_sw128(0, 2048, 1024, 0);     // pair = (row end, next row start): 16 pixels apart
    const uint64_t descB0 = smem_desc_sw128(0, kPairOBox, 1024, 0);
    int stage = 0; uint32_t phase = 0, ephase = 0;
    for (long long a = a0; a < a1;) {
      const int u = (int)(a / p.T);
      const int t0 = (int)(a - (long long)u * p.T);
      const int t1 = (int)min((long long)p.T, (long long)t0 + (a1 - a));
      mbar_wait(tmem_empty, ephase ^ 1);
      ephase ^= 1;
      tc_fence_after();
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::kStage);
        const uint32_t first = t > t0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const int w = 2 * j;                                         // first window of the pair
            const uint32_t woff = ((w / 3) * 18 + (w % 3)) * 128;
            const uint64_t dA = (j == 1 ? descA_wrap : descA_near) + (uint64_t)((sa + woff) >> 4);
            const uint64_t dB = descB0 + (uint64_t)((sa + kPairHalo) >> 4);
#pragma unroll
            for (int h = 0; h < 4; ++h)      // one tile row (16 pixels = K) per instruction
              umma_bf16(tmem_base + j * BN, dA + (uint64_t)((h * 18 * 128) >> 4), dB + (uint64_t)((h * 2048) >> 4), idesc,
                        h ? 1u : first);
          }
          umma_commit(&empty[stage]);
          if (t == t1 - 1) umma_commit(tmem_full);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      a += t1 - t0;
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int et = threadIdx.x - 128;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t fphase = 0;
    int sb = 0;
    for (long long a = a0; a < a1;) {
      const int u = (int)(a / p.T);
      const int t0 = (int)(a - (long long)u * p.T);
      const int t1 = (int)min((long long)p.T, (long long)t0 + (a1 - a));
      mbar_wait(tmem_full, fphase);
      fphase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 5; ++j) {
#pragma unroll 1
        for (int part = 0; part < BN / 32; ++part) {
          uint8_t* stg = sStg + sb * kStgTile;
          const bool last = j == 4 && part == BN / 32 - 1;
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          named_bar_sync(1, 128);
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + j * BN + part * 32, v);
          tmem_ld_wait();
          uint8_t* rowp = stg + row * 128;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<uint4*>(rowp + ((k ^ (row & 7)) << 4)) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          if (last) tc_fence_before();
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (et == 0) {
            if (last) mbar_arrive(tmem_empty);
            const int wa = 2 * j, wb = 2 * j + 1;
            const int col = p.ws_n0 + u * BN + part * 32;
            tma_reduce_add_3d(&tmWs, stg, col, p.ws_m0, p.tap_reversed ? 8 - wa : wa);
            if (wb < 9) tma_reduce_add_3d(&tmWs, stg + 8192, col, p.ws_m0, p.tap_reversed ? 8 - wb : wb);
            tma_commit_group();
          }
          sb ^= 1;
        }
      }
      a += t1 - t0;
    }
    if (et == 0) tma_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, int STAGES>
int launch_wgrad_pair(const WgradTcOp& op, const WgradPairParams& p, cudaStream_t st) {
  using S = WgradPairSmem<BN, STAGES>;
  static_assert(S::kBytes <= 232448, "shared memory budget exceeded");
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = wgrad3x3_tc_pair_kernel<BN, STAGES>;
  if (!attr_done[dev & 15]) {
    MAU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes));
    attr_done[dev & 15] = true;
  }
  kern<<<op.grid, kThreads, S::kBytes, st>>>(op.tmDy, op.tmX, op.tmWs, p);
  MAU_LAUNCHED();
  return 0;
}

// workspace [9][D1][ld0] -> OIHW.  normal: D1 = Cout, inner = ci; swapped: D1 = Cin, inner = co.
// A block owns 256 consecutive (co, ci) pairs of the OIHW tensor: nine coalesced plane reads into shared memory,
// then one contiguous 256 x 9 float run written out (the direct version wrote 36-byte pieces per thread).
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ ws, int swap, int Cout, int Cin,
                                                             int ld0, float* __restrict__ dw) {
  __shared__ float tile[256 * 9 + 8];
  const long long total = (long long)Cout * Cin;
  const long long plane = (long long)(swap ? Cin : Cout) * ld0;
  for (long long base = (long long)blockIdx.x * 256; base < total; base += (long long)gridDim.x * 256) {
    const long long i = base + threadIdx.x;
    if (i < total) {
      long long src;
      if (!swap) { const int ci = (int)(i % Cin), co = (int)(i / Cin); src = (long long)co * ld0 + ci; }
      else       { const int ci = (int)(i % Cin), co = (int)(i / Cin); src = (long long)ci * ld0 + co; }
#pragma unroll
      for (int t = 0; t < 9; ++t) tile[threadIdx.x * 9 + t] = ws[t * plane + src];
    }
    __syncthreads();
    const long long n = (total - base < 256 ? total - base : 256) * 9;
    float* dst = dw + base * 9;
    for (int e = threadIdx.x; e < n; e += 256) dst[e] = tile[e];
    __syncthreads();
  }
}

template <int BN, bool SWAP, int STAGES>
int launch_wgrad_v2(const WgradTcOp& op, const WgradV2Params& p, cudaStream_t st) {
  using S = WgradV2Smem<BN, SWAP, STAGES>;
  static_assert(S::kBytes <= 232448, "shared memory budget exceeded");
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = wgrad3x3_tc_v2_kernel<BN, SWAP, STAGES>;
  if (!attr_done[dev & 15]) {
    MAU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes));
    attr_done[dev & 15] = true;
  }
  kern<<<op.grid, kThreads, S::kBytes, st>>>(op.tmDy, op.tmX, op.tmWs, p);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace

// tile-padding cost of one orientation: sum over segments of (padded M x padded N), N = 64 instructions
// discounted for their shared-memory operand bandwidth limit
static double orient_cost(int Cout, int nseg, const int* seg_len, bool swap) {
  double c = 0;
  for (int s = 0; s < nseg; ++s) {
    const int Md = swap ? seg_len[s] : Cout, Nd = swap ? Cout : seg_len[s];
    const int bn = Nd <= 64 ? 64 : 128;
    c += (double)ceil_div(Md, 128) * 128 * ceil_div(Nd, bn) * bn / (bn == 64 ? 0.67 : 1.0);
  }
  return c;
}
int wgrad_tc_pick_swap(int Cout, int nseg, const int* seg_len) {
  if (const char* e = getenv("MAU_WGRAD_SWAP")) return atoi(e);
  // narrow side <= 64 channels: the tap-pair kernel (2 = dY carries the shift, 3 = X carries it); MAU_WGRAD_PAIR=0 disables
  static const bool pair_ok = [] { const char* e = getenv("MAU_WGRAD_PAIR"); return !e || atoi(e) != 0; }();
  if (pair_ok && Cout <= 64) return 2;
  if (pair_ok && nseg == 1 && seg_len[0] <= 64) return 3;
  return orient_cost(Cout, nseg, seg_len, true) < orient_cost(Cout, nseg, seg_len, false) ? 1 : 0;
}
size_t wgrad_tc_workspace_floats(int Cout, int Cin_w, int swap) {
  const int d0 = (swap & 1) ? Cout : Cin_w, d1 = (swap & 1) ? Cin_w : Cout;
  return (size_t)9 * d1 * round_up(d0, 4);
}
// N tile of the pair kernel: 96 columns (5 x 96 = 480 TMEM columns) when that pads less than 64-wide tiles
static int pair_bn(int Nd) {
  if (Nd <= 64) return 64;
  const double c96 = (double)ceil_div(Nd, 96) * 96, c64 = (double)ceil_div(Nd, 64) * 64 / 0.75;
  return c96 <= c64 ? 96 : 64;
}

int wgrad_tc_prepare(WgradTcOp* op, const View& x_seg, const View& dy, int ci_w0, int Cin_w, float* ws, int swap) {
  if (x_seg.B != dy.B || x_seg.H != dy.H || x_seg.W != dy.W) return fail("wgrad_tc: geometry mismatch");
  if (x_seg.cs % 8 || x_seg.c0 % 8 || dy.cs % 8 || dy.c0 % 8) return fail("wgrad_tc: views must be 8-channel aligned");
  op->B = dy.B; op->H = dy.H; op->W = dy.W;
  op->Cout = dy.C; op->Cin = x_seg.C; op->ci_w0 = ci_w0; op->Cin_w = Cin_w;
  op->ws = ws; op->swap = swap;
  if (!ws) {          // v1: one CTA per (tile, filter row, pixel split), atomics into the OIHW gradient
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
  if (swap >= 2) {      // tap-pair kernel: S = shifted narrow operand (halo box), O = the other operand
    const bool dy_shift = swap == 2;
    const View& sv = dy_shift ? dy : x_seg;
    const View& ov = dy_shift ? x_seg : dy;
    if (sv.C > 64) return fail("wgrad_tc pair: the shifted operand has %d > 64 channels", sv.C);
    op->bn = pair_bn(ov.C);
    op->m_tiles = 1;
    op->n_tiles = ceil_div(ov.C, op->bn);
    const int T = dy.B * ceil_div(dy.H, 4) * ceil_div(dy.W, 16);
    const int U = op->n_tiles;
    const int sms = sm_budget();
    const int G = 2 * U <= sms ? U * std::min(sms / U, T) : (int)std::min<long long>(sms, (long long)U * T);
    op->grid = dim3((unsigned)G, 1, 1);
    MAU_TRY(make_nhwc_map(&op->tmDy, DT_BF16, sv, 64, 18, 6));
    MAU_TRY(make_nhwc_map(&op->tmX, DT_BF16, ov, 64, 16, 4));
    const int d0 = dy_shift ? Cin_w : dy.C, d1 = dy_shift ? dy.C : Cin_w;      // workspace [9][narrow][wide]
    uint64_t dims[3] = {(uint64_t)d0, (uint64_t)d1, 9};
    uint64_t str[2] = {(uint64_t)round_up(d0, 4) * 4, (uint64_t)round_up(d0, 4) * 4 * (uint64_t)d1};
    uint32_t box[3] = {32, 64, 1};
    MAU_TRY(make_tensor_map(&op->tmWs, DT_F32, 3, ws, dims, str, box, true));
    return 0;
  }
  const int Md = swap ? x_seg.C : dy.C, Nd = swap ? dy.C : x_seg.C;
  op->bn = Nd <= 64 ? 64 : 128;
  op->m_tiles = ceil_div(Md, 128);
  op->n_tiles = ceil_div(Nd, op->bn);
  const int T = dy.B * ceil_div(dy.H, 4) * ceil_div(dy.W, 16);
  const int U = op->m_tiles * op->n_tiles * 3;
  const int sms = sm_budget();
  int G;
  if (2 * U <= sms) G = U * std::min(sms / U, T);                                   // split-K, units in lock step
  else G = (int)std::min<long long>(sms, (long long)U * T);                        // stream-K over all SMs
  op->grid = dim3((unsigned)G, 1, 1);
  MAU_TRY(make_nhwc_map(&op->tmDy, DT_BF16, dy, 64, 16, 4));
  MAU_TRY(make_nhwc_map(&op->tmX, DT_BF16, x_seg, 64, 18, 4));
  {   // workspace [9][D1][ld0] fp32, box {32, 128, 1}
    const int d0 = swap ? dy.C : Cin_w, d1 = swap ? Cin_w : dy.C;
    uint64_t dims[3] = {(uint64_t)d0, (uint64_t)d1, 9};
    uint64_t str[2] = {(uint64_t)round_up(d0, 4) * 4, (uint64_t)round_up(d0, 4) * 4 * (uint64_t)d1};
    uint32_t box[3] = {32, 128, 1};
    MAU_TRY(make_tensor_map(&op->tmWs, DT_F32, 3, ws, dims, str, box, true));
  }
  return 0;
}

int wgrad_tc_launch(const WgradTcOp& op, float* dw_oihw, cudaStream_t st) {
  if (!op.ws) {
    WgradParams p;
    p.B = op.B; p.H = op.H; p.W = op.W;
    p.tiles_w = ceil_div(op.W, 8); p.tiles_h = ceil_div(op.H, 8);
    p.Cout = op.Cout; p.Cin = op.Cin; p.ci_w0 = op.ci_w0; p.Cin_w = op.Cin_w;
    p.tiles_per_split = op.tiles_per_split;
    p.total_tiles = op.B * p.tiles_w * p.tiles_h;
    if (op.bn == 64) return launch_wgrad<64, 6>(op, p, dw_oihw, st);
    return launch_wgrad<128, 5>(op, p, dw_oihw, st);
  }
  if (op.swap >= 2) {
    WgradPairParams p;
    p.tiles_w = ceil_div(op.W, 16); p.tiles_h = ceil_div(op.H, 4);
    p.T = op.B * p.tiles_w * p.tiles_h;
    p.n_tiles = op.n_tiles;
    p.atoms = (long long)op.n_tiles * p.T;
    p.tap_reversed = op.swap == 2 ? 1 : 0;
    p.ws_n0 = op.swap == 2 ? op.ci_w0 : 0;
    p.ws_m0 = op.swap == 2 ? 0 : op.ci_w0;
    return op.bn == 64 ? launch_wgrad_pair<64, 6>(op, p, st) : launch_wgrad_pair<96, 5>(op, p, st);
  }
  WgradV2Params p;
  p.tiles_w = ceil_div(op.W, 16); p.tiles_h = ceil_div(op.H, 4);
  p.T = op.B * p.tiles_w * p.tiles_h;
  p.m_tiles = op.m_tiles; p.n_tiles = op.n_tiles;
  p.atoms = (long long)op.m_tiles * op.n_tiles * 3 * p.T;
  p.ws_m0 = op.swap ? op.ci_w0 : 0;
  p.ws_n0 = op.swap ? 0 : op.ci_w0;
  if (op.bn == 64) return op.swap ? launch_wgrad_v2<64, true, 6>(op, p, st) : launch_wgrad_v2<64, false, 6>(op, p, st);
  return op.swap ? launch_wgrad_v2<128, true, 5>(op, p, st) : launch_wgrad_v2<128, false, 5>(op, p, st);
}

int wgrad_tc_finalize(const float* ws, int swap, int Cout, int Cin_w, float* dw_oihw, cudaStream_t st) {
  swap &= 1;                // pair modes 2 / 3 use the layouts of 0 / 1
  const int d0 = swap ? Cout : Cin_w;
  const long long total = (long long)Cout * Cin_w;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  wgrad_finalize_kernel<<<blocks, 256, 0, st>>>(ws, swap, Cout, Cin_w, round_up(d0, 4), dw_oihw);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
