// conv_tc.cu -- 3x3 / pad 1 / stride 1 convolution as an implicit GEMM on the sm_100a tensor cores.
//
// Replaces: nn.Conv2d(cin, cout, 3, padding=1) forward (reference src/model.py:12,14) and, with a
// transposed + spatially flipped weight pack, its data gradient (the autograd of the same lines).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + d(tap), cin] * W[cout, cin, tap]
//   GEMM view: M = 128 output pixels (a TH x TW spatial box of one image), N = BN output channels,
//              K = 9 taps x Cin, consumed in chunks of 64 channels of one tap.
//
// Data movement: activations are NHWC bf16.  A-operand tiles are fetched by TMA as *shifted boxes*
// of the 4-D tensor {C, W, H, B}; out-of-image coordinates are zero-filled by the TMA unit, which is
// the convolution's zero padding, and partial tiles at the right/bottom edge are clipped by the TMA
// store.  Three main-loop variants trade L2->SMEM traffic for descriptor tricks:
//   MODE_TAP  : one {64, TW, TH} box per tap                (9 A loads per 64-channel chunk)
//   MODE_ROW3 : three {64, 8, TH+2} boxes (one per horizontal shift); the three vertical taps are
//               1024-byte-aligned sub-windows of a box     (3 A loads per chunk)
//   MODE_HALO : one {64, 10, TH+2} halo box; all nine taps are sub-windows addressed through the
//               UMMA descriptor (start offset + 1280-byte group stride)   (1 A load per chunk)
// B-operand tiles ({64 k, BN cout} of one tap) come from a packed bf16 weight tensor [9][Cout][Kp].
// Accumulation is fp32 in TMEM (tcgen05.mma kind::f16); the epilogue applies a per-channel affine
// (+ReLU) -- folded eval-mode BatchNorm, or the plain bias in training -- converts to bf16 and
// leaves through a swizzled SMEM staging tile and a TMA store (or TMA reduce-add for gradients
// that accumulate into a concat buffer).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (TMEM lane quarter = warp % 4).
#include "conv_tc.h"
#include "ptx.cuh"
#include "tma.h"
#include <algorithm>
#include <cstdlib>

namespace mau {
using namespace ptx;

namespace {

constexpr int kThreads = 256;

template <int MODE> struct ATile;
template <> struct ATile<MODE_TAP>  { static constexpr int kStage = 16384; static constexpr int kTx = 16384; };
template <> struct ATile<MODE_ROW3> { static constexpr int kStage = 3 * 18432; static constexpr int kTx = 3 * 18432; };
template <> struct ATile<MODE_HALO> { static constexpr int kStage = 23552; static constexpr int kTx = 23040; };

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) { stage = 0; phase ^= 1; }
  }
};

template <int BN, int MODE, int NA, int NB>
__global__ void __launch_bounds__(kThreads) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const __grid_constant__ CUtensorMap tmY,
                                                              const ConvTcParams p) {
  constexpr int A_STAGE = ATile<MODE>::kStage;
  constexpr int B_STAGE = BN * 128;
  static_assert((BN / 64) * 16384 <= NA * A_STAGE + NB * B_STAGE, "epilogue staging does not fit");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + NA * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NB * B_STAGE);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + NA;
  uint64_t* fullB = bars + 2 * NA;
  uint64_t* emptyB = bars + 2 * NA + NB;
  uint64_t* tmem_full = bars + 2 * NA + 2 * NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 2);
  float* s_shift = s_scale + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int b = blockIdx.x / tiles_per_img;
  const int trem = blockIdx.x - b * tiles_per_img;
  const int th = trem / p.tiles_w;
  const int h0 = th * p.TH;
  const int w0 = (trem - th * p.tiles_w) * p.TW;
  const int n0 = blockIdx.y * BN;

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    prefetch_tensormap(&tmY);
    for (int i = 0; i < NA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int total_chunks = 0;
  for (int s = 0; s < p.nseg; ++s) total_chunks += p.seg_chunks[s];

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      Ring ra, rb;
      int kB = 0;  // packed K offset (elements)
      for (int sg = 0; sg < p.nseg; ++sg) {
        for (int kc = 0; kc < p.seg_chunks[sg]; ++kc, kB += 64) {
          const int cA = p.seg_start[sg] + kc * 64;
          if (MODE == MODE_ROW3) {
            mbar_wait(&emptyA[ra.stage], ra.phase ^ 1);
            mbar_expect_tx(&fullA[ra.stage], ATile<MODE>::kTx);
#pragma unroll
            for (int s = 0; s < 3; ++s)
              tma_load_4d(sA + ra.stage * A_STAGE + s * 18432, &tmA, &fullA[ra.stage], cA, w0 + s - 1, h0 - 1, b);
            ra.advance(NA);
          } else if (MODE == MODE_HALO) {
            mbar_wait(&emptyA[ra.stage], ra.phase ^ 1);
            mbar_expect_tx(&fullA[ra.stage], ATile<MODE>::kTx);
            tma_load_4d(sA + ra.stage * A_STAGE, &tmA, &fullA[ra.stage], cA, w0 - 1, h0 - 1, b);
            ra.advance(NA);
          }
          for (int tap = 0; tap < 9; ++tap) {
            if (MODE == MODE_TAP) {
              const int r = tap / 3, s = tap - r * 3;
              mbar_wait(&emptyA[ra.stage], ra.phase ^ 1);
              mbar_expect_tx(&fullA[ra.stage], ATile<MODE>::kTx);
              tma_load_4d(sA + ra.stage * A_STAGE, &tmA, &fullA[ra.stage], cA, w0 + s - 1, h0 + r - 1, b);
              ra.advance(NA);
            }
            mbar_wait(&emptyB[rb.stage], rb.phase ^ 1);
            mbar_expect_tx(&fullB[rb.stage], B_STAGE);
            tma_load_3d(sB + rb.stage * B_STAGE, &tmB, &fullB[rb.stage], kB, n0, tap);
            rb.advance(NB);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(128, BN, 0, 0);
      Ring ra, rb;
      uint32_t acc = 0;
      for (int chunk = 0; chunk < total_chunks; ++chunk) {
        if (MODE != MODE_TAP) {
          mbar_wait(&fullA[ra.stage], ra.phase);
        }
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - r * 3;
          if (MODE == MODE_TAP) mbar_wait(&fullA[ra.stage], ra.phase);
          mbar_wait(&fullB[rb.stage], rb.phase);
          tc_fence_after();
          uint32_t a_addr = smem_u32(sA + ra.stage * A_STAGE);
          uint32_t sbo = 1024;
          if (MODE == MODE_ROW3) a_addr += s * 18432 + r * 1024;
          if (MODE == MODE_HALO) { a_addr += (r * 10 + s) * 128; sbo = 1280; }
          const uint32_t b_addr = smem_u32(sB + rb.stage * B_STAGE);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t aa = a_addr + k * 32;
            const uint32_t bo = (MODE == MODE_HALO && p.halo_base_offset) ? ((aa >> 7) & 7) : 0;
            const uint64_t da = smem_desc_sw128(aa, 16, sbo, bo);
            const uint64_t db = smem_desc_sw128(b_addr + k * 32, 16, 1024, 0);
            umma_bf16(tmem_base, da, db, idesc, acc);
            acc = 1;
          }
          umma_commit(&emptyB[rb.stage]);
          rb.advance(NB);
          if (MODE == MODE_TAP) {
            umma_commit(&emptyA[ra.stage]);
            ra.advance(NA);
          }
        }
        if (MODE != MODE_TAP) {
          umma_commit(&emptyA[ra.stage]);
          ra.advance(NA);
        }
      }
      umma_commit(tmem_full);
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int et = threadIdx.x - 128;  // 0..127
    for (int i = et; i < BN; i += 128) {
      const int c = n0 + i;
      const bool ok = c < p.Cout;
      s_scale[i] = ok ? (p.scale ? p.scale[c] : 1.f) : 0.f;
      s_shift[i] = ok ? (p.shift ? p.shift[c] : 0.f) : 0.f;
    }
    named_bar_sync(1, 128);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator row == TMEM lane == pixel index inside the tile
    uint8_t* stg = smem;          // pipeline buffers are idle now: reuse as the store staging tile
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + cb * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = cb * 32 + 2 * j;
        float y0 = __uint_as_float(v[2 * j]) * s_scale[c] + s_shift[c];
        float y1 = __uint_as_float(v[2 * j + 1]) * s_scale[c + 1] + s_shift[c + 1];
        if (p.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
        __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
      }
      uint8_t* rowp = stg + (cb >> 1) * 16384 + m * 128;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int chunk16 = (cb & 1) * 4 + jj;
        uint4 val = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
        *reinterpret_cast<uint4*>(rowp + ((chunk16 ^ (m & 7)) << 4)) = val;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (et == 0) {
#pragma unroll 1
      for (int g = 0; g < BN / 64; ++g) {
        if (n0 + g * 64 < p.Cout) {
          if (p.accumulate) tma_reduce_add_4d(&tmY, stg + g * 16384, n0 + g * 64, w0, h0, b);
          else              tma_store_4d(&tmY, stg + g * 16384, n0 + g * 64, w0, h0, b);
        }
      }
      tma_commit_group();
      tma_wait_group0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ==========================================================================================
// v2: persistent, halo-reuse main loop with MT sub-tiles per B tile and a double-buffered TMEM
// accumulator so that the epilogue of work item i overlaps the MMAs of item i+1.
//
//   work item = MT consecutive 128-pixel tiles (16 rows x 8 cols each, possibly from different
//               images) x one BN-wide output-channel tile; MT * BN = 256 TMEM columns per buffer.
//   per 64-channel chunk: MT halo boxes {64, 10, 18} (one TMA load each) feed 9 taps x MT
//               sub-tiles; every B tile ({64, BN} of one tap) is used by MT MMAs groups, which
//               divides the weight traffic from L2 by MT.
// ==========================================================================================
constexpr int kHaloBytes = 23040;    // 18 x 10 pixels x 128 B
constexpr int kHaloStride = 23552;   // padded to a multiple of 1024 B (swizzle atom alignment)

template <int BN, int MT, int NBUF, int NA, int NB, int NSTG>
struct V2Smem {
  static constexpr int A_STAGE = MT * kHaloStride;
  static constexpr int B_STAGE = BN * 128;
  static constexpr int SU = BN % 128 == 0 ? 128 : 64;   // columns staged per TMA-store round (BN = 64, 192: 64)
  static constexpr int STG = (SU / 64) * 16384;    // 128 pixels x SU bf16 channels
  static constexpr int kBars = 2 * NA + 2 * NB + 4;
  static constexpr int kHeadOC = 4;                 // fused 1x1 head: up to 4 output channels
  static constexpr size_t kBytes = 1024 + (size_t)NA * A_STAGE + (size_t)NB * B_STAGE + (size_t)NSTG * STG + 8 * kBars +
                                   32 + (2 + (BN <= 128 ? kHeadOC : 0)) * BN * sizeof(float);
};

// EM = epilogue mode: 0 none (data gradient), 1 + shift (training forward: bias), 2 * scale + shift (eval: folded
// BatchNorm), 3 = 2 + fused 1x1 head.  A template parameter so that every instance carries one epilogue only.
// BRES = weights resident: when the whole packed weight tensor of the launch (9 taps x K chunks, one N tile) fits the
// NB B stages it is loaded once per CTA instead of once per work item (the Cin <= 64, Cout <= 64 layers at 250x250:
// 72 KB of weights against 46 KB of activations per item).
// named barriers of the statistics warps (STATS): staging buffer s "full" (128 epilogue threads arrive, 128 statistics
// threads wait) and "free" (the reverse); id 0 is __syncthreads, id 1 the epilogue's own barrier.  STATS instances run
// with 384 threads: warps 8..11 are the statistics warps, one per SM sub-partition, so that each of the four epilogue
// warps shares its issue slots with a quarter of the statistics work (two statistics warps on warps 2 / 3 slowed two of
// the four epilogue warps -- and with them every tile -- twice as much)
constexpr int kBarFull = 2, kBarFree = 4, kStatThreads = 256, kStatsBlock = 384;

// per-channel sum / sum of squares of one 64-channel group of a staged bf16 tile (rows of 128 B, 16-byte chunks XOR-swizzled
// by row & 7): lane l owns channels 2l, 2l + 1; valid(row) masks pixels outside the image
template <typename F>
__device__ __forceinline__ void stat_rows(const uint8_t* tile, int row0, int row1, int lane, F&& valid, float (&s)[2], float (&q)[2]) {
  // eight rows per iteration, all eight loads issued before the first use (a one-row-at-a-time loop is bound by the
  // shared-memory latency and made the epilogue wait for its staging buffer); row0 and row1 are multiples of 8, so the
  // swizzle term of row r + i is i
  float sa[2] = {0.f, 0.f}, sb[2] = {0.f, 0.f}, qa[2] = {0.f, 0.f}, qb[2] = {0.f, 0.f};
  const uint8_t* base = tile + (lane & 3) * 4;
  const int c16 = lane >> 2;
  for (int r = row0; r < row1; r += 8) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = *reinterpret_cast<const uint32_t*>(base + (r + i) * 128 + ((c16 ^ i) << 4));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t x = valid(r + i) ? w[i] : 0u;
      const float v0 = __uint_as_float(x << 16), v1 = __uint_as_float(x & 0xffff0000u);
      if (i & 1) { sb[0] += v0; sb[1] += v1; qb[0] = fmaf(v0, v0, qb[0]); qb[1] = fmaf(v1, v1, qb[1]); }
      else       { sa[0] += v0; sa[1] += v1; qa[0] = fmaf(v0, v0, qa[0]); qa[1] = fmaf(v1, v1, qa[1]); }
    }
  }
  s[0] = sa[0] + sb[0]; s[1] = sa[1] + sb[1]; q[0] = qa[0] + qb[0]; q[1] = qa[1] + qb[1];
}

template <int BN, int MT, int NBUF, int NA, int NB, int NSTG, bool BT, int EM, bool BRES, bool STATS = false>
__global__ void __launch_bounds__(STATS ? kStatsBlock : kThreads, 1) conv3x3_tc_v2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const __grid_constant__ CUtensorMap tmY,
                                                                   const ConvTcParams p) {
  using S = V2Smem<BN, MT, NBUF, NA, NB, NSTG>;
  static_assert(NBUF * MT * BN <= 512 && (NBUF == 1 || NBUF == 2), "accumulators must fit the 512 TMEM columns");
  constexpr int kBufCols = MT * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + NA * S::A_STAGE;
  uint8_t* sStg = sB + NB * S::B_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + NSTG * S::STG);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + NA;
  uint64_t* fullB = emptyA + NA;
  uint64_t* emptyB = fullB + NB;
  uint64_t* tmem_full = emptyB + NB;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_scale = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~uintptr_t(15));   // float4 reads
  float* s_shift = s_scale + BN;
  float* s_hw = s_shift + BN;            // [kHeadOC][BN], only when BN <= 128
  constexpr bool kCanHead = BN <= 128 && EM == 3;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_img = p.tiles_w * p.tiles_h;
  const int total_pt = p.total_ptiles;
  const int groups = (total_pt + MT - 1) / MT;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int total_items = groups * n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); prefetch_tensormap(&tmY);
    for (int i = 0; i < NA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int total_chunks = 0;
  for (int s = 0; s < p.nseg; ++s) total_chunks += p.seg_chunks[s];

  if (warp == 0) {
    // =========================== TMA producer (whole warp converged, one elected lane issues) ==========
    Ring ra, rb;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int n0 = (it / groups) * BN;
      const int pt0 = (it % groups) * MT;
      const int nsub = min(MT, total_pt - pt0);
      int kB = 0;
      for (int sg = 0; sg < p.nseg; ++sg) {
        for (int kc = 0; kc < p.seg_chunks[sg]; ++kc, kB += 64) {
          const int cA = p.seg_start[sg] + kc * 64;
          mbar_wait(&emptyA[ra.stage], ra.phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&fullA[ra.stage], nsub * kHaloBytes);
            for (int j = 0; j < nsub; ++j) {
              const int pt = pt0 + j;
              const int b = pt / tiles_img;
              const int rem = pt - b * tiles_img;
              const int th = rem / p.tiles_w;
              tma_load_4d(sA + ra.stage * S::A_STAGE + j * kHaloStride, &tmA, &fullA[ra.stage], cA,
                          (rem - th * p.tiles_w) * 8 - 1, th * 16 - 1, b);
            }
          }
          __syncwarp();
          ra.advance(NA);
          for (int tap = 0; tap < 9; ++tap) {
            if (BRES && it != (int)blockIdx.x) break;      // resident weights: loaded with the first item only
            mbar_wait(&emptyB[rb.stage], rb.phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&fullB[rb.stage], S::B_STAGE);
              if (BT) {
                // data gradient: B = W^T of the flipped tap, read MN-major straight from the FORWARD weight pack
                // [9][Cout][Kp]: boxes {64 ci, 64 co} -> smem [64 k rows][128 B of n]; no transposed re-pack
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_3d(sB + rb.stage * S::B_STAGE + j * 8192, &tmB, &fullB[rb.stage], p.bt_col0 + n0 + 64 * j, kB, 8 - tap);
              } else {
                tma_load_3d(sB + rb.stage * S::B_STAGE, &tmB, &fullB[rb.stage], kB, n0, tap);
              }
            }
            __syncwarp();
            rb.advance(NB);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (whole warp converged, one elected lane issues) ============
    constexpr uint32_t idesc = idesc_bf16_f32(128, BN, 0, BT ? 1 : 0);
    // descriptor templates: only the 14-bit start-address field changes between MMAs
    const uint64_t descA0 = smem_desc_sw128(0, 16, 1280, 0);
    // B K-major: rows of 128 B = 64 k of one n.  B MN-major (BT): rows of 128 B = 64 n of one k, 64-wide n groups
    // at LBO = 8192 B, 8-row k groups at SBO = 1024 B; a K = 16 step advances 16 rows = 2048 B.
    const uint64_t descB0 = BT ? smem_desc_sw128(0, 8192, 1024, 0) : smem_desc_sw128(0, 16, 1024, 0);
    constexpr int kBStep = BT ? 2048 : 32;
    Ring ra, rb;
    int buf = 0;
    uint32_t ephase[2] = {0, 0};
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int pt0 = (it % groups) * MT;
      const int nsub = min(MT, total_pt - pt0);
      mbar_wait(&tmem_empty[buf], ephase[buf] ^ 1);   // epilogue has drained this accumulator buffer
      ephase[buf] ^= 1;
      tc_fence_after();
      const uint32_t d0 = tmem_base + buf * kBufCols;
      for (int chunk = 0; chunk < total_chunks; ++chunk) {
        mbar_wait(&fullA[ra.stage], ra.phase);
        const uint32_t a_stage = smem_u32(sA + ra.stage * S::A_STAGE);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - r * 3;
          if (!BRES || it == (int)blockIdx.x) mbar_wait(&fullB[rb.stage], rb.phase);
          tc_fence_after();
          const uint64_t db = descB0 + (uint64_t)(smem_u32(sB + rb.stage * S::B_STAGE) >> 4);
          const uint64_t da = descA0 + (uint64_t)((a_stage + (r * 10 + s) * 128) >> 4);
          const uint32_t first = (chunk | tap) ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              if (j < nsub) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(d0 + j * BN, da + (uint64_t)((j * kHaloStride + k * 32) >> 4), db + (uint64_t)((k * kBStep) >> 4),
                            idesc, k ? 1u : first);
              }
            }
            if (!BRES) umma_commit(&emptyB[rb.stage]);
            if (tap == 8) umma_commit(&emptyA[ra.stage]);
            if (tap == 8 && chunk == total_chunks - 1) umma_commit(&tmem_full[buf]);
          }
          __syncwarp();
          rb.advance(NB);
        }
        ra.advance(NA);
      }
      buf = (buf + 1) % NBUF;
    }
  } else if (STATS && warp >= 8) {
    // =========================== BatchNorm statistics (training forward) ===========================
    // Mirrors the epilogue's iteration space; reads every staged tile once while the epilogue warps work on the next.
    constexpr int NU = BN / S::SU;                  // store rounds per 128-pixel tile
    constexpr int NG = S::SU / 64;                  // 64-channel groups per round
    const int sw = warp - 8;                        // NG == 2: group sw & 1, rows (sw >> 1) * 64 ..; NG == 1: rows sw * 32 ..
    const int grp = NG == 2 ? (sw & 1) : 0;
    const int r0 = NG == 2 ? (sw >> 1) * 64 : sw * 32, r1 = r0 + (NG == 2 ? 64 : 32);
    double as[NU][2], aq[NU][2];
#pragma unroll
    for (int u = 0; u < NU; ++u) { as[u][0] = as[u][1] = aq[u][0] = aq[u][1] = 0.0; }
    auto flush = [&](int n0) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int c = n0 + u * S::SU + grp * 64 + 2 * lane;
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (c + e < p.Cout) { atomicAdd(p.stats + c + e, as[u][e]); atomicAdd(p.stats + p.Cout + c + e, aq[u][e]); }
        as[u][0] = as[u][1] = aq[u][0] = aq[u][1] = 0.0;
      }
    };
    for (int s = 0; s < NSTG; ++s) named_bar_arrive(kBarFree + s, kStatThreads);      // every staging buffer starts free
    int stg = 0, cur_n0 = -1;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int n0 = (it / groups) * BN;
      const int pt0 = (it % groups) * MT;
      const int nsub = min(MT, total_pt - pt0);
      if (n0 != cur_n0) { if (cur_n0 >= 0) flush(cur_n0); cur_n0 = n0; }
      for (int j = 0; j < nsub; ++j) {
        const int pt = pt0 + j;
        const int b = pt / tiles_img;
        const int rem = pt - b * tiles_img;
        const int th = rem / p.tiles_w;
        const int w0 = (rem - th * p.tiles_w) * 8, h0 = th * 16;
        auto valid = [&](int r) { return h0 + (r >> 3) < p.H && w0 + (r & 7) < p.W; };
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          named_bar_sync(kBarFull + stg, kStatThreads);
          float s2[2], q2[2];
          stat_rows(sStg + stg * S::STG + grp * 16384, r0, r1, lane, valid, s2, q2);
          named_bar_arrive(kBarFree + stg, kStatThreads);
          as[u][0] += (double)s2[0]; as[u][1] += (double)s2[1]; aq[u][0] += (double)q2[0]; aq[u][1] += (double)q2[1];
          stg = (stg + 1) % NSTG;
        }
      }
    }
    if (cur_n0 >= 0) flush(cur_n0);
  } else if (warp >= 4 && warp < 8) {
    // =========================== epilogue ===========================
    const int et = threadIdx.x - 128;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    constexpr int emode = EM >= 2 ? 2 : EM;
    int buf = 0, stg = 0, cur_n0 = -1;
    uint32_t fphase[2] = {0, 0};
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int n0 = (it / groups) * BN;
      const int pt0 = (it % groups) * MT;
      const int nsub = min(MT, total_pt - pt0);
      if (n0 != cur_n0) {
        named_bar_sync(1, 128);           // previous users of s_scale are done
        for (int i = et; i < BN; i += 128) {
          const int c = n0 + i;
          const bool ok = c < p.Cout;
          s_scale[i] = ok ? (p.scale ? p.scale[c] : 1.f) : 0.f;
          s_shift[i] = ok ? (p.shift ? p.shift[c] : 0.f) : 0.f;
          if constexpr (kCanHead)
            for (int o = 0; o < S::kHeadOC; ++o) s_hw[o * BN + i] = (ok && o < p.head_oc) ? p.head_w[o * p.Cout + c] : 0.f;
        }
        named_bar_sync(1, 128);
        cur_n0 = n0;
      }
      mbar_wait(&tmem_full[buf], fphase[buf]);
      fphase[buf] ^= 1;
      tc_fence_after();
      for (int j = 0; j < nsub; ++j) {
        const int pt = pt0 + j;
        const int b = pt / tiles_img;
        const int rem = pt - b * tiles_img;
        const int th = rem / p.tiles_w;
        const int w0 = (rem - th * p.tiles_w) * 8, h0 = th * 16;
        float hacc[S::kHeadOC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int u = 0; u < BN / S::SU; ++u) {          // store rounds of SU columns
          uint8_t* sbuf = sStg + stg * S::STG;
          const bool last = (j == nsub - 1) && (u == BN / S::SU - 1);
          // the TMA store that last read this staging buffer must have finished reading it
          if (p.store_y) {
            if (et == 0) { if (NSTG == 1) tma_wait_group_read0(); else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
            named_bar_sync(1, 128);
          }
          if constexpr (STATS) named_bar_sync(kBarFree + stg, kStatThreads);   // the statistics warps have read this buffer
          // TMEM -> registers is double-buffered: the load of column block cb + 1 is in flight while block cb is
          // converted and staged.  Per-channel affine: mode 0 = none (data gradient), 1 = + shift (training forward:
          // bias), 2 = * scale + shift (eval: folded BatchNorm); coefficients are read as float4 broadcasts.
          constexpr int kCB = S::SU / 32;
          uint32_t vv[2][32];
          const uint32_t tcol0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBufCols + j * BN + u * S::SU;
          tmem_ld_32x32(tcol0, vv[0]);
#pragma unroll
          for (int cb = 0; cb < kCB; ++cb) {
            uint32_t (&v)[32] = vv[cb & 1];
            const int col = u * S::SU + cb * 32;
            tmem_ld_wait();
            reg_fence32(v);
            if (cb + 1 < kCB) tmem_ld_32x32(tcol0 + (cb + 1) * 32, vv[(cb + 1) & 1]);
            uint32_t pk[16];
            if constexpr (emode == 0) {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                float y0 = __uint_as_float(v[2 * jj]), y1 = __uint_as_float(v[2 * jj + 1]);
                if (p.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
                pk[jj] = *reinterpret_cast<uint32_t*>(&h2);
              }
            } else if constexpr (emode == 1) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 sh = *reinterpret_cast<const float4*>(s_shift + col + 4 * j4);
                float y0 = __uint_as_float(v[4 * j4]) + sh.x, y1 = __uint_as_float(v[4 * j4 + 1]) + sh.y;
                float y2 = __uint_as_float(v[4 * j4 + 2]) + sh.z, y3 = __uint_as_float(v[4 * j4 + 3]) + sh.w;
                if (p.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
                __nv_bfloat162 ha = __floats2bfloat162_rn(y0, y1), hb = __floats2bfloat162_rn(y2, y3);
                pk[2 * j4] = *reinterpret_cast<uint32_t*>(&ha);
                pk[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&hb);
              }
            } else {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const int c = col + 4 * j4;
                const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
                const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
                float y0 = fmaf(__uint_as_float(v[4 * j4]), sc.x, sh.x), y1 = fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y);
                float y2 = fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), y3 = fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w);
                if (p.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
                __nv_bfloat162 ha = __floats2bfloat162_rn(y0, y1), hb = __floats2bfloat162_rn(y2, y3);
                pk[2 * j4] = *reinterpret_cast<uint32_t*>(&ha);
                pk[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&hb);
                if constexpr (kCanHead) {           // the head sees the bf16-rounded activation, like the unfused path
                  const float r0 = __low2float(ha), r1 = __high2float(ha), r2 = __low2float(hb), r3 = __high2float(hb);
#pragma unroll
                  for (int o = 0; o < S::kHeadOC; ++o) {
                    const float4 hw = *reinterpret_cast<const float4*>(s_hw + o * BN + c);
                    hacc[o] = fmaf(r3, hw.w, fmaf(r2, hw.z, fmaf(r1, hw.y, fmaf(r0, hw.x, hacc[o]))));
                  }
                }
              }
            }
            if (p.store_y) {
              uint8_t* rowp = sbuf + (cb >> 1) * 16384 + m * 128;
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int chunk16 = (cb & 1) * 4 + jj;
                *reinterpret_cast<uint4*>(rowp + ((chunk16 ^ (m & 7)) << 4)) =
                    make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
              }
            }
          }
          if (last) tc_fence_before();   // all TMEM reads of this buffer are complete (wait::ld above)
          if constexpr (STATS) named_bar_arrive(kBarFull + stg, kStatThreads);   // this thread's part of the tile is staged
          if (p.store_y) fence_proxy_async_smem();
          if (p.store_y || last) named_bar_sync(1, 128);
          if (et == 0) {
            if (last) mbar_arrive(&tmem_empty[buf]);   // hand the accumulator buffer back to the MMA warp
            if (p.store_y) {
#pragma unroll 1
              for (int g = 0; g < S::SU / 64; ++g) {
                const int c = n0 + u * S::SU + g * 64;
                if (c < p.Cout) {
                  if (p.accumulate) tma_reduce_add_4d(&tmY, sbuf + g * 16384, c, w0, h0, b);
                  else              tma_store_4d(&tmY, sbuf + g * 16384, c, w0, h0, b);
                }
              }
              tma_commit_group();
            }
          }
          if (p.store_y) stg = (stg + 1) % NSTG;
        }
        if constexpr (kCanHead) {
          const int h = h0 + (m >> 3), w = w0 + (m & 7);
          if (h < p.H && w < p.W) {
#pragma unroll
            for (int o = 0; o < S::kHeadOC; ++o)
              if (o < p.head_oc) {
                float v = hacc[o] + p.head_b[o];
                if (p.head_tanh && o == 0) v = tanhf(v);
                p.head_out[(((long long)b * p.head_oc + o) * p.H + h) * p.W + w] = v;
              }
          }
        }
      }
      buf = (buf + 1) % NBUF;
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

template <int BN, int MT, int NBUF, int NA, int NB, int NSTG, bool BT, int EM, bool BRES = false, bool STATS = false>
int launch_v2(const ConvTcOp& op, cudaStream_t st) {
  using S = V2Smem<BN, MT, NBUF, NA, NB, NSTG>;
  static_assert(S::kBytes <= 232448, "shared memory budget exceeded");
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = conv3x3_tc_v2_kernel<BN, MT, NBUF, NA, NB, NSTG, BT, EM, BRES, STATS>;
  if (!attr_done[dev & 15]) {
    MAU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes));
    attr_done[dev & 15] = true;
  }
  kern<<<op.grid, STATS ? kStatsBlock : kThreads, S::kBytes, st>>>(op.tmA, op.tmB, op.tmY, op.p);
  MAU_LAUNCHED();
  return 0;
}

// ==========================================================================================
// v3 ("col3"): Cout = 64 layers.  An M = 128, N = 64 MMA reads 6 KB of shared memory for 32 clocks of tensor work
// (192 B/clk against ~85 sustained): the v2 kernel runs those layers at 35-38 % tensor-pipe activity.  Here the
// three HORIZONTAL taps of a filter row are folded into N: one MMA  D[128 px, 192] += A(row r) * [W(r,0)|W(r,1)|W(r,2)]
// uses the same A window for all three, so A is read 3 instead of 9 times per chunk (10 KB per 96 clocks).  The
// accumulator group s holds  D_s[h, w] = sum_{r,k} X[h+r-1, w] * W[k, co, (r,s)]  (input column = centre column), and
//     Y[h, w] = D_0[h, w-1] + D_1[h, w] + D_2[h, w+1]
// is formed in the epilogue with two lane shuffles (a tile is 8 rows x 16 columns, a warp owns two tile rows).
// Columns 0 and 15 of a tile have no neighbour in the tile: tiles overlap by two columns (14 valid columns).  The A
// operand needs no horizontal halo: TMA box {64 ch, 16, 10}, row pitch 2048 B, the three vertical taps are the
// 1024-byte-aligned windows starting at halo rows 0, 1, 2.  With Cout = 64 the [9][64][Kp] weight pack *is*
// [3][192][Kp], so no second pack exists.
// ==========================================================================================
constexpr int kC3ABytes = 160 * 128;      // {64, 16, 10} bf16
constexpr int kC3BBytes = 192 * 128;      // 192 n x 64 k
constexpr int kC3Stg = 16384;             // 112 valid pixels x 128 B (padded)

template <int NA, int NB>
struct V3Smem {
  static constexpr int kBars = 2 * NA + 2 * NB + 4;
  static constexpr size_t kBytes = 1024 + (size_t)NA * kC3ABytes + (size_t)NB * kC3BBytes + 2 * kC3Stg + 8 * kBars + 32 +
                                   (2 + 4) * 64 * sizeof(float);
};

template <int NA, int NB, int EM, bool STATS = false>
__global__ void __launch_bounds__(STATS ? kStatsBlock : kThreads, 1) conv3x3_tc_col3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     const __grid_constant__ CUtensorMap tmY,
                                                                     const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + NA * kC3ABytes;
  uint8_t* sStg = sB + NB * kC3BBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 2 * kC3Stg);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + NA;
  uint64_t* fullB = emptyA + NA;
  uint64_t* emptyB = fullB + NB;
  uint64_t* tmem_full = emptyB + NB;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_scale = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~uintptr_t(15));
  float* s_shift = s_scale + 64;
  float* s_hw = s_shift + 64;            // [4][64] fused head weights (EM == 3)
  constexpr int kHeadOC = 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_img = p.tiles_w * p.tiles_h;
  const int total_items = p.total_ptiles;
  int total_chunks = 0;
  for (int s = 0; s < p.nseg; ++s) total_chunks += p.seg_chunks[s];
  const bool resident = 3 * total_chunks <= NB;     // all weight tiles fit the B stages: load them once per CTA

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); prefetch_tensormap(&tmY);
    for (int i = 0; i < NA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    Ring ra, rb;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int b = it / tiles_img;
      const int rem = it - b * tiles_img;
      const int th = rem / p.tiles_w;
      const int w0 = (rem - th * p.tiles_w) * 14 - 1, h0 = th * 8 - 1;     // top-left of the input window
      const bool load_b = !resident || it == (int)blockIdx.x;
      int kB = 0;
      for (int sg = 0; sg < p.nseg; ++sg) {
        for (int kc = 0; kc < p.seg_chunks[sg]; ++kc, kB += 64) {
          const int cA = p.seg_start[sg] + kc * 64;
          mbar_wait(&emptyA[ra.stage], ra.phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&fullA[ra.stage], kC3ABytes);
            tma_load_4d(sA + ra.stage * kC3ABytes, &tmA, &fullA[ra.stage], cA, w0, h0, b);
          }
          __syncwarp();
          ra.advance(NA);
          if (load_b) {
            for (int r = 0; r < 3; ++r) {
              mbar_wait(&emptyB[rb.stage], rb.phase ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&fullB[rb.stage], kC3BBytes);
                tma_load_3d(sB + rb.stage * kC3BBytes, &tmB, &fullB[rb.stage], kB, 0, r);
              }
              __syncwarp();
              rb.advance(NB);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = idesc_bf16_f32(128, 192, 0, 0);
    const uint64_t descA0 = smem_desc_sw128(0, 16, 1024, 0);
    const uint64_t descB0 = smem_desc_sw128(0, 16, 1024, 0);
    Ring ra, rb;
    int buf = 0;
    uint32_t ephase[2] = {0, 0};
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const bool wait_b = !resident || it == (int)blockIdx.x;
      mbar_wait(&tmem_empty[buf], ephase[buf] ^ 1);
      ephase[buf] ^= 1;
      tc_fence_after();
      const uint32_t d0 = tmem_base + buf * 192;
      if (resident) { rb.stage = 0; }
      for (int chunk = 0; chunk < total_chunks; ++chunk) {
        mbar_wait(&fullA[ra.stage], ra.phase);
        const uint32_t a_stage = smem_u32(sA + ra.stage * kC3ABytes);
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
          if (wait_b) mbar_wait(&fullB[rb.stage], rb.phase);
          tc_fence_after();
          const uint64_t db = descB0 + (uint64_t)(smem_u32(sB + rb.stage * kC3BBytes) >> 4);
          const uint64_t da = descA0 + (uint64_t)((a_stage + r * 2048) >> 4);      // halo rows r .. r+7
          const uint32_t first = (chunk | r) ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d0, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, k ? 1u : first);
            if (!resident) umma_commit(&emptyB[rb.stage]);
            if (r == 2) umma_commit(&emptyA[ra.stage]);
            if (r == 2 && chunk == total_chunks - 1) umma_commit(&tmem_full[buf]);
          }
          __syncwarp();
          rb.advance(NB);
        }
        ra.advance(NA);
      }
      buf ^= 1;
    }
  } else if (STATS && warp >= 8) {
    // =========================== BatchNorm statistics (training forward) ===========================
    // the staged tile is the packed 8 x 14 block of valid outputs (112 rows of 64 channels), split 32 / 32 / 24 / 24 over four warps
    const int sw = warp - 8;
    const int r0 = sw < 2 ? sw * 32 : 64 + (sw - 2) * 24, r1 = r0 + (sw < 2 ? 32 : 24);
    double as[2] = {0.0, 0.0}, aq[2] = {0.0, 0.0};
    for (int s = 0; s < 2; ++s) named_bar_arrive(kBarFree + s, kStatThreads);
    int stg = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int b = it / tiles_img;
      const int rem = it - b * tiles_img;
      const int th = rem / p.tiles_w;
      const int w0 = (rem - th * p.tiles_w) * 14, h0 = th * 8;
      auto valid = [&](int r) { const int hr = r / 14; return h0 + hr < p.H && w0 + (r - hr * 14) < p.W; };
      named_bar_sync(kBarFull + stg, kStatThreads);
      float s2[2], q2[2];
      stat_rows(sStg + stg * kC3Stg, r0, r1, lane, valid, s2, q2);
      named_bar_arrive(kBarFree + stg, kStatThreads);
      as[0] += (double)s2[0]; as[1] += (double)s2[1]; aq[0] += (double)q2[0]; aq[1] += (double)q2[1];
      stg ^= 1;
    }
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (2 * lane + e < p.Cout) { atomicAdd(p.stats + 2 * lane + e, as[e]); atomicAdd(p.stats + p.Cout + 2 * lane + e, aq[e]); }
  } else if (warp >= 4 && warp < 8) {
    // =========================== epilogue ===========================
    const int et = threadIdx.x - 128;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int hh = m >> 4, ww = m & 15;                 // position inside the 8 x 16 tile
    const bool valid_col = ww >= 1 && ww <= 14;
    const int ridx = hh * 14 + ww - 1;                  // row of the packed 8 x 14 staging tile
    for (int i = et; i < 64; i += 128) {
      const bool ok = i < p.Cout;
      s_scale[i] = ok ? (p.scale ? p.scale[i] : 1.f) : 0.f;
      s_shift[i] = ok ? (p.shift ? p.shift[i] : 0.f) : 0.f;
      if (EM == 3)
        for (int o = 0; o < kHeadOC; ++o) s_hw[o * 64 + i] = (ok && o < p.head_oc) ? p.head_w[o * p.Cout + i] : 0.f;
    }
    named_bar_sync(1, 128);
    int buf = 0, stg = 0;
    uint32_t fphase[2] = {0, 0};
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int b = it / tiles_img;
      const int rem = it - b * tiles_img;
      const int th = rem / p.tiles_w;
      const int w0 = (rem - th * p.tiles_w) * 14, h0 = th * 8;            // first VALID output pixel of the tile
      mbar_wait(&tmem_full[buf], fphase[buf]);
      fphase[buf] ^= 1;
      tc_fence_after();
      uint8_t* sbuf = sStg + stg * kC3Stg;
      if (p.store_y) {
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(1, 128);
      }
      if constexpr (STATS) named_bar_sync(kBarFree + stg, kStatThreads);    // the statistics warps have read this buffer
      float hacc[kHeadOC] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 192;
#pragma unroll 1
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t v0[32], v1[32], v2[32];
        tmem_ld_32x32(tcol + cb * 32, v0);
        tmem_ld_32x32(tcol + 64 + cb * 32, v1);
        tmem_ld_32x32(tcol + 128 + cb * 32, v2);
        tmem_ld_wait();
        reg_fence32(v0); reg_fence32(v1); reg_fence32(v2);
        uint32_t pk[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int c = cb * 32 + 4 * j4;
          float y[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[4 * j4 + e]), 1);       // D_0[h, w-1]
            const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[4 * j4 + e]), 1);    // D_2[h, w+1]
            y[e] = left + __uint_as_float(v1[4 * j4 + e]) + right;
          }
          if (EM == 1) {
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
            y[0] += sh.x; y[1] += sh.y; y[2] += sh.z; y[3] += sh.w;
          } else if (EM >= 2) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
            y[0] = fmaf(y[0], sc.x, sh.x); y[1] = fmaf(y[1], sc.y, sh.y); y[2] = fmaf(y[2], sc.z, sh.z); y[3] = fmaf(y[3], sc.w, sh.w);
          }
          if (p.relu) { y[0] = fmaxf(y[0], 0.f); y[1] = fmaxf(y[1], 0.f); y[2] = fmaxf(y[2], 0.f); y[3] = fmaxf(y[3], 0.f); }
          __nv_bfloat162 ha = __floats2bfloat162_rn(y[0], y[1]), hb = __floats2bfloat162_rn(y[2], y[3]);
          pk[2 * j4] = *reinterpret_cast<uint32_t*>(&ha);
          pk[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&hb);
          if (EM == 3) {
            const float r0 = __low2float(ha), r1 = __high2float(ha), r2 = __low2float(hb), r3 = __high2float(hb);
#pragma unroll
            for (int o = 0; o < kHeadOC; ++o) {
              const float4 hw = *reinterpret_cast<const float4*>(s_hw + o * 64 + c);
              hacc[o] = fmaf(r3, hw.w, fmaf(r2, hw.z, fmaf(r1, hw.y, fmaf(r0, hw.x, hacc[o]))));
            }
          }
        }
        if (p.store_y && valid_col) {
          uint8_t* rowp = sbuf + ridx * 128;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int chunk16 = cb * 4 + jj;
            *reinterpret_cast<uint4*>(rowp + ((chunk16 ^ (ridx & 7)) << 4)) = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
          }
        }
      }
      tc_fence_before();
      if constexpr (STATS) named_bar_arrive(kBarFull + stg, kStatThreads);
      if (p.store_y) fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (et == 0) {
        mbar_arrive(&tmem_empty[buf]);
        if (p.store_y) {
          tma_store_4d(&tmY, sbuf, 0, w0, h0, b);
          tma_commit_group();
        }
      }
      if (p.store_y) stg ^= 1;
      if (EM == 3) {
        const int h = h0 + hh, w = w0 + ww - 1;
        if (valid_col && h < p.H && w < p.W) {
#pragma unroll
          for (int o = 0; o < kHeadOC; ++o)
            if (o < p.head_oc) {
              float v = hacc[o] + p.head_b[o];
              if (p.head_tanh && o == 0) v = tanhf(v);
              p.head_out[(((long long)b * p.head_oc + o) * p.H + h) * p.W + w] = v;
            }
        }
      }
      buf ^= 1;
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

template <int NA, int NB, int EM, bool STATS = false>
int launch_col3(const ConvTcOp& op, cudaStream_t st) {
  using S = V3Smem<NA, NB>;
  static_assert(S::kBytes <= 232448, "shared memory budget exceeded");
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = conv3x3_tc_col3_kernel<NA, NB, EM, STATS>;
  if (!attr_done[dev & 15]) {
    MAU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes));
    attr_done[dev & 15] = true;
  }
  kern<<<op.grid, STATS ? kStatsBlock : kThreads, S::kBytes, st>>>(op.tmA, op.tmB, op.tmY, op.p);
  MAU_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [9][N][Kp] bf16
// fwd : out[t][n][kp] = W[n][kmap[kp]][t]            (kmap = -1 -> 0)
// dgrad: out[t][n][kp] = W[kp][ci0 + n][8 - t]       (kp >= Cout -> 0)
// one warp per (output channel n, 64-wide packed-K chunk): the 64 x 9 source floats of a chunk are contiguous
// in OIHW when the chunk maps to consecutive input channels, so they are read as one coalesced 2304-byte run,
// transposed through shared memory and written as nine 128-byte bf16 rows
__global__ void __launch_bounds__(256) pack_w_fwd_kernel(const float* __restrict__ w, int Cout, int Cin,
                                                         const int* __restrict__ kmap, int Kp,
                                                         __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[8][9][65];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = Kp >> 6;
  const long long total = (long long)Cout * chunks;
  for (long long item = (long long)blockIdx.x * 8 + warp; item < total; item += (long long)gridDim.x * 8) {
    const int n = (int)(item / chunks), kp0 = (int)(item % chunks) << 6;
    const float* src = w + (long long)n * Cin * 9;
    int ci[18];
    float v[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) ci[j] = kmap[kp0 + (lane + 32 * j) / 9];          // 18 index loads in flight ...
#pragma unroll
    for (int j = 0; j < 18; ++j) {                                                 // ... then 18 weight loads in flight
      const int e = lane + 32 * j, t = e - (e / 9) * 9;
      v[j] = ci[j] >= 0 ? src[ci[j] * 9 + t] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      const int e = lane + 32 * j, kl = e / 9;
      tile[warp][e - kl * 9][kl] = v[j];
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const __nv_bfloat162 v = __floats2bfloat162_rn(tile[warp][t][2 * lane], tile[warp][t][2 * lane + 1]);
      *reinterpret_cast<__nv_bfloat162*>(out + ((long long)t * Cout + n) * Kp + kp0 + 2 * lane) = v;
    }
    __syncwarp();
  }
}
__global__ void pack_w_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int ci0, int N, int Kp,
                                    __nv_bfloat16* __restrict__ out) {
  const long long total = 9LL * N * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kp = (int)(i % Kp);
    const long long r = i / Kp;
    const int n = (int)(r % N);
    const int t = (int)(r / N);
    const float v = kp < Cout ? w[((long long)kp * Cin + (ci0 + n)) * 9 + (8 - t)] : 0.f;
    out[i] = __float2bfloat16_rn(v);
  }
}

template <int BN, int MODE, int NA, int NB>
int launch_inst(const ConvTcOp& op, cudaStream_t st) {
  constexpr size_t smem = 1024 + NA * ATile<MODE>::kStage + NB * BN * 128 + 8 * (2 * NA + 2 * NB + 1) + 16 +
                          2 * BN * sizeof(float);
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = conv3x3_tc_kernel<BN, MODE, NA, NB>;
  if (!attr_done[dev & 15]) {
    MAU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev & 15] = true;
  }
  kern<<<op.grid, kThreads, smem, st>>>(op.tmA, op.tmB, op.tmY, op.p);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace

int conv_tc_pick_bn(int Cout) { return Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256); }
// v2 tiles are 64 or 128 channels wide: 128 unless that would leave more than half of the last tile empty
static int pick_bn_v2(int Cout) {
  // cost model: tiles x width / measured MMA efficiency of that instruction width (SMEM operand
  // bandwidth: N = 64 and N = 128 instructions cannot keep the tensor pipe full from one CTA)
  const int bn[4] = {256, 192, 128, 64};
  const double eff[4] = {1.0, 0.9, 0.8, 0.55};
  int best = 256; double best_cost = 1e30;
  for (int i = 0; i < 4; ++i) {
    const double cost = (double)ceil_div(Cout, bn[i]) * bn[i] / eff[i];
    if (cost < best_cost) { best_cost = cost; best = bn[i]; }
  }
  return best;
}

static int prepare_impl(ConvTcOp* op, const View& xbuf, int nseg, const int* seg_start, const int* seg_len,
                        const void* wpacked, int Kp, int n_rows, const View& y, int mode, const float* scale,
                        const float* shift, int relu, int accumulate, int halo_base_offset, int bt, int bt_col0,
                        int bt_kp_fwd);

int conv_tc_prepare(ConvTcOp* op, const View& xbuf, int nseg, const int* seg_start, const int* seg_len,
                    const void* wpacked, int Kp, int n_rows, const View& y, int mode, const float* scale,
                    const float* shift, int relu, int accumulate, int halo_base_offset) {
  return prepare_impl(op, xbuf, nseg, seg_start, seg_len, wpacked, Kp, n_rows, y, mode, scale, shift, relu, accumulate,
                      halo_base_offset, 0, 0, 0);
}

// data gradient of one input segment: dX[:, seg] = conv(dZ, W^T flipped), B operand read MN-major from the forward
// pack wpack_fwd [9][Cout][Kp_fwd]; col0 = packed column of the segment's first input channel
int conv_tc_prepare_dgrad(ConvTcOp* op, const View& dz, const void* wpack_fwd, int Kp_fwd, int col0, const View& gx,
                          int accumulate) {
  const int zero = 0, C = dz.C;
  return prepare_impl(op, dz, 1, &zero, &C, wpack_fwd, round_up(C, 64), C, gx, MODE_HALO, nullptr, nullptr, 0, accumulate,
                      0, 1, col0, Kp_fwd);
}

static int prepare_impl(ConvTcOp* op, const View& xbuf, int nseg, const int* seg_start, const int* seg_len,
                        const void* wpacked, int Kp, int n_rows, const View& y, int mode, const float* scale,
                        const float* shift, int relu, int accumulate, int halo_base_offset, int bt, int bt_col0,
                        int bt_kp_fwd) {
  if (nseg < 1 || nseg > 4) return fail("conv_tc: 1..4 input segments supported, got %d", nseg);
  if (xbuf.cs % 8 || y.cs % 8 || y.c0 % 8 || Kp % 64)
    return fail("conv_tc: channel strides/offsets must be multiples of 8 (cs=%d, ycs=%d, yc0=%d, Kp=%d)", xbuf.cs,
                y.cs, y.c0, Kp);
  if (xbuf.B != y.B || xbuf.H != y.H || xbuf.W != y.W) return fail("conv_tc: input/output geometry mismatch");
  ConvTcParams& p = op->p;
  p = ConvTcParams();
  op->mode = mode;
  op->bn = mode == MODE_HALO ? pick_bn_v2(y.C) : conv_tc_pick_bn(y.C);
  op->mt = op->bn == 64 ? 4 : (op->bn == 128 ? 2 : 1);     // measured best per width (tools/conv_bench.py); 192, 256: 1
  op->nbuf = 2;
  if (mode == MODE_HALO) {
    if (const char* e = getenv("MAU_CONV_CFG")) {      // experiment knob: "bn,mt,nbuf"
      int a = 0, b = 0, c = 0;
      if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3) { op->bn = a; op->mt = b; op->nbuf = c; }
    }
  }
  op->bres = 0;
  op->col3 = 0;
  // (measured: with a single K chunk the three-accumulator epilogue -- 3 tcgen05.ld + 2 shuffles per output -- costs
  //  more than the saved operand traffic, 104 vs 96 us on the 64 -> 64 layers; from two chunks up col3 wins, 199 vs 240 us
  //  at K = 192)
  if (mode == MODE_HALO && !bt && y.C == 64 && Kp >= 128 && !accumulate && !getenv("MAU_CONV_CFG") && !getenv("MAU_NO_COL3")) {
    op->col3 = 1; op->bn = 64; op->mt = 1; op->nbuf = 2;
  } else if (mode == MODE_HALO && op->bn == 64 && y.C <= 64 && Kp == 64 && !getenv("MAU_CONV_CFG") && !getenv("MAU_NO_BRES")) {
    op->bres = 1; op->mt = 2; op->nbuf = 2;
  }
  p.TW = (mode == MODE_TAP) ? 16 : 8;
  p.TH = 128 / p.TW;
  p.tiles_w = ceil_div(y.W, p.TW);
  p.tiles_h = ceil_div(y.H, p.TH);
  if (op->col3) { p.TW = 14; p.TH = 8; p.tiles_w = ceil_div(y.W, 14); p.tiles_h = ceil_div(y.H, 8); }
  p.Cout = y.C;
  p.nseg = nseg;
  int chunks = 0;
  for (int i = 0; i < nseg; ++i) {
    if (seg_start[i] % 8) return fail("conv_tc: segment start %d not a multiple of 8", seg_start[i]);
    p.seg_start[i] = seg_start[i];
    p.seg_chunks[i] = ceil_div(seg_len[i], 64);
    chunks += p.seg_chunks[i];
  }
  if (chunks * 64 != Kp) return fail("conv_tc: packed K %d does not match segments (%d chunks)", Kp, chunks);
  p.relu = relu;
  p.accumulate = accumulate;
  p.halo_base_offset = halo_base_offset;
  p.scale = scale;
  p.shift = shift;
  p.H = y.H; p.W = y.W;
  p.total_ptiles = y.B * p.tiles_w * p.tiles_h;
  op->grid = dim3((unsigned)p.total_ptiles, (unsigned)ceil_div(y.C, op->bn), 1);
  if (mode == MODE_HALO) {   // persistent: one CTA per SM (or fewer when there is less work)
    const int items = ceil_div(p.total_ptiles, op->mt) * ceil_div(y.C, op->bn);
    op->grid = dim3((unsigned)std::min(items, sm_budget()), 1, 1);
  }
  // A: whole input buffer {C, W, H, B}
  View xa = xbuf;
  int bw = p.TW, bh = p.TH;
  if (mode == MODE_ROW3) { bw = 8; bh = p.TH + 2; }
  if (mode == MODE_HALO) { bw = 10; bh = p.TH + 2; }
  if (op->col3) { bw = 16; bh = 10; }
  MAU_TRY(make_nhwc_map(&op->tmA, DT_BF16, xa, 64, bw, bh));
  // B: packed weights {Kp, n_rows, 9}; dgrad (bt) reads {64 ci, 64 co} boxes of the forward pack instead
  p.bt = bt; p.bt_col0 = bt_col0;
  if (bt) {
    uint64_t dims[3] = {(uint64_t)bt_kp_fwd, (uint64_t)n_rows, 9};
    uint64_t str[2] = {(uint64_t)bt_kp_fwd * 2, (uint64_t)bt_kp_fwd * 2 * (uint64_t)n_rows};
    uint32_t box[3] = {64, 64, 1};
    MAU_TRY(make_tensor_map(&op->tmB, DT_BF16, 3, const_cast<void*>(wpacked), dims, str, box, true));
  } else if (op->col3) {     // [9][64][Kp] == [3][192][Kp]: one box per filter row holds its three horizontal taps
    uint64_t dims[3] = {(uint64_t)Kp, 192, 3};
    uint64_t str[2] = {(uint64_t)Kp * 2, (uint64_t)Kp * 2 * 192};
    uint32_t box[3] = {64, 192, 1};
    MAU_TRY(make_tensor_map(&op->tmB, DT_BF16, 3, const_cast<void*>(wpacked), dims, str, box, true));
  } else {
    uint64_t dims[3] = {(uint64_t)Kp, (uint64_t)n_rows, 9};
    uint64_t str[2] = {(uint64_t)Kp * 2, (uint64_t)Kp * 2 * (uint64_t)n_rows};
    uint32_t box[3] = {64, (uint32_t)op->bn, 1};
    MAU_TRY(make_tensor_map(&op->tmB, DT_BF16, 3, const_cast<void*>(wpacked), dims, str, box, true));
  }
  // Y: the output *view* {Cout, W, H, B} (channel clipping at the view end, not the buffer end)
  MAU_TRY(make_nhwc_map(&op->tmY, DT_BF16, y, 64, p.TW, p.TH));
  return 0;
}

int conv_tc_launch(const ConvTcOp& op, cudaStream_t st) {
#define MAU_INST(BN_, MODE_, NA_, NB_) \
  if (op.bn == BN_ && op.mode == MODE_) return launch_inst<BN_, MODE_, NA_, NB_>(op, st);
  MAU_INST(64, MODE_TAP, 4, 4)
  MAU_INST(128, MODE_TAP, 3, 3)
  MAU_INST(256, MODE_TAP, 4, 4)
  MAU_INST(64, MODE_ROW3, 2, 6)
  MAU_INST(128, MODE_ROW3, 2, 4)
  MAU_INST(256, MODE_ROW3, 2, 3)
#undef MAU_INST
  if (op.mode == MODE_HALO) {
    const int em = op.p.head_out ? 3 : (op.p.scale ? 2 : (op.p.shift ? 1 : 0));
    if (op.p.bt && em != 0) return fail("conv_tc: the MN-major weight path has no epilogue affine");
    if (em == 3 && op.bn > 128) return fail("conv_tc: fused head needs BN <= 128");
#define MAU_V2(BN_, MT_, NBUF_, NA_, NB_, NSTG_)                                                        \
    if (op.bn == BN_ && op.mt == MT_ && op.nbuf == NBUF_) {                                             \
      if (op.p.bt) return launch_v2<BN_, MT_, NBUF_, NA_, NB_, NSTG_, true, 0>(op, st);                 \
      if (em == 0) return launch_v2<BN_, MT_, NBUF_, NA_, NB_, NSTG_, false, 0>(op, st);                \
      if (em == 1 && op.p.stats) return launch_v2<BN_, MT_, NBUF_, NA_, NB_, NSTG_, false, 1, false, true>(op, st); \
      if (em == 1) return launch_v2<BN_, MT_, NBUF_, NA_, NB_, NSTG_, false, 1>(op, st);                \
      if (em == 2) return launch_v2<BN_, MT_, NBUF_, NA_, NB_, NSTG_, false, 2>(op, st);                \
      if constexpr (BN_ <= 128) return launch_v2<BN_, MT_, NBUF_, NA_, NB_, NSTG_, false, 3>(op, st);   \
    }
    if (op.col3) {          // Cout == 64: three horizontal taps folded into N = 192
      if (em == 0) return launch_col3<3, 4, 0>(op, st);
      if (em == 1 && op.p.stats) return launch_col3<3, 4, 1, true>(op, st);
      if (em == 1) return launch_col3<3, 4, 1>(op, st);
      if (em == 2) return launch_col3<3, 4, 2>(op, st);
      return launch_col3<3, 4, 3>(op, st);
    }
    if (op.bres) {          // 64-wide, K = 64: weights resident in 9 B stages, MT = 2
      if (op.p.bt) return launch_v2<64, 2, 2, 2, 9, 2, true, 0, true>(op, st);
      if (em == 0) return launch_v2<64, 2, 2, 2, 9, 2, false, 0, true>(op, st);
      if (em == 1 && op.p.stats) return launch_v2<64, 2, 2, 2, 9, 2, false, 1, true, true>(op, st);
      if (em == 1) return launch_v2<64, 2, 2, 2, 9, 2, false, 1, true>(op, st);
      if (em == 2) return launch_v2<64, 2, 2, 2, 9, 2, false, 2, true>(op, st);
      return launch_v2<64, 2, 2, 2, 9, 2, false, 3, true>(op, st);
    }
    MAU_V2(64, 4, 2, 2, 3, 1)
    MAU_V2(128, 2, 2, 2, 4, 2)
    MAU_V2(192, 1, 2, 2, 4, 2)
    MAU_V2(256, 1, 2, 2, 4, 1)
#undef MAU_V2
    return fail("conv_tc: no v2 kernel instance for BN=%d MT=%d NBUF=%d", op.bn, op.mt, op.nbuf);
  }
  return fail("conv_tc: no kernel instance for BN=%d mode=%d", op.bn, op.mode);
}

int conv_tc_pack_fwd(const float* w_oihw, int Cout, int Cin, const int* kmap_dev, int Kp, void* out,
                     cudaStream_t st) {
  const long long items = (long long)Cout * (Kp / 64);
  int blocks = (int)((items + 7) / 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_w_fwd_kernel<<<blocks, 256, 0, st>>>(w_oihw, Cout, Cin, kmap_dev, Kp, static_cast<__nv_bfloat16*>(out));
  MAU_LAUNCHED();
  return 0;
}
int conv_tc_pack_dgrad(const float* w_oihw, int Cout, int Cin, int ci0, int N, int Kp, void* out,
                       cudaStream_t st) {
  const long long total = 9LL * N * Kp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_w_dgrad_kernel<<<blocks, 256, 0, st>>>(w_oihw, Cout, Cin, ci0, N, Kp, static_cast<__nv_bfloat16*>(out));
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
