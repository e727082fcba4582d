// optim.cu -- fused multi-tensor AdamW step (decoupled weight decay), one launch for every parameter.
//
// Replaces: torch.optim.AdamW(...).step() of the reference training loop (src/train.py:213-214,255), which
// issues several element-wise kernels per parameter tensor (138 tensors).  Same arithmetic, same order:
//   p *= 1 - lr*wd;  m += (g - m)*(1 - b1);  v = b2*v + (1 - b2)*g*g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: reads p, g, m, v and writes p, m, v once (28 B per parameter).
#include "ops.h"
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace mau {
namespace {

constexpr int kChunk = 4096;       // elements per block-iteration
struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; };
struct AdamChunk { int tensor; int chunk; };

__global__ void __launch_bounds__(256) adamw_kernel(const AdamTensor* __restrict__ tensors,
                                                    const AdamChunk* __restrict__ chunks, int n_chunks, float decay,
                                                    float one_minus_b1, float b2, float one_minus_b2, float step_size,
                                                    float inv_bc2_sqrt, float eps) {
  for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    const AdamChunk ch = chunks[ci];
    const AdamTensor t = tensors[ch.tensor];
    const long long base = (long long)ch.chunk * kChunk;
    const int n = (int)min((long long)kChunk, t.n - base);
    float* p = t.p + base; const float* g = t.g + base; float* m = t.m + base; float* v = t.v + base;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
      pp *= decay;
      mm = mm + (gg - mm) * one_minus_b1;
      vv = vv * b2 + one_minus_b2 * gg * gg;
      const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
      pp -= step_size * (mm / denom);
    };
    if (vec) {
      const int n4 = n >> 2;
      for (int i = threadIdx.x; i < n4; i += 256 * 4) {      // 4 independent float4 quads in flight per tensor
        float4 P[4], G[4], M[4], V[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = i + u * 256;
          if (j < n4) {
            P[u] = reinterpret_cast<float4*>(p)[j]; G[u] = reinterpret_cast<const float4*>(g)[j];
            M[u] = reinterpret_cast<float4*>(m)[j]; V[u] = reinterpret_cast<float4*>(v)[j];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = i + u * 256;
          if (j < n4) {
            upd(P[u].x, G[u].x, M[u].x, V[u].x); upd(P[u].y, G[u].y, M[u].y, V[u].y);
            upd(P[u].z, G[u].z, M[u].z, V[u].z); upd(P[u].w, G[u].w, M[u].w, V[u].w);
            reinterpret_cast<float4*>(p)[j] = P[u]; reinterpret_cast<float4*>(m)[j] = M[u];
            reinterpret_cast<float4*>(v)[j] = V[u];
          }
        }
      }
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) upd(p[i], g[i], m[i], v[i]);
    } else {
      for (int i = threadIdx.x; i < n; i += 256) upd(p[i], g[i], m[i], v[i]);
    }
  }
}

struct Scratch { void* dev = nullptr; size_t bytes = 0; };

// gradient bucket <-> bf16 wire format of the data-parallel all-reduce (parallel.py, grad_dtype="bf16")
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) | (reinterpret_cast<uintptr_t>(dst) & 7)) == 0;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (vec) {
    for (long long i = t0; i < n4; i += stride) {
      const float4 v = reinterpret_cast<const float4*>(src)[i];
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      reinterpret_cast<uint2*>(dst)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
    for (long long i = (n4 << 2) + t0; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
  } else {
    for (long long i = t0; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
  }
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) | (reinterpret_cast<uintptr_t>(src) & 7)) == 0;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (vec) {
    for (long long i = t0; i < n4; i += stride) {
      const uint2 u = reinterpret_cast<const uint2*>(src)[i];
      reinterpret_cast<float4*>(dst)[i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                                                      __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    }
    for (long long i = (n4 << 2) + t0; i < n; i += stride) dst[i] = __bfloat162float(src[i]);
  } else {
    for (long long i = t0; i < n; i += stride) dst[i] = __bfloat162float(src[i]);
  }
}

}  // namespace

int op_cast_f32_bf16(const float* src, void* dst_bf16, long long n, int max_blocks, cudaStream_t st) {
  if (n <= 0) return 0;
  const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, max_blocks > 0 ? max_blocks : 148 * 4);
  cast_f32_bf16_kernel<<<blocks, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst_bf16), n);
  MAU_LAUNCHED();
  return 0;
}
int op_cast_bf16_f32(const void* src_bf16, float* dst, long long n, int max_blocks, cudaStream_t st) {
  if (n <= 0) return 0;
  const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, max_blocks > 0 ? max_blocks : 148 * 4);
  cast_bf16_f32_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src_bf16), dst, n);
  MAU_LAUNCHED();
  return 0;
}

int op_adamw_step(int n_tensors, void* const* params, void* const* grads, void* const* exp_avg,
                  void* const* exp_avg_sq, const long long* numels, double lr, double beta1, double beta2, double eps,
                  double weight_decay, long long step, cudaStream_t st) {
  if (n_tensors <= 0) return 0;
  if (step < 1) return fail("adamw: step must be >= 1");
  std::vector<AdamTensor> ts(n_tensors);
  std::vector<AdamChunk> chunks;
  for (int i = 0; i < n_tensors; ++i) {
    if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i]) return fail("adamw: null tensor %d", i);
    ts[i] = AdamTensor{static_cast<float*>(params[i]), static_cast<const float*>(grads[i]),
                       static_cast<float*>(exp_avg[i]), static_cast<float*>(exp_avg_sq[i]), numels[i]};
    const int nc = (int)((numels[i] + kChunk - 1) / kChunk);
    for (int c = 0; c < nc; ++c) chunks.push_back(AdamChunk{i, c});
  }
  if (chunks.empty()) return 0;
  const size_t tb = (sizeof(AdamTensor) * ts.size() + 255) & ~size_t(255);
  const size_t need = tb + sizeof(AdamChunk) * chunks.size();
  // per-(device, stream) table scratch: copies and launches are ordered on the stream, so it can be reused
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, Scratch> pool;
  int dev = 0;
  cudaGetDevice(&dev);
  void* dptr = nullptr;
  {
    std::lock_guard<std::mutex> lock(mu);
    Scratch& s = pool[{dev, st}];
    if (s.bytes < need) {
      if (s.dev) { cudaStreamSynchronize(st); cudaFree(s.dev); }
      s.dev = nullptr; s.bytes = 0;
      if (cudaMalloc(&s.dev, need * 2) != cudaSuccess) return fail("adamw: table allocation failed");
      s.bytes = need * 2;
    }
    dptr = s.dev;
  }
  // pageable-host async copies are staged by the runtime before the call returns
  MAU_CUDA(cudaMemcpyAsync(dptr, ts.data(), sizeof(AdamTensor) * ts.size(), cudaMemcpyHostToDevice, st));
  MAU_CUDA(cudaMemcpyAsync(static_cast<char*>(dptr) + tb, chunks.data(), sizeof(AdamChunk) * chunks.size(),
                           cudaMemcpyHostToDevice, st));
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const int blocks = (int)std::min<size_t>(chunks.size(), 148 * 8);
  adamw_kernel<<<blocks, 256, 0, st>>>(static_cast<const AdamTensor*>(dptr),
                                       reinterpret_cast<const AdamChunk*>(static_cast<char*>(dptr) + tb),
                                       (int)chunks.size(), (float)(1.0 - lr * weight_decay), (float)(1.0 - beta1),
                                       (float)beta2, (float)(1.0 - beta2), (float)(lr / bc1), (float)(1.0 / sqrt(bc2)),
                                       (float)eps);
  MAU_LAUNCHED();
  return 0;
}

}  // namespace mau
