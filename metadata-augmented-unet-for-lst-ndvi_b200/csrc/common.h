// common.h -- shared host-side helpers of libmau_b200 (error plumbing, views, launch count)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <atomic>

namespace mau {

enum DType : int { DT_BF16 = 0, DT_F32 = 1 };
inline size_t dtype_size(int dt) { return dt == DT_BF16 ? 2 : 4; }

// ---- error plumbing: nothing throws across the C ABI -------------------------------------
std::string& last_error();                    // thread local
int fail(const char* fmt, ...);               // records message, returns -1
extern std::atomic<long long> g_launches;     // kernels launched by this library

#define MAU_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::mau::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define MAU_TRY(expr)                 \
  do {                                \
    int _rc = (expr);                 \
    if (_rc != 0) return _rc;         \
  } while (0)

// check the launch that just happened and count it
#define MAU_LAUNCHED()                                                                   \
  do {                                                                                   \
    ::mau::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ::mau::fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ---- NHWC channel-slice view ---------------------------------------------------------------
// element (b,h,w,c) lives at ptr[((b*H + h)*W + w)*cs + c0 + c]   (ptr is the buffer base)
struct View {
  void* ptr = nullptr;
  int B = 0, H = 0, W = 0;
  int cs = 0;   // channel stride of the underlying buffer (elements, multiple of 8)
  int c0 = 0;   // first channel of this view
  int C = 0;    // channels in this view
  long long pixels() const { return (long long)B * H * W; }
};

bool whole_waves_enabled();   // grids of equal-work element-wise kernels rounded down to whole waves (MAU_WHOLE_WAVES=0: off)
int sm_budget();              // SMs persistent kernels may occupy (device SM count - reserve)
void set_sm_reserve(int n);   // SMs left free for concurrent collective kernels (data-parallel training)
void set_sm_reserve_override(int n);   // per-thread scope: >= 0 replaces the global reserve, -1 ends the scope

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace mau
