// oracle/cuda_emu.h -- TEST INFRASTRUCTURE ONLY: enough of the CUDA execution model to run a simple kernel source file
// on the CPU, so that a kernel written without access to a GPU can have its indexing, guards and block reductions checked
// against the oracle before its first hardware run.  A launch runs the blocks of the grid one after another; every block
// is blockDim real threads (std::thread) with thread-local threadIdx / blockIdx, __syncthreads() is a std::barrier,
// __shared__ storage is a function-local static (shared by the threads of the one block in flight), atomicAdd on double
// takes a lock.  Used by a small .cpp under oracle/ that includes this header and then the .cu file itself
// (see oracle/ssim_kernels_emu.cpp), compiled with g++ -std=c++20.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu {
  unsigned x = 0, y = 0, z = 0;
};
inline thread_local uint3_emu threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
typedef void* cudaStream_t;
typedef int cudaError_t;

inline std::barrier<>*& emu_barrier() {
  static std::barrier<>* b = nullptr;
  return b;
}
inline void __syncthreads() { emu_barrier()->arrive_and_wait(); }
inline double atomicAdd(double* p, double v) {
  static std::mutex m;
  std::lock_guard<std::mutex> g(m);
  double old = *p;
  *p += v;
  return old;
}
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return 0;
}

template <typename Kernel, typename... Args>
void emu_launch(Kernel kernel, dim3 grid, dim3 block, Args... args) {
  gridDim = grid;
  blockDim = block;
  const unsigned nthreads = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nthreads);
        emu_barrier() = &bar;
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          th.emplace_back([&, t] {
            threadIdx.x = t % block.x;
            threadIdx.y = (t / block.x) % block.y;
            threadIdx.z = t / (block.x * block.y);
            blockIdx.x = bx;
            blockIdx.y = by;
            blockIdx.z = bz;
            kernel(args...);
            bar.arrive_and_drop();   // a thread that returned early no longer takes part in later barriers
          });
        for (auto& x : th) x.join();
      }
}

// the slice of csrc/ops.h + common.h a simple kernel file uses
namespace mau {
inline int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  return -1;
}
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
}  // namespace mau
#define MAU_KERNEL_ENV 1
#define MAU_CUDA(expr) (void)(expr)
#define MAU_LAUNCHED() (void)0
#define MAU_LAUNCH(kernel, grid, block, stream, ...) emu_launch(kernel, grid, block, __VA_ARGS__)
