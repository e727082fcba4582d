#include "common.h"
#include <cstdlib>
namespace mau {
std::atomic<long long> g_launches{0};
std::string& last_error() {
  thread_local std::string e;
  return e;
}
int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return -1;
}

// SMs the persistent kernels of this library size their grids for.  MAU_SM_RESERVE (or mau_set_sm_reserve)
// leaves SMs free for a concurrently running collective kernel: a persistent grid that does not fit
// next to it would otherwise run its last CTAs as a second wave.
static std::atomic<int> g_sm_reserve{-1};
static thread_local int t_reserve_override = -1;     // >= 0: used instead of the global reserve (plan build scopes)
void set_sm_reserve_override(int n) { t_reserve_override = n; }
void set_sm_reserve(int n) { g_sm_reserve.store(n < 0 ? 0 : n); }
// MAU_WHOLE_WAVES=0 keeps the grids of the equal-work element-wise kernels as they were before they were rounded
// down to whole waves of resident blocks (A/B measurements)
bool whole_waves_enabled() {
  static const bool on = [] {
    const char* e = getenv("MAU_WHOLE_WAVES");
    return !(e && e[0] == '0');
  }();
  return on;
}
int sm_budget() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (t_reserve_override >= 0) return sms - t_reserve_override > 8 ? sms - t_reserve_override : 8;
  int r = g_sm_reserve.load();
  if (r < 0) {
    const char* e = getenv("MAU_SM_RESERVE");
    r = e ? atoi(e) : 0;
    if (r < 0) r = 0;
    g_sm_reserve.store(r);
  }
  return sms - r > 8 ? sms - r : 8;
}
}  // namespace mau
