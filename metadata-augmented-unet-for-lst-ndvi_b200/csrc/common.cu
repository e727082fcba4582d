#include "common.h"
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
}  // namespace mau
