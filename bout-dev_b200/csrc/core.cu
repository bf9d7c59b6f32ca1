// Library-wide state: per-thread error message, launch counter, version.
#include "common.cuh"

#include <cstdarg>

namespace b200 {

std::string& lastError() {
  static thread_local std::string msg;
  return msg;
}

int fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  lastError() = buf;
  return status;
}

std::atomic<unsigned long long> g_launches{0};

} // namespace b200

extern "C" const char* bout_b200_last_error(void) { return b200::lastError().c_str(); }
extern "C" int bout_b200_version(void) { return 100; }
extern "C" unsigned long long bout_b200_launch_count(void) {
  return b200::g_launches.load(std::memory_order_relaxed);
}
