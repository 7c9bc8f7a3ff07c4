// Error string, version, launch counter.
#include "common.cuh"
#include <stdarg.h>

namespace viai {
static thread_local char t_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_sweep_end{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
}  // namespace viai

extern "C" const char* viai_last_error(void) { return viai::t_err; }
extern "C" int viai_version(void) { return 100; }
extern "C" long long viai_launch_count(void) { return viai::g_launches.load(); }
