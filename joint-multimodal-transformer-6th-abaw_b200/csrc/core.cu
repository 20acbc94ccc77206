// Library-wide state: error string (thread-local), launch counter, ABI version.
#include "common.cuh"

namespace jmt {
std::atomic<int64_t> g_launch_count{0};
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace jmt

extern "C" int jmt_abi_version(void) { return JMT_ABI_VERSION; }
extern "C" const char* jmt_last_error(void) { return jmt::g_err; }
extern "C" int64_t jmt_launch_count(void) { return jmt::g_launch_count.load(std::memory_order_relaxed); }
