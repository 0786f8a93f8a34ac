// libm4d: error reporting, ABI version, launch counter.
#include "common.cuh"

static thread_local char t_err[512] = "no error";
std::atomic<uint64_t> g_m4d_launches{0};

void m4d_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

extern "C" {
int m4d_abi_version(void) { return M4D_ABI_VERSION; }
const char* m4d_last_error_string(void) { return t_err; }
uint64_t m4d_launch_count(void) { return g_m4d_launches.load(std::memory_order_relaxed); }
}
