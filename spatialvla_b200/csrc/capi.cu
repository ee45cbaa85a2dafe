// Library-wide state of the C ABI: last-error text, launch counter, ABI version.
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "svla_common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

bool svla_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SVLA_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

void svla_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void svla_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* svla_last_error(void) { return g_err; }
extern "C" int svla_abi_version(void) { return SVLA_ABI_VERSION; }
extern "C" long long svla_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
