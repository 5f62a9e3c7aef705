// Error string + launch counter shared by all nrgbd translation units.
#include <atomic>
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void nrgbd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {
const char* nrgbd_last_error(void) { return g_err; }
void nrgbd_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long nrgbd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void nrgbd_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
int nrgbd_abi_version(void) { return 1; }
}
