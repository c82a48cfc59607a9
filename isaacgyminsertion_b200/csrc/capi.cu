// Library-wide pieces of the C-ABI: version and the thread-local error string.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "igi_common.cuh"
#include "../../include/igi_b200.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void igi_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void igi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int igi_version(void) { return IGI_B200_VERSION; }
extern "C" const char* igi_last_error(void) { return g_err; }
extern "C" long long igi_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
