// Library-level entry points: last-error text, version, launch counter.
#include <atomic>
#include "common.cuh"
#include "xmlb200.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void xmlb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void xmlb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* xmlb_last_error(void) { return g_err; }
extern "C" int xmlb_version(void) { return 100; }
extern "C" long long xmlb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
