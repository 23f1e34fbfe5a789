// Library-wide state: thread-local error text, launch counter, version.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace ppnet {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ppnet

extern "C" const char* ppnet_last_error(void) { return ppnet::g_err; }
extern "C" int ppnet_version(void) { return 100; }
extern "C" int64_t ppnet_launch_count(void) { return ppnet::g_launches.load(); }

// struct layout guard for ctypes / cgo / JNI mirrors: 0 -> sizeof(ppnet_gen_params), 1 -> sizeof(ppnet_path_params),
// 2 -> sizeof(ppnet_pipeline_io)
extern "C" int64_t ppnet_sizeof_params(int32_t which) {
    return which == 0 ? (int64_t)sizeof(ppnet_gen_params) : which == 1 ? (int64_t)sizeof(ppnet_path_params)
           : which == 2 ? (int64_t)sizeof(ppnet_pipeline_io) : -1;
}
