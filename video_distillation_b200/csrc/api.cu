// Error reporting and bookkeeping shared by every entry point of libvd_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vd {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace vd

extern "C" const char* vd_last_error(void) { return vd::g_err; }
extern "C" int vd_abi_version(void) { return 1; }
extern "C" int64_t vd_launch_count(void) { return vd::g_launches.load(); }
extern "C" void vd_launch_count_reset(void) { vd::g_launches.store(0); }
