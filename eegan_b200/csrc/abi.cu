// Error channel + version of the C ABI (include/eegan_b200.h).
#include <stdlib.h>

#include "common.cuh"

namespace eegan {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("EEGAN_PDL");
        return e ? atoi(e) != 0 : true;
    }();
    return on;
}
int num_sms_current() {
    static std::atomic<int> cache[EEGAN_MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    if (dev < EEGAN_MAX_DEVICES) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int v = 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    if (dev < EEGAN_MAX_DEVICES) cache[dev].store(v, std::memory_order_relaxed);
    return v;
}
}  // namespace eegan

extern "C" int eegan_abi_version(void) { return EEGAN_B200_ABI_VERSION; }
extern "C" const char* eegan_last_error(void) { return eegan::g_err; }
