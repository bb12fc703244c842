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
}  // namespace eegan

extern "C" int eegan_abi_version(void) { return EEGAN_B200_ABI_VERSION; }
extern "C" const char* eegan_last_error(void) { return eegan::g_err; }
