// Error reporting and version entry points of the C ABI.
#include "common.cuh"
#include <cstring>

namespace pb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

}  // namespace pb

extern "C" int pb_version(void) { return 100; }

extern "C" const char* pb_last_error(void) { return pb::g_error; }
