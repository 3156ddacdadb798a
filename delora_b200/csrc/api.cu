// Error channel + ABI version of the C ABI (include/delora_b200.h).
#include <stdarg.h>
#include "common.cuh"

namespace delora {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace delora

extern "C" int delora_abi_version(void) { return DELORA_B200_ABI_VERSION; }
extern "C" const char* delora_last_error(void) { return delora::g_error; }
