// ABI plumbing: version, thread-local error text, launch checking.
#include <cstdarg>
#include <cstring>
#include "common.cuh"

namespace aegis {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}
}  // namespace aegis

extern "C" int aegis_abi_version(void) { return AEGIS_ABI_VERSION; }
extern "C" const char* aegis_last_error(void) { return aegis::g_error; }
extern "C" int aegis_device_sm_count(void) { return aegis::sm_count(); }
