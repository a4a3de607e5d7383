// Error reporting shared by every entry point of the C-ABI.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace ngacf {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return NGACF_ERR_CUDA;
    }
    return NGACF_OK;
}
}  // namespace ngacf

extern "C" const char* ngacf_last_error(void) { return ngacf::g_err; }
extern "C" int ngacf_version(void) { return 100; }
