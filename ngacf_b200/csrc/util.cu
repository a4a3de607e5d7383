// Error reporting shared by every entry point of the C-ABI.
#include <stdarg.h>
#include <stdlib.h>
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

bool dense_on_tensor_cores() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("NGACF_DENSE");
        mode = (e && strcmp(e, "ffma") == 0) ? 0 : 1;
    }
    return mode == 1;
}
}  // namespace ngacf

extern "C" const char* ngacf_last_error(void) { return ngacf::g_err; }
extern "C" int ngacf_version(void) { return 100; }
