// Library-level entry points: version, last error, device check.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

static thread_local char g_last_error[512] = "";

void spnet_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int spnet_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        spnet_set_error("%s: %s", what, cudaGetErrorString(e));
        return SPNET_ERR_CUDA;
    }
    return SPNET_OK;
}

bool spnet_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        // measured on B200 inside the captured step graph: 9.94 ms with PDL edges vs 9.81 ms without (round 1); round 2,
        // no early trigger: 7.23 vs 7.26 ms (common.cuh)
        // (kernel boundaries in a graph are already ~1 us), so it is opt-in
        const char* e = getenv("SPNET_B200_PDL");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return v != 0;
}

extern "C" {

int spnet_version(void) { return 100; }

const char* spnet_last_error(void) { return g_last_error; }

// Returns 0 when the current device is compute capability 10.x (the only
// architecture this library carries code for), SPNET_ERR_ARCH otherwise.
int spnet_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        spnet_set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return SPNET_ERR_CUDA;
    }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        spnet_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return SPNET_ERR_CUDA;
    }
    if (p.major != 10) {
        spnet_set_error("spnet_b200 is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
        return SPNET_ERR_ARCH;
    }
    return SPNET_OK;
}

}  // extern "C"
