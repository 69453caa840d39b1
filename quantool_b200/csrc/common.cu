#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace qt {
static thread_local char g_err[512] = "";

void set_last_error(const char* what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error(what, e);
        return QT_ERR_CUDA;
    }
    return QT_OK;
}
const char* last_error() { return g_err; }
}  // namespace qt

extern "C" {
const char* qt_last_error(void) { return qt::last_error(); }
int qt_abi_version(void) { return 1; }
int qt_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
}
