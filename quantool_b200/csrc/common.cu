#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace qt {
static thread_local char g_err[512] = "";

void set_last_error(const char* what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
}
static unsigned long long g_launches = 0;
unsigned long long launch_count() { return g_launches; }
int check_launch(const char* what) {
    __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
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
int qt_abi_version(void) { return 2; }
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
unsigned long long qt_launch_count(void) { return qt::launch_count(); }
int qt_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
}
