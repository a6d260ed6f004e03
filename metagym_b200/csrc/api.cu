// Library-wide entry points of libmgb200: error string, version, device probe.
#include <string.h>

#include "mgb_common.cuh"

static thread_local char g_err[512] = "";

void mgb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *mgb_last_error(void) { return g_err; }

extern "C" const char *mgb_version(void) { return "metagym_b200 0.1 (sm_100a)"; }

extern "C" int mgb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        mgb_set_error("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
        cudaGetLastError();
        return MGB_ERR_CUDA;
    }
    return n;
}
