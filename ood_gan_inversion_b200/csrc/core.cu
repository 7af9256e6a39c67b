// Error state, version and device probe of libood_b200.
#include "common.cuh"

namespace ood {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;

int check_launch(const char *what, int kernels) {
    g_launches += (unsigned long long)kernels;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return OOD_ERR_CUDA;
    }
    return OOD_OK;
}

}  // namespace ood

extern "C" int ood_version(void) { return 100; }

extern "C" unsigned long long ood_launch_count(void) { return ood::g_launches; }

namespace ood { thread_local int g_conv_route = 0; }
extern "C" int ood_last_conv_route(void) { return ood::g_conv_route; }

extern "C" const char *ood_last_error(void) { return ood::g_err; }

extern "C" int ood_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}
