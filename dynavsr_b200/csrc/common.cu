// dynavsr_b200/csrc/common.cu -- error plumbing shared by every C-ABI entry point.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dvsr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// The reference only printf()s launch failures and carries on (deform_conv_cuda_kernel.cu:793-797);
// here a failed launch is an error code the Python side turns into a RuntimeError.
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
        return DVSR_ERR_CUDA;
    }
    return DVSR_OK;
}

// SMs of the current device, queried once per device (nothing assumes 148)
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
// CTAs one launch of a persistent kernel may use: the call's policy, clamped to the device (0 = every SM)
int cta_budget(const dvsr_policy& pol) {
    const int n = sm_count();
    return (pol.cta_budget < 1 || pol.cta_budget > n) ? n : pol.cta_budget;
}

}  // namespace dvsr

extern "C" int dvsr_sm_count(void) { return dvsr::sm_count(); }
extern "C" const char* dvsr_last_error(void) { return dvsr::g_err; }
extern "C" int dvsr_version(void) { return 100; }
