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

}  // namespace dvsr

extern "C" const char* dvsr_last_error(void) { return dvsr::g_err; }
extern "C" int dvsr_version(void) { return 100; }
