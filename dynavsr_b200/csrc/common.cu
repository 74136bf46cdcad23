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

static int g_cta_budget = 148;
int cta_budget() { return g_cta_budget; }

}  // namespace dvsr

// SM budget of ONE launch of the persistent kernels (conv_tc2, conv_wgrad_tc, mdcn_tc): with P frames in flight on P
// streams (adapt.AdaptationPool) every launch is held to ~148 / P CTAs, so the pipelines run side by side instead of a
// 148-CTA launch of one frame holding up the 30-CTA launches of the others.  148 (default) = the whole GPU.
extern "C" int dvsr_set_cta_budget(int n) { dvsr::g_cta_budget = n < 1 ? 1 : (n > 148 ? 148 : n); return 0; }
extern "C" int dvsr_get_cta_budget(void) { return dvsr::g_cta_budget; }
extern "C" const char* dvsr_last_error(void) { return dvsr::g_err; }
extern "C" int dvsr_version(void) { return 100; }
