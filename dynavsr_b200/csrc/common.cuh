// dynavsr_b200/csrc/common.cuh -- shared helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dvsr_b200.h"

namespace dvsr {

// ---- error plumbing: every extern "C" entry returns 0 or a negative code; message via dvsr_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError() -> code
int sm_count();                       // SMs of the current device (cached query)
int cta_budget(const dvsr_policy& pol);   // CTAs one launch of a persistent kernel may use (policy, clamped to the device)

#define DVSR_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            dvsr::set_error(__VA_ARGS__);       \
            return DVSR_ERR_INVALID;            \
        }                                       \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers
__device__ __forceinline__ float act_apply(float v, int act, float slope) {
    if (act == DVSR_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == DVSR_ACT_LRELU) return v > 0.f ? v : v * slope;
    return v;
}
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// 256-bit read-only load (LDG.E.256.CONSTANT, sm_100): p must be 32-byte aligned.  One request per 8-channel deformable
// group corner instead of two -- the DCN gather is bound by L1 wavefronts, not by bytes.
__device__ __forceinline__ void ldg8(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One modulated bilinear tap (reference semantics: deform_conv_cuda_kernel.cu:466-496 and :617):
// sample position (h, w) in an H x W image; corners outside contribute 0; the whole tap is 0
// unless -1 < h < H and -1 < w < W.
struct BilinTap {
    int o00, o01, o10, o11;   // pixel indices (h*W+w) of the 4 corners, -1 if that corner is outside
    float w00, w01, w10, w11; // bilinear weights
    float lh, lw;             // fractional parts
    bool inside;
};
__device__ __forceinline__ BilinTap make_tap(float h, float w, int H, int W) {
    BilinTap t;
    t.inside = (h > -1.f) && (w > -1.f) && (h < (float)H) && (w < (float)W);
    t.o00 = t.o01 = t.o10 = t.o11 = -1;
    t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
    t.lh = t.lw = 0.f;
    if (!t.inside) return t;
    float hf = floorf(h), wf = floorf(w);
    int h0 = (int)hf, w0 = (int)wf, h1 = h0 + 1, w1 = w0 + 1;
    float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
    t.lh = lh; t.lw = lw;
    bool h0ok = h0 >= 0, w0ok = w0 >= 0, h1ok = h1 <= H - 1, w1ok = w1 <= W - 1;
    if (h0ok && w0ok) t.o00 = h0 * W + w0;
    if (h0ok && w1ok) t.o01 = h0 * W + w1;
    if (h1ok && w0ok) t.o10 = h1 * W + w0;
    if (h1ok && w1ok) t.o11 = h1 * W + w1;
    t.w00 = hh * hw; t.w01 = hh * lw; t.w10 = lh * hw; t.w11 = lh * lw;
    return t;
}

}  // namespace dvsr
