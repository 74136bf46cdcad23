// dynavsr_b200/csrc/tsa.cu
//
// TSA fusion glue (EDVR_arch.py:163-203) as two fused HBM-bound kernels (+ their adjoints):
//   tsa_temporal : cor_f = sum_c emb_f * emb_ref ; prob_f = sigmoid(cor_f) ; out[b][p][f*C + c] = aligned_f * prob_f
//                  (output is pixel-major with F*C channels = exactly `aligned_fea.view(B, -1, H, W)` in NHWC)
//                  -- replaces the python loop :170-173, the sigmoid :174, the `repeat` that materialises
//                  a 73.7 MB probability tensor :175 and the multiply :176 (5 + 1 + 1 + 1 launches).
//   tsa_combine  : out = fea * sigmoid(att) * 2 + att_add        (:200-202, 4 launches in the reference)
// One warp per pixel; lanes stride over the channels, so every access is a coalesced NHWC row.
#include "common.cuh"

namespace dvsr {

constexpr int TSA_MAXF = 8;

__global__ void tsa_temporal_kernel(const float* __restrict__ aligned, const float* __restrict__ emb,
                                    const float* __restrict__ emb_ref, float* __restrict__ prob,
                                    float* __restrict__ out, int B, int F, long long HW, int C) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long px = warp; px < (long long)B * HW; px += nwarps) {
        const int b = (int)(px / HW);
        const long long p = px - (long long)b * HW;
        const float* ref = emb_ref + px * C;
        for (int f = 0; f < F; ++f) {
            const long long row = (((long long)b * F + f) * HW + p);
            const float* e = emb + row * C;
            float dot = 0.f;
            for (int c = lane; c < C; c += 32) dot = fmaf(__ldg(e + c), __ldg(ref + c), dot);
            dot = warp_sum(dot);
            const float pr = 1.f / (1.f + expf(-dot));
            if (lane == 0) prob[row] = pr;
            const float* a = aligned + row * C;
            float* o = out + (px * F + f) * C;   // pixel-major: [B][HW][F*C], the K layout of the 1x1 fusion convs
            for (int c = lane; c < C; c += 32) o[c] = __ldg(a + c) * pr;
        }
    }
}

__global__ void tsa_temporal_bwd_kernel(const float* __restrict__ aligned, const float* __restrict__ emb,
                                        const float* __restrict__ emb_ref, const float* __restrict__ prob,
                                        const float* __restrict__ gout, float* __restrict__ galigned,
                                        float* __restrict__ gemb, float* __restrict__ gemb_ref, int B, int F,
                                        long long HW, int C) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long px = warp; px < (long long)B * HW; px += nwarps) {
        const int b = (int)(px / HW);
        const long long p = px - (long long)b * HW;
        const float* ref = emb_ref + px * C;
        float gcor[TSA_MAXF];
        for (int f = 0; f < F; ++f) {
            const long long row = (((long long)b * F + f) * HW + p);
            const float pr = __ldg(prob + row);
            const float* a = aligned + row * C;
            const float* g = gout + (px * F + f) * C;
            float* ga = galigned + row * C;
            float gp = 0.f;
            for (int c = lane; c < C; c += 32) {
                const float gv = __ldg(g + c);
                gp = fmaf(gv, __ldg(a + c), gp);
                ga[c] = gv * pr;
            }
            gp = warp_sum(gp);
            gcor[f] = gp * pr * (1.f - pr);
            float* ge = gemb + row * C;
            for (int c = lane; c < C; c += 32) ge[c] = gcor[f] * __ldg(ref + c);
        }
        float* gr = gemb_ref + px * C;
        for (int c = lane; c < C; c += 32) {
            float s = 0.f;
            for (int f = 0; f < F; ++f) s = fmaf(gcor[f], __ldg(emb + (((long long)b * F + f) * HW + p) * C + c), s);
            gr[c] = s;
        }
    }
}

__global__ void tsa_combine_kernel(const float* __restrict__ fea, const float* __restrict__ att,
                                   const float* __restrict__ att_add, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float s = 1.f / (1.f + expf(-att[i]));
        out[i] = fea[i] * s * 2.f + att_add[i];
    }
}

__global__ void tsa_combine_bwd_kernel(const float* __restrict__ fea, const float* __restrict__ att,
                                       const float* __restrict__ gout, float* __restrict__ gfea,
                                       float* __restrict__ gatt, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float s = 1.f / (1.f + expf(-att[i]));
        const float g = gout[i];
        gfea[i] = g * s * 2.f;
        gatt[i] = g * fea[i] * 2.f * s * (1.f - s);
    }
}

static int grid_for(long long work_items, int per_block) {
    long long b = (work_items + per_block - 1) / per_block;
    if (b > sm_count() * 16) b = sm_count() * 16;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace dvsr

using namespace dvsr;
#define ST ((cudaStream_t)stream)

extern "C" int dvsr_tsa_temporal(const float* aligned, const float* emb, const float* emb_ref, float* prob, float* out,
                                 int B, int F, long long HW, int C, void* stream) {
    DVSR_REQUIRE(aligned && emb && emb_ref && prob && out && B > 0 && F > 0 && HW > 0 && C > 0, "tsa_temporal: bad arguments");
    tsa_temporal_kernel<<<grid_for((long long)B * HW, 8), 256, 0, ST>>>(aligned, emb, emb_ref, prob, out, B, F, HW, C);
    return check_launch("tsa_temporal");
}
extern "C" int dvsr_tsa_temporal_bwd(const float* aligned, const float* emb, const float* emb_ref, const float* prob,
                                     const float* gout, float* galigned, float* gemb, float* gemb_ref, int B, int F,
                                     long long HW, int C, void* stream) {
    DVSR_REQUIRE(aligned && emb && emb_ref && prob && gout && galigned && gemb && gemb_ref, "tsa_temporal_bwd: null pointer");
    DVSR_REQUIRE(B > 0 && F > 0 && F <= TSA_MAXF && HW > 0 && C > 0, "tsa_temporal_bwd: bad shape (F <= %d)", TSA_MAXF);
    tsa_temporal_bwd_kernel<<<grid_for((long long)B * HW, 8), 256, 0, ST>>>(aligned, emb, emb_ref, prob, gout, galigned, gemb,
                                                                          gemb_ref, B, F, HW, C);
    return check_launch("tsa_temporal_bwd");
}
extern "C" int dvsr_tsa_combine(const float* fea, const float* att, const float* att_add, float* out, long long n, void* stream) {
    DVSR_REQUIRE(fea && att && att_add && out && n > 0, "tsa_combine: bad arguments");
    tsa_combine_kernel<<<grid_for(n, 256), 256, 0, ST>>>(fea, att, att_add, out, n);
    return check_launch("tsa_combine");
}
extern "C" int dvsr_tsa_combine_bwd(const float* fea, const float* att, const float* gout, float* gfea, float* gatt,
                                    long long n, void* stream) {
    DVSR_REQUIRE(fea && att && gout && gfea && gatt && n > 0, "tsa_combine_bwd: bad arguments");
    tsa_combine_bwd_kernel<<<grid_for(n, 256), 256, 0, ST>>>(fea, att, gout, gfea, gatt, n);
    return check_launch("tsa_combine_bwd");
}
