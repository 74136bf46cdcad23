/*
 * oracle/mdcn_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement, in plain C, of the reference's modulated deformable convolution
 * (DCNv2) forward and backward.  NCHW fp32 tensors, caller-owned outputs, exactly the
 * reference operator boundary.  It follows the *algorithm* of
 *
 *   codes/models/archs/dcn/src/deform_conv_cuda.cpp:486-564   (forward host loop)
 *   codes/models/archs/dcn/src/deform_conv_cuda.cpp:566-679   (backward host loop)
 *   codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:466-496  (bilinear sample, per-corner bounds)
 *   codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:498-567  (gradient / coordinate weights)
 *   codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:569-632  (modulated im2col)
 *   codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:634-692  (col2im: grad wrt input)
 *   codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:694-766  (col2im_coord: grad wrt offset, mask)
 *
 * but is organised per output pixel instead of per "column" element, has no scratch
 * `columns` matrix in the forward pass and accumulates every reduction in double so
 * that it is a tighter truth than the reference's fp32 cuBLAS/atomicAdd order.
 *
 * Parity pin: the reference ships NO golden vectors for this op, so this
 * file is pinned against (a) torchvision.ops.deform_conv2d on CPU and (b) the unmodified
 * reference Python modules driven through that op; see tests/test_oracle.py and
 * oracle/make_golden.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int B, C, H, W;          /* input  [B, C, H, W]                                  */
    int Co;                  /* weight [Co, C/groups, kh, kw]                        */
    int kh, kw, sh, sw, ph, pw, dh, dw;
    int groups, dg;          /* conv groups, deformable groups                       */
    int Ho, Wo;              /* output spatial size                                  */
} mdcn_shape;

static void infer(mdcn_shape *s) {
    /* deform_conv_cuda.cpp:513-516 */
    s->Ho = (s->H + 2 * s->ph - (s->dh * (s->kh - 1) + 1)) / s->sh + 1;
    s->Wo = (s->W + 2 * s->pw - (s->dw * (s->kw - 1) + 1)) / s->sw + 1;
}

/* One bilinear tap: returns the 4 corner indices (-1 when that corner is outside the
 * image and contributes 0) and the 4 weights.  deform_conv_cuda_kernel.cu:466-496.
 * `inside` mirrors the (-1,H)x(-1,W) open-interval test of kernel.cu:617. */
typedef struct { int idx[4]; float w[4]; int inside; float lh, lw; int h_low, w_low; } tap_t;

static tap_t make_tap(float h, float w, int H, int W) {
    tap_t t;
    memset(&t, 0, sizeof(t));
    t.inside = (h > -1.f && w > -1.f && h < (float)H && w < (float)W);
    for (int i = 0; i < 4; ++i) t.idx[i] = -1;
    if (!t.inside) return t;
    int h_low = (int)floorf(h), w_low = (int)floorf(w);
    int h_high = h_low + 1, w_high = w_low + 1;
    float lh = h - h_low, lw = w - w_low, hh = 1.f - lh, hw = 1.f - lw;
    t.lh = lh; t.lw = lw; t.h_low = h_low; t.w_low = w_low;
    if (h_low >= 0 && w_low >= 0)          t.idx[0] = h_low * W + w_low;
    if (h_low >= 0 && w_high <= W - 1)     t.idx[1] = h_low * W + w_high;
    if (h_high <= H - 1 && w_low >= 0)     t.idx[2] = h_high * W + w_low;
    if (h_high <= H - 1 && w_high <= W - 1) t.idx[3] = h_high * W + w_high;
    t.w[0] = hh * hw; t.w[1] = hh * lw; t.w[2] = lh * hw; t.w[3] = lh * lw;
    return t;
}

static float tap_sample(const tap_t *t, const float *plane) {
    float v[4];
    for (int i = 0; i < 4; ++i) v[i] = t->idx[i] >= 0 ? plane[t->idx[i]] : 0.f;
    /* same association as kernel.cu:494 */
    return (t->w[0] * v[0] + t->w[1] * v[1] + t->w[2] * v[2] + t->w[3] * v[3]);
}

/* offset layout [B, dg*2*kh*kw, Ho, Wo], channel = (g*kh*kw + k)*2 + {0:dy,1:dx};
 * mask layout [B, dg*kh*kw, Ho, Wo]  (kernel.cu:596-607). */
static inline float off_at(const float *offset, const mdcn_shape *s, int b, int g, int k, int xy, int p) {
    int K = s->kh * s->kw;
    return offset[(((size_t)b * s->dg + g) * 2 * K + 2 * k + xy) * (size_t)(s->Ho * s->Wo) + p];
}
static inline float mask_at(const float *mask, const mdcn_shape *s, int b, int g, int k, int p) {
    int K = s->kh * s->kw;
    return mask[(((size_t)b * s->dg + g) * K + k) * (size_t)(s->Ho * s->Wo) + p];
}

/* column value for (channel c, tap k, output pixel p) of sample b: kernel.cu:609-623 */
static float column_value(const float *x, const float *offset, const float *mask,
                          const mdcn_shape *s, int b, int c, int k, int p) {
    int cpg = s->C / s->dg, g = c / cpg;
    int ho = p / s->Wo, wo = p % s->Wo, i = k / s->kw, j = k % s->kw;
    float h = (float)(ho * s->sh - s->ph + i * s->dh) + off_at(offset, s, b, g, k, 0, p);
    float w = (float)(wo * s->sw - s->pw + j * s->dw) + off_at(offset, s, b, g, k, 1, p);
    tap_t t = make_tap(h, w, s->H, s->W);
    if (!t.inside) return 0.f;
    const float *plane = x + ((size_t)b * s->C + c) * (size_t)(s->H * s->W);
    return tap_sample(&t, plane) * mask_at(mask, s, b, g, k, p);
}

int mdcn_oracle_forward(const float *x, const float *offset, const float *mask,
                        const float *weight, const float *bias, float *y,
                        int B, int C, int H, int W, int Co, int kh, int kw,
                        int sh, int sw, int ph, int pw, int dh, int dw, int groups, int dg) {
    mdcn_shape s = {B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, 0, 0};
    if (C % groups || Co % groups || C % dg) return -1;
    infer(&s);
    int K = kh * kw, P = s.Ho * s.Wo, Cg = C / groups, Cog = Co / groups;
    float *col = (float *)malloc(sizeof(float) * (size_t)C * K);
    if (!col) return -2;
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < K; ++k) col[c * K + k] = column_value(x, offset, mask, &s, b, c, k, p);
            for (int co = 0; co < Co; ++co) {
                int g = co / Cog;
                double acc = 0.0; /* cpp:545-550 (addmm) then cpp:561-563 (bias) */
                const float *wrow = weight + (size_t)co * Cg * K;
                for (int cc = 0; cc < Cg; ++cc)
                    for (int k = 0; k < K; ++k) acc += (double)wrow[cc * K + k] * (double)col[(g * Cg + cc) * K + k];
                if (bias) acc += (double)bias[co];
                y[((size_t)b * Co + co) * P + p] = (float)acc;
            }
        }
    free(col);
    return 0;
}

/* Backward.  All grad_* buffers are caller-owned and are OVERWRITTEN (the reference
 * caller passes zeros_like tensors, deform_conv.py:128-132, and the callee accumulates
 * into them; net effect is identical). */
int mdcn_oracle_backward(const float *x, const float *offset, const float *mask,
                         const float *weight, const float *gy,
                         float *gx, float *goffset, float *gmask, float *gweight, float *gbias,
                         int B, int C, int H, int W, int Co, int kh, int kw,
                         int sh, int sw, int ph, int pw, int dh, int dw, int groups, int dg) {
    mdcn_shape s = {B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, 0, 0};
    if (C % groups || Co % groups || C % dg) return -1;
    infer(&s);
    int K = kh * kw, P = s.Ho * s.Wo, Cg = C / groups, Cog = Co / groups, cpg = C / dg;
    size_t nx = (size_t)B * C * H * W, nw = (size_t)Co * Cg * K;
    double *gxd = (double *)calloc(nx, sizeof(double));
    double *gwd = (double *)calloc(nw, sizeof(double));
    double *gbd = (double *)calloc((size_t)Co, sizeof(double));
    double *gcol = (double *)malloc(sizeof(double) * (size_t)C * K);
    if (!gxd || !gwd || !gbd || !gcol) return -2;

    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            int ho = p / s.Wo, wo = p % s.Wo;
            /* grad_col = W^T * gy   (cpp:617-620) */
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < K; ++k) {
                    int g = c / Cg, cc = c % Cg;
                    double acc = 0.0;
                    for (int o = 0; o < Cog; ++o) {
                        int co = g * Cog + o;
                        acc += (double)weight[((size_t)co * Cg + cc) * K + k] * (double)gy[((size_t)b * Co + co) * P + p];
                    }
                    gcol[c * K + k] = acc;
                }
            /* per (deformable group, tap): coordinate + mask gradients (kernel.cu:694-766)
             * and the scatter into grad_input (kernel.cu:634-692). */
            for (int g = 0; g < dg; ++g)
                for (int k = 0; k < K; ++k) {
                    int i = k / kw, j = k % kw;
                    float oh = off_at(offset, &s, b, g, k, 0, p), ow = off_at(offset, &s, b, g, k, 1, p);
                    float m = mask_at(mask, &s, b, g, k, p);
                    float h = (float)(ho * sh - ph + i * dh) + oh;
                    float w = (float)(wo * sw - pw + j * dw) + ow;
                    tap_t t = make_tap(h, w, H, W);
                    double d_h = 0.0, d_w = 0.0, d_m = 0.0;
                    if (t.inside) {
                        float hh = 1.f - t.lh, hw = 1.f - t.lw;
                        /* d(bilinear)/dh and /dw per corner: kernel.cu:536-563 */
                        float dwh[4] = {-hw, -t.lw, hw, t.lw};
                        float dww[4] = {-hh, hh, -t.lh, t.lh};
                        for (int cc = 0; cc < cpg; ++cc) {
                            int c = g * cpg + cc;
                            const float *plane = x + ((size_t)b * C + c) * (size_t)(H * W);
                            double gc = gcol[c * K + k];
                            double val = 0.0, ch = 0.0, cw = 0.0;
                            for (int q = 0; q < 4; ++q) {
                                if (t.idx[q] < 0) continue;
                                float v = plane[t.idx[q]];
                                val += (double)t.w[q] * v;
                                ch += (double)dwh[q] * v;
                                cw += (double)dww[q] * v;
                                /* kernel.cu:676-690: weight * grad_col * mask scattered to the corner */
                                gxd[((size_t)b * C + c) * (size_t)(H * W) + t.idx[q]] += (double)t.w[q] * gc * (double)m;
                            }
                            d_m += gc * val;              /* kernel.cu:746-753 (unmasked sample)  */
                            d_h += ch * gc * (double)m;   /* kernel.cu:754-758                    */
                            d_w += cw * gc * (double)m;
                        }
                    }
                    size_t ob = (((size_t)b * dg + g) * 2 * K + 2 * k) * (size_t)P + p;
                    goffset[ob] = (float)d_h;
                    goffset[ob + P] = (float)d_w;
                    gmask[(((size_t)b * dg + g) * K + k) * (size_t)P + p] = (float)d_m;
                }
            /* grad_weight += gy * col^T, grad_bias += gy * 1  (cpp:641-666) */
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < K; ++k) {
                    float cv = column_value(x, offset, mask, &s, b, c, k, p);
                    if (cv == 0.f) continue;
                    int g = c / Cg, cc = c % Cg;
                    for (int o = 0; o < Cog; ++o) {
                        int co = g * Cog + o;
                        gwd[((size_t)co * Cg + cc) * K + k] += (double)gy[((size_t)b * Co + co) * P + p] * (double)cv;
                    }
                }
            for (int co = 0; co < Co; ++co) gbd[co] += (double)gy[((size_t)b * Co + co) * P + p];
        }
    for (size_t i = 0; i < nx; ++i) gx[i] = (float)gxd[i];
    for (size_t i = 0; i < nw; ++i) gweight[i] = (float)gwd[i];
    if (gbias) for (int co = 0; co < Co; ++co) gbias[co] = (float)gbd[co];
    free(gxd); free(gwd); free(gbd); free(gcol);
    return 0;
}
