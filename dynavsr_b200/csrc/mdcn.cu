// dynavsr_b200/csrc/mdcn.cu
//
// Modulated deformable convolution (DCNv2): data-side backward kernel and the reference's NCHW
// operator boundary.
//
//   mdcn_bwd_data_kernel  fuses, per 64-pixel tile and per tap, what the reference does with three
//   full-size passes over a 132.7 MB `columns` buffer (deform_conv_cuda.cpp:617-634):
//       grad_col = W^T . gy                      (addmm_,           cpp:617-620)
//       grad_offset / grad_mask                  (col2im_coord,     kernel.cu:694-766)
//       grad_input scatter                       (col2im + atomics, kernel.cu:634-692)
//   grad_col lives only in shared memory.  The scatter uses 16-byte vector reductions
//   (red.global.add.v4.f32) over the NHWC channel slice of each bilinear corner.
//
//   The forward pass and the weight gradient reuse the implicit-GEMM kernels of conv_simt.cu with the
//   deformable A-loader (no im2col in HBM).
#include "common.cuh"

namespace dvsr {

constexpr int DP = 64;   // pixels per CTA
constexpr int DCT = 64;  // input-channel tile
constexpr int DNT = 256;

__device__ __forceinline__ void red_add4(float* p, float4 v) {
#if __CUDA_ARCH__ >= 900
    atomicAdd(reinterpret_cast<float4*>(p), v);
#else
    atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
#endif
}

template <bool VEC4>
__global__ void __launch_bounds__(DNT)
mdcn_bwd_data_kernel(const dvsr_conv_desc d, const float* __restrict__ gy, int gy_pix_stride,
                     const float* __restrict__ wd, float* __restrict__ gx, int gx_pix_stride,
                     float* __restrict__ goff, int goff_pix_stride, float* __restrict__ gmask,
                     int gmask_pix_stride) {
    extern __shared__ __align__(16) float smem[];
    const int Co = d.Co, C = d.seg[0].C, KK = d.KH * d.KW;
    const int LG = DP + 4, LC = DCT + 4;
    float* Gs = smem;                 // [Co][LG]   gy tile, transposed
    float* Ws = Gs + Co * LG;         // [Co][DCT]  weight slice of one tap
    float* Cs = Ws + Co * DCT;        // [DP][LC]   grad_col tile

    const int tid = threadIdx.x;
    const long long M = (long long)d.N * d.Ho * d.Wo;
    const long long m0 = (long long)blockIdx.x * DP;
    const int hw = d.Ho * d.Wo;
    const dvsr_conv_seg& sg = d.seg[0];
    const int cpg = C / d.dg;

    // gy tile -> Gs[co][pix]
    for (int e = tid; e < DP * Co; e += DNT) {
        const int p = e / Co, co = e - p * Co;
        const long long m = m0 + p;
        Gs[co * LG + p] = (m < M) ? __ldg(gy + m * gy_pix_stride + co) : 0.f;
    }

    const int tx = tid & 15, ty = tid >> 4;  // GEMM: pixels ty*4..+3, channels tx*4..+3
    for (int ct = 0; ct < C; ct += DCT) {
        const int ctw = min(DCT, C - ct);      // valid channels in this tile
        const int ngt = ctw / cpg;             // deformable groups covered by this tile
        for (int tap = 0; tap < KK; ++tap) {
            __syncthreads();  // Gs ready (first iter) / previous phase 2 done with Ws, Cs
            for (int e = tid; e < Co * DCT; e += DNT) {
                const int co = e / DCT, c = e - co * DCT;
                Ws[e] = (c < ctw) ? __ldg(wd + ((long long)tap * Co + co) * C + ct + c) : 0.f;
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            for (int co = 0; co < Co; ++co) {
                const float4 a = *reinterpret_cast<const float4*>(&Gs[co * LG + ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Ws[co * DCT + tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(&Cs[(ty * 4 + i) * LC + tx * 4]) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            __syncthreads();

            // phase 2: one work item per (pixel, deformable group)
            const int kh = tap / d.KW, kw = tap - kh * d.KW;
            for (int item = tid; item < DP * ngt; item += DNT) {
                const int p = item / ngt, gl = item - p * ngt;
                const long long m = m0 + p;
                if (m >= M) continue;
                const int g = ct / cpg + gl;
                const int n = (int)(m / hw);
                const int r = (int)(m - (long long)n * hw);
                const int oh = r / d.Wo, ow = r - oh * d.Wo;
                const float* op = d.offset + m * d.off_pix_stride + (g * KK + tap) * 2;
                const float dy = __ldg(op), dx = __ldg(op + 1);
                const float mk = __ldg(d.mask + m * d.mask_pix_stride + g * KK + tap);
                const float h = (float)(oh * d.stride - d.pad + kh * d.dil) + dy;
                const float w = (float)(ow * d.stride - d.pad + kw * d.dil) + dx;
                const BilinTap t = make_tap(h, w, d.H, d.W);
                float d_m = 0.f, d_h = 0.f, d_w = 0.f;
                if (t.inside) {
                    const float hh = 1.f - t.lh, hwt = 1.f - t.lw;
                    int T = sg.T > 0 ? sg.T : 1;
                    int q = n / T, rr = n - q * T;
                    const long long img_i = (long long)q * sg.Tsrc + (sg.t_fixed >= 0 ? sg.t_fixed : rr + sg.dt);
                    const float* img = sg.ptr + img_i * sg.img_stride;
                    // gx uses the same image indexing as x but its own strides (dense NHWC of the source)
                    float* gimg = gx ? gx + img_i * ((long long)d.H * d.W * gx_pix_stride) : nullptr;
                    const float* crow = Cs + p * LC + gl * cpg;
                    const int cbase = ct + gl * cpg;
                    if (VEC4) {
                        for (int cc = 0; cc < cpg; cc += 4) {
                            const float4 gc = *reinterpret_cast<const float4*>(crow + cc);
                            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                            const float4 v00 = t.o00 >= 0 ? ldg4(img + (long long)t.o00 * sg.pix_stride + cbase + cc) : z;
                            const float4 v01 = t.o01 >= 0 ? ldg4(img + (long long)t.o01 * sg.pix_stride + cbase + cc) : z;
                            const float4 v10 = t.o10 >= 0 ? ldg4(img + (long long)t.o10 * sg.pix_stride + cbase + cc) : z;
                            const float4 v11 = t.o11 >= 0 ? ldg4(img + (long long)t.o11 * sg.pix_stride + cbase + cc) : z;
                            const float g4[4] = {gc.x, gc.y, gc.z, gc.w};
                            const float a4[4] = {v00.x, v00.y, v00.z, v00.w}, b4[4] = {v01.x, v01.y, v01.z, v01.w};
                            const float e4[4] = {v10.x, v10.y, v10.z, v10.w}, f4[4] = {v11.x, v11.y, v11.z, v11.w};
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) {
                                const float val = t.w00 * a4[q4] + t.w01 * b4[q4] + t.w10 * e4[q4] + t.w11 * f4[q4];
                                const float ch = -hwt * a4[q4] - t.lw * b4[q4] + hwt * e4[q4] + t.lw * f4[q4];
                                const float cw = -hh * a4[q4] + hh * b4[q4] - t.lh * e4[q4] + t.lh * f4[q4];
                                d_m = fmaf(g4[q4], val, d_m);
                                d_h = fmaf(ch, g4[q4] * mk, d_h);
                                d_w = fmaf(cw, g4[q4] * mk, d_w);
                            }
                            if (gimg) {
                                const float4 gm = make_float4(gc.x * mk, gc.y * mk, gc.z * mk, gc.w * mk);
                                if (t.o00 >= 0) red_add4(gimg + (long long)t.o00 * gx_pix_stride + cbase + cc,
                                                         make_float4(gm.x * t.w00, gm.y * t.w00, gm.z * t.w00, gm.w * t.w00));
                                if (t.o01 >= 0) red_add4(gimg + (long long)t.o01 * gx_pix_stride + cbase + cc,
                                                         make_float4(gm.x * t.w01, gm.y * t.w01, gm.z * t.w01, gm.w * t.w01));
                                if (t.o10 >= 0) red_add4(gimg + (long long)t.o10 * gx_pix_stride + cbase + cc,
                                                         make_float4(gm.x * t.w10, gm.y * t.w10, gm.z * t.w10, gm.w * t.w10));
                                if (t.o11 >= 0) red_add4(gimg + (long long)t.o11 * gx_pix_stride + cbase + cc,
                                                         make_float4(gm.x * t.w11, gm.y * t.w11, gm.z * t.w11, gm.w * t.w11));
                            }
                        }
                    } else {
                        for (int cc = 0; cc < cpg; ++cc) {
                            const float gc = crow[cc];
                            const int c = cbase + cc;
                            const float a = t.o00 >= 0 ? __ldg(img + (long long)t.o00 * sg.pix_stride + c) : 0.f;
                            const float b = t.o01 >= 0 ? __ldg(img + (long long)t.o01 * sg.pix_stride + c) : 0.f;
                            const float e = t.o10 >= 0 ? __ldg(img + (long long)t.o10 * sg.pix_stride + c) : 0.f;
                            const float f = t.o11 >= 0 ? __ldg(img + (long long)t.o11 * sg.pix_stride + c) : 0.f;
                            const float val = t.w00 * a + t.w01 * b + t.w10 * e + t.w11 * f;
                            const float ch = -hwt * a - t.lw * b + hwt * e + t.lw * f;
                            const float cw = -hh * a + hh * b - t.lh * e + t.lh * f;
                            d_m = fmaf(gc, val, d_m);
                            d_h = fmaf(ch, gc * mk, d_h);
                            d_w = fmaf(cw, gc * mk, d_w);
                            if (gimg) {
                                const float gm = gc * mk;
                                if (t.o00 >= 0) atomicAdd(gimg + (long long)t.o00 * gx_pix_stride + c, gm * t.w00);
                                if (t.o01 >= 0) atomicAdd(gimg + (long long)t.o01 * gx_pix_stride + c, gm * t.w01);
                                if (t.o10 >= 0) atomicAdd(gimg + (long long)t.o10 * gx_pix_stride + c, gm * t.w10);
                                if (t.o11 >= 0) atomicAdd(gimg + (long long)t.o11 * gx_pix_stride + c, gm * t.w11);
                            }
                        }
                    }
                }
                if (goff) {
                    float* go = goff + m * goff_pix_stride + (g * KK + tap) * 2;
                    go[0] = d_h;
                    go[1] = d_w;
                }
                if (gmask) gmask[m * gmask_pix_stride + g * KK + tap] = d_m;
            }
        }
    }
}

static size_t bwd_data_smem(int Co) { return sizeof(float) * ((size_t)Co * (DP + 4) + (size_t)Co * DCT + (size_t)DP * (DCT + 4)); }

}  // namespace dvsr

using namespace dvsr;

extern "C" int dvsr_mdcn_bwd_data(const dvsr_conv_desc* d, const float* gy, int gy_pix_stride, const float* wd,
                                  float* gx, int gx_pix_stride, float* goff, int goff_pix_stride, float* gmask,
                                  int gmask_pix_stride, void* stream) {
    DVSR_REQUIRE(d && gy && wd, "mdcn_bwd_data: null pointer");
    DVSR_REQUIRE(d->deform == 1 && d->nseg == 1 && !d->transposed, "mdcn_bwd_data: descriptor must describe a deformable forward op");
    const int C = d->seg[0].C;
    DVSR_REQUIRE(d->dg > 0 && C % d->dg == 0, "mdcn_bwd_data: C=%d not divisible by dg=%d", C, d->dg);
    const int cpg = C / d->dg;
    DVSR_REQUIRE((C % DCT == 0 || C < DCT) && (DCT % cpg == 0 || C < DCT), "mdcn_bwd_data: unsupported C=%d / channels-per-group=%d", C, cpg);
    const size_t smem = bwd_data_smem(d->Co);
    DVSR_REQUIRE(smem <= 200 * 1024, "mdcn_bwd_data: Co=%d needs %zu B of shared memory", d->Co, smem);
    const long long M = (long long)d->N * d->Ho * d->Wo;
    const bool v4 = (cpg % 4 == 0) && (d->seg[0].pix_stride % 4 == 0) && (d->seg[0].img_stride % 4 == 0) &&
                    (((uintptr_t)d->seg[0].ptr & 15) == 0) && (!gx || ((((uintptr_t)gx) & 15) == 0 && gx_pix_stride % 4 == 0));
    cudaStream_t st = (cudaStream_t)stream;
    if (v4) {
        cudaFuncSetAttribute(mdcn_bwd_data_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mdcn_bwd_data_kernel<true><<<cdiv(M, DP), DNT, smem, st>>>(*d, gy, gy_pix_stride, wd, gx, gx_pix_stride, goff,
                                                                  goff_pix_stride, gmask, gmask_pix_stride);
    } else {
        cudaFuncSetAttribute(mdcn_bwd_data_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mdcn_bwd_data_kernel<false><<<cdiv(M, DP), DNT, smem, st>>>(*d, gy, gy_pix_stride, wd, gx, gx_pix_stride, goff,
                                                                   goff_pix_stride, gmask, gmask_pix_stride);
    }
    return check_launch("mdcn_bwd_data");
}

// ------------------------------------------------------------------------------------------------
// Reference operator boundary (NCHW fp32).  Layout changes to NHWC happen in the caller-provided
// workspace; the arithmetic is done by the NHWC kernels above / in conv_simt.cu.
namespace {
struct Ws {
    long long x, off, mask, y, wp, gy, gx, goff, gmask, wd, total;
};
Ws plan(int B, int C, int H, int W, int Co, int kh, int kw, int stride, int pad, int dil, int dg, int backward,
        int* Ho_, int* Wo_) {
    const int Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    const int Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    if (Ho_) *Ho_ = Ho;
    if (Wo_) *Wo_ = Wo;
    const long long KK = (long long)kh * kw, P = (long long)B * Ho * Wo;
    auto al = [](long long n) { return (n + 63) / 64 * 64; };  // 256-byte granules (in floats)
    Ws w;
    long long o = 0;
    w.x = o; o += al((long long)B * H * W * C);
    w.off = o; o += al(P * 2 * dg * KK);
    w.mask = o; o += al(P * dg * KK);
    w.y = o; o += al(P * Co);
    w.wp = o; o += al(KK * C * Co);
    w.gy = w.gx = w.goff = w.gmask = w.wd = 0;
    if (backward) {
        w.gy = w.y;  // the forward y slot is reused for gy
        w.gx = o; o += al((long long)B * H * W * C);
        w.goff = o; o += al(P * 2 * dg * KK);
        w.gmask = o; o += al(P * dg * KK);
        w.wd = o; o += al(KK * C * Co);
    }
    w.total = o;
    return w;
}
}  // namespace

// The reference entry points take (h, w) pairs for stride / padding / dilation (deform_conv_cuda.cpp:486-492); every kernel of
// this library is isotropic (as is every call the reference's Python layer makes: deform_conv.py:104-106 passes the same value
// twice), so an anisotropic request is reported as unsupported instead of being silently squared.
static int iso(const char* who, int sh, int sw, int ph, int pw, int dh, int dw) {
    if (sh != sw || ph != pw || dh != dw) {
        set_error("%s: anisotropic stride/pad/dilation (%d,%d)/(%d,%d)/(%d,%d) is not supported", who, sh, sw, ph, pw, dh, dw);
        return DVSR_ERR_UNSUPPORTED;
    }
    return 0;
}

extern "C" long long dvsr_mdcn_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw, int stride_h, int stride_w,
                                               int pad_h, int pad_w, int dil_h, int dil_w, int dg, int backward) {
    if (iso("mdcn_workspace_bytes", stride_h, stride_w, pad_h, pad_w, dil_h, dil_w)) return DVSR_ERR_UNSUPPORTED;
    const int stride = stride_h, pad = pad_h, dil = dil_h;
    return plan(B, C, H, W, Co, kh, kw, stride, pad, dil, dg, backward, nullptr, nullptr).total * (long long)sizeof(float);
}

static int fill_desc(dvsr_conv_desc* d, const float* x_nhwc, const float* off, const float* mask, int B, int C, int H,
                     int W, int Ho, int Wo, int Co, int kh, int kw, int stride, int pad, int dil, int dg) {
    memset(d, 0, sizeof(*d));
    d->N = B; d->H = H; d->W = W; d->Ho = Ho; d->Wo = Wo;
    d->KH = kh; d->KW = kw; d->stride = stride; d->pad = pad; d->dil = dil;
    d->nseg = 1;
    d->seg[0].ptr = x_nhwc; d->seg[0].C = C; d->seg[0].pix_stride = C;
    d->seg[0].img_stride = (long long)H * W * C;
    d->seg[0].T = 1; d->seg[0].Tsrc = 1; d->seg[0].dt = 0; d->seg[0].t_fixed = -1;
    d->Co = Co;
    d->deform = 1; d->dg = dg;
    d->offset = off; d->off_pix_stride = 2 * dg * kh * kw;
    d->mask = mask; d->mask_pix_stride = dg * kh * kw;
    return 0;
}

static dvsr_wlayout conv2d_layout(int Co, int C, int kh, int kw) {
    dvsr_wlayout wl;
    memset(&wl, 0, sizeof(wl));
    wl.co_stride = (long long)C * kh * kw;
    wl.ci_stride = (long long)kh * kw;
    wl.nseg = 1; wl.seg_base[0] = 0; wl.seg_C[0] = C; wl.taps = kh * kw; wl.Co = Co;
    return wl;
}

extern "C" int dvsr_mdcn_forward_nchw(const float* x, const float* offset, const float* mask, const float* weight,
                                      const float* bias, float* y, int B, int C, int H, int W, int Co, int kh, int kw,
                                      int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int groups, int dg,
                                      void* workspace, long long workspace_bytes, void* stream) {
    DVSR_REQUIRE(x && offset && mask && weight && y && workspace, "mdcn_forward_nchw: null pointer");
    if (int rc0 = iso("mdcn_forward_nchw", stride_h, stride_w, pad_h, pad_w, dil_h, dil_w)) return rc0;
    const int stride = stride_h, pad = pad_h, dil = dil_h;
    if (groups != 1) { set_error("mdcn_forward_nchw: groups=%d is not supported (the reference never uses groups != 1)", groups); return DVSR_ERR_UNSUPPORTED; }
    DVSR_REQUIRE(dg > 0 && C % dg == 0, "mdcn_forward_nchw: channels %d not divisible by deformable groups %d", C, dg);
    int Ho, Wo;
    const Ws w = plan(B, C, H, W, Co, kh, kw, stride, pad, dil, dg, 0, &Ho, &Wo);
    DVSR_REQUIRE(Ho > 0 && Wo > 0, "mdcn_forward_nchw: empty output");
    DVSR_REQUIRE(workspace_bytes >= w.total * (long long)sizeof(float), "mdcn_forward_nchw: workspace too small (%lld < %lld)",
                 workspace_bytes, w.total * (long long)sizeof(float));
    float* ws = (float*)workspace;
    const int KK = kh * kw;
    int rc;
    if ((rc = dvsr_nchw_to_nhwc(x, ws + w.x, B, C, H, W, stream))) return rc;
    if ((rc = dvsr_nchw_to_nhwc(offset, ws + w.off, B, 2 * dg * KK, Ho, Wo, stream))) return rc;
    if ((rc = dvsr_nchw_to_nhwc(mask, ws + w.mask, B, dg * KK, Ho, Wo, stream))) return rc;
    const dvsr_wlayout wl = conv2d_layout(Co, C, kh, kw);
    if ((rc = dvsr_pack_weights(weight, ws + w.wp, &wl, 0, 0, stream))) return rc;
    dvsr_conv_desc d;
    fill_desc(&d, ws + w.x, ws + w.off, ws + w.mask, B, C, H, W, Ho, Wo, Co, kh, kw, stride, pad, dil, dg);
    d.bias = bias;
    d.y = ws + w.y; d.y_pix_stride = Co;
    if ((rc = dvsr_conv_fprop(&d, ws + w.wp, stream))) return rc;
    return dvsr_nhwc_to_nchw(ws + w.y, y, B, Co, Ho, Wo, stream);
}

extern "C" int dvsr_mdcn_backward_nchw(const float* x, const float* offset, const float* mask, const float* weight,
                                       const float* gy, float* gx, float* goffset, float* gmask, float* gweight,
                                       float* gbias, int B, int C, int H, int W, int Co, int kh, int kw, int stride_h,
                                       int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int groups, int dg,
                                       void* workspace, long long workspace_bytes, void* stream) {
    DVSR_REQUIRE(x && offset && mask && weight && gy && workspace, "mdcn_backward_nchw: null pointer");
    if (int rc0 = iso("mdcn_backward_nchw", stride_h, stride_w, pad_h, pad_w, dil_h, dil_w)) return rc0;
    const int stride = stride_h, pad = pad_h, dil = dil_h;
    DVSR_REQUIRE(gx && goffset && gmask && gweight, "mdcn_backward_nchw: null gradient buffer");
    if (groups != 1) { set_error("mdcn_backward_nchw: groups=%d is not supported", groups); return DVSR_ERR_UNSUPPORTED; }
    DVSR_REQUIRE(dg > 0 && C % dg == 0, "mdcn_backward_nchw: channels %d not divisible by deformable groups %d", C, dg);
    int Ho, Wo;
    const Ws w = plan(B, C, H, W, Co, kh, kw, stride, pad, dil, dg, 1, &Ho, &Wo);
    DVSR_REQUIRE(workspace_bytes >= w.total * (long long)sizeof(float), "mdcn_backward_nchw: workspace too small (%lld < %lld)",
                 workspace_bytes, w.total * (long long)sizeof(float));
    float* ws = (float*)workspace;
    const int KK = kh * kw;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = dvsr_nchw_to_nhwc(x, ws + w.x, B, C, H, W, stream))) return rc;
    if ((rc = dvsr_nchw_to_nhwc(offset, ws + w.off, B, 2 * dg * KK, Ho, Wo, stream))) return rc;
    if ((rc = dvsr_nchw_to_nhwc(mask, ws + w.mask, B, dg * KK, Ho, Wo, stream))) return rc;
    if ((rc = dvsr_nchw_to_nhwc(gy, ws + w.gy, B, Co, Ho, Wo, stream))) return rc;
    const dvsr_wlayout wl = conv2d_layout(Co, C, kh, kw);
    if ((rc = dvsr_pack_weights(weight, ws + w.wd, &wl, 1, 0, stream))) return rc;
    dvsr_conv_desc d;
    fill_desc(&d, ws + w.x, ws + w.off, ws + w.mask, B, C, H, W, Ho, Wo, Co, kh, kw, stride, pad, dil, dg);
    if (cudaMemsetAsync(ws + w.gx, 0, sizeof(float) * (size_t)B * H * W * C, st) != cudaSuccess) return check_launch("memset gx");
    if ((rc = dvsr_mdcn_bwd_data(&d, ws + w.gy, Co, ws + w.wd, ws + w.gx, C, ws + w.goff, 2 * dg * KK, ws + w.gmask,
                                 dg * KK, stream))) return rc;
    if (cudaMemsetAsync(gweight, 0, sizeof(float) * (size_t)Co * C * KK, st) != cudaSuccess) return check_launch("memset gw");
    if ((rc = dvsr_conv_wgrad(&d, ws + w.gy, Co, gweight, &wl, stream))) return rc;
    if (gbias) {
        if (cudaMemsetAsync(gbias, 0, sizeof(float) * (size_t)Co, st) != cudaSuccess) return check_launch("memset gb");
        if ((rc = dvsr_act_bwd(ws + w.gy, nullptr, nullptr, nullptr, gbias, (long long)B * Ho * Wo, Co, DVSR_ACT_NONE, 0.f, 0, 0, Ho, Wo, stream))) return rc;
    }
    if ((rc = dvsr_nhwc_to_nchw(ws + w.gx, gx, B, C, H, W, stream))) return rc;
    if ((rc = dvsr_nhwc_to_nchw(ws + w.goff, goffset, B, 2 * dg * KK, Ho, Wo, stream))) return rc;
    return dvsr_nhwc_to_nchw(ws + w.gmask, gmask, B, dg * KK, Ho, Wo, stream);
}
