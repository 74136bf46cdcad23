// dynavsr_b200/csrc/conv_simt.cu
//
// Exact-fp32 implicit-GEMM convolution family on CUDA cores (FFMA), NHWC.
//   * conv_fprop_kernel : y[pix][co] = epilogue( sum_k A[pix][k] * Wp[k][co] )
//                         A is gathered on the fly -- plain shifted taps, "transposed" taps (data gradient
//                         of a strided conv) or the modulated bilinear taps of DCNv2 -- so no im2col /
//                         `columns` buffer ever exists in HBM (the reference materialises 132.7 MB of it
//                         per L1 call: deform_conv_cuda.cpp:527-529).
//   * conv_wgrad_kernel : gw[k][co] += sum_pix A[pix][k] * gy[pix][co]   (same A loader)
//   * pack_weights_kernel, conv_small_co_kernel
// This is the parity path (bit-for-bit fp32 FMA accumulation, no tf32) and the fallback for the odd
// shapes (Cin = 3, Cout = 3, strided, 4x4, Conv3d); the tcgen05 path in conv_tc.cu takes the heavy
// 3x3 / 1x1 layers.
//
// Replaces: cuDNN convs at EDVR_arch.py:68-90,141-159,224-249, arch_util.py:42-43,
//           LRimg_estimator.py:77-88 and the DCN im2col + addmm_ pair deform_conv_cuda.cpp:534-563,
//           deform_conv_cuda_kernel.cu:569-632.
#include "pack_device.cuh"

namespace dvsr {

constexpr int BM = 128;      // output pixels per CTA
constexpr int BN = 64;       // output channels per CTA
constexpr int BK = 16;       // K slice (input channels of one tap)
constexpr int LDA = BM + 4;  // padded leading dim of the transposed A tile
constexpr int NT = 256;

struct PixSlot {
    int n, oh, ow;
    long long lin;  // linear output pixel index
    bool valid;
};

__device__ __forceinline__ PixSlot decode_pixel(long long m, long long M, int Ho, int Wo) {
    PixSlot s;
    s.valid = m < M;
    long long mm = s.valid ? m : 0;
    s.lin = mm;
    int hw = Ho * Wo;
    s.n = (int)(mm / hw);
    int r = (int)(mm - (long long)s.n * hw);
    s.oh = r / Wo;
    s.ow = r - s.oh * Wo;
    return s;
}

// source image of output image n, or -1 when the temporal tap falls outside the clip
__device__ __forceinline__ long long seg_image(const dvsr_conv_seg& sg, int n) {
    int T = sg.T > 0 ? sg.T : 1;
    int q = n / T, r = n - q * T;
    int t = sg.t_fixed >= 0 ? sg.t_fixed : r + sg.dt;
    if (t < 0 || t >= sg.Tsrc) return -1;
    return (long long)q * sg.Tsrc + t;
}

// Load VEC consecutive channels [c, c+VEC) of tap (kh, kw) of segment `sg` for output pixel `ps`.
template <int VEC, bool DEFORM>
__device__ __forceinline__ float4 load_a(const dvsr_conv_desc& d, const dvsr_conv_seg& sg, const PixSlot& ps,
                                         int kh, int kw, int c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ps.valid || c >= sg.C) return v;
    const long long img_i = seg_image(sg, ps.n);
    if (img_i < 0) return v;
    const float* img = sg.ptr + img_i * sg.img_stride;
    if (!DEFORM) {
        int ih, iw;
        bool ok;
        if (!d.transposed) {
            ih = ps.oh * d.stride - d.pad + kh * d.dil;
            iw = ps.ow * d.stride - d.pad + kw * d.dil;
            ok = (ih >= 0) && (ih < d.H) && (iw >= 0) && (iw < d.W);
        } else {
            int nh = ps.oh + d.pad - kh * d.dil, nw = ps.ow + d.pad - kw * d.dil;
            ok = (nh >= 0) && (nw >= 0);
            ih = nh / d.stride;
            iw = nw / d.stride;
            ok = ok && (ih * d.stride == nh) && (iw * d.stride == nw) && (ih < d.H) && (iw < d.W);
        }
        if (!ok) return v;
        const float* p = img + ((long long)ih * d.W + iw) * sg.pix_stride + c;
        if (VEC == 4) {
            v = ldg4(p);
        } else {
            v.x = __ldg(p);
            if (c + 1 < sg.C) v.y = __ldg(p + 1);
            if (c + 2 < sg.C) v.z = __ldg(p + 2);
            if (c + 3 < sg.C) v.w = __ldg(p + 3);
        }
        return v;
    } else {
        const int KK = d.KH * d.KW, k = kh * d.KW + kw;
        const int cpg = sg.C / d.dg;
        const int g = c / cpg;
        const float* op = d.offset + ps.lin * d.off_pix_stride + (g * KK + k) * 2;
        const float dy = __ldg(op), dx = __ldg(op + 1);
        const float m = __ldg(d.mask + ps.lin * d.mask_pix_stride + g * KK + k);
        const float h = (float)(ps.oh * d.stride - d.pad + kh * d.dil) + dy;
        const float w = (float)(ps.ow * d.stride - d.pad + kw * d.dil) + dx;
        BilinTap t = make_tap(h, w, d.H, d.W);
        if (!t.inside) return v;
        float4 a = v, b = v, e = v, f = v;
        if (VEC == 4) {
            if (t.o00 >= 0) a = ldg4(img + (long long)t.o00 * sg.pix_stride + c);
            if (t.o01 >= 0) b = ldg4(img + (long long)t.o01 * sg.pix_stride + c);
            if (t.o10 >= 0) e = ldg4(img + (long long)t.o10 * sg.pix_stride + c);
            if (t.o11 >= 0) f = ldg4(img + (long long)t.o11 * sg.pix_stride + c);
        } else {
            // scalar path: only channel c (callers with VEC == 1 and DEFORM step one channel at a time)
            if (t.o00 >= 0) a.x = __ldg(img + (long long)t.o00 * sg.pix_stride + c);
            if (t.o01 >= 0) b.x = __ldg(img + (long long)t.o01 * sg.pix_stride + c);
            if (t.o10 >= 0) e.x = __ldg(img + (long long)t.o10 * sg.pix_stride + c);
            if (t.o11 >= 0) f.x = __ldg(img + (long long)t.o11 * sg.pix_stride + c);
        }
        // same association as the reference: (w1*v1 + w2*v2 + w3*v3 + w4*v4) * mask
        v.x = (t.w00 * a.x + t.w01 * b.x + t.w10 * e.x + t.w11 * f.x) * m;
        v.y = (t.w00 * a.y + t.w01 * b.y + t.w10 * e.y + t.w11 * f.y) * m;
        v.z = (t.w00 * a.z + t.w01 * b.z + t.w10 * e.z + t.w11 * f.z) * m;
        v.w = (t.w00 * a.w + t.w01 * b.w + t.w10 * e.w + t.w11 * f.w) * m;
        return v;
    }
}

// Uniform K iterator over (segment, tap, 16-channel chunk).
struct KIter {
    int s, tap, c0, kbase;  // kbase = row of Wp for (s, tap, c0)
};
// dense variant: c0 = offset inside the segment's KH*KW*C dense K range
__device__ __forceinline__ bool kiter_next_dense(KIter& it, const dvsr_conv_desc& d) {
    const int kmax = d.KH * d.KW * d.seg[it.s].C;
    it.c0 += BK;
    if (it.c0 < kmax) { it.kbase += BK; return true; }
    it.kbase += kmax - (it.c0 - BK);
    it.c0 = 0;
    it.s++;
    if (d.wshare) it.kbase = 0;
    return it.s < d.nseg;
}
__device__ __forceinline__ bool kiter_next(KIter& it, const dvsr_conv_desc& d) {
    const int KK = d.KH * d.KW;
    int C = d.seg[it.s].C;
    it.c0 += BK;
    if (it.c0 < C) { it.kbase += BK; return true; }
    it.kbase += C - (it.c0 - BK);
    it.c0 = 0;
    it.tap++;
    if (it.tap < KK) return true;
    it.tap = 0;
    it.s++;
    if (d.wshare) it.kbase = 0;
    return it.s < d.nseg;
}

// DENSE (small input-channel counts, e.g. RGB): the K loop runs over the dense index tap*C + c in slices of 16
// instead of one 16-channel slice per tap (which would be 81 % padding for C = 3).
template <int VEC, bool DEFORM, bool DENSE = false>
__global__ void __launch_bounds__(NT, 2) conv_fprop_kernel(const dvsr_conv_desc d, const float* __restrict__ wp) {
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const long long M = (long long)d.N * d.Ho * d.Wo;
    const long long m0 = (long long)blockIdx.x * BM;
    const int co0 = blockIdx.y * BN;

    // A loader mapping: 2 float4 per thread
    const int a_pix = tid >> 2, a_cv = (tid & 3) * 4;
    PixSlot ps[2];
    ps[0] = decode_pixel(m0 + a_pix, M, d.Ho, d.Wo);
    ps[1] = decode_pixel(m0 + a_pix + 64, M, d.Ho, d.Wo);
    // B loader mapping: 1 float4 per thread
    const int b_row = tid >> 4, b_col = (tid & 15) * 4;
    const bool co_vec = ((d.Co & 3) == 0);

    // compute mapping
    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    KIter it = {0, 0, 0, 0};
    float4 ra[2], rb;

    auto load_tiles = [&](const KIter& k) {
        const dvsr_conv_seg& sg = d.seg[k.s];
        if (DENSE) {
            // k.c0 is the dense offset tap*C + c of this slice inside segment k.s
            const int kmax = d.KH * d.KW * sg.C;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float t4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = k.c0 + a_cv + j;
                    t4[j] = 0.f;
                    if (kk < kmax) {
                        const int tap = kk / sg.C, c = kk - tap * sg.C;
                        const int kh = tap / d.KW, kw = tap - kh * d.KW;
                        t4[j] = load_a<1, false>(d, sg, ps[i], kh, kw, c).x;
                    }
                }
                ra[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
            }
            rb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k.c0 + b_row < kmax) {
                const float* p = wp + (long long)(k.kbase + b_row) * d.Co + co0 + b_col;
                if (co_vec) {
                    if (co0 + b_col < d.Co) rb = ldg4(p);
                } else {
                    if (co0 + b_col + 0 < d.Co) rb.x = __ldg(p + 0);
                    if (co0 + b_col + 1 < d.Co) rb.y = __ldg(p + 1);
                    if (co0 + b_col + 2 < d.Co) rb.z = __ldg(p + 2);
                    if (co0 + b_col + 3 < d.Co) rb.w = __ldg(p + 3);
                }
            }
            return;
        }
        const int kh = k.tap / d.KW, kw = k.tap - kh * d.KW;
        if (VEC == 4 || !DEFORM) {
            ra[0] = load_a<VEC, DEFORM>(d, sg, ps[0], kh, kw, k.c0 + a_cv);
            ra[1] = load_a<VEC, DEFORM>(d, sg, ps[1], kh, kw, k.c0 + a_cv);
        } else {
            // scalar deformable path (channels-per-group not a multiple of 4)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float t0 = load_a<1, true>(d, sg, ps[i], kh, kw, k.c0 + a_cv + 0).x;
                float t1 = load_a<1, true>(d, sg, ps[i], kh, kw, k.c0 + a_cv + 1).x;
                float t2 = load_a<1, true>(d, sg, ps[i], kh, kw, k.c0 + a_cv + 2).x;
                float t3 = load_a<1, true>(d, sg, ps[i], kh, kw, k.c0 + a_cv + 3).x;
                ra[i] = make_float4(t0, t1, t2, t3);
            }
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k.c0 + b_row < sg.C) {
            const float* p = wp + (long long)(k.kbase + b_row) * d.Co + co0 + b_col;
            if (co_vec) {
                if (co0 + b_col < d.Co) rb = ldg4(p);
            } else {
                if (co0 + b_col + 0 < d.Co) rb.x = __ldg(p + 0);
                if (co0 + b_col + 1 < d.Co) rb.y = __ldg(p + 1);
                if (co0 + b_col + 2 < d.Co) rb.z = __ldg(p + 2);
                if (co0 + b_col + 3 < d.Co) rb.w = __ldg(p + 3);
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p = a_pix + 64 * i;
            As[buf][a_cv + 0][p] = ra[i].x;
            As[buf][a_cv + 1][p] = ra[i].y;
            As[buf][a_cv + 2][p] = ra[i].z;
            As[buf][a_cv + 3][p] = ra[i].w;
        }
        *reinterpret_cast<float4*>(&Bs[buf][b_row][b_col]) = rb;
    };

    load_tiles(it);
    store_tiles(0);
    __syncthreads();
    int buf = 0;
    bool more = DENSE ? kiter_next_dense(it, d) : kiter_next(it, d);
    while (true) {
        if (more) load_tiles(it);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (!more) break;
        store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
        more = DENSE ? kiter_next_dense(it, d) : kiter_next(it, d);
    }

    // ---- epilogue
    const int co = co0 + tx * 4;
    if (co >= d.Co) return;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (d.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (co + j < d.Co) bv[j] = __ldg(d.bias + co + j);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = acc[i][j] + bv[j];
            if (d.act == DVSR_ACT_SIGMOID_SPLIT) t = (co + j >= d.sig_split) ? sigmoidf_(t) : t;
            else t = act_apply(t, d.act, d.slope);
            v[j] = t;
        }
        if (d.res) {
            const float* r = d.res + m * d.res_pix_stride + co;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (co + j < d.Co) v[j] += __ldg(r + j);
        }
        if (d.shuffle == 2) {
            // PixelShuffle(2): out[n][2*oh + i2][2*ow + j2][c] = v[4c + 2*i2 + j2]   (EDVR_arch.py:247)
            const int hw = d.Ho * d.Wo;
            const int n = (int)(m / hw);
            const int r = (int)(m - (long long)n * hw);
            const int oh = r / d.Wo, ow = r - oh * d.Wo;
            const int c = co >> 2;  // co is a multiple of 4
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (co + j >= d.Co) continue;
                const long long op = ((long long)n * (2 * d.Ho) + 2 * oh + (j >> 1)) * (2 * d.Wo) + 2 * ow + (j & 1);
                float* yp = d.y + op * d.y_pix_stride + c;
                *yp = d.accumulate ? (*yp + v[j]) : v[j];
            }
        } else {
            float* yp = d.y + m * d.y_pix_stride + co;
            if (co_vec && ((d.y_pix_stride & 3) == 0) && !d.accumulate) {
                *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (co + j < d.Co) yp[j] = d.accumulate ? (yp[j] + v[j]) : v[j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: gw[(s,tap,ci)][co] += sum_pix A[pix][(s,tap,ci)] * gy[pix][co]
// CTA tile: 128 k-rows x 64 co, split over pixel ranges (gridDim.z), fp32 atomics into the
// PyTorch-layout gradient.
constexpr int WK = 128;  // k rows per CTA (8 chunks of 16 channels)
constexpr int WP = 16;   // pixels per smem stage

template <int VEC, bool DEFORM, bool DENSE = false>
__global__ void __launch_bounds__(NT, 2)
conv_wgrad_kernel(const dvsr_conv_desc d, const float* __restrict__ gy, int gy_pix_stride, float* __restrict__ gw,
                  const dvsr_wlayout wl, int chunks_total, long long pix_per_split) {
    __shared__ __align__(16) float As[2][WP][WK];
    __shared__ __align__(16) float Gs[2][WP][BN];

    const int tid = threadIdx.x;
    const long long M = (long long)d.N * d.Ho * d.Wo;
    const long long p_begin = (long long)blockIdx.z * pix_per_split;
    const long long p_end = min(M, p_begin + pix_per_split);
    const int co0 = blockIdx.y * BN;
    const int KK = d.KH * d.KW;

    // This thread's fixed A column: chunk index -> (segment, tap, c0)
    const int a_kv = tid & 31;             // float4 column within the 128-wide k tile
    const int a_p = tid >> 5;              // pixel rows a_p and a_p + 8
    const int chunk = blockIdx.x * (WK / BK) + (a_kv >> 2);
    int cs = 0, ctap = 0, cc0 = 0;
    bool chunk_ok = chunk < chunks_total;
    {
        int rem = chunk_ok ? chunk : 0;
        for (int s = 0; s < d.nseg; ++s) {
            if (DENSE) {
                // chunks of 16 dense K indices (tap*C + c); cc0 = dense offset of this chunk
                int n_in_seg = (KK * d.seg[s].C + BK - 1) / BK;
                if (rem < n_in_seg) { cs = s; cc0 = rem * BK; break; }
                rem -= n_in_seg;
            } else {
                int per_tap = (d.seg[s].C + BK - 1) / BK;
                int n_in_seg = per_tap * KK;
                if (rem < n_in_seg) { cs = s; ctap = rem / per_tap; cc0 = (rem - ctap * per_tap) * BK; break; }
                rem -= n_in_seg;
            }
        }
    }
    const int a_c = cc0 + (a_kv & 3) * 4;
    const int kh = ctap / d.KW, kw = ctap - kh * d.KW;
    const int g_p = tid >> 4, g_c = (tid & 15) * 4;  // G loader: 16 pixels x 16 float4
    const bool co_vec = ((d.Co & 3) == 0) && ((gy_pix_stride & 3) == 0);

    const int tx = tid & 15, ty = tid >> 4;  // compute: k rows ty*8..+7, co tx*4..+3
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra[2], rg;
    auto load_tiles = [&](long long p0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const long long p = p0 + a_p + 8 * i;
            if (chunk_ok && p < p_end) {
                PixSlot ps = decode_pixel(p, M, d.Ho, d.Wo);
                if (DENSE) {
                    const int C = d.seg[cs].C, kmax = KK * C;
                    float t4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int kk = a_c + j;
                        t4[j] = 0.f;
                        if (kk < kmax) {
                            const int tap = kk / C, c = kk - tap * C;
                            t4[j] = load_a<1, false>(d, d.seg[cs], ps, tap / d.KW, tap % d.KW, c).x;
                        }
                    }
                    ra[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                } else if (VEC == 4 || !DEFORM) {
                    ra[i] = load_a<VEC, DEFORM>(d, d.seg[cs], ps, kh, kw, a_c);
                } else {
                    ra[i].x = load_a<1, true>(d, d.seg[cs], ps, kh, kw, a_c + 0).x;
                    ra[i].y = load_a<1, true>(d, d.seg[cs], ps, kh, kw, a_c + 1).x;
                    ra[i].z = load_a<1, true>(d, d.seg[cs], ps, kh, kw, a_c + 2).x;
                    ra[i].w = load_a<1, true>(d, d.seg[cs], ps, kh, kw, a_c + 3).x;
                }
            }
        }
        rg = make_float4(0.f, 0.f, 0.f, 0.f);
        const long long p = p0 + g_p;
        if (p < p_end && co0 + g_c < d.Co) {
            const float* q = gy + p * gy_pix_stride + co0 + g_c;
            if (co_vec) rg = ldg4(q);
            else {
                rg.x = __ldg(q);
                if (co0 + g_c + 1 < d.Co) rg.y = __ldg(q + 1);
                if (co0 + g_c + 2 < d.Co) rg.z = __ldg(q + 2);
                if (co0 + g_c + 3 < d.Co) rg.w = __ldg(q + 3);
            }
        }
    };
    auto store_tiles = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][a_p][a_kv * 4]) = ra[0];
        *reinterpret_cast<float4*>(&As[buf][a_p + 8][a_kv * 4]) = ra[1];
        *reinterpret_cast<float4*>(&Gs[buf][g_p][g_c]) = rg;
    };

    if (p_begin < p_end) {
        load_tiles(p_begin);
        store_tiles(0);
        __syncthreads();
        int buf = 0;
        for (long long p0 = p_begin; p0 < p_end; p0 += WP) {
            const bool more = (p0 + WP) < p_end;
            if (more) load_tiles(p0 + WP);
#pragma unroll
            for (int pp = 0; pp < WP; ++pp) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][pp][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][pp][ty * 8 + 4]);
                const float4 g = *reinterpret_cast<const float4*>(&Gs[buf][pp][tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
            }
            if (more) {
                store_tiles(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }

    // scatter-add into the PyTorch-layout gradient.  Row r of the tile = chunk (r/16), channel r%16.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = ty * 8 + i;
        const int ch = blockIdx.x * (WK / BK) + (r >> 4);
        if (ch >= chunks_total) continue;
        int rem = ch, s = 0, tap = 0, c0 = 0;
        for (int q = 0; q < d.nseg; ++q) {
            if (DENSE) {
                int n_in_seg = (KK * d.seg[q].C + BK - 1) / BK;
                if (rem < n_in_seg) { s = q; c0 = rem * BK; break; }
                rem -= n_in_seg;
            } else {
                int per_tap = (d.seg[q].C + BK - 1) / BK;
                int n_in_seg = per_tap * KK;
                if (rem < n_in_seg) { s = q; tap = rem / per_tap; c0 = (rem - tap * per_tap) * BK; break; }
                rem -= n_in_seg;
            }
        }
        int ci = c0 + (r & 15);
        if (DENSE) {
            if (ci >= KK * d.seg[s].C) continue;
            tap = ci / d.seg[s].C;
            ci -= tap * d.seg[s].C;
        } else if (ci >= d.seg[s].C) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co >= d.Co) continue;
            atomicAdd(gw + (long long)co * wl.co_stride + wl.seg_base[s] + (long long)ci * wl.ci_stride + tap, acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Small-Cout direct convolution (conv_last 64 -> 3, EDVR_arch.py:249,307): one thread per output pixel,
// weights in shared memory, plain (non-deformable, non-transposed) taps, single segment.
constexpr int SMALL_CO_MAX = 4;
__global__ void __launch_bounds__(128) conv_small_co_kernel(const dvsr_conv_desc d, const float* __restrict__ wp) {
    extern __shared__ float4 ws4[];  // [Ktotal] rows of (up to) 4 output channels, zero padded
    const dvsr_conv_seg& sg = d.seg[0];
    const int KK = d.KH * d.KW;
    const int Kt = KK * sg.C;
    for (int i = threadIdx.x; i < Kt; i += blockDim.x) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < d.Co && o < SMALL_CO_MAX; ++o) t[o] = wp[(long long)i * d.Co + o];
        ws4[i] = make_float4(t[0], t[1], t[2], t[3]);
    }
    __syncthreads();
    const long long M = (long long)d.N * d.Ho * d.Wo;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    PixSlot ps = decode_pixel(m, M, d.Ho, d.Wo);
    const long long img_i = seg_image(sg, ps.n);
    const float* img = sg.ptr + (img_i < 0 ? 0 : img_i) * sg.img_stride;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int tap = 0; tap < KK; ++tap) {
        const int kh = tap / d.KW, kw = tap - kh * d.KW;
        const int ih = ps.oh * d.stride - d.pad + kh * d.dil, iw = ps.ow * d.stride - d.pad + kw * d.dil;
        if (ih < 0 || ih >= d.H || iw < 0 || iw >= d.W) continue;
        const float* p = img + ((long long)ih * d.W + iw) * sg.pix_stride;
        const float4* wk = ws4 + tap * sg.C;
        for (int c = 0; c < sg.C; c += 4) {
            const float4 x = ldg4(p + c);
            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = wk[c + q];
                a0 = fmaf(xv[q], w.x, a0);
                a1 = fmaf(xv[q], w.y, a1);
                a2 = fmaf(xv[q], w.z, a2);
                a3 = fmaf(xv[q], w.w, a3);
            }
        }
    }
    const float accv[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int o = 0; o < SMALL_CO_MAX; ++o) {
        if (o >= d.Co) break;
        float v = accv[o] + (d.bias ? __ldg(d.bias + o) : 0.f);
        v = act_apply(v, d.act, d.slope);
        if (d.res) v += __ldg(d.res + m * d.res_pix_stride + o);
        float* yp = d.y + m * d.y_pix_stride + o;
        *yp = d.accumulate ? (*yp + v) : v;
    }
}

// Coalesced variant: 4 lanes share one output pixel (each takes a quarter of the input channels), so a warp reads
// 8 pixels x 256 B = 2 KiB contiguous per tap; partial sums are combined with two shuffles.  C % 16 == 0.
__global__ void __launch_bounds__(256) conv_small_co4_kernel(const dvsr_conv_desc d, const float* __restrict__ wp) {
    extern __shared__ float4 ws4[];
    const dvsr_conv_seg& sg = d.seg[0];
    const int KK = d.KH * d.KW;
    const int Kt = KK * sg.C;
    for (int i = threadIdx.x; i < Kt; i += blockDim.x) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < d.Co && o < SMALL_CO_MAX; ++o) t[o] = wp[(long long)i * d.Co + o];
        ws4[i] = make_float4(t[0], t[1], t[2], t[3]);
    }
    __syncthreads();
    const long long M = (long long)d.N * d.Ho * d.Wo;
    const int sub = threadIdx.x & 3;                       // quarter of the channels
    // persistent over pixel groups: the 9 KiB weight staging above is paid once per block, not once per 64 pixels
    for (long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2; m < ((M + 63) / 64) * 64;
         m += ((long long)gridDim.x * blockDim.x) >> 2) {
    const bool valid = m < M;
    PixSlot ps = decode_pixel(m, M, d.Ho, d.Wo);
    const long long img_i = seg_image(sg, ps.n);
    const float* img = sg.ptr + (img_i < 0 ? 0 : img_i) * sg.img_stride;
    const int cq = sg.C >> 2, cb = sub * cq;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (valid) {
        for (int tap = 0; tap < KK; ++tap) {
            const int kh = tap / d.KW, kw = tap - kh * d.KW;
            const int ih = ps.oh * d.stride - d.pad + kh * d.dil, iw = ps.ow * d.stride - d.pad + kw * d.dil;
            if (ih < 0 || ih >= d.H || iw < 0 || iw >= d.W) continue;
            const float* p = img + ((long long)ih * d.W + iw) * sg.pix_stride + cb;
            const float4* wk = ws4 + tap * sg.C + cb;
            for (int c = 0; c < cq; c += 4) {
                const float4 x = ldg4(p + c);
                const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w = wk[c + q];
                    a0 = fmaf(xv[q], w.x, a0); a1 = fmaf(xv[q], w.y, a1); a2 = fmaf(xv[q], w.z, a2); a3 = fmaf(xv[q], w.w, a3);
                }
            }
        }
    }
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (!valid || sub >= d.Co) continue;
    // lane `sub` finishes output channel `sub`
    float v = (sub == 0 ? a0 : sub == 1 ? a1 : sub == 2 ? a2 : a3) + (d.bias ? __ldg(d.bias + sub) : 0.f);
    v = act_apply(v, d.act, d.slope);
    if (d.res) v += __ldg(d.res + m * d.res_pix_stride + sub);
    float* yp = d.y + m * d.y_pix_stride + sub;
    *yp = d.accumulate ? (*yp + v) : v;
    }
}

// ------------------------------------------------------------------------------------------------
static int validate_desc(const dvsr_conv_desc* d) {
    DVSR_REQUIRE(d != nullptr, "conv: null descriptor");
    DVSR_REQUIRE(d->nseg >= 1 && d->nseg <= DVSR_MAX_SEG, "conv: nseg=%d out of range", d->nseg);
    DVSR_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0 && d->Co > 0, "conv: bad shape");
    DVSR_REQUIRE(d->KH > 0 && d->KW > 0 && d->stride > 0 && d->dil > 0, "conv: bad kernel geometry");
    for (int s = 0; s < d->nseg; ++s) {
        DVSR_REQUIRE(d->seg[s].ptr != nullptr && d->seg[s].C > 0 && d->seg[s].pix_stride >= d->seg[s].C,
                     "conv: bad segment %d", s);
    }
    if (d->deform) {
        DVSR_REQUIRE(d->nseg == 1 && !d->transposed, "conv: deformable sampling needs one plain segment");
        DVSR_REQUIRE(d->dg > 0 && d->seg[0].C % d->dg == 0, "conv: channels %d not divisible by deformable groups %d",
                     d->seg[0].C, d->dg);
        DVSR_REQUIRE(d->offset && d->mask, "conv: deformable sampling needs offset and mask");
    }
    if (d->shuffle) DVSR_REQUIRE(d->shuffle == 2 && d->Co % 4 == 0, "conv: only PixelShuffle(2) with Co %% 4 == 0");
    DVSR_REQUIRE(d->out_step == 0, "conv: strided output placement is only implemented by the tensor-core path");
    return 0;
}

// every segment has fewer than 16 channels (RGB inputs): use the dense-K kernels
// ------------------------------------------------------------------------------------------------------------------------
// Weight gradient of a conv with very few OUTPUT channels (conv_last 64 -> 3, MFDN conv6): gw[co][ci][tap] += sum_o x[o + tap][ci]
// gy[o][co].  The implicit-GEMM kernel above spends a 64-wide N tile on 3 columns (120 us at 176 x 320, 3.6 % of the SM time of
// an adapted frame; this kernel: 68 us); this is a memory-bound reduction instead: one thread per input channel (coalesced 256-byte pixel rows), 4
// pixel lanes per CTA, KK x Co accumulators in registers, one block-level reduction and one red.add per (co, ci, tap) and CTA.
constexpr int WSC_MAX_KK = 9, WSC_MAX_CO = 4, WSC_PIX = 192;    // 192 pixels per CTA: ~294 CTAs at 176 x 320 (2 per SM), 48 per pixel lane
__global__ void __launch_bounds__(256) conv_wgrad_small_co_kernel(const dvsr_conv_desc d, const float* __restrict__ gy, int gy_pix_stride,
                                                                  float* __restrict__ gw, const dvsr_wlayout wl) {
    __shared__ float red[4][WSC_MAX_KK * WSC_MAX_CO][64];
    const int c = threadIdx.x & 63, lane_p = threadIdx.x >> 6;
    const dvsr_conv_seg& sg = d.seg[0];
    const int KK = d.KH * d.KW;
    const long long M = (long long)d.N * d.Ho * d.Wo;
    const long long m0 = (long long)blockIdx.x * WSC_PIX, m1 = m0 + WSC_PIX < M ? m0 + WSC_PIX : M;
    const long long img_stride = sg.img_stride > 0 ? sg.img_stride : (long long)d.H * d.W * sg.pix_stride;
    for (int cb = 0; cb < sg.C; cb += 64) {
        const int ci = cb + c;
        float acc[WSC_MAX_KK][WSC_MAX_CO];
#pragma unroll
        for (int t = 0; t < WSC_MAX_KK; ++t)
#pragma unroll
            for (int o = 0; o < WSC_MAX_CO; ++o) acc[t][o] = 0.f;
        if (ci < sg.C) {
            // pixel coordinates advance incrementally (the 64-bit divisions of a per-pixel decode were most of the instruction stream)
            long long m = m0 + lane_p;
            int n = (int)(m / ((long long)d.Ho * d.Wo));
            int r0 = (int)(m - (long long)n * d.Ho * d.Wo);
            int oy = r0 / d.Wo, ox = r0 - oy * d.Wo;
#pragma unroll 2
            for (; m < m1; m += 4, ox += 4) {
                while (ox >= d.Wo) { ox -= d.Wo; if (++oy == d.Ho) { oy = 0; ++n; } }
                float g[WSC_MAX_CO];
#pragma unroll
                for (int o = 0; o < WSC_MAX_CO; ++o) g[o] = o < d.Co ? __ldg(gy + m * gy_pix_stride + o) : 0.f;
                const float* img = sg.ptr + (long long)n * img_stride + ci;
                // all tap loads first, then the FMAs: with load -> FMA per tap the in-order issue serialises nine L2 round trips per
                // pixel (measured 125 us for 56 320 pixels)
                float xv[WSC_MAX_KK];
                int kh = 0, kw = 0;
#pragma unroll
                for (int t = 0; t < WSC_MAX_KK; ++t) {
                    const int iy = oy - d.pad + kh, ix = ox - d.pad + kw;
                    const bool inb = t < KK && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
                    const float* px = img + ((long long)(inb ? iy : 0) * d.W + (inb ? ix : 0)) * sg.pix_stride;
                    xv[t] = __ldg(px);
                    if (!inb) xv[t] = 0.f;
                    if (++kw == d.KW) { kw = 0; ++kh; }
                }
#pragma unroll
                for (int t = 0; t < WSC_MAX_KK; ++t)
#pragma unroll
                    for (int o = 0; o < WSC_MAX_CO; ++o) acc[t][o] = fmaf(xv[t], g[o], acc[t][o]);
            }
        }
        // reduce the 4 pixel lanes in shared memory, then one red.add per element -- issued by all 256 threads in a CTA-dependent
        // rotated order: every CTA of the launch adds into the same KK x Co x C addresses, and in a fixed order they all hit the same
        // address at the same time (measured: 590 CTAs x 1 728 same-order atomics = 149 us, slower than the GEMM kernel it replaces)
#pragma unroll
        for (int t = 0; t < WSC_MAX_KK; ++t)
#pragma unroll
            for (int o = 0; o < WSC_MAX_CO; ++o) red[lane_p][t * WSC_MAX_CO + o][c] = acc[t][o];
        __syncthreads();
        const int total = KK * d.Co * 64;
        const int rot = (int)((blockIdx.x * 997u) % (unsigned)total);
        for (int e0 = threadIdx.x; e0 < total; e0 += 256) {
            int e = e0 + rot;
            if (e >= total) e -= total;
            const int cc = e & 63, to = e >> 6;              // to = t * Co + o
            const int t = to / d.Co, o = to - t * d.Co;
            const int cig = cb + cc;
            if (cig < sg.C) {
                const int si = t * WSC_MAX_CO + o;
                const float v = red[0][si][cc] + red[1][si][cc] + red[2][si][cc] + red[3][si][cc];
                atomicAdd(gw + wl.seg_base[0] + (long long)o * wl.co_stride + (long long)cig * wl.ci_stride + t, v);
            }
        }
        __syncthreads();
    }
}

static bool small_c(const dvsr_conv_desc* d) {
    for (int s = 0; s < d->nseg; ++s)
        if (d->seg[s].C >= 16) return false;
    return true;
}

static bool all_vec4(const dvsr_conv_desc* d) {
    for (int s = 0; s < d->nseg; ++s) {
        const dvsr_conv_seg& g = d->seg[s];
        if ((g.C & 3) || (g.pix_stride & 3) || (g.img_stride & 3) || ((uintptr_t)g.ptr & 15)) return false;
    }
    if (d->deform && ((d->seg[0].C / d->dg) & 3)) return false;
    return true;
}

}  // namespace dvsr

using namespace dvsr;

extern "C" int dvsr_pack_weights(const float* w, float* wp, const dvsr_wlayout* wl, int mode, int seg, void* stream) {
    DVSR_REQUIRE(w && wp && wl, "pack_weights: null pointer");
    DVSR_REQUIRE(wl->nseg >= 1 && wl->nseg <= DVSR_MAX_SEG && wl->taps > 0 && wl->Co > 0, "pack_weights: bad layout");
    long long total = 0;
    if (mode == 0) {
        for (int s = 0; s < wl->nseg; ++s) total += (long long)wl->seg_C[s] * wl->taps * wl->Co;
    } else {
        DVSR_REQUIRE(seg >= 0 && seg < wl->nseg, "pack_weights: bad segment");
        total = (long long)wl->seg_C[seg] * wl->taps * wl->Co;
    }
    dvsr_pack_job j;
    memset(&j, 0, sizeof(j));
    j.w = w; j.wp = wp; j.wl = *wl; j.mode = mode; j.seg = seg; j.total = total;
    return dvsr_pack_job_run(&j, stream);      // the kernel lives in pack_table.cu (no cross-TU device linking)
}

extern "C" int dvsr_conv_fprop(const dvsr_conv_desc* d, const float* wp, void* stream) {
    if (int rc = validate_desc(d)) return rc;
    DVSR_REQUIRE(wp && d->y, "conv_fprop: null weight/output");
    const long long M = (long long)d->N * d->Ho * d->Wo;
    dim3 grid(cdiv(M, BM), cdiv(d->Co, BN));
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = all_vec4(d);
    if (d->deform) {
        if (v4) conv_fprop_kernel<4, true><<<grid, NT, 0, st>>>(*d, wp);
        else conv_fprop_kernel<1, true><<<grid, NT, 0, st>>>(*d, wp);
    } else {
        if (v4) conv_fprop_kernel<4, false><<<grid, NT, 0, st>>>(*d, wp);
        else if (small_c(d)) conv_fprop_kernel<1, false, true><<<grid, NT, 0, st>>>(*d, wp);
        else conv_fprop_kernel<1, false><<<grid, NT, 0, st>>>(*d, wp);
    }
    return check_launch("conv_fprop");
}

extern "C" int dvsr_conv_wgrad(const dvsr_conv_desc* d, const float* gy, int gy_pix_stride, float* gw,
                               const dvsr_wlayout* wl, void* stream) {
    if (int rc = validate_desc(d)) return rc;
    DVSR_REQUIRE(gy && gw && wl, "conv_wgrad: null pointer");
    DVSR_REQUIRE(!d->transposed, "conv_wgrad: descriptor must describe the forward op");
    const int KK = d->KH * d->KW;
    if (!d->deform && d->nseg == 1 && d->Co <= WSC_MAX_CO && KK <= WSC_MAX_KK && d->stride == 1 && d->dil == 1 && wl->ci_bits == 0 &&
        d->seg[0].T <= 1 && d->seg[0].t_fixed < 0 && d->seg[0].C >= 16) {
        const long long Mpix = (long long)d->N * d->Ho * d->Wo;
        conv_wgrad_small_co_kernel<<<(unsigned)((Mpix + WSC_PIX - 1) / WSC_PIX), 256, 0, (cudaStream_t)stream>>>(*d, gy, gy_pix_stride, gw, *wl);
        return check_launch("conv_wgrad (small Co)");
    }
    const bool dense = !d->deform && !all_vec4(d) && small_c(d);
    int chunks = 0;
    for (int s = 0; s < d->nseg; ++s) chunks += dense ? (KK * d->seg[s].C + BK - 1) / BK : ((d->seg[s].C + BK - 1) / BK) * KK;
    const long long M = (long long)d->N * d->Ho * d->Wo;
    const int gx = cdiv(chunks, WK / BK), gyd = cdiv(d->Co, BN);
    // enough pixel splits for ~4 waves of 2 CTAs per SM, at least 256 pixels per split
    long long splits = (8LL * sm_count() + (long long)gx * gyd - 1) / ((long long)gx * gyd);
    long long max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long per = (M + splits - 1) / splits;
    per = (per + WP - 1) / WP * WP;
    splits = (M + per - 1) / per;
    dim3 grid(gx, gyd, (unsigned)splits);
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = all_vec4(d);
    if (d->deform) {
        if (v4) conv_wgrad_kernel<4, true><<<grid, NT, 0, st>>>(*d, gy, gy_pix_stride, gw, *wl, chunks, per);
        else conv_wgrad_kernel<1, true><<<grid, NT, 0, st>>>(*d, gy, gy_pix_stride, gw, *wl, chunks, per);
    } else {
        if (v4) conv_wgrad_kernel<4, false><<<grid, NT, 0, st>>>(*d, gy, gy_pix_stride, gw, *wl, chunks, per);
        else if (dense) conv_wgrad_kernel<1, false, true><<<grid, NT, 0, st>>>(*d, gy, gy_pix_stride, gw, *wl, chunks, per);
        else conv_wgrad_kernel<1, false><<<grid, NT, 0, st>>>(*d, gy, gy_pix_stride, gw, *wl, chunks, per);
    }
    return check_launch("conv_wgrad");
}

extern "C" int dvsr_conv_small_co(const dvsr_conv_desc* d, const float* wp, void* stream) {
    if (int rc = validate_desc(d)) return rc;
    DVSR_REQUIRE(d->nseg == 1 && !d->deform && !d->transposed && !d->shuffle, "conv_small_co: plain single-segment conv only");
    DVSR_REQUIRE(d->Co <= SMALL_CO_MAX, "conv_small_co: Co=%d > %d", d->Co, SMALL_CO_MAX);
    DVSR_REQUIRE(all_vec4(d), "conv_small_co: input channels must be a multiple of 4 and 16B aligned");
    const long long M = (long long)d->N * d->Ho * d->Wo;
    const size_t smem = (size_t)d->KH * d->KW * d->seg[0].C * 4 * sizeof(float);
    DVSR_REQUIRE(smem <= 48 * 1024, "conv_small_co: weights do not fit in shared memory");
    if (d->seg[0].C % 16 == 0) {
        int blocks = cdiv(M * 4, 256);
        if (blocks > sm_count() * 8) blocks = sm_count() * 8;
        conv_small_co4_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(*d, wp);
    }
    else conv_small_co_kernel<<<cdiv(M, 128), 128, smem, (cudaStream_t)stream>>>(*d, wp);
    return check_launch("conv_small_co");
}
