// dynavsr_b200/csrc/elementwise.cu
//
// HBM-bound helpers of the hot path, all NHWC: layout changes at the NCHW boundary, bilinear
// resampling (F.interpolate(..., mode='bilinear', align_corners=False): EDVR_arch.py:107-120,192,197,311),
// the fused max+avg 3x3/s2 pooling of TSA (EDVR_arch.py:149-150,184-185,190-191), the reflection /
// replication pads of MFDN (LRimg_estimator.py:75-76), per-frame mean removal (LRimg_estimator.py:99-100)
// and the activation-derivative + bias-gradient pass that precedes every conv backward.
// Each adjoint is written in gather form (deterministic, no atomics) unless stated.
#include "common.cuh"

namespace dvsr {

// ---------------------------------------------------------------- batched 2-D transpose
// dst[b][c][r] = src[b][r][c]   (src: rows x cols, dst: cols x rows)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    __shared__ float tile[32][33];
    const long long boff = (long long)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[boff + (long long)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[boff + (long long)c * rows + r] = tile[threadIdx.x][i];
    }
}

static int launch_transpose(const float* src, float* dst, int batch, int rows, int cols, cudaStream_t st) {
    // gridDim.y/z limits: rows/32 can exceed 65535 only for > 2M rows; fold the batch into z.
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32), batch), block(32, 8);
    if (grid.y > 65535 || grid.z > 65535) { set_error("transpose: tensor too large (%d x %d x %d)", batch, rows, cols); return DVSR_ERR_INVALID; }
    transpose_kernel<<<grid, block, 0, st>>>(src, dst, rows, cols);
    return check_launch("transpose");
}

// ---------------------------------------------------------------- bilinear resampling
__device__ __forceinline__ void src_index(int o, int scale, int size, int& i0, int& i1, float& l) {
    // PyTorch area_pixel_compute_source_index, align_corners=False, scale_factor given
    float s = ((float)o + 0.5f) * (1.0f / (float)scale) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + ((i0 < size - 1) ? 1 : 0);
    l = s - (float)i0;
}

template <int VEC>
__global__ void upsample_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                int scale, float mul, int accumulate) {
    const int Cv = C / VEC, Ho = H * scale, Wo = W * scale;
    const long long total = (long long)N * Ho * Wo * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long p = i / Cv;
        const int ow = (int)(p % Wo); p /= Wo;
        const int oh = (int)(p % Ho);
        const int n = (int)(p / Ho);
        int h0, h1, w0, w1; float lh, lw;
        src_index(oh, scale, H, h0, h1, lh);
        src_index(ow, scale, W, w0, w1, lw);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float* b = x + (long long)n * H * W * C + cv * VEC;
        float* yo = y + (((long long)n * Ho + oh) * Wo + ow) * C + cv * VEC;
        if (VEC == 4) {
            const float4 a = ldg4(b + ((long long)h0 * W + w0) * C), bb = ldg4(b + ((long long)h0 * W + w1) * C);
            const float4 c = ldg4(b + ((long long)h1 * W + w0) * C), dd = ldg4(b + ((long long)h1 * W + w1) * C);
            float4 r;
            r.x = mul * (hh * (hw * a.x + lw * bb.x) + lh * (hw * c.x + lw * dd.x));
            r.y = mul * (hh * (hw * a.y + lw * bb.y) + lh * (hw * c.y + lw * dd.y));
            r.z = mul * (hh * (hw * a.z + lw * bb.z) + lh * (hw * c.z + lw * dd.z));
            r.w = mul * (hh * (hw * a.w + lw * bb.w) + lh * (hw * c.w + lw * dd.w));
            if (accumulate) { const float4 o = *reinterpret_cast<float4*>(yo); r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
            *reinterpret_cast<float4*>(yo) = r;
        } else {
            const float a = __ldg(b + ((long long)h0 * W + w0) * C), bb = __ldg(b + ((long long)h0 * W + w1) * C);
            const float c = __ldg(b + ((long long)h1 * W + w0) * C), dd = __ldg(b + ((long long)h1 * W + w1) * C);
            float r = mul * (hh * (hw * a + lw * bb) + lh * (hw * c + lw * dd));
            if (accumulate) r += *yo;
            *yo = r;
        }
    }
}

// adjoint in gather form: input pixel i collects from every output o whose stencil touches i
template <int VEC>
__global__ void upsample_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int N, int H, int W, int C,
                                    int scale, float mul) {
    const int Cv = C / VEC, Ho = H * scale, Wo = W * scale;
    const long long total = (long long)N * H * W * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long p = i / Cv;
        const int iw = (int)(p % W); p /= W;
        const int ih = (int)(p % H);
        const int n = (int)(p / H);
        // candidate outputs: src(o) in [i-1, i+1)  <=>  o in [s*i - s/2 - 0.5, s*i + 1.5 s - 0.5); widen by 1 and test exactly
        const int oh_lo = max(0, scale * ih - scale / 2 - 1), oh_hi = min(Ho - 1, scale * ih + (3 * scale) / 2);
        const int ow_lo = max(0, scale * iw - scale / 2 - 1), ow_hi = min(Wo - 1, scale * iw + (3 * scale) / 2);
        float acc[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
        for (int oh = oh_lo; oh <= oh_hi; ++oh) {
            int h0, h1; float lh;
            src_index(oh, scale, H, h0, h1, lh);
            float wh = 0.f;
            if (h0 == ih) wh += 1.f - lh;
            if (h1 == ih) wh += lh;
            if (wh == 0.f) continue;
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                int w0, w1; float lw;
                src_index(ow, scale, W, w0, w1, lw);
                float ww = 0.f;
                if (w0 == iw) ww += 1.f - lw;
                if (w1 == iw) ww += lw;
                if (ww == 0.f) continue;
                const float* g = gy + (((long long)n * Ho + oh) * Wo + ow) * C + cv * VEC;
                if (VEC == 4) {
                    const float4 v = ldg4(g);
                    acc[0] = fmaf(wh * ww, v.x, acc[0]); acc[1 % VEC] = fmaf(wh * ww, v.y, acc[1 % VEC]);
                    acc[2 % VEC] = fmaf(wh * ww, v.z, acc[2 % VEC]); acc[3 % VEC] = fmaf(wh * ww, v.w, acc[3 % VEC]);
                } else {
                    acc[0] = fmaf(wh * ww, __ldg(g), acc[0]);
                }
            }
        }
        float* o = gx + (((long long)n * H + ih) * W + iw) * C + cv * VEC;
#pragma unroll
        for (int q = 0; q < VEC; ++q) o[q] = mul * acc[q];
    }
}

// ---------------------------------------------------------------- fused 3x3/s2/p1 max + avg pooling
template <int VEC>
__global__ void pool_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo) {
    const int Cv = C / VEC;
    const long long total = (long long)N * Ho * Wo * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long p = i / Cv;
        const int ow = (int)(p % Wo); p /= Wo;
        const int oh = (int)(p % Ho);
        const int n = (int)(p / Ho);
        float mx[VEC], sm[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) { mx[q] = -INFINITY; sm[q] = 0.f; }
        for (int dh = 0; dh < 3; ++dh) {
            const int ih = oh * 2 - 1 + dh;
            if (ih < 0 || ih >= H) continue;
            for (int dw = 0; dw < 3; ++dw) {
                const int iw = ow * 2 - 1 + dw;
                if (iw < 0 || iw >= W) continue;
                const float* s = x + (((long long)n * H + ih) * W + iw) * C + cv * VEC;
                float v[VEC];
                if (VEC == 4) { const float4 t = ldg4(s); v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w; }
                else v[0] = __ldg(s);
#pragma unroll
                for (int q = 0; q < VEC; ++q) { mx[q] = v[q] > mx[q] ? v[q] : mx[q]; sm[q] += v[q]; }
            }
        }
        float* o = y + (((long long)n * Ho + oh) * Wo + ow) * (2 * C) + cv * VEC;
#pragma unroll
        for (int q = 0; q < VEC; ++q) { o[q] = mx[q]; o[C + q] = sm[q] * (1.f / 9.f); }   // count_include_pad=True
    }
}

template <int VEC>
__global__ void pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                int N, int H, int W, int C, int Ho, int Wo) {
    const int Cv = C / VEC;
    const long long total = (long long)N * H * W * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long p = i / Cv;
        const int iw = (int)(p % W); p /= W;
        const int ih = (int)(p % H);
        const int n = (int)(p / H);
        float acc[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
        // windows containing (ih, iw): oh*2-1 <= ih <= oh*2+1
        for (int oh = max(0, ih / 2); oh <= min(Ho - 1, (ih + 1) / 2); ++oh) {
            for (int ow = max(0, iw / 2); ow <= min(Wo - 1, (iw + 1) / 2); ++ow) {
                const float* g = gy + (((long long)n * Ho + oh) * Wo + ow) * (2 * C) + cv * VEC;
                // first maximum in (h, w) scan order wins, as in PyTorch's max_pool2d
                float mx[VEC]; int arg[VEC];
#pragma unroll
                for (int q = 0; q < VEC; ++q) { mx[q] = -INFINITY; arg[q] = -1; }
                for (int dh = 0; dh < 3; ++dh) {
                    const int h = oh * 2 - 1 + dh;
                    if (h < 0 || h >= H) continue;
                    for (int dw = 0; dw < 3; ++dw) {
                        const int w = ow * 2 - 1 + dw;
                        if (w < 0 || w >= W) continue;
                        const float* s = x + (((long long)n * H + h) * W + w) * C + cv * VEC;
                        float v[VEC];
                        if (VEC == 4) { const float4 t = ldg4(s); v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w; }
                        else v[0] = __ldg(s);
#pragma unroll
                        for (int q = 0; q < VEC; ++q)
                            if (v[q] > mx[q]) { mx[q] = v[q]; arg[q] = h * W + w; }
                    }
                }
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    acc[q] += __ldg(g + C + q) * (1.f / 9.f);
                    if (arg[q] == ih * W + iw) acc[q] += __ldg(g + q);
                }
            }
        }
        float* o = gx + (((long long)n * H + ih) * W + iw) * C + cv * VEC;
#pragma unroll
        for (int q = 0; q < VEC; ++q) o[q] = acc[q];
    }
}

// ---------------------------------------------------------------- padding
__device__ __forceinline__ int pad_map(int o, int p, int size, int mode) {
    int i = o - p;
    if (mode == 0) {  // reflect
        if (i < 0) i = -i;
        if (i >= size) i = 2 * (size - 1) - i;
    } else {          // replicate
        i = i < 0 ? 0 : (i >= size ? size - 1 : i);
    }
    return i;
}

template <int VEC>
__global__ void pad2d_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int p, int mode) {
    const int Cv = C / VEC, Ho = H + 2 * p, Wo = W + 2 * p;
    const long long total = (long long)N * Ho * Wo * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long q = i / Cv;
        const int ow = (int)(q % Wo); q /= Wo;
        const int oh = (int)(q % Ho);
        const int n = (int)(q / Ho);
        const float* s = x + (((long long)n * H + pad_map(oh, p, H, mode)) * W + pad_map(ow, p, W, mode)) * C + cv * VEC;
        float* o = y + i * VEC;
        if (VEC == 4) *reinterpret_cast<float4*>(o) = ldg4(s); else *o = __ldg(s);
    }
}

// candidate padded positions o with pad_map(o) == i (at most 1 + 2p of them)
__device__ __forceinline__ int pad_candidates(int i, int p, int size, int mode, int* out) {
    int n = 0;
    out[n++] = i + p;
    if (mode == 0) {  // reflect: o - p = -i  or  o - p = 2(size-1) - i
        if (i >= 1 && i <= p) out[n++] = p - i;
        if (i <= size - 2 && i >= size - 1 - p) out[n++] = p + 2 * (size - 1) - i;
    } else {          // replicate: the whole border band collapses onto the edge pixel
        if (i == 0) for (int o = 0; o < p; ++o) out[n++] = o;
        if (i == size - 1) for (int o = size + p; o < size + 2 * p; ++o) out[n++] = o;
    }
    return n;
}

// adjoint (gather): sum over every padded position that maps to input (ih, iw)
template <int VEC>
__global__ void pad2d_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int N, int H, int W, int C, int p, int mode) {
    const int Cv = C / VEC, Ho = H + 2 * p, Wo = W + 2 * p;
    const long long total = (long long)N * H * W * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long q = i / Cv;
        const int iw = (int)(q % W); q /= W;
        const int ih = (int)(q % H);
        const int n = (int)(q / H);
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
        int ch[8], cw_[8];
        const int nh = pad_candidates(ih, p, H, mode, ch), nw = pad_candidates(iw, p, W, mode, cw_);
        for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b) {
                const float* g = gy + (((long long)n * Ho + ch[a]) * Wo + cw_[b]) * C + cv * VEC;
                if (VEC == 4) { const float4 t = ldg4(g); acc[0] += t.x; acc[1 % VEC] += t.y; acc[2 % VEC] += t.z; acc[3 % VEC] += t.w; }
                else acc[0] += __ldg(g);
            }
        float* o = gx + i * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = acc[k];
    }
}

template <int VEC>
__global__ void pad3d_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int H, int W, int C) {
    const int Cv = C / VEC, To = T + 2, Ho = H + 2, Wo = W + 2;
    const long long total = (long long)B * To * Ho * Wo * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long q = i / Cv;
        const int ow = (int)(q % Wo); q /= Wo;
        const int oh = (int)(q % Ho); q /= Ho;
        const int ot = (int)(q % To);
        const int b = (int)(q / To);
        const float* s = x + ((((long long)b * T + pad_map(ot, 1, T, 1)) * H + pad_map(oh, 1, H, 1)) * W + pad_map(ow, 1, W, 1)) * C + cv * VEC;
        float* o = y + i * VEC;
        if (VEC == 4) *reinterpret_cast<float4*>(o) = ldg4(s); else *o = __ldg(s);
    }
}

template <int VEC>
__global__ void pad3d_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int B, int T, int H, int W, int C) {
    const int Cv = C / VEC, To = T + 2, Ho = H + 2, Wo = W + 2;
    const long long total = (long long)B * T * H * W * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        long long q = i / Cv;
        const int iw = (int)(q % W); q /= W;
        const int ih = (int)(q % H); q /= H;
        const int it = (int)(q % T);
        const int b = (int)(q / T);
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
        for (int ot = it; ot <= it + 2; ++ot) {
            if (pad_map(ot, 1, T, 1) != it) continue;
            for (int oh = ih; oh <= ih + 2; ++oh) {
                if (pad_map(oh, 1, H, 1) != ih) continue;
                for (int ow = iw; ow <= iw + 2; ++ow) {
                    if (pad_map(ow, 1, W, 1) != iw) continue;
                    const float* g = gy + ((((long long)b * To + ot) * Ho + oh) * Wo + ow) * C + cv * VEC;
                    if (VEC == 4) { const float4 t = ldg4(g); acc[0] += t.x; acc[1 % VEC] += t.y; acc[2 % VEC] += t.z; acc[3 % VEC] += t.w; }
                    else acc[0] += __ldg(g);
                }
            }
        }
        float* o = gx + i * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = acc[k];
    }
}

// ---------------------------------------------------------------- per-image channel means
// m[n][c] += sum over a slice of the image / HW (m zeroed by the wrapper); grid = (slices, N), coalesced reads
__global__ void spatial_mean_kernel(const float* __restrict__ x, float* __restrict__ m, int HW, int C, int per_block) {
    const int n = blockIdx.y;
    const long long total = (long long)HW * C;
    const long long e0 = (long long)blockIdx.x * per_block * C;
    const long long e1 = min(total, e0 + (long long)per_block * C);
    const float* s = x + (long long)n * total;
    // every thread keeps one running sum per channel residue it meets: element e has channel e % C; with a stride of
    // blockDim.x*... simplest exact scheme: thread t accumulates elements t, t + T, ... and the channel of element e
    // is e % C, so use T = blockDim.x rounded down to a multiple of C to keep each thread on ONE channel
    const int T = (blockDim.x / C) * C;
    float acc = 0.f;
    if ((int)threadIdx.x < T)
        for (long long e = e0 + threadIdx.x; e < e1; e += T) acc += __ldg(s + e);
    extern __shared__ float red[];
    red[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < C) {
        float v = 0.f;
        for (int t = threadIdx.x; t < T; t += C) v += red[t];
        atomicAdd(m + n * C + threadIdx.x, v / (float)HW);
    }
}

__global__ void add_channel_bias_kernel(const float* __restrict__ x, const float* __restrict__ m, float* __restrict__ y,
                                        long long total, long long per_img, int C, float sign) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / per_img), c = (int)(i % C);
        y[i] = x[i] + sign * __ldg(m + n * C + c);
    }
}

// ---------------------------------------------------------------- activation derivative + bias gradient
// gpre[pix][c] = gy[..] * act'(y[..]);  gbias[c] += sum_pix gpre[pix][c]
// shuffle == 2: gy / y are in the PixelShuffled layout [n][2Ho][2Wo][C/4]; gpre is [n][Ho][Wo][C].
template <int VEC>
__global__ void act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ res,
                               float* __restrict__ gpre, float* __restrict__ gbias, long long npix, int C, int act,
                               float slope, int sig_split, int shuffle, int Ho, int Wo, long long pix_per_block) {
    const int cw = C / VEC;                 // channel vectors per pixel
    const int rows = blockDim.x / cw;       // pixels handled per block iteration
    const int t = threadIdx.x;
    const bool active = t < rows * cw;
    const int cv = active ? t % cw : 0, prow = active ? t / cw : 0;
    const long long p0 = (long long)blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
    float bsum[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) bsum[q] = 0.f;
    if (active) {
        // unrolled so that the loads of several pixel rows are in flight together (in-order issue: one row per L2 round trip otherwise)
#pragma unroll 4
        for (long long p = p0 + prow; p < p1; p += rows) {
            float g[VEC], yv[VEC];
            if (shuffle == 2) {
                // channel c = 4*cc + 2*i + j  <-  shuffled pixel (2*oh+i, 2*ow+j), channel cc
                const int hw = Ho * Wo;
                const int n = (int)(p / hw);
                const int r = (int)(p - (long long)n * hw);
                const int oh = r / Wo, ow = r - oh * Wo;
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const int c = cv * VEC + q, cc = c >> 2, ii = (c >> 1) & 1, jj = c & 1;
                    const long long sp = (((long long)n * (2 * Ho) + 2 * oh + ii) * (2 * Wo) + 2 * ow + jj) * (C / 4) + cc;
                    g[q] = __ldg(gy + sp);
                    yv[q] = y ? __ldg(y + sp) : 0.f;
                }
            } else {
                const long long e = p * C + cv * VEC;
                if (VEC == 4) {
                    const float4 a = ldg4(gy + e);
                    g[0] = a.x; g[1 % VEC] = a.y; g[2 % VEC] = a.z; g[3 % VEC] = a.w;
                if (y) { const float4 b = ldg4(y + e); yv[0] = b.x; yv[1 % VEC] = b.y; yv[2 % VEC] = b.z; yv[3 % VEC] = b.w; }
                    // epilogue order is act(v) + res: recover act(v) when a residual was fused after the activation
                    if (y && res) { const float4 r4 = ldg4(res + e); yv[0] -= r4.x; yv[1 % VEC] -= r4.y; yv[2 % VEC] -= r4.z; yv[3 % VEC] -= r4.w; }
                } else {
                    g[0] = __ldg(gy + e);
                    if (y) yv[0] = __ldg(y + e);
                    if (y && res) yv[0] -= __ldg(res + e);
                }
            }
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                float d = 1.f;
                if (act == DVSR_ACT_RELU) d = yv[q] > 0.f ? 1.f : 0.f;
                else if (act == DVSR_ACT_LRELU) d = yv[q] > 0.f ? 1.f : slope;
                else if (act == DVSR_ACT_SIGMOID_SPLIT) d = (cv * VEC + q >= sig_split) ? yv[q] * (1.f - yv[q]) : 1.f;
                g[q] *= d;
                bsum[q] += g[q];
            }
            if (gpre) {
                float* o = gpre + p * C + cv * VEC;
                if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(g[0], g[1 % VEC], g[2 % VEC], g[3 % VEC]);
                else o[0] = g[0];
            }
        }
    }
    if (gbias) {
        extern __shared__ float red[];  // [blockDim.x][VEC]
#pragma unroll
        for (int q = 0; q < VEC; ++q) red[t * VEC + q] = active ? bsum[q] : 0.f;
        __syncthreads();
        if (t < cw) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                float s = 0.f;
                for (int r = 0; r < rows; ++r) s += red[(r * cw + t) * VEC + q];
                atomicAdd(gbias + t * VEC + q, s);
            }
        }
    }
}

// ---------------------------------------------------------------- frame -> 8-bit image (+ squared error vs a reference image)
// utils/util.py:112-142 (tensor2img: clamp to [0, 1], * 255, round half to even, uint8, HWC, RGB or BGR) fused with the
// integer part of calculate_psnr (:262-269): sse += sum (img - ref)^2, exact in 64-bit integers.  One thread makes four
// consecutive output bytes (one 32-bit store).
__global__ void frame_to_u8_kernel(const float* __restrict__ x, unsigned char* __restrict__ out,
                                   const unsigned char* __restrict__ ref, unsigned long long* __restrict__ sse,
                                   long long total, int C, int reverse) {
    unsigned int err = 0;           // <= 4 * 255^2 per iteration; a thread makes < 2^14 iterations for any sane frame
    const long long quads = (total + 3) / 4;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
        unsigned int word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long i = q * 4 + e;
            if (i >= total) break;
            const long long pix = i / C;
            const int c = (int)(i - pix * C);
            const float v = __ldg(x + pix * C + (reverse ? C - 1 - c : c));
            const int b = __float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * 255.f);
            word |= (unsigned int)b << (8 * e);
            if (ref) {
                const int d = b - (int)__ldg(ref + i);
                err += (unsigned int)(d * d);
            }
        }
        if (q * 4 + 3 < total) {
            *reinterpret_cast<unsigned int*>(out + q * 4) = word;
        } else {
            for (int e = 0; q * 4 + e < total; ++e) out[q * 4 + e] = (unsigned char)(word >> (8 * e));
        }
    }
    if (sse) {
        unsigned long long e64 = err;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e64 += __shfl_xor_sync(0xffffffffu, e64, o);
        if ((threadIdx.x & 31) == 0 && e64) atomicAdd(sse, e64);
    }
}

template <typename F>
static int launch_1d(long long total, F f) {
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;  // grid-stride; 16 CTAs of 256 threads per SM
    if (blocks < 1) blocks = 1;
    f(blocks);
    return 0;
}

}  // namespace dvsr

using namespace dvsr;
#define ST ((cudaStream_t)stream)
#define A16(p) ((((uintptr_t)(p)) & 15) == 0)

extern "C" int dvsr_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, void* stream) {
    DVSR_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad arguments");
    return launch_transpose(x, y, N, C, H * W, ST);
}
extern "C" int dvsr_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int W, void* stream) {
    DVSR_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad arguments");
    return launch_transpose(x, y, N, H * W, C, ST);
}

extern "C" int dvsr_upsample_bilinear(const float* x, float* y, int N, int H, int W, int C, int scale, float mul,
                                      int accumulate, void* stream) {
    DVSR_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && scale >= 1, "upsample_bilinear: bad arguments");
    const bool v4 = (C % 4 == 0) && A16(x) && A16(y);
    const long long total = (long long)N * H * scale * W * scale * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) upsample_kernel<4><<<b, 256, 0, ST>>>(x, y, N, H, W, C, scale, mul, accumulate);
        else upsample_kernel<1><<<b, 256, 0, ST>>>(x, y, N, H, W, C, scale, mul, accumulate);
    });
    return check_launch("upsample_bilinear");
}
extern "C" int dvsr_upsample_bilinear_bwd(const float* gy, float* gx, int N, int H, int W, int C, int scale, float mul,
                                          void* stream) {
    DVSR_REQUIRE(gy && gx && N > 0 && H > 0 && W > 0 && C > 0 && scale >= 1, "upsample_bilinear_bwd: bad arguments");
    const bool v4 = (C % 4 == 0) && A16(gy) && A16(gx);
    const long long total = (long long)N * H * W * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) upsample_bwd_kernel<4><<<b, 256, 0, ST>>>(gy, gx, N, H, W, C, scale, mul);
        else upsample_bwd_kernel<1><<<b, 256, 0, ST>>>(gy, gx, N, H, W, C, scale, mul);
    });
    return check_launch("upsample_bilinear_bwd");
}

extern "C" int dvsr_pool_maxavg(const float* x, float* y, int N, int H, int W, int C, void* stream) {
    DVSR_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "pool_maxavg: bad arguments");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const bool v4 = (C % 4 == 0) && A16(x) && A16(y);
    const long long total = (long long)N * Ho * Wo * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) pool_kernel<4><<<b, 256, 0, ST>>>(x, y, N, H, W, C, Ho, Wo);
        else pool_kernel<1><<<b, 256, 0, ST>>>(x, y, N, H, W, C, Ho, Wo);
    });
    return check_launch("pool_maxavg");
}
extern "C" int dvsr_pool_maxavg_bwd(const float* x, const float* gy, float* gx, int N, int H, int W, int C, void* stream) {
    DVSR_REQUIRE(x && gy && gx && N > 0 && H > 0 && W > 0 && C > 0, "pool_maxavg_bwd: bad arguments");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const bool v4 = (C % 4 == 0) && A16(x) && A16(gx);
    const long long total = (long long)N * H * W * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) pool_bwd_kernel<4><<<b, 256, 0, ST>>>(x, gy, gx, N, H, W, C, Ho, Wo);
        else pool_bwd_kernel<1><<<b, 256, 0, ST>>>(x, gy, gx, N, H, W, C, Ho, Wo);
    });
    return check_launch("pool_maxavg_bwd");
}

extern "C" int dvsr_pad2d(const float* x, float* y, int N, int H, int W, int C, int p, int mode, void* stream) {
    DVSR_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && p >= 0 && (mode == 0 || mode == 1), "pad2d: bad arguments");
    DVSR_REQUIRE(mode == 1 || (p < H && p < W), "pad2d: reflection pad %d needs H, W > pad", p);
    const bool v4 = (C % 4 == 0) && A16(x) && A16(y);
    const long long total = (long long)N * (H + 2 * p) * (W + 2 * p) * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) pad2d_kernel<4><<<b, 256, 0, ST>>>(x, y, N, H, W, C, p, mode);
        else pad2d_kernel<1><<<b, 256, 0, ST>>>(x, y, N, H, W, C, p, mode);
    });
    return check_launch("pad2d");
}
extern "C" int dvsr_pad2d_bwd(const float* gy, float* gx, int N, int H, int W, int C, int p, int mode, void* stream) {
    DVSR_REQUIRE(gy && gx && N > 0 && H > 0 && W > 0 && C > 0 && p >= 0 && p <= 3 && (mode == 0 || mode == 1), "pad2d_bwd: bad arguments (pad <= 3)");
    const bool v4 = (C % 4 == 0) && A16(gy) && A16(gx);
    const long long total = (long long)N * H * W * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) pad2d_bwd_kernel<4><<<b, 256, 0, ST>>>(gy, gx, N, H, W, C, p, mode);
        else pad2d_bwd_kernel<1><<<b, 256, 0, ST>>>(gy, gx, N, H, W, C, p, mode);
    });
    return check_launch("pad2d_bwd");
}
extern "C" int dvsr_pad3d_replicate(const float* x, float* y, int B, int T, int H, int W, int C, void* stream) {
    DVSR_REQUIRE(x && y && B > 0 && T > 0 && H > 0 && W > 0 && C > 0, "pad3d_replicate: bad arguments");
    const bool v4 = (C % 4 == 0) && A16(x) && A16(y);
    const long long total = (long long)B * (T + 2) * (H + 2) * (W + 2) * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) pad3d_kernel<4><<<b, 256, 0, ST>>>(x, y, B, T, H, W, C);
        else pad3d_kernel<1><<<b, 256, 0, ST>>>(x, y, B, T, H, W, C);
    });
    return check_launch("pad3d_replicate");
}
// Replication pad + temporal taps folded into channels (see dvsr_tcat_pad3 in the header): one thread per output pixel.
__global__ void tcat_pad3_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int H, int W, int C) {
    const int Ho = H + 2, Wo = W + 2;
    const long long total = (long long)B * T * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long q = i;
        const int ow = (int)(q % Wo); q /= Wo;
        const int oh = (int)(q % Ho); q /= Ho;
        const int t = (int)(q % T);
        const int b = (int)(q / T);
        const int ih = min(max(oh - 1, 0), H - 1), iw = min(max(ow - 1, 0), W - 1);
        float4* o = reinterpret_cast<float4*>(y + i * 12);
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const int it = min(max(t + kt - 1, 0), T - 1);
            const float* s = x + ((((long long)b * T + it) * H + ih) * W + iw) * C;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < C; ++c) v[c] = __ldg(s + c);
            o[kt] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// Blur-and-subsample degradation HR -> LR (random_kernel_generator.py:83-130): reflection pad L/2, depthwise L x L kernel
// (the same for every channel, optionally one per frame), stride `scale`, optional 8-bit quantisation (vsrbase.py:188).
// One thread per output pixel; the kernel taps sit in shared memory, the HR reads are served by L1 / L2.
__global__ void degrade_kernel(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ y, int T, int H, int W,
                               int C, int L, int scale, int Tk, int kmode, int Ho, int Wo, int quantize) {
    extern __shared__ float ks[];            // [L*L] of this block's frame
    const int t = blockIdx.z;
    const int ki = kmode == 0 ? 0 : (kmode == 1 ? t : ((t - 1) % Tk + Tk) % Tk);
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < L * L; i += blockDim.x * blockDim.y) ks[i] = __ldg(k + (long long)ki * L * L + i);
    __syncthreads();
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ox >= Wo || oy >= Ho) return;
    const int p = L / 2;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* img = x + (long long)t * H * W * C;
    for (int a = 0; a < L; ++a) {
        int iy = oy * scale - p + a;
        iy = iy < 0 ? -iy : (iy >= H ? 2 * (H - 1) - iy : iy);          // ReflectionPad2d
        const float* row = img + (long long)iy * W * C;
        for (int b = 0; b < L; ++b) {
            int ix = ox * scale - p + b;
            ix = ix < 0 ? -ix : (ix >= W ? 2 * (W - 1) - ix : ix);
            const float w = ks[a * L + b];
            const float* q = row + ix * C;
            for (int c = 0; c < C; ++c) acc[c] = fmaf(w, __ldg(q + c), acc[c]);
        }
    }
    float* o = y + (((long long)t * Ho + oy) * Wo + ox) * C;
    for (int c = 0; c < C; ++c) o[c] = quantize ? rintf(acc[c] * 255.f) / 255.f : acc[c];
}

extern "C" int dvsr_degrade(const float* x, const float* k, float* y, int T, int H, int W, int C, int L, int scale, int Tk,
                            int kmode, int quantize, void* stream) {
    DVSR_REQUIRE(x && k && y && T > 0 && H > 0 && W > 0 && C > 0 && C <= 4 && L > 0 && scale > 0 && Tk > 0, "degrade: bad arguments");
    DVSR_REQUIRE(L / 2 < H && L / 2 < W, "degrade: reflection padding %d needs an image larger than that", L / 2);
    DVSR_REQUIRE(kmode >= 0 && kmode <= 2 && (size_t)L * L * sizeof(float) <= 48 * 1024, "degrade: bad kernel arguments");
    const int Ho = (H + 2 * (L / 2) - L) / scale + 1, Wo = (W + 2 * (L / 2) - L) / scale + 1;
    DVSR_REQUIRE(Ho > 0 && Wo > 0 && T <= 65535, "degrade: empty output");
    dim3 block(32, 8), grid((Wo + 31) / 32, (Ho + 7) / 8, T);
    degrade_kernel<<<grid, block, (size_t)L * L * sizeof(float), ST>>>(x, k, y, T, H, W, C, L, scale, Tk, kmode, Ho, Wo, quantize);
    return check_launch("degrade");
}

extern "C" int dvsr_tcat_pad3(const float* x, float* y, int B, int T, int H, int W, int C, void* stream) {
    DVSR_REQUIRE(x && y && B > 0 && T > 0 && H > 0 && W > 0 && C > 0 && C <= 4, "tcat_pad3: bad arguments (C <= 4)");
    DVSR_REQUIRE(A16(y), "tcat_pad3: output must be 16-byte aligned");
    const long long total = (long long)B * T * (H + 2) * (W + 2);
    long long blocks = (total + 255) / 256;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    tcat_pad3_kernel<<<(int)blocks, 256, 0, ST>>>(x, y, B, T, H, W, C);
    return check_launch("tcat_pad3");
}

extern "C" int dvsr_pad3d_replicate_bwd(const float* gy, float* gx, int B, int T, int H, int W, int C, void* stream) {
    DVSR_REQUIRE(gy && gx && B > 0 && T > 0 && H > 0 && W > 0 && C > 0, "pad3d_replicate_bwd: bad arguments");
    const bool v4 = (C % 4 == 0) && A16(gy) && A16(gx);
    const long long total = (long long)B * T * H * W * (v4 ? C / 4 : C);
    launch_1d(total, [&](int b) {
        if (v4) pad3d_bwd_kernel<4><<<b, 256, 0, ST>>>(gy, gx, B, T, H, W, C);
        else pad3d_bwd_kernel<1><<<b, 256, 0, ST>>>(gy, gx, B, T, H, W, C);
    });
    return check_launch("pad3d_replicate_bwd");
}

extern "C" int dvsr_spatial_mean(const float* x, float* m, int N, int HW, int C, void* stream) {
    DVSR_REQUIRE(x && m && N > 0 && HW > 0 && C > 0, "spatial_mean: bad arguments");
    DVSR_REQUIRE(C <= 256, "spatial_mean: C=%d > 256", C);
    if (cudaMemsetAsync(m, 0, sizeof(float) * (size_t)N * C, ST) != cudaSuccess) return check_launch("spatial_mean memset");
    const int per_block = 2048;     // pixels per block
    dim3 grid(cdiv(HW, per_block), N);
    spatial_mean_kernel<<<grid, 256, 256 * sizeof(float), ST>>>(x, m, HW, C, per_block);
    return check_launch("spatial_mean");
}
extern "C" int dvsr_add_channel_bias(const float* x, const float* m, float* y, int N, int HW, int C, float sign, void* stream) {
    DVSR_REQUIRE(x && m && y && N > 0 && HW > 0 && C > 0, "add_channel_bias: bad arguments");
    const long long total = (long long)N * HW * C;
    launch_1d(total, [&](int b) { add_channel_bias_kernel<<<b, 256, 0, ST>>>(x, m, y, total, (long long)HW * C, C, sign); });
    return check_launch("add_channel_bias");
}

extern "C" int dvsr_frame_to_u8(const float* x, unsigned char* out, const unsigned char* ref, unsigned long long* sse,
                                long long npix, int C, int reverse, void* stream) {
    DVSR_REQUIRE(x && out && npix > 0 && C > 0, "frame_to_u8: bad arguments");
    DVSR_REQUIRE((((uintptr_t)out) & 3) == 0, "frame_to_u8: the image buffer must be 4-byte aligned");
    DVSR_REQUIRE(!sse || ref, "frame_to_u8: a squared-error accumulator needs a reference image");
    const long long total = npix * C;
    launch_1d((total + 3) / 4, [&](int b) { frame_to_u8_kernel<<<b, 256, 0, ST>>>(x, out, ref, sse, total, C, reverse); });
    return check_launch("frame_to_u8");
}

extern "C" int dvsr_act_bwd(const float* gy, const float* y, const float* res, float* gpre, float* gbias, long long npix,
                            int C, int act, float slope, int sig_split, int shuffle, int Ho, int Wo, void* stream) {
    DVSR_REQUIRE(gy && npix > 0 && C > 0, "act_bwd: bad arguments");
    DVSR_REQUIRE(act == DVSR_ACT_NONE || y, "act_bwd: activation derivative needs the saved output");
    DVSR_REQUIRE(shuffle == 0 || (shuffle == 2 && C % 4 == 0 && gpre && gpre != gy && !res), "act_bwd: bad shuffle arguments");
    const bool v4 = (C % 4 == 0) && (C / 4 <= 256) && A16(gy) && (!y || A16(y)) && (!res || A16(res)) && (!gpre || A16(gpre)) && shuffle == 0;
    const int vec = v4 ? 4 : 1;
    const int cw = C / vec;
    const int threads = cw <= 256 ? 256 : (cw + 31) / 32 * 32;      // one thread per channel vector of a pixel row
    DVSR_REQUIRE(threads <= 1024, "act_bwd: C=%d too wide", C);
    const int rows = threads / cw;
    long long blocks = (npix + (long long)rows * 8 - 1) / ((long long)rows * 8);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    if (blocks < 1) blocks = 1;
    long long per = (npix + blocks - 1) / blocks;
    per = (per + rows - 1) / rows * rows;
    blocks = (npix + per - 1) / per;
    const size_t smem = gbias ? sizeof(float) * threads * vec : 0;
    if (v4) act_bwd_kernel<4><<<(int)blocks, threads, smem, ST>>>(gy, y, res, gpre, gbias, npix, C, act, slope, sig_split, shuffle, Ho, Wo, per);
    else {
        act_bwd_kernel<1><<<(int)blocks, threads, smem, ST>>>(gy, y, res, gpre, gbias, npix, C, act, slope, sig_split, shuffle, Ho, Wo, per);
    }
    return check_launch("act_bwd");
}
