// dynavsr_b200/csrc/mdcn_bwd_tc.cu
//
// Backward of the modulated deformable convolution (DCNv2) with BOTH of its GEMMs on the tcgen05 tensor cores, in ONE kernel:
// gradients w.r.t. the sampled input, the offsets, the mask AND the weight.  Replaces what the reference does with five passes
// over a 132.7 MB `columns` buffer (deform_conv_cuda.cpp:617-666: addmm_ -> col2im_coord -> col2im -> im2col again -> addmm_;
// deform_conv_cuda_kernel.cu:634-766) and, in this library, the CUDA-core pair mdcn_bwd_data_kernel + deformable conv_wgrad_kernel.
//
// Per 128-pixel tile and tap t (EDVR geometry: 64 -> 64 channels, 8 deformable groups of 8 channels):
//   1. data GEMM      grad_col_t [128 pix x 64 ci] = gy [128 x 64 co] . W_t^T        tcgen05.mma, BF16x3, accumulator in TMEM
//   2. 16 worker warps (one thread per (pixel, group), lane = pixel = TMEM lane) read their 8 grad_col values with tcgen05.ld,
//      gather the four bilinear corners of x (256-bit loads, reference bounds rule kernel.cu:466-496), and produce
//        grad_offset / grad_mask            (col2im_coord, kernel.cu:694-766)     -> plain stores
//        grad_input                          (col2im, kernel.cu:634-692)           -> red.global.add.v4.f32 per corner
//        col_t = mask * bilinear(x)          (the forward's im2col row)            -> shared memory, TRANSPOSED [ci][pix], bf16 hi|lo
//   3. weight GEMM    gw_t [64 ci x 64 co] += col_t^T [64 x 128 pix] . gy^T [128 pix x 64 co]^T   tcgen05.mma, BF16x3, accumulators
//      resident in TMEM for the whole kernel (persistent CTAs), added to the PyTorch-layout gradient once at the end.
// grad_col and col never exist outside TMEM / shared memory.  The taps are split over blockIdx.y (5 + 4 for a 3x3 kernel) so that
// the weight-gradient accumulators (64 TMEM columns per tap) and the double-buffered grad_col accumulator fit the 512 columns.
// Warp roles: 0 = TMA producer (one 16 KiB weight block per tap, pack mode 10: rows 0-63 bf16 hi parts of W_t[ci][co], rows 64-127
// lo parts), 1 = MMA issuer / TMEM owner, 2-17 = workers.
#include "tc_common.cuh"

namespace dvsr {

constexpr int MB_THREADS = 64 + 512;
constexpr int MB_WBYTES = 16384;          // one tap's weight block: 128 rows x 128 B
constexpr int MB_TILE = 16384;            // a [128 rows x 128 B] or 2 x [64 rows x 128 B] bf16 operand tile

struct MbParams {
    const float* x; int pix_stride; long long img_stride;
    int N, H, W, Ho, Wo, KH, KW, stride, pad, dil;
    const float* offset; int off_pix_stride;
    const float* mask; int mask_pix_stride;
    const float* gy; int gy_pix_stride;
    float* gx; int gx_pix_stride;
    float* goff; int goff_pix_stride;
    float* gmask; int gmask_pix_stride;
    float* gw; long long co_stride, ci_stride, seg_base;
    int tiles_total, taps_a;              // taps handled by blockIdx.y = 0 (the rest by blockIdx.y = 1)
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// byte offset of bf16 element (row, k) inside a K-major 128B-swizzled operand made of 64-element K-blocks of `rows` rows each
__device__ __forceinline__ uint32_t kmajor_off(int row, int k, int rows) {
    const int kb = k >> 6, kk = k & 63;
    return (uint32_t)(kb * rows * 128 + row * 128 + ((((kk >> 3) ^ (row & 7))) << 4) + ((kk & 7) << 1));
}

__global__ void __launch_bounds__(MB_THREADS, 1)
mdcn_bwd_tc_kernel(const __grid_constant__ CUtensorMap wmap, const MbParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                           // [2][16 KiB] weight block of a tap
    uint8_t* s_gy = s_w + 2 * MB_WBYTES;           // gy  [128 pix rows][64 co] hi, then lo           (A of the data GEMM)
    uint8_t* s_gyt = s_gy + 2 * MB_TILE;           // gy^T 2 K-blocks x [64 co rows][64 pix] hi, lo    (B of the weight GEMM)
    uint8_t* s_col = s_gyt + 2 * MB_TILE;          // [2 buffers] col^T 2 K-blocks x [64 ci rows][64 pix] hi, lo (A of the weight GEMM)
    uint64_t* bars = (uint64_t*)(s_col + 4 * MB_TILE);
    uint64_t* w_full = bars;            // [2]
    uint64_t* w_empty = bars + 2;       // [2]
    uint64_t* acc_full = bars + 4;      // [2]
    uint64_t* acc_empty = bars + 6;     // [2] 512 arrivals
    uint64_t* col_ready = bars + 8;     // [2] 512 arrivals
    uint64_t* col_free = bars + 10;     // [2]
    uint64_t* gy_ready = bars + 12;     // 512 arrivals
    uint64_t* gy_free = bars + 13;
    uint64_t* wg_done = bars + 14;
    uint32_t* tmem_slot = (uint32_t*)(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KK = p.KH * p.KW;
    const int tap0 = blockIdx.y * p.taps_a;
    const int ntaps = min(p.taps_a, KK - tap0);
    const long long M = (long long)p.N * p.Ho * p.Wo;
    const int hw = p.Ho * p.Wo;

    if (warp == 0 && elect_one()) prefetch_tmap(&wmap);
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
                mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 512);
                mbar_init(&col_ready[i], 512); mbar_init(&col_free[i], 1);
            }
            mbar_init(gy_ready, 512); mbar_init(gy_free, 1); mbar_init(wg_done, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t wg_col0 = 128;                  // weight-gradient accumulators: 64 columns per local tap from here

    if (warp == 0) {
        // ===================== producer: one weight block per (tile, tap) =====================
        if (elect_one()) {
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x)
                for (int tl = 0; tl < ntaps; ++tl, ++it) {
                    const int b = it & 1;
                    mbar_wait_relaxed(&w_empty[b], ((it >> 1) & 1) ^ 1, 64);
                    mbar_expect_tx(&w_full[b], MB_WBYTES);
                    tma_load_2d(&wmap, &w_full[b], s_w + b * MB_WBYTES, 0, (tap0 + tl) * 128);
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_d = make_idesc_bf16(128, 64);       // data GEMM   : M = 128 pixels, N = 64 ci
        const uint32_t idesc_w = make_idesc_bf16(64, 64);        // weight GEMM : M = 64 ci,     N = 64 co
        const uint64_t dc = make_desc(0, 16, 1024, 2);
        const uint64_t gy_hi = dc + (uint64_t)(smem_u32(s_gy) >> 4), gy_lo = gy_hi + (MB_TILE >> 4);
        const uint64_t gyt_hi = dc + (uint64_t)(smem_u32(s_gyt) >> 4), gyt_lo = gyt_hi + (MB_TILE >> 4);
        auto weight_gemm = [&](int s, int tl, int tile_i) {      // stage s produced col^T of local tap tl
            const int b = s & 1;
            mbar_wait(&col_ready[b], (s >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t c_hi = dc + (uint64_t)(smem_u32(s_col + b * 2 * MB_TILE) >> 4), c_lo = c_hi + (MB_TILE >> 4);
                const uint32_t dcol = tmem_base + wg_col0 + (uint32_t)tl * 64u;
#pragma unroll
                for (int k = 0; k < 8; ++k) {                    // K = 128 pixels: two 64-pixel K-blocks of 8 KiB, 4 x 32 B each
                    const uint64_t ko = (uint64_t)((k >> 2) * (8192 >> 4) + (k & 3) * 2);
                    mma_bf16(dcol, c_hi + ko, gyt_hi + ko, idesc_w, (tile_i > 0 || k > 0) ? 1u : 0u);
                    mma_bf16(dcol, c_lo + ko, gyt_hi + ko, idesc_w, 1u);
                    mma_bf16(dcol, c_hi + ko, gyt_lo + ko, idesc_w, 1u);
                }
                mma_commit(&col_free[b]);
            }
            __syncwarp();
        };
        int it = 0, tile_i = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++tile_i) {
            mbar_wait(gy_ready, tile_i & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int tl = 0; tl < ntaps; ++tl, ++it) {
                const int b = it & 1, ph = (it >> 1) & 1;
                mbar_wait(&w_full[b], ph);
                mbar_wait(&acc_empty[b], ph ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t w_hi = dc + (uint64_t)(smem_u32(s_w + b * MB_WBYTES) >> 4), w_lo = w_hi + (8192 >> 4);
                    const uint32_t dcol = tmem_base + (uint32_t)b * 64u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                // K = 64 co: 4 x 16
                        mma_bf16(dcol, gy_hi + 2 * k, w_hi + 2 * k, idesc_d, k > 0 ? 1u : 0u);
                        mma_bf16(dcol, gy_lo + 2 * k, w_hi + 2 * k, idesc_d, 1u);
                        mma_bf16(dcol, gy_hi + 2 * k, w_lo + 2 * k, idesc_d, 1u);
                    }
                    mma_commit(&w_empty[b]);
                    mma_commit(&acc_full[b]);
                }
                __syncwarp();
                if (tl > 0) weight_gemm(it - 1, tl - 1, tile_i);     // one stage behind: the workers need grad_col first
            }
            weight_gemm(it - 1, ntaps - 1, tile_i);
            if (elect_one()) mma_commit(gy_free);                    // every MMA that reads this tile's gy / gy^T is issued
            __syncwarp();
        }
        if (elect_one()) mma_commit(wg_done);
        __syncwarp();
    } else {
        // ===================== workers: one thread per (pixel, deformable-group pair) =====================
        const int q = warp & 3;                         // TMEM lane quarter this warp may read (= warp id % 4)
        const int gl = (warp - 2) >> 2;                 // 0..3: handles groups gl and gl + 4
        const int prow = q * 32 + lane;                 // pixel row of the tile = TMEM lane of the data-GEMM accumulator
        const uint32_t sgy = smem_u32(s_gy), sgyt = smem_u32(s_gyt), scol = smem_u32(s_col);
        int it = 0, tile_i = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++tile_i) {
            const long long m = (long long)tile * 128 + prow;
            const bool ok = m < M;
            const long long mm = ok ? m : 0;
            const int n = (int)(mm / hw);
            const int r = (int)(mm - (long long)n * hw);
            const int oy = r / p.Wo, ox = r - oy * p.Wo;
            // ---- gy tile: 16 output channels of this thread's pixel -> K-major rows (data GEMM) and transposed (weight GEMM)
            float g16[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 t = ok ? ldg4(p.gy + mm * p.gy_pix_stride + gl * 16 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                g16[j] = t.x; g16[j + 1] = t.y; g16[j + 2] = t.z; g16[j + 3] = t.w;
            }
            if (tile_i > 0) mbar_wait(gy_free, (tile_i - 1) & 1);
            {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) split_bf16x2(g16[2 * j], g16[2 * j + 1], hi[j], lo[j]);
                const uint32_t rowb = (uint32_t)prow * 128u, sw = (uint32_t)(prow & 7);
                const uint32_t c0 = ((uint32_t)(2 * gl) ^ sw) << 4, c1 = ((uint32_t)(2 * gl + 1) ^ sw) << 4;
                sts_v4(sgy + rowb + c0, hi[0], hi[1], hi[2], hi[3]);
                sts_v4(sgy + rowb + c1, hi[4], hi[5], hi[6], hi[7]);
                sts_v4(sgy + MB_TILE + rowb + c0, lo[0], lo[1], lo[2], lo[3]);
                sts_v4(sgy + MB_TILE + rowb + c1, lo[4], lo[5], lo[6], lo[7]);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t o = kmajor_off(gl * 16 + j, prow, 64);
                    const uint32_t h = (j & 1) ? (hi[j >> 1] >> 16) : (hi[j >> 1] & 0xffffu);
                    const uint32_t l = (j & 1) ? (lo[j >> 1] >> 16) : (lo[j >> 1] & 0xffffu);
                    sts_u16(sgyt + o, h);
                    sts_u16(sgyt + MB_TILE + o, l);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(gy_ready);

            const float* img = p.x + (long long)n * p.img_stride;
            float* gimg = p.gx ? p.gx + (long long)n * ((long long)p.H * p.W * p.gx_pix_stride) : nullptr;
            for (int tl = 0; tl < ntaps; ++tl, ++it) {
                const int b = it & 1, ph = (it >> 1) & 1;
                const int tap = tap0 + tl;
                const int kh = tap / p.KW, kw = tap - kh * p.KW;
                // offsets / mask of both items first: their latency hides behind the accumulator wait
                float dy[2], dx[2], mk[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int g = gl + 4 * i;
                    dy[i] = dx[i] = mk[i] = 0.f;
                    if (ok) {
                        const float2 o2 = __ldg(reinterpret_cast<const float2*>(p.offset + mm * p.off_pix_stride + (g * KK + tap) * 2));
                        dy[i] = o2.x; dx[i] = o2.y;
                        mk[i] = __ldg(p.mask + mm * p.mask_pix_stride + g * KK + tap);
                    }
                }
                mbar_wait(&acc_full[b], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float gc[2][8];
                const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 64);
                tmem_ld8(tacc + gl * 8, gc[0]);
                tmem_ld8(tacc + (gl + 4) * 8, gc[1]);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&acc_empty[b]);
                mbar_wait(&col_free[b], ph ^ 1);
                const uint32_t colb = scol + (uint32_t)b * 2u * MB_TILE;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int g = gl + 4 * i;
                    float col[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    float d_m = 0.f, d_h = 0.f, d_w = 0.f;
                    const float h = (float)(oy * p.stride - p.pad + kh * p.dil) + dy[i];
                    const float w = (float)(ox * p.stride - p.pad + kw * p.dil) + dx[i];
                    if (ok && h > -1.f && w > -1.f && h < (float)p.H && w < (float)p.W) {
                        const float hf = floorf(h), wf = floorf(w);
                        const int h0 = (int)hf, w0 = (int)wf;
                        const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hwt = 1.f - lw;
                        const bool v_h0 = h0 >= 0, v_h1 = h0 + 1 <= p.H - 1, v_w0 = w0 >= 0, v_w1 = w0 + 1 <= p.W - 1;
                        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 a0 = z, a1 = z, b0 = z, b1 = z, e0 = z, e1 = z, f0 = z, f1 = z;
                        const long long o00 = ((long long)h0 * p.W + w0) * p.pix_stride + g * 8;
                        const long long rs = (long long)p.W * p.pix_stride;
                        if (v_h0 && v_w0) ldg8(img + o00, a0, a1);
                        if (v_h0 && v_w1) ldg8(img + o00 + p.pix_stride, b0, b1);
                        if (v_h1 && v_w0) ldg8(img + o00 + rs, e0, e1);
                        if (v_h1 && v_w1) ldg8(img + o00 + rs + p.pix_stride, f0, f1);
                        const float w00 = hh * hwt, w01 = hh * lw, w10 = lh * hwt, w11 = lh * lw;
                        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                        const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                        float gm[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float val = w00 * av[c] + w01 * bv[c] + w10 * ev[c] + w11 * fv[c];
                            const float chh = -hwt * av[c] - lw * bv[c] + hwt * ev[c] + lw * fv[c];
                            const float cww = -hh * av[c] + hh * bv[c] - lh * ev[c] + lh * fv[c];
                            gm[c] = gc[i][c] * mk[i];
                            d_m = fmaf(gc[i][c], val, d_m);
                            d_h = fmaf(chh, gm[c], d_h);
                            d_w = fmaf(cww, gm[c], d_w);
                            col[c] = val * mk[i];
                        }
                        if (gimg) {
                            float* gp = gimg + ((long long)h0 * p.W + w0) * p.gx_pix_stride + g * 8;
                            const long long grs = (long long)p.W * p.gx_pix_stride;
                            auto scatter = [&](float* dst, float wk) {
                                atomicAdd(reinterpret_cast<float4*>(dst), make_float4(gm[0] * wk, gm[1] * wk, gm[2] * wk, gm[3] * wk));
                                atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(gm[4] * wk, gm[5] * wk, gm[6] * wk, gm[7] * wk));
                            };
                            if (v_h0 && v_w0) scatter(gp, w00);
                            if (v_h0 && v_w1) scatter(gp + p.gx_pix_stride, w01);
                            if (v_h1 && v_w0) scatter(gp + grs, w10);
                            if (v_h1 && v_w1) scatter(gp + grs + p.gx_pix_stride, w11);
                        }
                    }
                    if (ok) {
                        if (p.goff) *reinterpret_cast<float2*>(p.goff + mm * p.goff_pix_stride + (g * KK + tap) * 2) = make_float2(d_h, d_w);
                        if (p.gmask) p.gmask[mm * p.gmask_pix_stride + g * KK + tap] = d_m;
                    }
                    // col^T: row = input channel g*8 + c, column = this pixel; bf16 hi and lo parts
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        uint32_t hi, lo;
                        split_bf16x2(col[c], col[c + 1], hi, lo);
                        const uint32_t o0 = kmajor_off(g * 8 + c, prow, 64), o1 = kmajor_off(g * 8 + c + 1, prow, 64);
                        sts_u16(colb + o0, hi & 0xffffu);
                        sts_u16(colb + o1, hi >> 16);
                        sts_u16(colb + MB_TILE + o0, lo & 0xffffu);
                        sts_u16(colb + MB_TILE + o1, lo >> 16);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&col_ready[b]);
            }
        }
        // ===================== epilogue: weight-gradient accumulators -> PyTorch-layout gradient =====================
        if (p.gw && warp < 6) {                       // warps 2-5 cover the four TMEM lane quarters
            mbar_wait_relaxed(wg_done, 0, 256);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // M = 64 accumulator: row m lives in TMEM lane (m / 16) * 32 + m % 16 (tools/umma_probe.cu)
            const int ci = lane < 16 ? q * 16 + lane : -1;
            for (int tl = 0; tl < ntaps; ++tl) {
                const int tap = tap0 + tl;
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + wg_col0 + (uint32_t)(tl * 64 + c0), v);
                    if (ci < 0) continue;
                    float* dst = p.gw + p.seg_base + (long long)ci * p.ci_stride + tap;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(dst + (long long)(c0 + j) * p.co_stride, v[j]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace dvsr

using namespace dvsr;

// EDVR geometry only (every DCN of EDVR-M): 64 -> 64 channels, 8 deformable groups of 8 channels, at most 9 taps.
extern "C" int dvsr_mdcn_bwd_tc_supported(const dvsr_conv_desc* d) {
    if (!d || !d->deform || d->nseg != 1 || d->transposed) return 0;
    const dvsr_conv_seg& g = d->seg[0];
    if (g.C != 64 || d->dg != 8 || d->Co != 64) return 0;
    if (d->KH * d->KW > 9 || d->KH * d->KW < 1) return 0;
    if ((g.pix_stride & 7) || (g.img_stride & 7) || ((uintptr_t)g.ptr & 31)) return 0;      // 256-bit corner loads
    if (g.T > 1 || g.t_fixed >= 0) return 0;
    if (((uintptr_t)d->offset & 7) || (d->off_pix_stride & 1)) return 0;                    // float2 offset loads
    return 1;
}

// Gradients w.r.t. input (gx: zero-filled or holding a gradient to accumulate into; may be NULL), offsets, mask (plain stores; may be
// NULL) and -- when gw != NULL -- the weight (ADDED to gw, PyTorch layout through wl).  wp10: dvsr_pack_weights_tc2 mode 10.
extern "C" int dvsr_mdcn_bwd_tc(const dvsr_conv_desc* d, const float* gy, int gy_pix_stride, const float* wp10, float* gx,
                                int gx_pix_stride, float* goff, int goff_pix_stride, float* gmask, int gmask_pix_stride, float* gw,
                                const dvsr_wlayout* wl, void* stream) {
    DVSR_REQUIRE(d && gy && wp10, "mdcn_bwd_tc: null pointer");
    DVSR_REQUIRE(dvsr_mdcn_bwd_tc_supported(d), "mdcn_bwd_tc: unsupported shape (use dvsr_mdcn_bwd_data + dvsr_conv_wgrad)");
    DVSR_REQUIRE((gy_pix_stride & 3) == 0 && (((uintptr_t)gy) & 15) == 0, "mdcn_bwd_tc: gy must be 16-byte aligned");
    DVSR_REQUIRE(!gx || ((((uintptr_t)gx) & 15) == 0 && (gx_pix_stride & 3) == 0), "mdcn_bwd_tc: gx must be 16-byte aligned");
    DVSR_REQUIRE(!goff || ((((uintptr_t)goff) & 7) == 0 && (goff_pix_stride & 1) == 0), "mdcn_bwd_tc: goff must be 8-byte aligned");
    DVSR_REQUIRE(!gw || wl, "mdcn_bwd_tc: weight gradient requested without a weight layout");
    EncodeTiledFn encode = get_encode_tiled();
    DVSR_REQUIRE(encode != nullptr, "mdcn_bwd_tc: cuTensorMapEncodeTiled is unavailable");
    const dvsr_conv_seg& g = d->seg[0];
    MbParams p;
    memset(&p, 0, sizeof(p));
    p.x = g.ptr; p.pix_stride = g.pix_stride;
    p.img_stride = g.img_stride > 0 ? g.img_stride : (long long)d->H * d->W * g.pix_stride;
    p.N = d->N; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo;
    p.KH = d->KH; p.KW = d->KW; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
    p.offset = d->offset; p.off_pix_stride = d->off_pix_stride; p.mask = d->mask; p.mask_pix_stride = d->mask_pix_stride;
    p.gy = gy; p.gy_pix_stride = gy_pix_stride;
    p.gx = gx; p.gx_pix_stride = gx_pix_stride;
    p.goff = goff; p.goff_pix_stride = goff_pix_stride; p.gmask = gmask; p.gmask_pix_stride = gmask_pix_stride;
    p.gw = gw;
    if (wl) { p.co_stride = wl->co_stride; p.ci_stride = wl->ci_stride; p.seg_base = wl->seg_base[0]; }
    const long long M = (long long)d->N * d->Ho * d->Wo;
    p.tiles_total = (int)((M + 127) / 128);
    const int KK = d->KH * d->KW;
    const int ygroups = KK > 1 ? 2 : 1;
    p.taps_a = (KK + ygroups - 1) / ygroups;
    CUtensorMap wmap;
    {
        cuuint64_t dims[2] = {32, (cuuint64_t)KK * 128};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp10, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "mdcn_bwd_tc: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    }
    const size_t smem = 1024 + 2 * MB_WBYTES + 8 * (size_t)MB_TILE + 256;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(mdcn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("mdcn_bwd_tc: cudaFuncSetAttribute");
        smem_set = smem;
    }
    int ctas = cta_budget(d->policy) / ygroups;
    if (ctas < 1) ctas = 1;
    if (ctas > p.tiles_total) ctas = p.tiles_total;
    dim3 grid(ctas, ygroups);
    mdcn_bwd_tc_kernel<<<grid, MB_THREADS, smem, (cudaStream_t)stream>>>(wmap, p);
    return check_launch("mdcn_bwd_tc");
}
