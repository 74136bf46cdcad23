// dynavsr_b200/csrc/conv_tc.cu
//
// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
//
//   y[pix][co] = epilogue( sum_{seg, tap, ci} x_seg[pix + tap][ci] * W[(seg, tap, ci)][co] )
//
// * NHWC fp32 activations are read as TF32 (kind::tf32, fp32 accumulation in TMEM) -- no conversion pass,
//   no im2col: for every (segment, tap, 32-channel chunk) the TMA engine drops the tap-shifted
//   8 x 16-pixel x 32-channel box straight into shared memory in the 128-byte-swizzled K-major layout the
//   MMA descriptor expects; out-of-image pixels (the conv padding and tile overhang) are zero-filled by
//   the TMA unit itself.
// * CTA tile: M = 128 output pixels (8 rows x 16 columns of one image) x N = all output channels
//   (<= 256, padded to a multiple of 16), so activations are fetched once per tile regardless of Co.
// * Warp-specialised: warp 0 = TMA producer (one elected lane), warp 1 = TMEM owner + MMA issuer (one
//   elected lane issues tcgen05.mma, tcgen05.commit releases smem stages / publishes the accumulator),
//   warps 2-5 = operand rounding during the main loop, then the epilogue (tcgen05.ld -> bias / ReLU /
//   LeakyReLU / sigmoid-split / residual / PixelShuffle(2) store).  4-stage mbarrier ring.
// * Numerics: the tensor core TRUNCATES fp32 operands to TF32 (a systematic shrink of ~2^-11 per operand
//   that compounds layer after layer).  Weights are therefore rounded to nearest TF32 when packed and the
//   activation tile is rounded to nearest (cvt.rna.tf32.f32) in shared memory by warps 2-5 between the TMA
//   arrival and the MMA, which makes the per-layer error unbiased (~2e-4 relative).
// * The same kernel computes data gradients of stride-1 convs (taps mirrored: src = o + pad - k) and
//   sums over broadcast segments (`wshare`).
//
// Replaces cuDNN for the heavy layers of EDVR_arch.py:68-90,141-159,224-249 / arch_util.py:42-43.
#include "tc_common.cuh"
#include "pack_device.cuh"

namespace dvsr {

constexpr int TC_TH = 8, TC_TW = 16;          // pixel tile
constexpr int TC_KC = 32;                     // channels per K chunk (128 bytes of fp32)
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_TH * TC_TW * TC_KC * 4;   // 16 KiB
constexpr int TC_THREADS = 192;

struct TcSeg { int C, T, Tsrc, dt, t_fixed; };
struct TcParams {
    int N, Ho, Wo;                 // output images / size
    int KH, KW, tap_sign, tap_base;  // src = o * stride + tap_base + tap_sign * k
    int stride;
    int out_step, out_off_y, out_off_x, out_H, out_W;   // output placement (dense when out_step == 1, offsets 0)
    int nseg, wshare;
    TcSeg seg[DVSR_MAX_SEG];
    int Co, Co_pad;
    const float* bias;
    int act; float slope; int sig_split;
    const float* res; int res_pix_stride;
    int shuffle;
    float* y; int y_pix_stride;
};
struct __align__(64) TcMaps { CUtensorMap x[DVSR_MAX_SEG]; CUtensorMap w; };

__device__ __forceinline__ int tc_seg_image(const TcSeg& sg, int n) {
    const int T = sg.T > 0 ? sg.T : 1;
    const int q = n / T, r = n - q * T;
    const int t = sg.t_fixed >= 0 ? sg.t_fixed : r + sg.dt;
    if (t < 0 || t >= sg.Tsrc) return -1;
    return q * sg.Tsrc + t;
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.Co_pad * TC_KC * 4;
    uint8_t* smem_a = smem;                               // [stages][16 KiB]
    uint8_t* smem_b = smem + TC_STAGES * TC_A_BYTES;      // [stages][Co_pad * 128 B]
    uint64_t* bars = (uint64_t*)(smem_b + TC_STAGES * b_bytes);
    uint64_t* full_bar = bars;                            // [stages]
    uint64_t* empty_bar = bars + TC_STAGES;               // [stages]
    uint64_t* ready_bar = bars + 2 * TC_STAGES;           // [stages] activation tile rounded to TF32
    uint64_t* accum_bar = bars + 3 * TC_STAGES;           // [1]
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * TC_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = (p.Wo + TC_TW - 1) / TC_TW, tiles_h = (p.Ho + TC_TH - 1) / TC_TH;
    int tile = blockIdx.x;
    const int n_img = tile / (tiles_w * tiles_h);
    tile -= n_img * tiles_w * tiles_h;
    const int oy0 = (tile / tiles_w) * TC_TH, ox0 = (tile % tiles_w) * TC_TW;

    // TMEM columns: power of two >= max(32, Co_pad)
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.Co_pad) tmem_cols <<= 1;

    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.nseg; ++s) prefetch_tmap(&maps.x[s]);
        prefetch_tmap(&maps.w);
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&ready_bar[i], 128); }
            mbar_init(accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int KK = p.KH * p.KW;
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0, phase = 0, wrow = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const int img = tc_seg_image(p.seg[s], n_img);
                if (p.wshare) wrow = 0;
                const int chunks = (p.seg[s].C + TC_KC - 1) / TC_KC;
                for (int tap = 0; tap < KK; ++tap) {
                    const int kh = tap / p.KW, kw = tap - kh * p.KW;
                    const int dy = p.tap_base + p.tap_sign * kh, dx = p.tap_base + p.tap_sign * kw;
                    for (int c = 0; c < chunks; ++c, ++wrow) {
                        mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 64);
                        mbar_expect_tx(&full_bar[stage], TC_A_BYTES + b_bytes);
                        // a temporal tap outside the clip reads image index N (fully out of bounds -> zeros)
                        tma_load_4d(&maps.x[s], &full_bar[stage], smem_a + stage * TC_A_BYTES, c * TC_KC, ox0 * p.stride + dx, oy0 * p.stride + dy,
                                    img < 0 ? 0x3fffffff : img);
                        tma_load_2d(&maps.w, &full_bar[stage], smem_b + stage * b_bytes, 0, wrow * p.Co_pad);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_tf32(128, p.Co_pad);
        int stage = 0, phase = 0;
        int total = 0;
        for (int s = 0; s < p.nseg; ++s) total += KK * ((p.seg[s].C + TC_KC - 1) / TC_KC);
        for (int it = 0; it < total; ++it) {
            mbar_wait(&ready_bar[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t adesc = make_desc_k128(smem_u32(smem_a + stage * TC_A_BYTES));
                const uint64_t bdesc = make_desc_k128(smem_u32(smem_b + stage * b_bytes));
#pragma unroll
                for (int k = 0; k < TC_KC / 8; ++k) {
                    // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle span: +2 in the (addr >> 4) field
                    mma_tf32(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                mma_commit(&empty_bar[stage]);                 // frees this smem stage when the MMAs retire
                if (it == total - 1) mma_commit(accum_bar);    // accumulator complete
            }
            __syncwarp();
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===================== main loop: round the activation tile to nearest TF32 in place =====================
        {
            const int t = threadIdx.x - 64;      // 0..127
            int stage = 0, phase = 0, total = 0;
            for (int s = 0; s < p.nseg; ++s) total += KK * ((p.seg[s].C + TC_KC - 1) / TC_KC);
            for (int it = 0; it < total; ++it) {
                mbar_wait_relaxed(&full_bar[stage], phase, 32);
                float4* a4 = reinterpret_cast<float4*>(smem_a + stage * TC_A_BYTES);
#pragma unroll
                for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
                    float4 v = a4[t + i * 128];
                    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
                    a4[t + i * 128] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (MMA)
                mbar_arrive(&ready_bar[stage]);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        // ===================== epilogue: 4 warps, one TMEM lane (= output pixel) per thread =====================
        mbar_wait_relaxed(accum_bar, 0, 256);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;             // pixel index inside the tile: row = ty * 16 + tx
        const int oy = oy0 + row / TC_TW, ox = ox0 + row % TC_TW;
        const int py = oy * p.out_step + p.out_off_y, px = ox * p.out_step + p.out_off_x;   // placement in the output image
        const bool valid = (oy < p.Ho) && (ox < p.Wo) && (py >= 0) && (py < p.out_H) && (px >= 0) && (px < p.out_W);
        const long long pix = ((long long)n_img * p.out_H + py) * p.out_W + px;
        for (int c0 = 0; c0 < p.Co_pad; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);   // warp-collective: no divergence before it
            if (c0 >= p.Co) continue;                  // (warp-uniform)
            if (p.shuffle != 2 && p.out_step == 1 && p.out_off_y == 0 && p.out_off_x == 0) {
                // dense output: quad transpose, then every quad writes full 128-byte lines (tc_common.cuh)
                quad_transpose32(v, lane);
                const int rb = row & ~3, i4 = lane & 3;
                const int oyb = oy0 + rb / TC_TW, oxb = ox0 + rb % TC_TW;
                const int nok = (oyb < p.Ho && oyb < p.out_H) ? max(0, min(4, min(p.Wo, p.out_W) - oxb)) : 0;
                const long long pixb = ((long long)n_img * p.out_H + oyb) * p.out_W + oxb;
                epilogue_store_t(v, p.bias ? p.bias + c0 + 8 * i4 : nullptr, c0 + 8 * i4, p.Co, pixb, nok, nullptr, 0, p.res, p.res_pix_stride,
                                 p.y, p.y_pix_stride, false, p.act, p.slope, p.sig_split);
                continue;
            }
            if (!valid) continue;
            epilogue_chunk(v, c0, p.Co, p.bias ? p.bias + c0 : nullptr, nullptr,
                           p.res ? p.res + pix * p.res_pix_stride + c0 : nullptr, p.act, p.slope, p.sig_split);
            if (p.shuffle == 2) {
                // out[n][2*oy + i][2*ox + jj][c] = v[4c + 2i + jj]: 8 output channels per 32-column chunk
                const int cq = c0 >> 2;
#pragma unroll
                for (int sub = 0; sub < 4; ++sub) {
                    const long long op = ((long long)n_img * (2 * p.Ho) + 2 * oy + (sub >> 1)) * (2 * p.Wo) + 2 * ox + (sub & 1);
                    float* yo = p.y + op * p.y_pix_stride + cq;
                    *reinterpret_cast<float4*>(yo) = make_float4(v[sub], v[4 + sub], v[8 + sub], v[12 + sub]);
                    *reinterpret_cast<float4*>(yo + 4) = make_float4(v[16 + sub], v[20 + sub], v[24 + sub], v[28 + sub]);
                }
            } else {
                float* yo = p.y + pix * p.y_pix_stride + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (c0 + j < p.Co) *reinterpret_cast<float4*>(yo + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ weights
// wp rows of 32 floats: row = ((s, tap, chunk), co_pad index); mode 2: K = input channels of segment s;
// mode 3 (data gradient of segment `seg`): K = forward output channels, rows = input channels of `seg`.
// ------------------------------------------------------------------------------------------------ host
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

static int round16(int v) { return (v + 15) / 16 * 16; }

}  // namespace dvsr

using namespace dvsr;

extern "C" int dvsr_conv_tc_supported(const dvsr_conv_desc* d) {
    if (!d || d->deform || d->dil != 1 || d->accumulate) return 0;
    if (d->stride != 1 && (d->stride != 2 || d->transposed)) return 0;   // strided forward convs via TMA element strides
    if (d->out_step && (d->shuffle || d->res)) return 0;
    if (d->Co > 256 || d->Co < 16 || (d->Co & 3)) return 0;
    if (d->shuffle && (d->Co % 32)) return 0;
    for (int s = 0; s < d->nseg; ++s) {
        const dvsr_conv_seg& g = d->seg[s];
        // a partial last K chunk is zero-filled by TMA (channel coordinate out of bounds) and by the weight packer
        if ((g.C & 3) || g.C < 16 || (g.pix_stride & 3) || (g.img_stride & 3) || ((uintptr_t)g.ptr & 15)) return 0;
    }
    if (((uintptr_t)d->y & 15) || (d->y_pix_stride & 3)) return 0;
    if (d->res && ((((uintptr_t)d->res) & 15) || (d->res_pix_stride & 3))) return 0;
    if (d->bias && (((uintptr_t)d->bias) & 15)) return 0;
    return 1;
}

extern "C" long long dvsr_conv_tc_packed_floats(const dvsr_wlayout* wl, int mode, int seg) {
    if (!wl) return 0;
    if (mode == 2) {
        long long blocks = 0;
        for (int s = 0; s < wl->nseg; ++s) blocks += (long long)wl->taps * ((wl->seg_C[s] + 31) / 32);
        return blocks * round16(wl->Co) * 32;
    }
    return (long long)wl->taps * ((wl->Co + 31) / 32) * round16(wl->seg_C[seg]) * 32;
}

extern "C" int dvsr_pack_weights_tc(const float* w, float* wp, const dvsr_wlayout* wl, int mode, int seg, void* stream) {
    DVSR_REQUIRE(w && wp && wl && (mode == 2 || mode == 3), "pack_weights_tc: bad arguments");
    if (mode == 3) DVSR_REQUIRE(seg >= 0 && seg < wl->nseg, "pack_weights_tc: bad segment");
    const long long total = dvsr_conv_tc_packed_floats(wl, mode, seg);
    const int rows_pad = mode == 2 ? round16(wl->Co) : round16(wl->seg_C[seg]);
    dvsr_pack_job j;
    memset(&j, 0, sizeof(j));
    j.w = w; j.wp = wp; j.wl = *wl; j.mode = mode; j.seg = seg; j.a0 = rows_pad; j.total = total;
    return dvsr_pack_job_run(&j, stream);      // the kernel lives in pack_table.cu (no cross-TU device linking)
}

extern "C" int dvsr_pack_weights_tc_parity(const float* w, float* wp, const dvsr_wlayout* wl, int seg, int KHf, int KWf,
                                           int a, int b, void* stream) {
    DVSR_REQUIRE(w && wp && wl && seg >= 0 && seg < wl->nseg, "pack_weights_tc_parity: bad arguments");
    DVSR_REQUIRE(KHf * KWf == wl->taps && a >= 0 && a < 2 && b >= 0 && b < 2, "pack_weights_tc_parity: bad kernel geometry");
    const int KHs = (KHf - a + 1) / 2, KWs = (KWf - b + 1) / 2;
    DVSR_REQUIRE(KHs > 0 && KWs > 0, "pack_weights_tc_parity: empty parity class");
    const int rows_pad = round16(wl->seg_C[seg]);
    const long long total = (long long)KHs * KWs * ((wl->Co + 31) / 32) * rows_pad * 32;
    dvsr_pack_job j;
    memset(&j, 0, sizeof(j));
    j.w = w; j.wp = wp; j.wl = *wl; j.mode = 4; j.seg = seg; j.a0 = rows_pad; j.a1 = KWf; j.a2 = KWs; j.a3 = 2 * a + b; j.total = total;
    return dvsr_pack_job_run(&j, stream);      // the kernel lives in pack_table.cu (no cross-TU device linking)
}

extern "C" int dvsr_conv_tc_fprop(const dvsr_conv_desc* d, const float* wp, void* stream) {
    DVSR_REQUIRE(d && wp && d->y, "conv_tc_fprop: null pointer");
    DVSR_REQUIRE(dvsr_conv_tc_supported(d), "conv_tc_fprop: unsupported shape (use dvsr_conv_fprop)");
    EncodeTiledFn encode = get_encode_tiled();
    DVSR_REQUIRE(encode != nullptr, "conv_tc_fprop: cuTensorMapEncodeTiled is unavailable");
    TcMaps maps;
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.KH = d->KH; p.KW = d->KW;
    p.stride = d->stride;
    p.out_step = d->out_step ? d->out_step : 1;
    p.out_off_y = d->out_step ? d->out_off_y : 0; p.out_off_x = d->out_step ? d->out_off_x : 0;
    p.out_H = d->out_step ? d->out_H : d->Ho; p.out_W = d->out_step ? d->out_W : d->Wo;
    p.tap_sign = d->transposed ? -1 : 1;
    p.tap_base = d->transposed ? d->pad : -d->pad;
    p.nseg = d->nseg; p.wshare = d->wshare;
    p.Co = d->Co; p.Co_pad = round16(d->Co);
    p.bias = d->bias; p.act = d->act; p.slope = d->slope; p.sig_split = d->sig_split;
    p.res = d->res; p.res_pix_stride = d->res_pix_stride; p.shuffle = d->shuffle;
    p.y = d->y; p.y_pix_stride = d->y_pix_stride;
    long long wrows = 0;
    for (int s = 0; s < d->nseg; ++s) {
        const dvsr_conv_seg& g = d->seg[s];
        p.seg[s].C = g.C; p.seg[s].T = g.T; p.seg[s].Tsrc = g.Tsrc; p.seg[s].dt = g.dt; p.seg[s].t_fixed = g.t_fixed;
        // number of source images addressable through this segment
        const int T = g.T > 0 ? g.T : 1;
        long long nsrc = ((long long)(d->N + T - 1) / T) * g.Tsrc;
        if (nsrc < 1) nsrc = 1;
        cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)nsrc};
        long long img_stride = g.img_stride > 0 ? g.img_stride : (long long)d->H * d->W * g.pix_stride;
        cuuint64_t strides[3] = {(cuuint64_t)g.pix_stride * 4, (cuuint64_t)d->W * g.pix_stride * 4, (cuuint64_t)img_stride * 4};
        cuuint32_t box[4] = {TC_KC, (cuuint32_t)(TC_TW * d->stride), (cuuint32_t)(TC_TH * d->stride), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        CUresult r = encode(&maps.x[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)g.ptr, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "conv_tc_fprop: cuTensorMapEncodeTiled(activation seg %d) failed with %d", s, (int)r);
        if (!d->wshare || s == 0) wrows += (long long)d->KH * d->KW * ((g.C + TC_KC - 1) / TC_KC) * p.Co_pad;
    }
    {
        cuuint64_t dims[2] = {TC_KC, (cuuint64_t)wrows};
        cuuint64_t strides[1] = {TC_KC * 4};
        cuuint32_t box[2] = {TC_KC, (cuuint32_t)p.Co_pad};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "conv_tc_fprop: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    }
    const int tiles = d->N * ((d->Ho + TC_TH - 1) / TC_TH) * ((d->Wo + TC_TW - 1) / TC_TW);
    const size_t smem = 1024 + (size_t)TC_STAGES * (TC_A_BYTES + p.Co_pad * TC_KC * 4) + 256;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("conv_tc_fprop: cudaFuncSetAttribute");
        smem_set = smem;
    }
    conv_tc_kernel<<<tiles, TC_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    return check_launch("conv_tc_fprop");
}
