// dynavsr_b200/csrc/mdcn_tc.cu
//
// Modulated deformable convolution forward (DCNv2, EDVR's PCD alignment) with the GEMM on the tcgen05 tensor
// cores.  Replaces `modulated_deformable_im2col_gpu_kernel` + `addmm_` + bias (+ LeakyReLU)
// (deform_conv_cuda_kernel.cu:569-632, deform_conv_cuda.cpp:534-563) without any `columns` buffer in HBM:
//
//   * 8 gather warps build, per 128-pixel tile, tap and 32-channel chunk, the modulated bilinear samples
//     directly in shared memory in the K-major 128-byte-swizzled operand layout -- one thread per
//     (pixel, deformable group): 2 offsets + 1 mask (coalesced from the fused [N,H,W,216] offset/mask tensor),
//     4 corners x 32 contiguous bytes (8 channels of the group, NHWC), blend, split into bf16 hi + lo;
//   * one warp issues the BF16x3 MMAs (x_hi.w_hi + x_lo.w_hi + x_hi.w_lo, fp32 accumulation in TMEM) against
//     the weights that stay resident in shared memory for the whole kernel (persistent CTAs, one per SM);
//   * 4 epilogue warps drain the double-buffered TMEM accumulator (bias, LeakyReLU, 256-bit stores).
// Exact reference sampling semantics: zero outside (-1,H)x(-1,W), per-corner bounds (kernel.cu:466-496,617).
// Two gather variants: mdcn_tc_kernel reads the corners straight from global memory (256-bit loads, next stage's offsets
// prefetched; any geometry, small launches); mdcn_tcs_kernel (below) stages an input window per tile in shared memory with
// TMA and is used for the 3x3 / stride-1 launches of EDVR with at least two tiles per SM.
#include "tc_common.cuh"

namespace dvsr {

constexpr int MD_ASTAGES = 4;
constexpr int MD_A_BYTES = 128 * 128;     // 128 pixel rows x 128 B
constexpr int MD_GATHER = 512;            // gather threads: one (pixel, deformable group of the chunk) each -- 16 warps, because the
                                          // gather is a long dependent instruction stream per thread and needs warps, not ILP, to hide it
constexpr int MD_PPT = 512 / MD_GATHER;   // pixels per gather thread (128 pixels x 4 groups per stage)
constexpr int MD_THREADS = 64 + MD_GATHER + 128;   // 1 producer + 1 MMA + 16 gather + 4 epilogue warps

struct MdParams {
    const float* x; int C, pix_stride; long long img_stride;
    int N, H, W, Ho, Wo, KH, KW, stride, pad, dil, dg;
    const float* offset; int off_pix_stride;
    const float* mask; int mask_pix_stride;
    int Co;
    const float* bias; int act; float slope;
    float* y; int y_pix_stride; int y_vec8;
    int nblocks, tiles_total;
    // staged variant (mdcn_tcs_kernel): 2-D output tiles of 16 x 8 pixels, input window of win_h x win_w pixels per chunk
    int tiles_w, tiles_h, margin, win_h, win_w, win_bytes;
};

__global__ void __launch_bounds__(MD_THREADS, 1)
mdcn_tc_kernel(const __grid_constant__ CUtensorMap wmap, const MdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem;                                    // [nblocks][64 x 128 B]
    uint8_t* smem_a = smem + p.nblocks * 8192;                 // [MD_ASTAGES][16 KiB]
    uint64_t* bars = (uint64_t*)(smem_a + MD_ASTAGES * MD_A_BYTES);
    uint64_t* b_full = bars;                       // [1]
    uint64_t* a_ready = bars + 1;                  // [4] 256 gather arrivals
    uint64_t* a_empty = bars + 5;                  // [4] MMA commit
    uint64_t* acc_full = bars + 9;                 // [2]
    uint64_t* acc_empty = bars + 11;               // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 13);
    float* bias_s = (float*)(bars + 16);           // [64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KK = p.KH * p.KW;
    const int chunks = p.C / 32;
    const long long M = (long long)p.N * p.Ho * p.Wo;
    const int hw = p.Ho * p.Wo;

    if (warp == 0 && elect_one()) prefetch_tmap(&wmap);
    if (warp == 1) {
        if (elect_one()) {
            mbar_init(b_full, 1);
            for (int i = 0; i < MD_ASTAGES; ++i) { mbar_init(&a_ready[i], MD_GATHER); mbar_init(&a_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        const int co = threadIdx.x - 64;
        bias_s[co] = (p.bias && co < p.Co) ? __ldg(p.bias + co) : 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(b_full, (uint32_t)p.nblocks * 8192u);
            for (int b = 0; b < p.nblocks; ++b) tma_load_2d(&wmap, b_full, smem_b + b * 8192, 0, b * 64);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (BF16x3) =====================
        const uint32_t idesc = make_idesc_bf16(128, 64);
        const uint64_t d_const = make_desc(0, 16, 1024, 2);
        mbar_wait(b_full, 0);
        int stage = 0, phase = 0, local = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            mbar_wait_relaxed(&acc_empty[acc], ((local >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dcol = tmem_base + acc * 64;
            for (int it = 0; it < chunks * KK; ++it) {        // block order = chunk-major, then tap (pack mode 7)
                mbar_wait_relaxed(&a_ready[stage], phase, 64);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t ad = d_const + (uint64_t)(smem_u32(smem_a + stage * MD_A_BYTES) >> 4);
                    const uint64_t bd = d_const + (uint64_t)(smem_u32(smem_b + it * 8192) >> 4);
                    mma_bf16(dcol, ad, bd, idesc, it > 0 ? 1u : 0u);          // x_hi . w_hi
                    mma_bf16(dcol, ad + 2, bd + 2, idesc, 1u);
                    mma_bf16(dcol, ad + 4, bd, idesc, 1u);                    // x_lo . w_hi
                    mma_bf16(dcol, ad + 6, bd + 2, idesc, 1u);
                    mma_bf16(dcol, ad, bd + 4, idesc, 1u);                    // x_hi . w_lo
                    mma_bf16(dcol, ad + 2, bd + 6, idesc, 1u);
                    mma_commit(&a_empty[stage]);
                    if (it == chunks * KK - 1) mma_commit(&acc_full[acc]);
                }
                __syncwarp();
                if (++stage == MD_ASTAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 2 + MD_GATHER / 32) {
        // ===================== gather: modulated bilinear samples -> swizzled bf16 hi|lo operand rows =====================
        const int t = threadIdx.x - 64;                 // 0..255
        const int gl = t & 3;                           // deformable group inside the 32-channel chunk (8 channels each)
        int stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            // the two pixels this thread serves in every stage of this tile
            long long mlin[MD_PPT]; int n_[MD_PPT], oy[MD_PPT], ox[MD_PPT]; bool ok[MD_PPT];
#pragma unroll
            for (int i = 0; i < MD_PPT; ++i) {
                const int prow = (t >> 2) + (MD_GATHER / 4) * i;
                mlin[i] = (long long)tile * 128 + prow;
                ok[i] = mlin[i] < M;
                const long long mm = ok[i] ? mlin[i] : 0;
                n_[i] = (int)(mm / hw);
                const int r = (int)(mm - (long long)n_[i] * hw);
                oy[i] = r / p.Wo; ox[i] = r - oy[i] * p.Wo;
            }
            // offsets / mask of stage it + 1 are fetched while stage it gathers: one dependent memory round trip per stage
            // (offset -> corner addresses) instead of two
            float ody[MD_PPT], odx[MD_PPT], omk[MD_PPT], ndy[MD_PPT] = {}, ndx[MD_PPT] = {}, nmk[MD_PPT] = {};
            auto load_off = [&](int c, int tap, float (&dy)[MD_PPT], float (&dx)[MD_PPT], float (&mk)[MD_PPT]) {
                const int g = c * 4 + gl;
#pragma unroll
                for (int i = 0; i < MD_PPT; ++i) {
                    dy[i] = dx[i] = mk[i] = 0.f;
                    if (ok[i]) {
                        const float* op = p.offset + mlin[i] * p.off_pix_stride + (g * KK + tap) * 2;
                        dy[i] = __ldg(op); dx[i] = __ldg(op + 1);
                        mk[i] = __ldg(p.mask + mlin[i] * p.mask_pix_stride + g * KK + tap);
                    }
                }
            };
            load_off(0, 0, ody, odx, omk);
            for (int c = 0; c < chunks; ++c) {
                const int g = c * 4 + gl;
                for (int tap = 0; tap < KK; ++tap) {
                    const int kh = tap / p.KW, kw = tap - kh * p.KW;
                    {
                        const int tn = tap + 1 < KK ? tap + 1 : 0, cn = tap + 1 < KK ? c : c + 1;
                        if (cn < chunks) load_off(cn, tn, ndy, ndx, nmk);
                    }
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    uint8_t* tile_a = smem_a + stage * MD_A_BYTES;
#pragma unroll
                    for (int i = 0; i < MD_PPT; ++i) {
                        const int prow = (t >> 2) + (MD_GATHER / 4) * i;
                        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (ok[i]) {
                            const float dy = ody[i], dx = odx[i], mk = omk[i];
                            const float h = (float)(oy[i] * p.stride - p.pad + kh * p.dil) + dy;
                            const float w = (float)(ox[i] * p.stride - p.pad + kw * p.dil) + dx;
                            const BilinTap bt = make_tap(h, w, p.H, p.W);
                            if (bt.inside) {
                                const float* img = p.x + (long long)n_[i] * p.img_stride + g * 8;
                                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                                float4 a0 = z, a1 = z, b0 = z, b1 = z, e0 = z, e1 = z, f0 = z, f1 = z;
                                if (bt.o00 >= 0) ldg8(img + (long long)bt.o00 * p.pix_stride, a0, a1);
                                if (bt.o01 >= 0) ldg8(img + (long long)bt.o01 * p.pix_stride, b0, b1);
                                if (bt.o10 >= 0) ldg8(img + (long long)bt.o10 * p.pix_stride, e0, e1);
                                if (bt.o11 >= 0) ldg8(img + (long long)bt.o11 * p.pix_stride, f0, f1);
                                // same association as the reference: (w1*v1 + w2*v2 + w3*v3 + w4*v4) * mask
                                v[0] = (bt.w00 * a0.x + bt.w01 * b0.x + bt.w10 * e0.x + bt.w11 * f0.x) * mk;
                                v[1] = (bt.w00 * a0.y + bt.w01 * b0.y + bt.w10 * e0.y + bt.w11 * f0.y) * mk;
                                v[2] = (bt.w00 * a0.z + bt.w01 * b0.z + bt.w10 * e0.z + bt.w11 * f0.z) * mk;
                                v[3] = (bt.w00 * a0.w + bt.w01 * b0.w + bt.w10 * e0.w + bt.w11 * f0.w) * mk;
                                v[4] = (bt.w00 * a1.x + bt.w01 * b1.x + bt.w10 * e1.x + bt.w11 * f1.x) * mk;
                                v[5] = (bt.w00 * a1.y + bt.w01 * b1.y + bt.w10 * e1.y + bt.w11 * f1.y) * mk;
                                v[6] = (bt.w00 * a1.z + bt.w01 * b1.z + bt.w10 * e1.z + bt.w11 * f1.z) * mk;
                                v[7] = (bt.w00 * a1.w + bt.w01 * b1.w + bt.w10 * e1.w + bt.w11 * f1.w) * mk;
                            }
                        }
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) split_bf16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
                        // row prow = [hi: 4 chunks of 8 channels | lo: 4 chunks]; 128B swizzle: chunk ^ (row & 7)
                        uint4* row = reinterpret_cast<uint4*>(tile_a + prow * 128);
                        const int ph = prow & 7;
                        row[gl ^ ph] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        row[(4 + gl) ^ ph] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&a_ready[stage]);
                    if (++stage == MD_ASTAGES) { stage = 0; phase ^= 1; }
#pragma unroll
                    for (int i = 0; i < MD_PPT; ++i) { ody[i] = ndy[i]; odx[i] = ndx[i]; omk[i] = nmk[i]; }
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int local = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            const long long pix = (long long)tile * 128 + row;
            const bool valid = pix < M;
            mbar_wait_relaxed(&acc_full[acc], (local >> 1) & 1, 256);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                // coalesced stores: quad transpose, then each quad writes one full 128-byte line per instruction (tc_common.cuh)
                const long long pixb = (long long)tile * 128 + (row & ~3);
                const int nok = (int)min(4LL, max(0LL, M - pixb));
                const int i4 = lane & 3;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    // one 32-column half at a time (register budget: 80 per thread in this 22-warp CTA); the accumulator is
                    // handed back to the MMA warp as soon as its second half is in registers
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64 + h * 32), v);
                    if (h == 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(&acc_empty[acc]);
                    }
                    const int c0 = h * 32;
                    if (c0 >= p.Co) continue;
                    quad_transpose32(v, lane);
                    epilogue_store_t(v, bias_s + c0 + 8 * i4, c0 + 8 * i4, p.Co, pixb, nok, nullptr, 0, nullptr, 0, p.y, p.y_pix_stride,
                                     p.y_vec8 != 0, p.act, p.slope, 0);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// Staged variant (3x3, stride 1, dilation 1 -- every DCN of EDVR): per 16 x 8 output tile and 32-channel chunk ONE TMA box
// brings the input window (tile + 1-pixel tap ring + `margin` pixels for the learned offsets; 26 x 18 pixels x 128 B = 60 KB)
// into shared memory, and the 9 taps x 4 corners x 128 pixels x 4 groups bilinear reads of that chunk are served from there
// (swizzled LDS.128 pairs) instead of L1/L2 -- 48 KB of L2 traffic per chunk-tile instead of ~590 KB.  Corners that a large
// offset pushes outside the window fall back to the global 256-bit load, so the result is exact for any offset.
// Weights are NOT resident here: the gather, not the GEMM, bounds this kernel, and a measurement with the gather switched off
// showed the 2-stage operand ring that resident weights left room for to be latency-bound (MMA completion -> slot-free
// round trip: 267 us of 417 us).  Each stage therefore carries its own 8 KiB weight block, streamed by TMA from L2 (317 MB per
// full-resolution call), which buys a 4-stage ring and a double-buffered input window.
// Offsets / mask of a (tile, chunk) staged in shared memory by TMA (round 2): per pixel the 4 groups x 9 taps x (dy, dx) of a
// 32-channel chunk are 288 contiguous bytes of the fused [N,H,W,216] tensor and the masks 144 -- two boxes (72 x 8 x 16 and
// 36 x 8 x 16 floats, 54 KB) per chunk.  The default path fetches (dy, dx) / mask per thread and tap with 8- and 4-byte global loads:
// 40.5 M L1 sectors per call for 7.6 M of data at a 29 % hit rate (the ~30 KB of L1 beside this kernel's shared memory hold a
// fraction of the 16 warps' working set) = 26 % of the L1 data pipe and the kernel's top stall (long scoreboard).
// Measured on one box (tools/gpu_r2_18.sh, 5x176x320, offsets ~N(0, s^2), s = 1.0 / 1.5 / 3.0): staged offsets with a 2-deep
// operand ring (all that fits beside them) 333 / 339 / 458 us vs 302 / 322 / 550 us for the per-tap global loads with a 3-deep
// ring -- the rolling prefetch queue already hides those loads on the fast path, and the ring stage they cost is worth more;
// only large offsets (the global fallback path then has the L1 to itself) profit.  Compile-time option, OFF by default.
#ifndef DVSR_MDS_OFFSMEM
#define DVSR_MDS_OFFSMEM 0
#endif
#ifndef DVSR_MDS_ASTAGES
#define DVSR_MDS_ASTAGES (DVSR_MDS_OFFSMEM ? 2 : 3)      // with staged offsets: 2 x 60 KB windows + 54 KB + 2 x 24 KB = 222 KB
#endif
#ifndef DVSR_MDS_MARGIN
#define DVSR_MDS_MARGIN 4
#endif
constexpr int MDS_ASTAGES = DVSR_MDS_ASTAGES;             // (A operand tile 16 KiB + streamed weight block 8 KiB) per stage
constexpr int MDS_STAGE_BYTES = MD_A_BYTES + 8192;
constexpr int MDS_OFF_BYTES = 128 * 72 * 4, MDS_MSK_BYTES = 128 * 36 * 4;
constexpr int MDS_MARGIN = DVSR_MDS_MARGIN;              // window margin for the learned offsets, pixels (4: 26 x 18 window = 60 KB x 2 buffers + 4 x 24 KB stages = 216 KB)
constexpr int MDS_WIN_H = 16 + 2 + 2 * MDS_MARGIN, MDS_WIN_W = 8 + 2 + 2 * MDS_MARGIN;     // 26 x 18 pixels
constexpr int MDS_WINBUFS = 2;             // input window double-buffered: the TMA of chunk c+1 flies while chunk c is gathered

__device__ __forceinline__ void lds8(uint32_t addr0, uint32_t addr1, float4& a, float4& b) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(addr0));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(addr1));
}

__global__ void __launch_bounds__(MD_THREADS, 1)
mdcn_tcs_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap omap,
                const __grid_constant__ CUtensorMap mmap, const MdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_w = smem;                                    // [MDS_WINBUFS] input windows [win_h * win_w][128 B], NOT swizzled
    uint8_t* smem_off = smem_w + MDS_WINBUFS * p.win_bytes;    // [128 pixels][72]: (dy, dx) of 4 groups x 9 taps   (DVSR_MDS_OFFSMEM)
    uint8_t* smem_msk = smem_off + (DVSR_MDS_OFFSMEM ? MDS_OFF_BYTES : 0);      // [128 pixels][36]: masks of 4 groups x 9 taps
    uint8_t* smem_a = smem_msk + (DVSR_MDS_OFFSMEM ? MDS_MSK_BYTES : 0);        // [MDS_ASTAGES][A tile 16 KiB | weight block 8 KiB]
    uint64_t* bars = (uint64_t*)(smem_a + MDS_ASTAGES * MDS_STAGE_BYTES);
    uint64_t* w_full = bars;                       // [4] TMA weight block landed
    uint64_t* a_ready = bars + 4;                  // [4] 16 arrivals (one per gather warp)
    uint64_t* a_empty = bars + 8;                  // [4] MMA commit
    uint64_t* acc_full = bars + 12;                // [2]
    uint64_t* acc_empty = bars + 14;               // [2]
    uint64_t* win_full = bars + 16;                // [2] TMA
    uint64_t* win_empty = bars + 18;               // [2] 512 gather arrivals
    uint64_t* off_full = bars + 20;                // [1] TMA: offsets + masks of the chunk landed
    uint64_t* off_empty = bars + 21;               // [1] 16 gather-warp arrivals: the chunk's last tap has read its offsets
    uint32_t* tmem_slot = (uint32_t*)(bars + 22);
    float* bias_s = (float*)(bars + 24);           // [64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KK = 9;
    const int chunks = p.C / 32;
    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (warp == 0 && elect_one()) { prefetch_tmap(&wmap); prefetch_tmap(&xmap); if (DVSR_MDS_OFFSMEM) { prefetch_tmap(&omap); prefetch_tmap(&mmap); } }
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < MDS_ASTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&a_ready[i], MD_GATHER / 32); mbar_init(&a_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
            for (int i = 0; i < MDS_WINBUFS; ++i) { mbar_init(&win_full[i], 1); mbar_init(&win_empty[i], MD_GATHER / 32); }
            mbar_init(off_full, 1); mbar_init(off_empty, MD_GATHER / 32);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        const int co = threadIdx.x - 64;
        bias_s[co] = (p.bias && co < p.Co) ? __ldg(p.bias + co) : 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: input window per (tile, chunk), one weight block per stage =====================
        if (elect_one()) {
            auto load_window = [&](int tile, int c, int cc) {          // cc = running chunk count of this CTA
                const int wb = cc & 1;
                const int tn = tile / tiles_per_img, tr = tile - tn * tiles_per_img;
                const int oy0 = (tr / p.tiles_w) * 16, ox0 = (tr % p.tiles_w) * 8;
                mbar_wait_relaxed(&win_empty[wb], ((cc >> 1) & 1) ^ 1);
                mbar_expect_tx(&win_full[wb], (uint32_t)(MDS_WIN_H * MDS_WIN_W * 128));      // the box, not the padded buffer
                tma_load_4d(&xmap, &win_full[wb], smem_w + wb * p.win_bytes, c * 32, ox0 - 1 - p.margin, oy0 - 1 - p.margin, tn);
            };
            auto load_offsets = [&](int tile, int c, int cc) {         // single buffer: waits until chunk cc - 1 has read its last tap
                if (!DVSR_MDS_OFFSMEM) return;
                const int tn = tile / tiles_per_img, tr = tile - tn * tiles_per_img;
                const int oy0 = (tr / p.tiles_w) * 16, ox0 = (tr % p.tiles_w) * 8;
                mbar_wait_relaxed(off_empty, (cc & 1) ^ 1);
                mbar_expect_tx(off_full, (uint32_t)(MDS_OFF_BYTES + MDS_MSK_BYTES));
                tma_load_4d(&omap, off_full, smem_off, c * 72, ox0, oy0, tn);
                tma_load_4d(&mmap, off_full, smem_msk, c * 36, ox0, oy0, tn);
            };
            int stage = 0, phase = 0, cc = 0;
            if ((int)blockIdx.x < p.tiles_total) { load_window(blockIdx.x, 0, 0); load_offsets(blockIdx.x, 0, 0); }
            for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
                for (int c = 0; c < chunks; ++c, ++cc) {
                    // window of the NEXT chunk (its buffer was freed when chunk cc - 1 finished)
                    if (c + 1 < chunks) load_window(tile, c + 1, cc + 1);
                    else if (tile + (int)gridDim.x < p.tiles_total) load_window(tile + (int)gridDim.x, 0, cc + 1);
                    for (int tap = 0; tap < KK; ++tap) {
                        mbar_wait_relaxed(&a_empty[stage], phase ^ 1, 32);
                        mbar_expect_tx(&w_full[stage], 8192u);
                        tma_load_2d(&wmap, &w_full[stage], smem_a + stage * MDS_STAGE_BYTES + MD_A_BYTES, 0, (c * KK + tap) * 64);
                        if (++stage == MDS_ASTAGES) { stage = 0; phase ^= 1; }
                    }
                    // offsets / masks of the NEXT chunk: the buffer is free as soon as this chunk's last tap has read its offsets
                    if (c + 1 < chunks) load_offsets(tile, c + 1, cc + 1);
                    else if (tile + (int)gridDim.x < p.tiles_total) load_offsets(tile + (int)gridDim.x, 0, cc + 1);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (BF16x3) =====================
        const uint32_t idesc = make_idesc_bf16(128, 64);
        const uint64_t d_const = make_desc(0, 16, 1024, 2);
        int stage = 0, phase = 0, local = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            mbar_wait_relaxed(&acc_empty[acc], ((local >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dcol = tmem_base + acc * 64;
            for (int it = 0; it < chunks * KK; ++it) {        // block order = chunk-major, then tap (pack mode 7)
                mbar_wait(&w_full[stage], phase);
                mbar_wait(&a_ready[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t ad = d_const + (uint64_t)(smem_u32(smem_a + stage * MDS_STAGE_BYTES) >> 4);
                    const uint64_t bd = d_const + (uint64_t)(smem_u32(smem_a + stage * MDS_STAGE_BYTES + MD_A_BYTES) >> 4);
                    mma_bf16(dcol, ad, bd, idesc, it > 0 ? 1u : 0u);          // x_hi . w_hi
                    mma_bf16(dcol, ad + 2, bd + 2, idesc, 1u);
                    mma_bf16(dcol, ad + 4, bd, idesc, 1u);                    // x_lo . w_hi
                    mma_bf16(dcol, ad + 6, bd + 2, idesc, 1u);
                    mma_bf16(dcol, ad, bd + 4, idesc, 1u);                    // x_hi . w_lo
                    mma_bf16(dcol, ad + 2, bd + 6, idesc, 1u);
                    mma_commit(&a_empty[stage]);
                    if (it == chunks * KK - 1) mma_commit(&acc_full[acc]);
                }
                __syncwarp();
                if (++stage == MDS_ASTAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 2 + MD_GATHER / 32) {
        // ===================== gather from the staged window =====================
        // One thread per (pixel, deformable group of the chunk), 9 taps unrolled.  Restructured after the round-1 ncu capture
        // (141 SASS instructions per item, a third of them bounds branches and generic-address math; 18 offset loads stalling
        // every chunk start; 512-thread mbarrier polls):
        //   * the TMA window is zero-filled outside the image, so a sample whose 2 x 2 corner block lies inside the window needs NO
        //     bounds test at all -- out-of-image corners read zeros, which is exactly the reference rule (kernel.cu:466-496); one
        //     window-containment test per item selects this branch-free path (4 swizzled LDS.128 pairs at immediate offsets);
        //     anything else (offsets beyond the 4-pixel margin) takes the exact per-corner path with global 256-bit loads;
        //   * (dy, dx, mask) travel in a 3-deep rolling prefetch queue (9 registers instead of 27, latency hidden behind 3 stages);
        //   * blend and hi|lo split in packed fp32x2 math (FFMA2), mask folded into the four bilinear weights;
        //   * explicit shared-memory stores at per-thread constant offsets; one lane per warp polls / arrives on the mbarriers.
        const int t = threadIdx.x - 64;                 // 0..511
        const int gl = t & 3;                           // deformable group inside the 32-channel chunk
        const int prow = t >> 2;                        // pixel row of the 128-pixel tile
        const uint32_t ph = (uint32_t)(prow & 7);
        // Bank-conflict-free shared-memory traffic (the round-2 ncu capture showed 35.5 M of 65.1 M shared wavefronts of this
        // kernel to be conflicts, L1TEX at 85 %): a 128-bit access is served per quarter-warp = 2 pixels x 4 groups here.
        //   * window reads: the window is NOT swizzled, so the 16-byte chunk a lane reads is (2 * group + half) whatever pixel
        //     its sample falls on; lanes of odd pixels read the upper half of their group's 32 bytes first, lanes of even pixels
        //     the lower half -> the 8 lanes always cover 8 distinct chunks;
        //   * operand-row stores (K-major 128B swizzle, required by the MMA): even pixels store hi then lo, odd pixels lo then hi
        //     -> chunks {0..3}^s and {4..7}^(s^1) are disjoint.
        // The price is 8 register selects per item (the two halves change places in odd lanes).
        const bool par = (prow & 1) != 0;
        const uint32_t st_hi = (uint32_t)prow * 128u + ((((uint32_t)gl) ^ ph) << 4);
        const uint32_t st_lo = (uint32_t)prow * 128u + ((((uint32_t)(4 + gl)) ^ ph) << 4);
        const uint32_t st_first = par ? st_lo : st_hi, st_second = par ? st_hi : st_lo;
        const uint32_t sa0 = smem_u32(smem_a);
        const uint32_t lane_off = (uint32_t)(gl * 32 + (par ? 16 : 0));
        struct Geo { long long mlin; int oy, ox, tn, wy0, wx0; bool ok; };
        auto geo_of = [&](int tile) {
            Geo gq;
            gq.tn = tile / tiles_per_img;
            const int tr = tile - gq.tn * tiles_per_img;
            const int oy0 = (tr / p.tiles_w) * 16, ox0 = (tr % p.tiles_w) * 8;
            gq.wy0 = oy0 - 1 - p.margin; gq.wx0 = ox0 - 1 - p.margin;
            gq.oy = oy0 + (prow >> 3); gq.ox = ox0 + (prow & 7);
            gq.ok = gq.oy < p.Ho && gq.ox < p.Wo;
            gq.mlin = ((long long)gq.tn * p.Ho + (gq.ok ? gq.oy : 0)) * p.Wo + (gq.ok ? gq.ox : 0);
            return gq;
        };
#if DVSR_MDS_OFFSMEM
        // this thread's (dy, dx) pairs / masks of the chunk in the staged boxes: [pixel][group][tap] -- LDS.64 / LDS.32 at immediate
        // offsets; conflict-free (a half-warp = 4 pixels x 4 groups reads words 72 p + 18 g + 2 tap (+1): 16 distinct even words mod 32)
        const uint32_t off_t = smem_u32(smem_off) + (uint32_t)(prow * 288 + gl * 72);
        const uint32_t msk_t = smem_u32(smem_msk) + (uint32_t)(prow * 144 + gl * 36);
#endif
        float qdy[3], qdx[3], qmk[3];
        auto load_q = [&](int slot, const Geo& gq, int c, int tap) {
            if (DVSR_MDS_OFFSMEM) return;
            qdy[slot] = qdx[slot] = qmk[slot] = 0.f;
            if (gq.ok) {
                const int g = c * 4 + gl;
                // (dy, dx) pairs are 8-byte aligned ([dg][k][2] layout, even pixel stride): one 64-bit request per tap
                const float2 o2 = __ldg(reinterpret_cast<const float2*>(p.offset + gq.mlin * p.off_pix_stride + (g * 9 + tap) * 2));
                qdy[slot] = o2.x; qdx[slot] = o2.y;
                qmk[slot] = __ldg(p.mask + gq.mlin * p.mask_pix_stride + g * 9 + tap);
            }
        };
        int stage = 0, phase = 0, cc = 0;
        Geo cur = geo_of(blockIdx.x < (unsigned)p.tiles_total ? (int)blockIdx.x : 0);
        if ((int)blockIdx.x < p.tiles_total) {
#pragma unroll
            for (int i = 0; i < 3; ++i) load_q(i, cur, 0, i);
        }
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            const bool has_next = tile + (int)gridDim.x < p.tiles_total;
            const Geo nxt = geo_of(has_next ? tile + (int)gridDim.x : tile);
            const float* img0 = p.x + (long long)cur.tn * p.img_stride;
            const float fy = (float)(cur.oy - 1), fx = (float)(cur.ox - 1);
            for (int c = 0; c < chunks; ++c) {
                const int g = c * 4 + gl;
                const int wb = cc & 1;
                const uint32_t win_s = smem_u32(smem_w + wb * p.win_bytes);
                if (lane == 0) { mbar_wait(&win_full[wb], (cc >> 1) & 1); if (DVSR_MDS_OFFSMEM) mbar_wait(off_full, cc & 1); }
                __syncwarp();
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int kh = tap / 3, kw = tap - kh * 3;
                    const int slot = tap % 3;
#if DVSR_MDS_OFFSMEM
                    float dy, dx, mk;
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(dy), "=f"(dx) : "r"(off_t + (uint32_t)(tap * 8)));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mk) : "r"(msk_t + (uint32_t)(tap * 4)));
#else
                    const float dy = qdy[slot], dx = qdx[slot], mk = qmk[slot];
#endif
                    // refill the slot with the stage three ahead (same chunk, next chunk, or chunk 0 of this CTA's next tile)
                    if (tap + 3 < 9) load_q(slot, cur, c, tap + 3);
                    else if (c + 1 < chunks) load_q(slot, cur, c + 1, tap + 3 - 9);
                    else if (has_next) load_q(slot, nxt, 0, tap + 3 - 9);
                    const float h = fy + (float)kh + dy, w = fx + (float)kw + dx;
                    const float hf = floorf(h), wf = floorf(w);
                    const int h0 = (int)hf, w0 = (int)wf;
                    const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
                    const int wy = h0 - cur.wy0, wx = w0 - cur.wx0;
                    // (Tried in round 2: L1 prefetches of the NEXT tap's corners for samples about to leave the window -- slower at every
                    // offset spread (std 1.5: 400 -> 583 us): the extra address math sits on every item's path and the prefetched lines do
                    // not survive in the ~30 KB of L1 left beside 198 KB of shared memory.)
                    // vF* = the 4 channels of the half this lane reads first (lower half in even-pixel lanes, upper in odd), vS* = the other
                    float2 vF01, vF23, vS01, vS23;
                    if ((unsigned)wy < (unsigned)(MDS_WIN_H - 1) && (unsigned)wx < (unsigned)(MDS_WIN_W - 1)) {
                        // ---- fast path: the whole 2 x 2 block is inside the (zero-filled) window
                        const uint32_t b0 = win_s + (uint32_t)(wy * MDS_WIN_W + wx) * 128u + lane_off;
                        const uint32_t b1 = b0 ^ 16u;
                        constexpr uint32_t dn = MDS_WIN_W * 128u;              // next window row
                        float4 A0, A1, B0, B1, E0, E1, F0, F1;
                        lds8(b0, b1, A0, A1);
                        lds8(b0 + 128u, b1 + 128u, B0, B1);
                        lds8(b0 + dn, b1 + dn, E0, E1);
                        lds8(b0 + dn + 128u, b1 + dn + 128u, F0, F1);
                        const float mh = hh * mk, ml = lh * mk;
                        const float m00 = mh * hw, m01 = mh * lw, m10 = ml * hw, m11 = ml * lw;
                        const float2 w00 = make_float2(m00, m00), w01 = make_float2(m01, m01), w10 = make_float2(m10, m10), w11 = make_float2(m11, m11);
#define DVSR_BLEND2(lo_, hi_, A, B, E, F) __ffma2_rn(make_float2(F.lo_, F.hi_), w11, __ffma2_rn(make_float2(E.lo_, E.hi_), w10, \
                            __ffma2_rn(make_float2(B.lo_, B.hi_), w01, __fmul2_rn(make_float2(A.lo_, A.hi_), w00))))
                        vF01 = DVSR_BLEND2(x, y, A0, B0, E0, F0);
                        vF23 = DVSR_BLEND2(z, w, A0, B0, E0, F0);
                        vS01 = DVSR_BLEND2(x, y, A1, B1, E1, F1);
                        vS23 = DVSR_BLEND2(z, w, A1, B1, E1, F1);
#undef DVSR_BLEND2
                    } else {
                        // ---- exact path for samples whose corner block leaves the window (or the image by more than the margin)
                        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (cur.ok && h > -1.f && w > -1.f && h < (float)p.H && w < (float)p.W) {
                            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                            float4 q0[4], q1[4];
#pragma unroll
                            for (int cr = 0; cr < 4; ++cr) {
                                const int hc = h0 + (cr >> 1), wc = w0 + (cr & 1);
                                q0[cr] = z; q1[cr] = z;
                                if (hc >= 0 && hc <= p.H - 1 && wc >= 0 && wc <= p.W - 1)      // kernel.cu:478-490 corner rule
                                    ldg8(img0 + ((long long)hc * p.W + wc) * p.pix_stride + g * 8, q0[cr], q1[cr]);
                            }
                            const float w00 = hh * hw, w01 = hh * lw, w10 = lh * hw, w11 = lh * lw;
                            v[0] = (w00 * q0[0].x + w01 * q0[1].x + w10 * q0[2].x + w11 * q0[3].x) * mk;
                            v[1] = (w00 * q0[0].y + w01 * q0[1].y + w10 * q0[2].y + w11 * q0[3].y) * mk;
                            v[2] = (w00 * q0[0].z + w01 * q0[1].z + w10 * q0[2].z + w11 * q0[3].z) * mk;
                            v[3] = (w00 * q0[0].w + w01 * q0[1].w + w10 * q0[2].w + w11 * q0[3].w) * mk;
                            v[4] = (w00 * q1[0].x + w01 * q1[1].x + w10 * q1[2].x + w11 * q1[3].x) * mk;
                            v[5] = (w00 * q1[0].y + w01 * q1[1].y + w10 * q1[2].y + w11 * q1[3].y) * mk;
                            v[6] = (w00 * q1[0].z + w01 * q1[1].z + w10 * q1[2].z + w11 * q1[3].z) * mk;
                            v[7] = (w00 * q1[0].w + w01 * q1[1].w + w10 * q1[2].w + w11 * q1[3].w) * mk;
                        }
                        vF01 = par ? make_float2(v[4], v[5]) : make_float2(v[0], v[1]);
                        vF23 = par ? make_float2(v[6], v[7]) : make_float2(v[2], v[3]);
                        vS01 = par ? make_float2(v[0], v[1]) : make_float2(v[4], v[5]);
                        vS23 = par ? make_float2(v[2], v[3]) : make_float2(v[6], v[7]);
                    }
                    uint32_t hiF[2], loF[2], hiS[2], loS[2];
                    split_bf16x2_packed(vF01, hiF[0], loF[0]);
                    split_bf16x2_packed(vF23, hiF[1], loF[1]);
                    split_bf16x2_packed(vS01, hiS[0], loS[0]);
                    split_bf16x2_packed(vS23, hiS[1], loS[1]);
                    // even pixel: first store = hi chunk (F, S), second = lo chunk (F, S); odd pixel: first = lo chunk (S, F), second = hi chunk (S, F)
                    const uint32_t s1a = par ? loS[0] : hiF[0], s1b = par ? loS[1] : hiF[1], s1c = par ? loF[0] : hiS[0], s1d = par ? loF[1] : hiS[1];
                    const uint32_t s2a = par ? hiS[0] : loF[0], s2b = par ? hiS[1] : loF[1], s2c = par ? hiF[0] : loS[0], s2d = par ? hiF[1] : loS[1];
                    if (DVSR_MDS_OFFSMEM && tap == 8) {
                        // every lane has consumed its last (dy, dx, mask) of this chunk: the producer may overwrite the boxes
                        __syncwarp();
                        if (lane == 0) mbar_arrive(off_empty);
                    }
                    if (lane == 0) mbar_wait(&a_empty[stage], phase ^ 1);
                    __syncwarp();
                    const uint32_t sa = sa0 + (uint32_t)stage * MDS_STAGE_BYTES;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + st_first), "r"(s1a), "r"(s1b), "r"(s1c), "r"(s1d) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + st_second), "r"(s2a), "r"(s2b), "r"(s2c), "r"(s2d) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_ready[stage]);
                    if (++stage == MDS_ASTAGES) { stage = 0; phase ^= 1; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&win_empty[wb]);            // this warp is done reading the window of (tile, chunk)
                ++cc;
            }
            cur = nxt;
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int local = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            const int tn = tile / tiles_per_img, tr = tile - tn * tiles_per_img;
            const int eoy = (tr / p.tiles_w) * 16 + (row >> 3), eox = (tr % p.tiles_w) * 8 + (row & 7);
            const bool valid = eoy < p.Ho && eox < p.Wo;
            const long long pix = ((long long)tn * p.Ho + eoy) * p.Wo + eox;
            mbar_wait_relaxed(&acc_full[acc], (local >> 1) & 1, 256);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                // coalesced stores: quad transpose, then each quad writes one full 128-byte line per instruction (tc_common.cuh)
                const int rb = row & ~3;
                const int oyb = (tr / p.tiles_w) * 16 + (rb >> 3), oxb = (tr % p.tiles_w) * 8 + (rb & 7);
                const int nok = oyb < p.Ho ? min(4, p.Wo - oxb) : 0;
                const long long pixb = ((long long)tn * p.Ho + oyb) * p.Wo + oxb;
                const int i4 = lane & 3;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    // one 32-column half at a time (register budget: 80 per thread in this 22-warp CTA); the accumulator is
                    // handed back to the MMA warp as soon as its second half is in registers
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64 + h * 32), v);
                    if (h == 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(&acc_empty[acc]);
                    }
                    const int c0 = h * 32;
                    if (c0 >= p.Co) continue;
                    quad_transpose32(v, lane);
                    epilogue_store_t(v, bias_s + c0 + 8 * i4, c0 + 8 * i4, p.Co, pixb, nok, nullptr, 0, nullptr, 0, p.y, p.y_pix_stride,
                                     p.y_vec8 != 0, p.act, p.slope, 0);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

}  // namespace dvsr

using namespace dvsr;

extern "C" int dvsr_mdcn_tc_supported(const dvsr_conv_desc* d) {
    if (!d || !d->deform || d->nseg != 1 || d->transposed || d->accumulate || d->shuffle || d->res || d->out_step) return 0;
    const dvsr_conv_seg& g = d->seg[0];
    if (g.C % 32 || d->dg <= 0 || g.C != d->dg * 8) return 0;               // 8 channels (32 bytes) per deformable group
    if (d->KH * d->KW * (g.C / 32) > 18) return 0;                          // weights must fit in shared memory
    if (d->Co > 64 || d->Co < 16 || (d->Co & 3)) return 0;
    if ((g.pix_stride & 7) || (g.img_stride & 7) || ((uintptr_t)g.ptr & 31)) return 0;       // 256-bit corner loads
    if (g.T > 1 || g.t_fixed >= 0) return 0;
    if (((uintptr_t)d->y & 15) || (d->y_pix_stride & 3)) return 0;
    if (d->act == DVSR_ACT_SIGMOID_SPLIT) return 0;
    return 1;
}

// d->policy.mdcn_staged -- 0 (default): large 3x3 / stride 1 launches use the staged-window kernel; 1: always the direct-gather
// kernel; 2: staged at every size (A/B measurements, tests)

// wp: dvsr_pack_weights_tc2 mode 7 (BF16x3 rows) over the single segment
extern "C" int dvsr_mdcn_tc_fprop(const dvsr_conv_desc* d, const float* wp, void* stream) {
    DVSR_REQUIRE(d && wp && d->y, "mdcn_tc_fprop: null pointer");
    DVSR_REQUIRE(dvsr_mdcn_tc_supported(d), "mdcn_tc_fprop: unsupported shape (use dvsr_conv_fprop)");
    EncodeTiledFn encode = get_encode_tiled();
    DVSR_REQUIRE(encode != nullptr, "mdcn_tc_fprop: cuTensorMapEncodeTiled is unavailable");
    const dvsr_conv_seg& g = d->seg[0];
    MdParams p;
    memset(&p, 0, sizeof(p));
    p.x = g.ptr; p.C = g.C; p.pix_stride = g.pix_stride;
    p.img_stride = g.img_stride > 0 ? g.img_stride : (long long)d->H * d->W * g.pix_stride;
    p.N = d->N; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo;
    p.KH = d->KH; p.KW = d->KW; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil; p.dg = d->dg;
    p.offset = d->offset; p.off_pix_stride = d->off_pix_stride; p.mask = d->mask; p.mask_pix_stride = d->mask_pix_stride;
    p.Co = d->Co; p.bias = d->bias; p.act = d->act; p.slope = d->slope;
    p.y = d->y; p.y_pix_stride = d->y_pix_stride;
    p.y_vec8 = ((((uintptr_t)d->y) & 31) == 0) && (d->y_pix_stride % 8 == 0) && (d->Co % 8 == 0);
    p.nblocks = d->KH * d->KW * (g.C / 32);
    const long long M = (long long)d->N * d->Ho * d->Wo;
    p.tiles_total = (int)((M + 127) / 128);
    CUtensorMap wmap;
    {
        cuuint64_t dims[2] = {32, (cuuint64_t)p.nblocks * 64};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {32, 64};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "mdcn_tc_fprop: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    }
    const size_t smem = 1024 + (size_t)p.nblocks * 8192 + (size_t)MD_ASTAGES * MD_A_BYTES + 512;
    DVSR_REQUIRE(smem <= 232448, "mdcn_tc_fprop: %zu B of shared memory needed", smem);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(mdcn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("mdcn_tc_fprop: cudaFuncSetAttribute");
        smem_set = smem;
    }
    // staged-window variant for the EDVR geometry (3x3, stride 1, pad 1, dilation 1) from 32 tiles up.  (Round 1 kept the direct kernel
    // below two tiles per SM; measured in round 2 at the inner-loop sizes, whole GPU / 37-CTA budget: 5x44x80 35 vs 36 / 84 vs 129 us,
    // 5x88x160 98 vs 149 / 324 vs 570 us, 5x22x40 (50 tiles) 20 vs 33 / 36 vs 33 us -- tools/gpu_r2_21.sh.)
    const int staged_mode = d->policy.mdcn_staged;
    if (staged_mode != 1 && d->KH == 3 && d->KW == 3 && d->stride == 1 && d->dil == 1 && d->pad == 1 && d->Ho == d->H && d->Wo == d->W &&
        (staged_mode == 2 || (long long)d->N * ((d->Wo + 7) / 8) * ((d->Ho + 15) / 16) >= 32) &&
        (((uintptr_t)d->offset & 7) == 0) && ((d->off_pix_stride & 1) == 0)) {
        p.margin = MDS_MARGIN;
        p.win_h = MDS_WIN_H;
        p.win_w = MDS_WIN_W;
        p.win_bytes = (p.win_h * p.win_w * 128 + 1023) / 1024 * 1024;
        p.tiles_w = (d->Wo + 7) / 8;
        p.tiles_h = (d->Ho + 15) / 16;
        p.tiles_total = d->N * p.tiles_w * p.tiles_h;
        CUtensorMap xmap;
        cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)g.pix_stride * 4, (cuuint64_t)d->W * g.pix_stride * 4, (cuuint64_t)p.img_stride * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.win_w, (cuuint32_t)p.win_h, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)g.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "mdcn_tc_fprop: cuTensorMapEncodeTiled(input window) failed with %d", (int)r);
        // offsets [N,H,W,216] and masks [N,H,W,72] (views of the fused tensor): boxes of 72 / 36 channels x 8 x 16 pixels per chunk
        CUtensorMap omap, mmap;
        memset(&omap, 0, sizeof(omap)); memset(&mmap, 0, sizeof(mmap));
        bool off_ok = !DVSR_MDS_OFFSMEM;
        if (DVSR_MDS_OFFSMEM && (d->off_pix_stride & 3) == 0 && (d->mask_pix_stride & 3) == 0 &&
            (((uintptr_t)d->offset) & 15) == 0 && (((uintptr_t)d->mask) & 15) == 0) {
            cuuint32_t es4[4] = {1, 1, 1, 1};
            cuuint64_t odims[4] = {(cuuint64_t)(2 * 9 * d->dg), (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->N};
            cuuint64_t ostr[3] = {(cuuint64_t)d->off_pix_stride * 4, (cuuint64_t)d->Wo * d->off_pix_stride * 4, (cuuint64_t)d->Ho * d->Wo * d->off_pix_stride * 4};
            cuuint32_t obox[4] = {72, 8, 16, 1};
            CUresult r1 = encode(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)d->offset, odims, ostr, obox, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t mdims[4] = {(cuuint64_t)(9 * d->dg), (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->N};
            cuuint64_t mstr[3] = {(cuuint64_t)d->mask_pix_stride * 4, (cuuint64_t)d->Wo * d->mask_pix_stride * 4, (cuuint64_t)d->Ho * d->Wo * d->mask_pix_stride * 4};
            cuuint32_t mbox[4] = {36, 8, 16, 1};
            CUresult r2 = encode(&mmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)d->mask, mdims, mstr, mbox, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            off_ok = r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS;
        }
        const size_t smem_s = 1024 + (size_t)MDS_WINBUFS * p.win_bytes + (DVSR_MDS_OFFSMEM ? MDS_OFF_BYTES + MDS_MSK_BYTES : 0) +
                              (size_t)MDS_ASTAGES * MDS_STAGE_BYTES + 512;
        if (smem_s <= 232448 && off_ok) {
            static size_t smem_set_s = 0;
            if (smem_s > smem_set_s) {
                if (cudaFuncSetAttribute(mdcn_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s) != cudaSuccess)
                    return check_launch("mdcn_tc_fprop: cudaFuncSetAttribute");
                smem_set_s = smem_s;
            }
            const int ctas_s = p.tiles_total < cta_budget(d->policy) ? p.tiles_total : cta_budget(d->policy);
            mdcn_tcs_kernel<<<ctas_s, MD_THREADS, smem_s, (cudaStream_t)stream>>>(wmap, xmap, omap, mmap, p);
            return check_launch("mdcn_tc_fprop (staged)");
        }
        const long long M2 = (long long)d->N * d->Ho * d->Wo;
        p.tiles_total = (int)((M2 + 127) / 128);
    }
    int ctas = p.tiles_total < cta_budget(d->policy) ? p.tiles_total : cta_budget(d->policy);
    mdcn_tc_kernel<<<ctas, MD_THREADS, smem, (cudaStream_t)stream>>>(wmap, p);
    return check_launch("mdcn_tc_fprop");
}
