// dynavsr_b200/csrc/conv_wgrad_tc.cu
//
// Weight gradient of a stride-1 (or, by parity decomposition, stride-2) convolution on the tcgen05 tensor cores,
// straight from the NHWC tensors:
//
//   gw[co][ci][tap] += sum_{pixels p} x[p + tap][ci] * gy[p][co]
//
// is a GEMM whose reduction dimension is the PIXEL index, so both operands are "MN-major" as they lie in
// memory (channels contiguous).  Per 8 x CH-pixel chunk the producer TMA-loads ONE halo tile of x
// ((CH+KH-1) x (8+KW-1) pixels, 32-channel blocks, 32-byte-atom 128B swizzle) and the matching gy tile; every
// tap's operand is then just a shifted start address inside that halo tile (tools/umma_probe.cu P1/P2), i.e.
// x is read from L2 once instead of KH*KW times and nothing like the reference's 132.7 MB `columns` re-gather
// (deform_conv_cuda.cpp:641-666 / ATen's cudnn wgrad) exists.
//   D_tap[ci][co] (M = ci tile of 64/128, N = Co padded to 16, one K=8 MMA per image row of 8 pixels)
// accumulates in TMEM over all pixel chunks of the CTA; the epilogue adds the partial sums into the PyTorch
// -layout gradient with red.global.add.f32.  Grid = (pixel splits, tap groups, ci tiles).
// TF32 operands are used as stored (hardware truncation): a weight gradient is a leaf, its ~1e-3 relative
// error does not compound through layers.
// Stride 2 (MFDN's 4x4/s2 convs LRimg_estimator.py:79-82, fea_L{2,3}_conv1 EDVR_arch.py:229,231): tap kh = a + 2t reads
// input row 2(oy + t) + a - pad, so each parity class (a, b) is a stride-1 weight gradient with ceil((KH-a)/2) x
// ceil((KW-b)/2) taps over the 2x-subsampled input x[2i + a - pad][2j + b - pad] -- which TMA delivers directly
// (element strides 2, negative / out-of-range coordinates zero-filled).  One launch per parity class.
#include "tc_common.cuh"

namespace dvsr {

constexpr int WG_THREADS = 192;
constexpr int WG_TW = 8;                 // chunk width in pixels (one K=8 MMA per image row)

struct WgParams {
    int N, Ho, Wo;                       // gy images / size
    int KH, KW;                          // taps of this launch (a parity class of the full kernel when xs == 2)
    int xs, x_org_y, x_org_x;            // input pixel of output (oy, ox), tap (kh, kw): (oy + kh) * xs + x_org_y, ...
    int tap_a, tap_b, KWf;               // full-kernel tap of local tap (kh, kw): (tap_a + xs*kh) * KWf + tap_b + xs*kw
    int T, Tsrc, dt, t_fixed;            // source-image rule of the segment
    int C, c_tile0, Mtile;               // segment channels; first channel and size (64/128) of this CTA's M tile
    int Co, Npad;
    int CH;                              // chunk rows
    int taps_per_group;
    int chunks_total, chunks_per_cta;
    int a_blk_bytes, b_blk_bytes, stages;
    long long co_stride, ci_stride, seg_base;   // dvsr_wlayout addressing
    int ci_bits, ci_lo_valid; long long ci_hi_stride;   // interleaved channels (0 = plain)
    float* gw;
};
struct __align__(64) WgMaps { CUtensorMap x; CUtensorMap gy; };

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int mblk = p.Mtile / 32, nblk = (p.Npad + 31) / 32;
    const int a_bytes = mblk * p.a_blk_bytes, b_bytes = nblk * p.b_blk_bytes;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = (uint64_t*)(smem + p.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + p.stages;
    uint64_t* accum_bar = bars + 2 * p.stages;
    uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KK = p.KH * p.KW;
    const int tap0 = blockIdx.y * p.taps_per_group;
    const int ntaps = min(p.taps_per_group, KK - tap0);
    const int c_tile = p.c_tile0 + blockIdx.z * 128;
    const int Mt = min(p.Mtile, ((p.C - blockIdx.z * 128) + 31) / 32 * 32);   // last ci tile may be 64 (or 32-padded)
    const int chunk_begin = blockIdx.x * p.chunks_per_cta;
    const int chunk_end = min(p.chunks_total, chunk_begin + p.chunks_per_cta);
    const int tiles_w = (p.Wo + WG_TW - 1) / WG_TW, tiles_h = (p.Ho + p.CH - 1) / p.CH;

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.taps_per_group * p.Npad) tmem_cols <<= 1;

    if (warp == 0 && elect_one()) { prefetch_tmap(&maps.x); prefetch_tmap(&maps.gy); }
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
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
    const int halo_w = WG_TW + p.KW - 1;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0, phase = 0;
            for (int ch = chunk_begin; ch < chunk_end; ++ch) {
                const int n = ch / (tiles_w * tiles_h);
                const int r = ch - n * tiles_w * tiles_h;
                const int oy0 = (r / tiles_w) * p.CH, ox0 = (r % tiles_w) * WG_TW;
                const int T = p.T > 0 ? p.T : 1;
                const int q = n / T, rr = n - q * T;
                const int t = p.t_fixed >= 0 ? p.t_fixed : rr + p.dt;
                const int img = (t < 0 || t >= p.Tsrc) ? 0x3fffffff : q * p.Tsrc + t;
                mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 64);
                // a halo block is (CH+KH-1) x halo_w pixels x 128 B; TMA always delivers the full box
                mbar_expect_tx(&full_bar[stage], (uint32_t)(mblk * (p.CH + p.KH - 1) * halo_w * 128 + nblk * p.CH * WG_TW * 128));
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                for (int b = 0; b < mblk; ++b)
                    tma_load_4d(&maps.x, &full_bar[stage], sa + b * p.a_blk_bytes, c_tile + b * 32, ox0 * p.xs + p.x_org_x,
                                oy0 * p.xs + p.x_org_y, img);
                for (int b = 0; b < nblk; ++b)
                    tma_load_4d(&maps.gy, &full_bar[stage], sb + b * p.b_blk_bytes, b * 32, ox0, oy0, n);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_tf32_major(Mt == 128 ? 128 : 64, p.Npad, 1, 1);
        int stage = 0, phase = 0;
        for (int ch = chunk_begin; ch < chunk_end; ++ch) {
            mbar_wait(&full_bar[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + stage * stage_bytes), sb = sa + a_bytes;
                for (int tl = 0; tl < ntaps; ++tl) {
                    const int tap = tap0 + tl, kh = tap / p.KW, kw = tap - kh * p.KW;
                    for (int row = 0; row < p.CH; ++row) {
                        // A: 8 consecutive pixels of halo row (row + kh), starting at column kw; B: row `row` of the gy tile
                        const uint64_t ad = make_desc(sa + ((row + kh) * halo_w + kw) * 128, p.a_blk_bytes, 512, 1);
                        const uint64_t bd = make_desc(sb + row * WG_TW * 128, p.b_blk_bytes, 512, 1);
                        mma_tf32(tmem_base + tl * p.Npad, ad, bd, idesc, (ch > chunk_begin || row > 0) ? 1u : 0u);
                    }
                }
                mma_commit(&empty_bar[stage]);
                if (ch == chunk_end - 1) mma_commit(accum_bar);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (chunk_begin < chunk_end) {
        mbar_wait_relaxed(accum_bar, 0, 256);      // the epilogue warps idle for the whole main loop
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int tl_lane = q * 32 + lane;
        // accumulator row of this TMEM lane: M = 128 -> row = lane; M = 64 -> lanes 0-15 of each quarter hold rows q*16..
        int m = -1;
        if (Mt == 128) m = tl_lane;
        else if (lane < 16) m = q * 16 + lane;
        const int ci = c_tile + m;
        bool row_ok = (m >= 0) && (m < Mt) && (ci < p.C);
        long long ci_off = (long long)ci * p.ci_stride;
        if (p.ci_bits) {
            const int lo = ci & ((1 << p.ci_bits) - 1);
            row_ok = row_ok && lo < p.ci_lo_valid;
            ci_off = (long long)lo * p.ci_stride + (long long)(ci >> p.ci_bits) * p.ci_hi_stride;
        }
        for (int tl = 0; tl < ntaps; ++tl) {
            const int ltap = tap0 + tl, lkh = ltap / p.KW, lkw = ltap - lkh * p.KW;
            const int tap = (p.tap_a + p.xs * lkh) * p.KWf + p.tap_b + p.xs * lkw;
            for (int c0 = 0; c0 < p.Npad; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * p.Npad + c0), v);
                if (!row_ok) continue;
                float* dst = p.gw + p.seg_base + ci_off + tap;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = c0 + j;
                    if (co < p.Co) atomicAdd(dst + (long long)co * p.co_stride, v[j]);
                }
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

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace dvsr

using namespace dvsr;

// Split-K policy (d->policy.min_chunks): at least n pixel chunks (8 x CH pixels each) per CTA.  4 (default) favours the latency
// of one launch; larger values mean fewer CTAs and fewer partial sums per launch -- less SM-time when streams share the GPU.

// Is segment `seg` of forward descriptor `d` eligible for the tensor-core weight gradient?
extern "C" int dvsr_conv_wgrad_tc_supported(const dvsr_conv_desc* d, int seg) {
    if (!d || d->deform || d->transposed || (d->stride != 1 && d->stride != 2) || d->dil != 1) return 0;
    if (seg < 0 || seg >= d->nseg) return 0;
    if (d->Co < 16 || d->Co > 256 || (d->Co & 3)) return 0;
    const dvsr_conv_seg& g = d->seg[seg];
    if ((g.C % 64 && (g.C > 32 || (g.C & 3))) || (g.pix_stride & 3) || (g.img_stride & 3) || ((uintptr_t)g.ptr & 15)) return 0;
    if (d->KH > 5 || d->KW > 5) return 0;
    return 1;
}

extern "C" int dvsr_conv_wgrad_tc(const dvsr_conv_desc* d, int seg, const float* gy, int gy_pix_stride, float* gw,
                                  const dvsr_wlayout* wl, void* stream) {
    DVSR_REQUIRE(d && gy && gw && wl, "conv_wgrad_tc: null pointer");
    DVSR_REQUIRE(dvsr_conv_wgrad_tc_supported(d, seg), "conv_wgrad_tc: unsupported shape (use dvsr_conv_wgrad)");
    DVSR_REQUIRE((gy_pix_stride & 3) == 0 && (((uintptr_t)gy) & 15) == 0, "conv_wgrad_tc: gy must be 16-byte aligned");
    EncodeTiledFn encode = get_encode_tiled();
    DVSR_REQUIRE(encode != nullptr, "conv_wgrad_tc: cuTensorMapEncodeTiled is unavailable");
    const dvsr_conv_seg& g = d->seg[seg];
    const int xs = d->stride;
    static size_t smem_set = 0;
    for (int pa = 0; pa < xs; ++pa)
    for (int pb = 0; pb < xs; ++pb) {
    WgParams p;
    memset(&p, 0, sizeof(p));
    const int KHs = (d->KH - pa + xs - 1) / xs, KWs = (d->KW - pb + xs - 1) / xs;   // taps of this parity class
    if (KHs <= 0 || KWs <= 0) continue;
    p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.KH = KHs; p.KW = KWs;
    p.xs = xs; p.x_org_y = pa - d->pad; p.x_org_x = pb - d->pad;
    p.tap_a = pa; p.tap_b = pb; p.KWf = d->KW;
    p.T = g.T; p.Tsrc = g.Tsrc; p.dt = g.dt; p.t_fixed = g.t_fixed;
    p.C = g.C; p.c_tile0 = 0;
    p.Mtile = g.C >= 128 ? 128 : 64;
    p.Co = d->Co; p.Npad = round_up(d->Co, 16);
    const int KK = KHs * KWs;
    // TMEM budget: taps_per_group * Npad <= 512 columns
    p.taps_per_group = 512 / p.Npad;
    if (p.taps_per_group > KK) p.taps_per_group = KK;
    // balance the groups (9 taps, 8 per group max -> 5 + 4)
    const int groups = (KK + p.taps_per_group - 1) / p.taps_per_group;
    p.taps_per_group = (KK + groups - 1) / groups;
    const int mblk = p.Mtile / 32, nblk = (p.Npad + 31) / 32;
    // chunk rows: 8 if three stages fit in ~200 KB, else 4
    int CH = 8, stages = 0;
    for (;; CH = 4) {
        p.a_blk_bytes = round_up((CH + KHs - 1) * (WG_TW + KWs - 1) * 128, 1024);
        p.b_blk_bytes = CH * WG_TW * 128;
        const int stage_bytes = mblk * p.a_blk_bytes + nblk * p.b_blk_bytes;
        stages = (200 * 1024) / stage_bytes;
        if (stages >= 3 || CH == 4) break;
    }
    DVSR_REQUIRE(stages >= 2, "conv_wgrad_tc: tile does not fit in shared memory (C=%d Co=%d)", g.C, d->Co);
    if (stages > 6) stages = 6;
    p.CH = CH; p.stages = stages;
    const int tiles_w = (d->Wo + WG_TW - 1) / WG_TW, tiles_h = (d->Ho + CH - 1) / CH;
    p.chunks_total = d->N * tiles_w * tiles_h;
    const int ztiles = (g.C + 127) / 128;
    // enough pixel splits to fill the GPU once (1 CTA per SM), at least 4 chunks per CTA
    int splits = cta_budget(d->policy) / (groups * ztiles);
    if (splits < 1) splits = 1;
    int per = (p.chunks_total + splits - 1) / splits;
    const int min_chunks = d->policy.min_chunks < 1 ? 4 : d->policy.min_chunks;
    if (per < min_chunks) per = min_chunks;
    p.chunks_per_cta = per;
    splits = (p.chunks_total + per - 1) / per;
    p.co_stride = wl->co_stride; p.ci_stride = wl->ci_stride; p.seg_base = wl->seg_base[seg];
    p.ci_bits = wl->ci_bits; p.ci_lo_valid = wl->ci_lo_valid; p.ci_hi_stride = wl->ci_hi_stride;
    p.gw = gw;

    WgMaps maps;
    {
        const int T = g.T > 0 ? g.T : 1;
        long long nsrc = ((long long)(d->N + T - 1) / T) * g.Tsrc;
        if (nsrc < 1) nsrc = 1;
        long long img_stride = g.img_stride > 0 ? g.img_stride : (long long)d->H * d->W * g.pix_stride;
        cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)nsrc};
        cuuint64_t strides[3] = {(cuuint64_t)g.pix_stride * 4, (cuuint64_t)d->W * g.pix_stride * 4, (cuuint64_t)img_stride * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)((WG_TW + KWs - 1) * xs), (cuuint32_t)((CH + KHs - 1) * xs), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)xs, (cuuint32_t)xs, 1};
        CUresult r = encode(&maps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)g.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "conv_wgrad_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Co, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)gy_pix_stride * 4, (cuuint64_t)d->Wo * gy_pix_stride * 4,
                                 (cuuint64_t)d->Ho * d->Wo * gy_pix_stride * 4};
        cuuint32_t box[4] = {32, WG_TW, (cuuint32_t)CH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&maps.gy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)gy, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "conv_wgrad_tc: cuTensorMapEncodeTiled(gy) failed with %d", (int)r);
    }
    const size_t smem = 1024 + (size_t)stages * (mblk * p.a_blk_bytes + nblk * p.b_blk_bytes) + 256;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("conv_wgrad_tc: cudaFuncSetAttribute");
        smem_set = smem;
    }
    dim3 grid(splits, groups, ztiles);
    conv_wgrad_tc_kernel<<<grid, WG_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    if (int rc = check_launch("conv_wgrad_tc")) return rc;
    }
    return check_launch("conv_wgrad_tc");
}
