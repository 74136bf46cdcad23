// dynavsr_b200/csrc/conv_tc2.cu
//
// Persistent tcgen05 implicit-GEMM convolution with SHARED-MEMORY-RESIDENT weights and halo reuse -- the
// workhorse for EDVR's 64-channel 3x3 layers (feature extraction, PCD offset/feature convs, TSA, the
// reconstruction trunk, upconv / HRconv / conv_last), the 1x1 fusion convs and MFDN's stride-1 layers.
//
// conv_tc.cu (v1) re-fetches, for every 128-pixel tile, nine tap-shifted copies of the activations plus the
// whole weight tensor from L2: 432 KB per tile, which pins it at the L2 bandwidth (~5 TB/s, ~110 TFLOP/s).
// Here
//   * each CTA (one per SM, persistent over tiles) loads the weights of its 64 output channels ONCE
//     (<= 144 KiB) and keeps them in shared memory;
//   * per tile and 32-channel chunk ONE halo tile ((16+KH-1) x (8+KW-1) pixels) is TMA-loaded and converted in
//     place; the KH*KW tap operands are descriptors that start at different 128-byte rows of that tile
//     (stride-byte-offset = halo row pitch; validated by tools/umma_probe.cu P1)  => 46 KB of L2 traffic per tile;
//   * two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// Precision (dvsr_conv_tc2_set_precision):
//   BF16x3 (default) -- the split warps rewrite every 128-byte pixel row (32 fp32 channels) as [32 x bf16 hi | 32 x bf16 lo]
//     (packed cvt.rn.bf16x2); weights are packed (pack modes 9 / 10) as 16 KiB blocks per (64-channel pair, tap) whose rows
//     0-63 hold the hi parts and rows 64-127 the lo parts of the 64 output channels.  Per tap and k-step the MMA warp issues
//     x_hi . [w_hi | w_lo] as ONE N = 128 MMA (columns [0,64) and [64,128) of the accumulator) and x_lo . w_hi as an N = 64 MMA
//     into columns [0,64); the epilogue adds the two halves.  fp32-class accuracy (4.6e-6 per layer) at the speed of the
//     single-pass TF32 mode, because one read of the activation operand serves two products.
//   TF32 -- operands rounded to nearest in shared memory / at pack time (modes 5 / 6), one product (2.9e-4 per layer).
//   BF16 (mode 2) -- the BF16x3 layouts and packs unchanged, but only x_hi . w_hi is issued (N = n_mma, rows 0-63 of a weight
//     block): plain bf16 operands with fp32 accumulation, ~3e-3 per layer.  Meant for the inner adaptation steps, whose errors
//     reach the output frame attenuated by how little the adaptation moves it (profiles/r1_precision_study.md).
// The kernel is bounded by shared-memory bandwidth (both MMA operands come from shared memory: 321 KB per chunk-tile at
// 128 B/clk, see profiles/r1_conv_tc2_timeline.txt).
// Round 2 rebuilt the epilogue and the ring of this kernel and measured every variant against this one on the same box
// (tools/gpu_r2_14.sh): 8 epilogue warps + quad-transposed full-line stores (stores' share of the L1 data pipe
// 24 % -> 4 %), a dry-run warm-up pass of the epilogue, hi-only resident weights with a 6-deep ring in the bf16 mode, rotated weight
// loads, k-steps alternating between accumulator column ranges, two CTAs per SM in the bf16 mode (hi-only weights + a 1-deep ring,
// <= 96 registers: 62 vs 46 us alone, 100.2 vs 101.3 frames/s in the pool).  None beat it: BF16x3 56.5 us here vs 61-69 us, bf16 45.4 vs 45.4 us
// (3x3 64->64 @5x176x320, graph-timed), 102.5 vs 91-100 adapted frames/s.  The kernel is paced by the MMA chain's operand fetch and its
// ~10 us start-up, not by the stores, so this version stays.
// Warp roles: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2-5 = operand conversion of the halo tile,
// 6-9 = epilogue (bias, ReLU / LeakyReLU / sigmoid-split, residual, PixelShuffle(2), K-split accumulation; scalar path for
// narrow / unaligned outputs such as conv_last 64 -> 3).  Output channels are processed in groups of 64 (blockIdx.y); inputs
// whose weights do not fit are K-split over several launches by the host wrapper (`accum_in`).  Grid policy: cta_budget() CTAs
// at most and >= min_tiles tiles per CTA (throughput mode of adapt.AdaptationPool).
#include "tc_common.cuh"
#include "pack_device.cuh"

namespace dvsr {

constexpr int T2_TH = 16, T2_TW = 8;      // pixel tile (M = 128): 16 rows of 8 pixels
constexpr int T2_NG = 64;                 // output channels per CTA
constexpr int T2_ASTAGES = 3;
constexpr int T2_PF = 6;                  // L2 prefetch distance in chunks
constexpr int T2_THREADS = 320;
constexpr int T2_MAX_BLOCKS = 18;         // resident weight blocks of 64 x 32 fp32 (8 KiB) -> 144 KiB

struct T2Seg { int C, T, Tsrc, dt, t_fixed; };
struct T2Params {
    int N, Ho, Wo;
    int KH, KW, tap_sign, tap_base;       // src = o + tap_base + tap_sign * k
    int nseg, wshare;
    T2Seg seg[DVSR_MAX_SEG];
    int Co;                               // real output channels
    int halo_h, halo_w, a_bytes;          // halo tile geometry / padded bytes per stage
    int nblocks;                          // resident weight blocks of this launch: TF32 -- ((seg, chunk), tap) blocks of 64 rows
                                          // (8 KiB); BF16x3 -- ((seg, 64-channel pair), tap) blocks of 128 rows (16 KiB)
    int wblk_bytes, wblk_rows;
    int seg_blk0[DVSR_MAX_SEG];           // first weight block of each segment
    int tiles_total;
    const float* bias;
    int act; float slope; int sig_split;
    const float* res; int res_pix_stride;
    const float* accum_in; int accum_pix_stride;   // pre-activation addend (K-split partial sums)
    int shuffle;
    float* y; int y_pix_stride;
    int y_vec8;                           // y rows are 32-byte aligned: 256-bit stores
    int bf16x3;                           // 1: operands split into bf16 hi + lo, 3 products (fp32-class accuracy);
                                          // 2: same layouts, only x_hi . w_hi is issued (plain bf16 operands, fp32 accumulate)
    int n_mma;                            // MMA N (16..64): output channels of this launch's widest group, rounded up to 16
    int scalar_out;                       // narrow / unaligned outputs (conv_last 64 -> 3): scalar epilogue, Co <= 32
    long long* trace;                     // optional per-event clock64 trace of CTA (0,0): [event][chunk]
};
struct __align__(64) T2Maps { CUtensorMap x[DVSR_MAX_SEG]; CUtensorMap w; };

__device__ __forceinline__ int t2_seg_image(const T2Seg& sg, int n) {
    const int T = sg.T > 0 ? sg.T : 1;
    const int q = n / T, r = n - q * T;
    const int t = sg.t_fixed >= 0 ? sg.t_fixed : r + sg.dt;
    if (t < 0 || t >= sg.Tsrc) return -1;
    return q * sg.Tsrc + t;
}

#define T2_TRACE(ev, idx) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (idx) < 64) p.trace[(ev) * 64 + (idx)] = clock64(); } while (0)

__global__ void __launch_bounds__(T2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ T2Maps maps, const T2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem;                                        // [nblocks][wblk_rows x 128 B]
    uint8_t* smem_a = smem + p.nblocks * p.wblk_bytes;             // [T2_ASTAGES][a_bytes]
    uint64_t* bars = (uint64_t*)(smem_a + T2_ASTAGES * p.a_bytes);
    uint64_t* b_full = bars;                  // [1]
    uint64_t* a_full = bars + 1;              // [3]
    uint64_t* a_ready = bars + 4;             // [3]
    uint64_t* a_empty = bars + 7;             // [3]
    uint64_t* acc_full = bars + 10;           // [2]
    uint64_t* acc_empty = bars + 12;          // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 14);
    float* bias_s = (float*)(bars + 16);      // [64] bias of this CTA's output-channel group (zeros when absent)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KK = p.KH * p.KW;
    const int ngrp = blockIdx.y;
    const int tiles_w = (p.Wo + T2_TW - 1) / T2_TW, tiles_h = (p.Ho + T2_TH - 1) / T2_TH;
    int chunks_total = 0;
    for (int s = 0; s < p.nseg; ++s) chunks_total += (p.seg[s].C + 31) / 32;
    // window offset of tap (kh, kw) inside the halo tile
    const int min_d = p.tap_base + (p.tap_sign < 0 ? -(p.KH - 1) : 0);
    const int min_dx = p.tap_base + (p.tap_sign < 0 ? -(p.KW - 1) : 0);

    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.nseg; ++s) prefetch_tmap(&maps.x[s]);
        prefetch_tmap(&maps.w);
    }
    if (warp == 1) {
        if (elect_one()) {
            mbar_init(b_full, 1);
            for (int i = 0; i < T2_ASTAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_ready[i], 128); mbar_init(&a_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 192 && threadIdx.x < 256) {
        const int co = blockIdx.y * T2_NG + (threadIdx.x - 192);
        bias_s[threadIdx.x - 192] = (p.bias && co < p.Co) ? __ldg(p.bias + co) : 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer =====================
        if (elect_one()) {
            mbar_expect_tx(b_full, (uint32_t)(p.nblocks * p.wblk_bytes));
            for (int b = 0; b < p.nblocks; ++b)
                tma_load_2d(&maps.w, b_full, smem_b + b * p.wblk_bytes, 0, (ngrp * p.nblocks + b) * p.wblk_rows);
            int stage = 0, phase = 0, trace_i = 0;
            // L2 prefetch cursor running T2_PF chunks ahead of the shared-memory ring (the ring is only 3 deep)
            int pf_tile = blockIdx.x, pf_s = 0, pf_c = 0;
            auto prefetch_next = [&]() {
                if (pf_tile >= p.tiles_total) return;
                const int n = pf_tile / (tiles_w * tiles_h);
                const int r = pf_tile - n * tiles_w * tiles_h;
                const int img = t2_seg_image(p.seg[pf_s], n);
                if (img >= 0)
                    tma_prefetch_4d(&maps.x[pf_s], pf_c * 32, (r % tiles_w) * T2_TW + min_dx, (r / tiles_w) * T2_TH + min_d, img);
                if (++pf_c == (p.seg[pf_s].C + 31) / 32) { pf_c = 0; if (++pf_s == p.nseg) { pf_s = 0; pf_tile += gridDim.x; } }
            };
            for (int i = 0; i < T2_PF; ++i) prefetch_next();
            for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
                const int n = tile / (tiles_w * tiles_h);
                const int r = tile - n * tiles_w * tiles_h;
                const int oy0 = (r / tiles_w) * T2_TH, ox0 = (r % tiles_w) * T2_TW;
                for (int s = 0; s < p.nseg; ++s) {
                    const int img = t2_seg_image(p.seg[s], n);
                    const int chunks = (p.seg[s].C + 31) / 32;
                    for (int c = 0; c < chunks; ++c) {
                        prefetch_next();
                        mbar_wait_relaxed(&a_empty[stage], phase ^ 1, 64);
                        T2_TRACE(0, trace_i); ++trace_i;
                        mbar_expect_tx(&a_full[stage], (uint32_t)(p.halo_h * p.halo_w * 128));
                        tma_load_4d(&maps.x[s], &a_full[stage], smem_a + stage * p.a_bytes, c * 32, ox0 + min_dx, oy0 + min_d,
                                    img < 0 ? 0x3fffffff : img);
                        if (++stage == T2_ASTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = p.bf16x3 ? make_idesc_bf16(128, p.n_mma) : make_idesc_tf32(128, p.n_mma);
        const uint32_t idesc_wide = make_idesc_bf16(128, 2 * T2_NG);      // [w_hi | w_lo] stacked along N (BF16x3, full groups)
        const bool wide = p.bf16x3 == 1 && p.n_mma == T2_NG;
        const bool single = p.bf16x3 == 2;
        // descriptor templates: only the 14-bit (address >> 4) field changes per tap / k-step
        const uint64_t ad_const = make_desc(0, 16, (uint32_t)p.halo_w * 128u, 2);
        const uint64_t bd_const = make_desc(0, 16, 1024, 2);
        mbar_wait(b_full, 0);
        int stage = 0, phase = 0, local = 0, trace_i = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            mbar_wait(&acc_empty[acc], ((local >> 1) & 1) ^ 1);      // epilogue drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int cidx = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const int chunks = (p.seg[s].C + 31) / 32;
                for (int c = 0; c < chunks; ++c, ++cidx) {
                    mbar_wait(&a_ready[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        T2_TRACE(3, trace_i);
                        const uint64_t ad0 = ad_const + (uint64_t)(smem_u32(smem_a + stage * p.a_bytes) >> 4);
                        // first weight block of this chunk; BF16x3 blocks hold a PAIR of chunks (64 K-channels per row)
                        const int blk = p.seg_blk0[s] + (p.bf16x3 ? (c >> 1) : c) * KK;
                        uint64_t bd = bd_const + (uint64_t)((smem_u32(smem_b + blk * p.wblk_bytes) + (p.bf16x3 ? (c & 1) * 64 : 0)) >> 4);
                        const uint32_t dcol = tmem_base + acc * (2 * T2_NG);
                        int wy = p.tap_sign < 0 ? p.KH - 1 : 0, wx0 = p.tap_sign < 0 ? p.KW - 1 : 0, wx = wx0, kw = 0;
                        for (int tap = 0; tap < KK; ++tap) {
                            const uint64_t ad = ad0 + (uint64_t)((wy * p.halo_w + wx) * 8);     // 128 B per pixel = 8 x 16 B
                            const uint32_t first = (cidx > 0 || tap > 0) ? 1u : 0u;
                            if (!p.bf16x3) {
                                mma_tf32(dcol, ad, bd, idesc, first);
                                mma_tf32(dcol, ad + 2, bd + 2, idesc, 1u);
                                mma_tf32(dcol, ad + 4, bd + 4, idesc, 1u);
                                mma_tf32(dcol, ad + 6, bd + 6, idesc, 1u);
                            } else if (single) {
                                mma_bf16(dcol, ad, bd, idesc, first);                 // x_hi . w_hi only
                                mma_bf16(dcol, ad + 2, bd + 2, idesc, 1u);
                            } else if (wide) {
                                // activation row = [hi(32 bf16) | lo(32 bf16)]; weight rows 0-63 = hi, 64-127 = lo; K = 16 = +2
                                mma_bf16(dcol, ad, bd, idesc_wide, first);            // x_hi . [w_hi | w_lo] -> columns [0,64) | [64,128)
                                mma_bf16(dcol, ad + 2, bd + 2, idesc_wide, 1u);
                                mma_bf16(dcol, ad + 4, bd, idesc, 1u);                // x_lo . w_hi         -> columns [0,64)
                                mma_bf16(dcol, ad + 6, bd + 2, idesc, 1u);
                            } else {
                                // narrow output group (N < 64): three N-wide products; the lo rows start 64 rows (8 KiB) further
                                mma_bf16(dcol, ad, bd, idesc, first);                 // x_hi . w_hi
                                mma_bf16(dcol, ad + 2, bd + 2, idesc, 1u);
                                mma_bf16(dcol, ad + 4, bd, idesc, 1u);                // x_lo . w_hi
                                mma_bf16(dcol, ad + 6, bd + 2, idesc, 1u);
                                mma_bf16(dcol, ad, bd + 512, idesc, 1u);              // x_hi . w_lo
                                mma_bf16(dcol, ad + 2, bd + 514, idesc, 1u);
                            }
                            bd += (uint64_t)(p.wblk_bytes >> 4);                      // next tap's weight block
                            wx += p.tap_sign;
                            if (++kw == p.KW) { kw = 0; wx = wx0; wy += p.tap_sign; }
                        }
                        mma_commit(&a_empty[stage]);
                        if (cidx == chunks_total - 1) mma_commit(&acc_full[acc]);
                        T2_TRACE(4, trace_i);
                    }
                    ++trace_i;
                    __syncwarp();
                    if (++stage == T2_ASTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 6) {
        // ===================== round the halo tile to nearest TF32, in place =====================
        const int t = threadIdx.x - 64;
        const int n16 = p.halo_h * p.halo_w * 8;        // float4 elements of the halo tile
        int stage = 0, phase = 0, trace_i = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            for (int c = 0; c < chunks_total; ++c) {
                mbar_wait_relaxed(&a_full[stage], phase, 32);
                if (t == 0) T2_TRACE(1, trace_i);
                float4* a4 = reinterpret_cast<float4*>(smem_a + stage * p.a_bytes);
                if (!p.bf16x3) {
                    for (int i = t; i < n16; i += 128) {
                        float4 v = a4[i];
                        v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
                        a4[i] = v;
                    }
                } else {
                    // BF16x3: rewrite every 128-byte pixel row (32 fp32 channels) in place as [32 x bf16 hi | 32 x bf16 lo].
                    // Rows keep the 128B swizzle: logical 16-byte chunk c of row r lives at chunk c ^ ((addr >> 7) & 7).
                    const int nrows = p.halo_h * p.halo_w;
                    for (int r = t; r < nrows; r += 128) {
                        uint4* row = reinterpret_cast<uint4*>(a4 + r * 8);
                        const uint32_t ph = (smem_u32(row) >> 7) & 7;
                        float f[32];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 v = *reinterpret_cast<const float4*>(row + (c ^ ph));
                            f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
                        }
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int q2 = 0; q2 < 16; ++q2) split_bf16x2(f[2 * q2], f[2 * q2 + 1], hi[q2], lo[q2]);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            row[c ^ ph] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                            row[(4 + c) ^ ph] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&a_ready[stage]);
                if (t == 0) T2_TRACE(2, trace_i);
                ++trace_i;
                if (++stage == T2_ASTAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int local = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            const int n = tile / (tiles_w * tiles_h);
            const int r = tile - n * tiles_w * tiles_h;
            const int oy = (r / tiles_w) * T2_TH + row / T2_TW, ox = (r % tiles_w) * T2_TW + row % T2_TW;
            const bool valid = (oy < p.Ho) && (ox < p.Wo);
            const long long pix = ((long long)n * p.Ho + oy) * p.Wo + ox;
            mbar_wait_relaxed(&acc_full[acc], (local >> 1) & 1, 128);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 192) T2_TRACE(5, local);
            float v0[32], v1[32];
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * T2_NG);
            const bool wide = p.bf16x3 == 1 && p.n_mma == T2_NG;     // columns [64,128) hold the x_hi . w_lo partial sums
            tmem_ld32(tacc, v0);
            if (p.scalar_out) {
                // narrow output (Co <= 32, any alignment): one 32-column read, per-channel loads / stores
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&acc_empty[acc]);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j >= p.Co) break;
                        float t = v0[j] + bias_s[j];
                        if (p.accum_in) t += __ldg(p.accum_in + pix * p.accum_pix_stride + j);
                        t = (p.act == DVSR_ACT_SIGMOID_SPLIT) ? (j >= p.sig_split ? sigmoidf_(t) : t) : act_apply(t, p.act, p.slope);
                        if (p.res) t += __ldg(p.res + pix * p.res_pix_stride + j);
                        p.y[pix * p.y_pix_stride + j] = t;
                    }
                }
                continue;
            }
            tmem_ld32(tacc + 32, v1);
            if (wide) {
                float t[32];
                tmem_ld32(tacc + 64, t);
#pragma unroll
                for (int j = 0; j < 32; ++j) v0[j] += t[j];
                tmem_ld32(tacc + 96, t);
#pragma unroll
                for (int j = 0; j < 32; ++j) v1[j] += t[j];
            }
            // the accumulator is in registers: hand the TMEM buffer back to the MMA warp before the global stores
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[acc]);
            if (threadIdx.x == 192) T2_TRACE(7, local);
            if (valid) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float (&v)[32] = h == 0 ? v0 : v1;
                    const int c0 = ngrp * T2_NG + h * 32;          // first global output channel of this chunk
                    if (c0 >= p.Co) continue;
                    epilogue_chunk(v, c0, p.Co, bias_s + h * 32, p.accum_in ? p.accum_in + pix * p.accum_pix_stride + c0 : nullptr,
                                   p.res ? p.res + pix * p.res_pix_stride + c0 : nullptr, p.act, p.slope, p.sig_split);
                    if (p.shuffle == 2) {
                        const int cq = c0 >> 2;
#pragma unroll
                        for (int sub = 0; sub < 4; ++sub) {
                            const long long op = ((long long)n * (2 * p.Ho) + 2 * oy + (sub >> 1)) * (2 * p.Wo) + 2 * ox + (sub & 1);
                            float* yo = p.y + op * p.y_pix_stride + cq;
                            if (p.y_vec8) {
                                st_global_v8(yo, v[sub], v[4 + sub], v[8 + sub], v[12 + sub], v[16 + sub], v[20 + sub], v[24 + sub], v[28 + sub]);
                            } else {
                                *reinterpret_cast<float4*>(yo) = make_float4(v[sub], v[4 + sub], v[8 + sub], v[12 + sub]);
                                *reinterpret_cast<float4*>(yo + 4) = make_float4(v[16 + sub], v[20 + sub], v[24 + sub], v[28 + sub]);
                            }
                        }
                    } else {
                        float* yo = p.y + pix * p.y_pix_stride + c0;
                        const int nvalid = min(32, p.Co - c0);
                        if (p.y_vec8) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8)
                                if (j < nvalid) st_global_v8(yo + j, v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (j < nvalid) *reinterpret_cast<float4*>(yo + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                }
            }
            if (threadIdx.x == 192) T2_TRACE(6, local);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

}  // namespace dvsr

using namespace dvsr;

static long long* g_t2_trace = nullptr;
// Grid policy (d->policy.min_tiles): 1 (default) = one tile per CTA until the GPU is full (lowest latency of a single launch);
// n > 1 = at least n tiles per CTA (fewer, longer-lived CTAs: less SM-time per launch when several streams share the GPU).
// debugging aid: device buffer of 8 x 64 clock64 stamps written by CTA (0,0) of the next launches (nullptr = off)
extern "C" int dvsr_conv_tc2_set_trace(long long* dev_buffer) { g_t2_trace = dev_buffer; return 0; }

// kernel-side precision code (T2Params.bf16x3): 0 = TF32, 1 = BF16x3, 2 = plain bf16 on the BF16x3 layouts
static int t2_prec_code(const dvsr_conv_desc* d) {
    return d->policy.precision == DVSR_PREC_TF32 ? 0 : (d->policy.precision == DVSR_PREC_BF16 ? 2 : 1);
}
// resident weight footprint in 8 KiB units: TF32 -- one 64-row block per (32-channel chunk, tap); BF16x3 -- one 128-row
// (16 KiB) block per (64-channel pair, tap)
static int t2_blocks(const dvsr_conv_desc* d) {
    int n = 0;
    for (int s = 0; s < (d->wshare ? 1 : d->nseg); ++s)
        n += t2_prec_code(d) ? 2 * ((d->seg[s].C + 63) / 64) : (d->seg[s].C + 31) / 32;
    return n * d->KH * d->KW;
}

// 1 if the WHOLE descriptor can run in one launch of the resident-weight kernel
extern "C" int dvsr_conv_tc2_supported(const dvsr_conv_desc* d) {
    if (!d || d->deform || d->stride != 1 || d->dil != 1 || d->accumulate || d->out_step) return 0;
    const bool narrow = d->Co <= 32 && (d->Co < 16 || (d->Co & 3) || (d->y_pix_stride & 3) || ((uintptr_t)d->y & 15) ||
                                        (d->res && ((((uintptr_t)d->res) & 15) || (d->res_pix_stride & 3))));
    if (narrow && d->shuffle) return 0;
    if (!narrow && (d->Co < 16 || (d->Co & 3))) return 0;
    if (d->shuffle && (d->Co % 32)) return 0;
    if (d->KH > 5 || d->KW > 5) return 0;
    for (int s = 0; s < d->nseg; ++s) {
        const dvsr_conv_seg& g = d->seg[s];
        if ((g.C & 3) || g.C < 4 || (g.pix_stride & 3) || (g.img_stride & 3) || ((uintptr_t)g.ptr & 15)) return 0;
        if (d->wshare && g.C != d->seg[0].C) return 0;
    }
    if (!narrow) {
        if (((uintptr_t)d->y & 15) || (d->y_pix_stride & 3)) return 0;
        if (d->res && ((((uintptr_t)d->res) & 15) || (d->res_pix_stride & 3))) return 0;
    }
    return t2_blocks(d) <= T2_MAX_BLOCKS;
}

extern "C" long long dvsr_conv_tc2_packed_floats(const dvsr_wlayout* wl, int mode, int seg_lo, int seg_hi) {
    if (!wl) return 0;
    if (mode == 9) {
        long long nb = 0;
        for (int s = seg_lo; s < seg_hi; ++s) nb += (long long)wl->taps * ((wl->seg_C[s] + 63) / 64);
        return nb * ((wl->Co + T2_NG - 1) / T2_NG) * 128 * 32;
    }
    if (mode == 10) return (long long)wl->taps * ((wl->Co + 63) / 64) * ((wl->seg_C[seg_lo] + T2_NG - 1) / T2_NG) * 128 * 32;
    if (mode == 5 || mode == 7) {
        long long nb = 0;
        for (int s = seg_lo; s < seg_hi; ++s) nb += (long long)wl->taps * ((wl->seg_C[s] + 31) / 32);
        return nb * ((wl->Co + T2_NG - 1) / T2_NG) * T2_NG * 32;
    }
    return (long long)wl->taps * ((wl->Co + 31) / 32) * ((wl->seg_C[seg_lo] + T2_NG - 1) / T2_NG) * T2_NG * 32;
}

// mode 5: forward weights of segments [seg_lo, seg_hi); mode 6: data-gradient weights of segment seg_lo
extern "C" int dvsr_pack_weights_tc2(const float* w, float* wp, const dvsr_wlayout* wl, int mode, int seg_lo, int seg_hi, void* stream) {
    DVSR_REQUIRE(w && wp && wl && mode >= 5 && mode <= 10, "pack_weights_tc2: bad arguments");
    const bool fwd = (mode == 5 || mode == 7 || mode == 9);
    const int kdiv = mode >= 9 ? 64 : 32, rows = mode >= 9 ? 128 : T2_NG;
    DVSR_REQUIRE(seg_lo >= 0 && seg_lo < wl->nseg && (!fwd || (seg_hi > seg_lo && seg_hi <= wl->nseg)), "pack_weights_tc2: bad segment range");
    int nblocks, ngroups;
    if (fwd) {
        nblocks = 0;
        for (int s = seg_lo; s < seg_hi; ++s) nblocks += wl->taps * ((wl->seg_C[s] + kdiv - 1) / kdiv);
        ngroups = (wl->Co + T2_NG - 1) / T2_NG;
    } else {
        nblocks = wl->taps * ((wl->Co + kdiv - 1) / kdiv);
        ngroups = (wl->seg_C[seg_lo] + T2_NG - 1) / T2_NG;
    }
    const long long total = (long long)nblocks * ngroups * rows * 32;
    dvsr_pack_job j;
    memset(&j, 0, sizeof(j));
    j.w = w; j.wp = wp; j.wl = *wl; j.mode = mode; j.seg = seg_lo; j.seg_hi = seg_hi; j.a0 = nblocks; j.total = total;
    return dvsr_pack_job_run(&j, stream);      // the kernel lives in pack_table.cu (no cross-TU device linking)
}

extern "C" int dvsr_conv_tc2_fprop(const dvsr_conv_desc* d, const float* wp, const float* accum_in, int accum_pix_stride, void* stream) {
    DVSR_REQUIRE(d && wp && d->y, "conv_tc2_fprop: null pointer");
    DVSR_REQUIRE(dvsr_conv_tc2_supported(d), "conv_tc2_fprop: unsupported shape");
    EncodeTiledFn encode = get_encode_tiled();
    DVSR_REQUIRE(encode != nullptr, "conv_tc2_fprop: cuTensorMapEncodeTiled is unavailable");
    T2Maps maps;
    T2Params p;
    memset(&p, 0, sizeof(p));
    p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.KH = d->KH; p.KW = d->KW;
    p.tap_sign = d->transposed ? -1 : 1;
    p.tap_base = d->transposed ? d->pad : -d->pad;
    p.nseg = d->nseg; p.wshare = d->wshare;
    p.Co = d->Co;
    p.halo_h = T2_TH + d->KH - 1; p.halo_w = T2_TW + d->KW - 1;
    p.a_bytes = (p.halo_h * p.halo_w * 128 + 1023) / 1024 * 1024;
    p.bf16x3 = t2_prec_code(d);
    p.wblk_rows = p.bf16x3 ? 128 : T2_NG;
    p.wblk_bytes = p.wblk_rows * 128;
    {
        int b = 0;
        for (int s = 0; s < d->nseg; ++s) {
            const int nb = (p.bf16x3 ? (d->seg[s].C + 63) / 64 : (d->seg[s].C + 31) / 32) * d->KH * d->KW;
            p.seg_blk0[s] = d->wshare ? 0 : b;          // wshare: every segment reads segment 0's blocks
            if (!d->wshare || s == 0) b += nb;
        }
        p.nblocks = b;
    }
    p.bias = d->bias; p.act = d->act; p.slope = d->slope; p.sig_split = d->sig_split;
    p.res = d->res; p.res_pix_stride = d->res_pix_stride; p.shuffle = d->shuffle;
    p.accum_in = accum_in; p.accum_pix_stride = accum_pix_stride;
    p.y = d->y; p.y_pix_stride = d->y_pix_stride;
    p.y_vec8 = ((((uintptr_t)d->y) & 31) == 0) && (d->y_pix_stride % 8 == 0) && (d->Co % 8 == 0);
    p.trace = g_t2_trace;
    p.scalar_out = d->Co <= 32 && (d->Co < 16 || (d->Co & 3) || (d->y_pix_stride & 3) || ((uintptr_t)d->y & 15) ||
                                   (d->res && ((((uintptr_t)d->res) & 15) || (d->res_pix_stride & 3))));
    p.n_mma = d->Co >= T2_NG ? T2_NG : (d->Co + 15) / 16 * 16;
    const int tiles_w = (d->Wo + T2_TW - 1) / T2_TW, tiles_h = (d->Ho + T2_TH - 1) / T2_TH;
    p.tiles_total = d->N * tiles_w * tiles_h;
    const int ngroups = (d->Co + T2_NG - 1) / T2_NG;
    for (int s = 0; s < d->nseg; ++s) {
        const dvsr_conv_seg& g = d->seg[s];
        p.seg[s].C = g.C; p.seg[s].T = g.T; p.seg[s].Tsrc = g.Tsrc; p.seg[s].dt = g.dt; p.seg[s].t_fixed = g.t_fixed;
        const int T = g.T > 0 ? g.T : 1;
        long long nsrc = ((long long)(d->N + T - 1) / T) * g.Tsrc;
        if (nsrc < 1) nsrc = 1;
        cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)nsrc};
        long long img_stride = g.img_stride > 0 ? g.img_stride : (long long)d->H * d->W * g.pix_stride;
        cuuint64_t strides[3] = {(cuuint64_t)g.pix_stride * 4, (cuuint64_t)d->W * g.pix_stride * 4, (cuuint64_t)img_stride * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.halo_w, (cuuint32_t)p.halo_h, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&maps.x[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)g.ptr, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "conv_tc2_fprop: cuTensorMapEncodeTiled(activation seg %d) failed with %d", s, (int)r);
    }
    {
        // the packed rows are 128 bytes either way: 32 x tf32, or [32 x bf16 hi | 32 x bf16 lo]
        cuuint64_t dims[2] = {32, (cuuint64_t)p.nblocks * ngroups * p.wblk_rows};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {32, (cuuint32_t)p.wblk_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DVSR_REQUIRE(r == CUDA_SUCCESS, "conv_tc2_fprop: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    }
    const size_t smem = 1024 + (size_t)p.nblocks * p.wblk_bytes + (size_t)T2_ASTAGES * p.a_bytes + 512;
    DVSR_REQUIRE(smem <= 232448, "conv_tc2_fprop: %zu B of shared memory needed", smem);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("conv_tc2_fprop: cudaFuncSetAttribute");
        smem_set = smem;
    }
    int ctas_x = cta_budget(d->policy) / ngroups;
    if (ctas_x < 1) ctas_x = 1;
    // throughput mode: at least min_tiles tiles per CTA, so that the per-CTA fixed cost (147 KB of weights, pipeline
    // fill) is amortised and the SMs left free serve the other frames in flight (adapt.AdaptationPool)
    const int min_tiles = d->policy.min_tiles < 1 ? 1 : d->policy.min_tiles;
    const int want = (p.tiles_total + min_tiles - 1) / min_tiles;
    if (ctas_x > want) ctas_x = want;
    if (ctas_x > p.tiles_total) ctas_x = p.tiles_total;
    dim3 grid(ctas_x, ngroups);
    conv_tc2_kernel<<<grid, T2_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    return check_launch("conv_tc2_fprop");
}
