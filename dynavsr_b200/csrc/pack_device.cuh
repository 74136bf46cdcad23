// dynavsr_b200/csrc/pack_device.cuh -- one definition of every packed-weight layout (used by the single-weight
// pack entry points and by the table-driven dvsr_pack_table that re-packs a whole model in one launch).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace dvsr {

__device__ __forceinline__ float pack_round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Offset of kernel channel `ci` inside a weight row (dvsr_wlayout: plain or interleaved channels); -1 = no such element.
__device__ __forceinline__ long long wl_ci_offset(const dvsr_wlayout& wl, int ci) {
    if (wl.ci_bits == 0) return (long long)ci * wl.ci_stride;
    const int lo = ci & ((1 << wl.ci_bits) - 1), hi = ci >> wl.ci_bits;
    if (lo >= wl.ci_lo_valid) return -1;
    return (long long)lo * wl.ci_stride + (long long)hi * wl.ci_hi_stride;
}

// Element `i` of the packed buffer described by job `j` (see dvsr_pack_job in include/dvsr_b200.h).
//   mode 0 / 1 : CUDA-core layouts            (conv_simt.cu)   exact fp32
//   mode 2 / 3 : streaming tcgen05 layouts     (conv_tc.cu)     a0 = padded rows;  mode 4 = mode 3 restricted to the
//                taps (a + 2t, b + 2u): a1 = KWf, a2 = KWs, a3 = 2a + b
//   mode 5 / 6 : resident-weight tcgen05 layouts (conv_tc2.cu)  a0 = blocks per output group, seg..seg_hi = segments
//   mode 7 / 8 : same block structure as 5 / 6 for the BF16x3 split: each 128-byte row holds 32 channels as
//                [hi: 32 x bf16 | lo: 32 x bf16]; one float slot of wp carries two consecutive bf16 values (mdcn_tc.cu)
//   mode 9 / 10: BF16x3 layout of conv_tc2.cu with hi and lo parts stacked along N: 16 KiB blocks of 128 rows x 128 B per
//                (64-channel pair of K, tap); rows 0-63 = hi part of output channel g*64 + r, rows 64-127 = lo part; a row
//                holds 64 K-channels as bf16.  One N = 128 MMA then yields x_hi.w_hi and x_hi.w_lo from a single read of
//                the activation operand.  a0 = blocks per output group.
__device__ __forceinline__ float pack_value(const dvsr_pack_job& j, long long i) {
    const dvsr_wlayout& wl = j.wl;
    const float* __restrict__ w = j.w;
    if (j.mode == 0) {
        const int co = (int)(i % wl.Co);
        long long k = i / wl.Co;
        int s = 0;
        for (; s < wl.nseg; ++s) {
            const long long n = (long long)wl.seg_C[s] * wl.taps;
            if (k < n) break;
            k -= n;
        }
        const int tap = (int)(k / wl.seg_C[s]), ci = (int)(k - (long long)tap * wl.seg_C[s]);
        return w[(long long)co * wl.co_stride + wl.seg_base[s] + (long long)ci * wl.ci_stride + tap];
    }
    if (j.mode == 1) {
        const int C = wl.seg_C[j.seg];
        const int ci = (int)(i % C);
        const long long r = i / C;
        const int co = (int)(r % wl.Co), tap = (int)(r / wl.Co);
        return w[(long long)co * wl.co_stride + wl.seg_base[j.seg] + (long long)ci * wl.ci_stride + tap];
    }
    if (j.mode >= 9) {
        const int k2 = (int)(i & 31);
        long long r = i >> 5;
        const int row = (int)(r % 128);
        r /= 128;
        const int nblocks = j.a0;
        int blk = (int)(r % nblocks);
        const int g = (int)(r / nblocks);
        const int n = g * 64 + (row & 63);
        const bool want_lo = row >= 64;
        uint32_t out = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int ch = 2 * k2 + e;            // channel inside the 64-channel pair
            float v = 0.f;
            if (j.mode == 9) {
                int s = j.seg, b2 = blk;
                for (; s < j.seg_hi; ++s) {
                    const int nb = wl.taps * ((wl.seg_C[s] + 63) / 64);
                    if (b2 < nb) break;
                    b2 -= nb;
                }
                const int pair = b2 / wl.taps, tap = b2 - pair * wl.taps;
                const int ci = pair * 64 + ch;
                const long long co_ = (n < wl.Co && ci < wl.seg_C[s]) ? wl_ci_offset(wl, ci) : -1;
                if (co_ >= 0) v = w[(long long)n * wl.co_stride + wl.seg_base[s] + co_ + tap];
            } else {
                const int pair = blk / wl.taps, tap = blk - pair * wl.taps;
                const int co = pair * 64 + ch;
                if (n < wl.seg_C[j.seg] && co < wl.Co)
                    v = w[(long long)co * wl.co_stride + wl.seg_base[j.seg] + (long long)n * wl.ci_stride + tap];
            }
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
            out |= (uint32_t)__bfloat16_as_ushort(want_lo ? l : h) << (16 * e);
        }
        return __uint_as_float(out);
    }
    if (j.mode >= 7) {
        // slot k2 of the 32-slot row: elements 2*k2, 2*k2+1 of [hi(32) | lo(32)]
        const int k2 = (int)(i & 31);
        long long r = i >> 5;
        const int nblocks = j.a0;
        const int rr = (int)(r % 64);
        r /= 64;
        int blk = (int)(r % nblocks);
        const int g = (int)(r / nblocks);
        const int n = g * 64 + rr;
        uint32_t out = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int el = 2 * k2 + e;            // 0..63
            const int ch = el & 31;
            const bool want_lo = el >= 32;
            float v = 0.f;
            if (j.mode == 7) {
                int s = j.seg, b2 = blk;
                for (; s < j.seg_hi; ++s) {
                    const int nb = wl.taps * ((wl.seg_C[s] + 31) / 32);
                    if (b2 < nb) break;
                    b2 -= nb;
                }
                const int chunk = b2 / wl.taps, tap = b2 - chunk * wl.taps;
                const int ci = chunk * 32 + ch;
                if (n < wl.Co && ci < wl.seg_C[s])
                    v = w[(long long)n * wl.co_stride + wl.seg_base[s] + (long long)ci * wl.ci_stride + tap];
            } else {
                const int chunk = blk / wl.taps, tap = blk - chunk * wl.taps;
                const int co = chunk * 32 + ch;
                if (n < wl.seg_C[j.seg] && co < wl.Co)
                    v = w[(long long)co * wl.co_stride + wl.seg_base[j.seg] + (long long)n * wl.ci_stride + tap];
            }
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
            out |= (uint32_t)__bfloat16_as_ushort(want_lo ? l : h) << (16 * e);
        }
        return __uint_as_float(out);
    }
    const int k = (int)(i & 31);
    long long r = i >> 5;
    float v = 0.f;
    if (j.mode <= 4) {
        const int rows_pad = j.a0;
        const int nrow = (int)(r % rows_pad);
        r /= rows_pad;
        if (j.mode == 2) {
            int s = 0, blk = (int)r;
            for (; s < wl.nseg; ++s) {
                const int n = wl.taps * ((wl.seg_C[s] + 31) / 32);
                if (blk < n) break;
                blk -= n;
            }
            const int chunks = (wl.seg_C[s] + 31) / 32;
            const int tap = blk / chunks, chunk = blk - tap * chunks;
            if (nrow < wl.Co && chunk * 32 + k < wl.seg_C[s])
                v = w[(long long)nrow * wl.co_stride + wl.seg_base[s] + (long long)(chunk * 32 + k) * wl.ci_stride + tap];
        } else {
            const int chunks = (wl.Co + 31) / 32;
            int tap = (int)(r / chunks);
            const int chunk = (int)(r - (long long)tap * chunks);
            if (j.mode == 4) tap = ((j.a3 >> 1) + 2 * (tap / j.a2)) * j.a1 + ((j.a3 & 1) + 2 * (tap % j.a2));
            if (nrow < wl.seg_C[j.seg] && chunk * 32 + k < wl.Co)
                v = w[(long long)(chunk * 32 + k) * wl.co_stride + wl.seg_base[j.seg] + (long long)nrow * wl.ci_stride + tap];
        }
    } else {
        const int nblocks = j.a0;
        const int rr = (int)(r % 64);
        r /= 64;
        int blk = (int)(r % nblocks);
        const int g = (int)(r / nblocks);
        const int n = g * 64 + rr;
        if (j.mode == 5) {
            int s = j.seg;
            for (; s < j.seg_hi; ++s) {
                const int nb = wl.taps * ((wl.seg_C[s] + 31) / 32);
                if (blk < nb) break;
                blk -= nb;
            }
            const int chunk = blk / wl.taps, tap = blk - chunk * wl.taps;
            const int ci = chunk * 32 + k;
            const long long co_ = (n < wl.Co && ci < wl.seg_C[s]) ? wl_ci_offset(wl, ci) : -1;
            if (co_ >= 0) v = w[(long long)n * wl.co_stride + wl.seg_base[s] + co_ + tap];
        } else {
            const int chunk = blk / wl.taps, tap = blk - chunk * wl.taps;
            const int co = chunk * 32 + k;
            if (n < wl.seg_C[j.seg] && co < wl.Co)
                v = w[(long long)co * wl.co_stride + wl.seg_base[j.seg] + (long long)n * wl.ci_stride + tap];
        }
    }
    return pack_round_tf32(v);   // the MMA would truncate; weights are rounded to nearest TF32 once, here
}


}  // namespace dvsr
